#!/usr/bin/env python
"""Per-call latency of the reference-facing single-frame path (BASELINE config 1: 640x480, batch 1, reference defaults):
apriltag(family).detect(gray), TagDetector.detect(bgr) + get_pose per tag.  python tools/latency_c1.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aprilslam_b200 import synth
from aprilslam_b200.detector import Detector, TagDetector, apriltag

sc = synth.sim_settings_scene(640, 480, cam_pos=(0.0, 0.0, 0.0))
gray = synth.render(sc)
bgr = synth.gray_to_bgr(gray)
det = apriltag("tagStandard41h12")
for _ in range(20):
    r = det.detect(gray)
t0 = time.perf_counter()
N = 200
for _ in range(N):
    r = det.detect(gray)
dt = (time.perf_counter() - t0) / N
print("apriltag.detect(gray 640x480): %.3f ms per call, %d tags" % (dt * 1e3, len(r)))
td = TagDetector({"camera_matrix": sc.K, "dist_coeffs": np.zeros((4, 1))}, tag_type="tagStandard41h12", tag_size=10.0)
for _ in range(20):
    ds = td.detect(bgr)
t0 = time.perf_counter()
for _ in range(N):
    ds = td.detect(bgr)
    for d in ds:
        td.get_pose(d)
dt = (time.perf_counter() - t0) / N
print("TagDetector.detect(bgr) + get_pose x%d: %.3f ms per frame" % (len(ds), dt * 1e3))
raw = Detector("tagStandard41h12", decimate=2.0)
raw.set_profiling(True)
for _ in range(5):
    raw.detect_pose_batch(gray, sc.K, None, 10.0)
t0 = time.perf_counter()
for _ in range(N):
    raw.detect_pose_batch(gray, sc.K, None, 10.0)
dt = (time.perf_counter() - t0) / N
print("Detector.detect_pose_batch(1 frame): %.3f ms per call; stages (ms):" % (dt * 1e3), {k: round(v, 3) for k, v in raw.stage_ms().items()})
