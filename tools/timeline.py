#!/usr/bin/env python
"""Timeline of the chunks in flight (CUDA events on every slot's streams, ms since the start of the call):
python tools/timeline.py [frames] [slots] [chunk]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from aprilslam_b200 import synth
from aprilslam_b200._lib import STAGE_NAMES
from aprilslam_b200.detector import Detector
from aprilslam_b200.render import render_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 3
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
W, H = 1920, 1080
K = synth.intrinsics(W, H, 45.0)
det = Detector("tag36h11", decimate=1.0, chunk_frames=chunk, pipeline_slots=slots)
frames = render_batch(det, [synth.grid_scene(W, H, i, (10, 5)) for i in range(B)])
torch.cuda.synchronize()
for _ in range(3):
    det.detect_pose_batch(frames, K, None, 0.2, cap_per_frame=64)
det.set_profiling(True)
torch.cuda.synchronize()
t0 = time.perf_counter()
det.detect_pose_batch(frames, K, None, 0.2, cap_per_frame=64)
dt = time.perf_counter() - t0
tl = det.timeline()
tl = tl[np.argsort(tl[:, 0])]
print("frames=%d slots=%d: %.2f ms wall (%.0f frames/s)" % (B, slots, dt * 1e3, B / dt))
print("%5s %4s %4s | %s" % ("b0", "n", "slot", " ".join("%8s" % s[:8] for s in ["start"] + STAGE_NAMES)))
for r in tl:
    print("%5d %4d %4d | %s" % (r[0], r[1], r[2], " ".join("%8.3f" % v for v in r[3:])))
dur = np.diff(tl[:, 3:], axis=1)
print("mean stage durations (ms):", {s: round(float(v), 3) for s, v in zip(STAGE_NAMES, dur.mean(0))})
print("sum of chunk spans %.2f ms vs wall %.2f ms" % (float((tl[:, -1] - tl[:, 3]).sum()), dt * 1e3))
