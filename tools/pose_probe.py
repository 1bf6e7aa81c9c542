#!/usr/bin/env python
"""k_pose probe: iteration counts and time per pose on detected corners of the bench workload (run on a B200 box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aprilslam_b200 import synth
from aprilslam_b200.detector import Detector

frames = np.stack([synth.render(synth.grid_scene(1920, 1080, i, (10, 5))) for i in range(8)])
K = synth.intrinsics(1920, 1080, 45.0)
det = Detector("tag36h11", decimate=1.0)
dets, poses = det.detect_pose_batch(frames, K, None, 0.2)
it = np.concatenate([p["iters"] for p in poses])
print("poses", len(it), "iters: min %d median %d mean %.1f max %d" % (it.min(), np.median(it), it.mean(), it.max()), np.bincount(it)[:40])
corners = np.concatenate([d["p"] for d in dets])
big = np.tile(corners, (16, 1, 1))
for _ in range(3):
    t0 = time.perf_counter()
    out = det.estimate_pose(big, K, None, 0.2)
    dt = time.perf_counter() - t0
print("agpu_pose: %d poses in %.3f ms (incl. copies)" % (len(big), dt * 1e3))
