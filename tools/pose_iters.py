#!/usr/bin/env python
"""LM iteration counts of k_pose on bench-like frames (distinct scenes): python tools/pose_iters.py [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aprilslam_b200 import synth
from aprilslam_b200.detector import Detector
from aprilslam_b200.render import render_batch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
K = synth.intrinsics(1920, 1080, 45.0)
det = Detector("tag36h11", decimate=1.0)
frames = render_batch(det, [synth.grid_scene(1920, 1080, i, (10, 5)) for i in range(n)])
dets, poses = det.detect_pose_batch(frames, K, None, 0.2)
it = np.concatenate([np.asarray(p["iters"]) for p in poses])
err = np.concatenate([np.asarray(p["err"]) for p in poses])
print("tags", len(it), "iters: mean %.1f median %d p90 %d p99 %d max %d" % (it.mean(), np.median(it), np.percentile(it, 90), np.percentile(it, 99), it.max()))
print("histogram", np.bincount(it)[:40].tolist())
# per warp of 8 tags (the kernel's granularity): the slowest tag decides
w = it[: len(it) // 8 * 8].reshape(-1, 8).max(1)
print("per-warp max: mean %.1f max %d" % (w.mean(), w.max()))
