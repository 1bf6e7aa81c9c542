#!/usr/bin/env python
"""End-to-end (pinned host frames -> host lists) throughput vs chunk size / slots: python tools/e2e_sweep.py [frames] [steps]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from aprilslam_b200 import synth
from aprilslam_b200.detector import Detector
from aprilslam_b200.render import render_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
W, H = 1920, 1080
K = synth.intrinsics(W, H, 45.0)
d0 = Detector("tag36h11", decimate=1.0)
frames = render_batch(d0, [synth.grid_scene(W, H, i, (10, 5)) for i in range(B)])
pinned = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
pinned.copy_(frames)
torch.cuda.synchronize()
del frames
d0.close()
host = pinned.numpy()
for chunk, slots in [(0, 3), (86, 3), (64, 3), (64, 4), (43, 4), (32, 4), (32, 6), (16, 8)]:
    det = Detector("tag36h11", decimate=1.0, chunk_frames=chunk, pipeline_slots=slots)
    for _ in range(2):
        det.detect_pose_batch(host, K, None, 0.2, cap_per_frame=64)
    t0 = time.perf_counter()
    for _ in range(steps):
        dets, poses = det.detect_pose_batch(host, K, None, 0.2, cap_per_frame=64)
    dt = (time.perf_counter() - t0) / steps
    det.close()
    print(json.dumps({"chunk": chunk, "slots": slots, "ms_per_step": dt * 1e3, "frames_per_s": B / dt,
                      "h2d_gbs": B * W * H / dt / 1e9}), flush=True)
