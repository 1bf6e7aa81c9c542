#!/usr/bin/env python
"""Scheduling sweep on the bench workload (1080p, ~50 tags/frame, frames rendered on the GPU): stream priorities,
persistent CTAs per SM of the quad-fit tiers / decode, chunk size and pipeline slots.
Usage: python tools/sweep_sched.py [frames] [steps] [configs.json]"""
import json
import os
import sys
import time

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

CONFIGS = [
    # name, env, chunk, slots
    ("base_noprio", {"AGPU_PRIO": "0"}, 0, 3),
    ("prio", {"AGPU_PRIO": "1"}, 0, 3),
    ("prio_t2-8-4-1_d2", {"AGPU_PRIO": "1", "AGPU_TIER_CTAS": "2,8,4,1", "AGPU_DECODE_CTAS": "2"}, 0, 3),
    ("prio_t1-4-2-1_d1", {"AGPU_PRIO": "1", "AGPU_TIER_CTAS": "1,4,2,1", "AGPU_DECODE_CTAS": "1"}, 0, 3),
    ("prio_t2-6-3-1_d2_s4", {"AGPU_PRIO": "1", "AGPU_TIER_CTAS": "2,6,3,1", "AGPU_DECODE_CTAS": "2"}, 0, 4),
    ("prio_c64_s4", {"AGPU_PRIO": "1"}, 64, 4),
    ("prio_c64_s6_t2-8-4-1_d2", {"AGPU_PRIO": "1", "AGPU_TIER_CTAS": "2,8,4,1", "AGPU_DECODE_CTAS": "2"}, 64, 6),
    ("noprio_c64_s6", {"AGPU_PRIO": "0"}, 64, 6),
    ("prio_c32_s8_t2-8-4-1_d2", {"AGPU_PRIO": "1", "AGPU_TIER_CTAS": "2,8,4,1", "AGPU_DECODE_CTAS": "2"}, 32, 8),
    ("prio_s2", {"AGPU_PRIO": "1"}, 0, 2),
]
if len(sys.argv) > 3:
    CONFIGS = [tuple(c) for c in json.load(open(sys.argv[3]))]

d0 = Detector("tag36h11", decimate=1.0)
frames = render_batch(d0, [synth.grid_scene(W, H, i, (10, 5)) for i in range(B)])
torch.cuda.synchronize()
d0.close()
results = []
for name, env, chunk, slots in CONFIGS:
    for k in ("AGPU_PRIO", "AGPU_TIER_CTAS", "AGPU_DECODE_CTAS", "AGPU_TAIL_THREADS", "AGPU_EDGE_WARPS", "AGPU_BOUNDARY_WARPS", "AGPU_TIER_CAP", "AGPU_MASKS", "AGPU_SLOTS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    det = Detector("tag36h11", decimate=1.0, chunk_frames=chunk, pipeline_slots=slots)
    for _ in range(2):
        dets, poses = det.detect_pose_batch(frames, K, None, 0.2, cap_per_frame=64)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        dets, poses = det.detect_pose_batch(frames, K, None, 0.2, cap_per_frame=64)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    ntags = float(np.mean([len(x) for x in dets]))
    det.close()
    r = {"name": name, "env": env, "chunk": chunk, "slots": slots, "ms_per_step": dt * 1e3, "frames_per_s": B / dt,
         "tags_per_frame": ntags}
    results.append(r)
    print(json.dumps(r), flush=True)
best = max(results, key=lambda r: r["frames_per_s"])
print("BEST", json.dumps(best))
