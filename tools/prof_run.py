#!/usr/bin/env python
"""One chunk of the bench workload (24 frames of 1080p, decimate 1, detect + pose), 3 warm-up calls and one
measured call -- the command profiled under ncu (every call issues the same 30 kernel launches, so
`ncu -s 90 -c 30` captures exactly the measured call).  Usage: python tools/prof_run.py [frames] [decimate] [chunk] [slots]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from aprilslam_b200 import synth
from aprilslam_b200.detector import Detector

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
d = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else n
slots = int(sys.argv[4]) if len(sys.argv) > 4 else 1
frames = np.stack([synth.render(synth.grid_scene(1920, 1080, i, (10, 5))) for i in range(8)])
t = torch.from_numpy(np.tile(frames, ((n + 7) // 8, 1, 1))[:n]).cuda()
K = synth.intrinsics(1920, 1080, 45.0)
det = Detector("tag36h11", decimate=d, chunk_frames=chunk, pipeline_slots=slots)
det.set_profiling(True)
for it in range(4):
    torch.cuda.synchronize()
    t0 = time.time()
    dets, poses = det.detect_pose_batch(t, K, None, 0.2)
    dt = time.time() - t0
print("frames=%d chunk=%d slots=%d decimate=%g: %.3f ms (%.1f frames/s), launches=%d, tags/frame=%.1f" % (
    n, chunk, slots, d, dt * 1e3, n / dt, det.launch_count(), np.mean([len(x) for x in dets])))
print({k: round(v, 3) for k, v in det.stage_ms().items()})
