#!/usr/bin/env python
"""One 16-frame chunk of the noise regime (BASELINE configs[4]: 1080p, tag25h9 + tagStandard41h12, synth.augment), for
`ncu --metrics gpu__time_duration.sum`: python tools/prof_c5.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aprilslam_b200 import synth
from aprilslam_b200.detector import Detector
W, H, B = 1920, 1080, 16
fs = (("tag25h9", range(35)), ("tagStandard41h12", range(5)))
frames = np.stack([synth.augment(synth.render(synth.grid_scene(W, H, 5000 + i, (10, 5), families=fs)), 7000 + i) for i in range(4)])
t = torch.from_numpy(np.tile(frames, (B // 4, 1, 1))).cuda()
det = Detector("tag25h9 tagStandard41h12", decimate=1.0, chunk_frames=B, pipeline_slots=1)
K = synth.intrinsics(W, H, 45.0)
det.set_profiling(True)
for _ in range(4):
    dets, poses = det.detect_pose_batch(t, K, None, 0.2)
print("launches", det.launch_count(), {k: round(v, 3) for k, v in det.stage_ms().items()}, det.counters())
