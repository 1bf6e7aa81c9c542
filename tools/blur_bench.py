#!/usr/bin/env python
"""U1 + U2 fused front end (k_decimate_blur) on the bench frames: 128 frames of 1080p, quad_sigma 0.8 / 1.5 / -0.8,
decimate 1 and 2.  Prints the kernel's own CUDA-event time, its algorithmic bytes (F*F*N_d in -- what the decimation
touches, whole 32-byte sectors for F = 2 -- plus N_d out) and the fraction of the measured copy peak.
Usage: python tools/blur_bench.py [frames]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from aprilslam_b200 import synth
from aprilslam_b200.detector import Detector

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
peak = 6550.7
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
frames = np.stack([synth.render(synth.grid_scene(1920, 1080, i, (10, 5))) for i in range(8)])
t = torch.from_numpy(np.tile(frames, ((n + 7) // 8, 1, 1))[:n]).cuda()
out = []
for d in (1.0, 2.0):
    for sigma in (0.8, 1.5, -0.8):
        det = Detector("tag36h11", decimate=d, blur=sigma, chunk_frames=n, pipeline_slots=1)
        det.set_profiling(True)
        for _ in range(4):
            lists = det.detect_batch(t)
        tab = det.kernel_table()
        ms, launches = tab["k_decimate_blur"]
        us = ms * 1e3 / max(1, launches)
        f = int(d)
        nd = (1920 // f) * (1080 // f)
        algo = (1920 * 1080 + nd) * n          # every source sector is touched at F <= 2
        rec = dict(decimate=d, sigma=sigma, us_per_launch=round(us, 1), frames=n, algorithmic_gb=round(algo / 1e9, 3),
                   achieved_gbs=round(algo / us / 1e3, 1), frac_of_measured_peak=round(algo / us / 1e3 / peak, 3),
                   tags_per_frame=float(np.mean([len(x) for x in lists])))
        print(json.dumps(rec))
        out.append(rec)
        det.close()
