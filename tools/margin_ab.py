#!/usr/bin/env python
"""Largest decision-margin / corner differences GPU vs oracle on the 24 augmented 1080p frames of tools/stress_check.py,
per detection.  AGPU_LIB=<another build of libaprilgpu.so> runs the same frames through that build (A/B check: the values
must not move when a kernel is restructured).  python tools/margin_ab.py"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from aprilslam_b200 import synth
from aprilslam_b200.detector import Detector
from oracle import binding as ob
famspec = (("tag36h11", range(587)),)
frames = np.stack([synth.augment(synth.render(synth.grid_scene(1920, 1080, 300 + s, (10, 5), families=famspec, px_range=(60, 110))), 900 + s) for s in range(24)])
g = Detector("tag36h11", decimate=1.0)
o = ob.OracleDetector("tag36h11", decimate=1.0)
dets = g.detect_batch(frames, cap_per_frame=128)
worst = []
for b in range(24):
    ref = o.detect_records(frames[b])
    r = dets[b]
    assert len(r) == len(ref)
    dm = np.abs(r["margin"] - ref["margin"]); dc = np.abs(r["p"] - ref["p"]).reshape(len(ref), -1).max(1)
    for i in np.argsort(-dm)[:2]:
        worst.append((float(dm[i]), b, int(ref["id"][i]), float(dc[i]), float(r["margin"][i]), float(ref["margin"][i])))
worst.sort(reverse=True)
print(os.environ.get("AGPU_LIB", "HEAD"), worst[:4])
