#!/usr/bin/env python
"""Throughput on the OTHER named shapes of BASELINE.json (bench.py covers configs[2]): configs[1] 1280x720 batch 256
decimate 2 ~10 tags; configs[3] 3840x2160 ~200 tags (decimate 2, 64 frames per GPU); configs[4] 1080p mixed families
(tag25h9 + tagStandard41h12 ids 0-4) under the seeded blur / noise / lighting augmentation of synth.augment.
Device-resident uint8 frames in, host lists out (detect + per-tag pose), one GPU.  python tools/bench_shapes.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from aprilslam_b200 import synth
from aprilslam_b200.detector import Detector
from aprilslam_b200.render import render_batch

CASES = [
    ("C2 1280x720 b256 decimate=2 ~10 tag36h11", (1280, 720), 256, 2.0, (5, 2), "tag36h11", (("tag36h11", range(587)),), False, 64),
    ("C4 3840x2160 b64 decimate=2 ~200 tag36h11", (3840, 2160), 64, 2.0, (20, 10), "tag36h11", (("tag36h11", range(587)),), False, 256),
    ("C5 1920x1080 b256 decimate=1 mixed 25h9+41h12, augmented", (1920, 1080), 256, 1.0, (10, 5), "tag25h9 tagStandard41h12",
     (("tag25h9", range(35)), ("tagStandard41h12", range(5))), True, 64),
]
for name, (W, H), B, d, grid, fams, famspec, aug, cap in CASES:
    det = Detector(fams, decimate=d)
    K = synth.intrinsics(W, H, 45.0)
    distinct = 16 if aug else B
    frames = render_batch(det, [synth.grid_scene(W, H, 5000 + i, grid, families=famspec) for i in range(distinct)])
    if aug:
        host = frames.cpu().numpy()
        host = np.stack([synth.augment(host[i], 7000 + i) for i in range(distinct)])
        frames = torch.from_numpy(np.tile(host, (B // distinct, 1, 1))).to(frames.device)
    torch.cuda.synchronize()
    for _ in range(2):
        dets, poses = det.detect_pose_batch(frames, K, None, 0.2, cap_per_frame=cap)
    steps = 3
    t0 = time.perf_counter()
    for _ in range(steps):
        dets, poses = det.detect_pose_batch(frames, K, None, 0.2, cap_per_frame=cap)
    dt = (time.perf_counter() - t0) / steps
    det.set_profiling(True)
    det.detect_pose_batch(frames, K, None, 0.2, cap_per_frame=cap)
    stage = {k: round(v, 2) for k, v in det.stage_ms().items()}
    c = det.counters()
    print(json.dumps({"workload": name, "frames_per_s": round(B / dt, 1), "ms_per_step": round(dt * 1e3, 3),
                      "Mpx_per_s": round(B * W * H / dt / 1e6, 1), "tags_per_frame": round(float(np.mean([len(x) for x in dets])), 2),
                      "pose_ok": round(float(np.mean([p["ok"].mean() if len(p) else 1.0 for p in poses])), 4),
                      "oversize_clusters": c["oversize_clusters"], "edge_points_per_frame": round(c["edge_points"] / B),
                      "clusters_per_frame": round(c["clusters"] / B), "quads_per_frame": round(c["quads"] / B, 1),
                      "stage_ms_sum_over_chunks": stage}), flush=True)
    det.close()
    del frames
    torch.cuda.empty_cache()
