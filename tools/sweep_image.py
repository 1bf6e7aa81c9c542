#!/usr/bin/env python
"""Image-stage sweep on a B200: stage parity spot check + achieved GB/s vs chunk size / segment height."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    from aprilslam_b200 import synth
    from aprilslam_b200.detector import Detector
    n = int(sys.argv[2]); d = float(sys.argv[3])
    frames = np.stack([synth.render(synth.grid_scene(1920, 1080, i, (10, 5))) for i in range(4)])
    t = torch.from_numpy(np.tile(frames, ((n + 3) // 4, 1, 1))[:n]).cuda()
    det = Detector("tag36h11", decimate=d, chunk_frames=n)
    det.set_profiling(True)
    best = 1e9
    for it in range(6):
        det.detect_batch(t)
        best = min(best, det.stage_ms()["image"])
    N = 1920 * 1080
    alg = (N + N) if d == 1 else (N // int(d) + 2 * (N // int(d) ** 2))
    print(json.dumps({"frames": n, "decimate": d, "seg": os.environ.get("AGPU_SEG_TILES", "8"), "image_ms": best,
                      "us_per_frame": best * 1e3 / n, "GBs": alg * n / (best / 1e3) / 1e9}))
else:
    for d in (1.0, 2.0):
        for n in (24, 64, 128):
            for seg in (4, 8, 16, 34):
                env = dict(os.environ, AGPU_SEG_TILES=str(seg))
                out = subprocess.run([sys.executable, __file__, "child", str(n), str(d)], env=env, capture_output=True, text=True)
                print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:], flush=True)
