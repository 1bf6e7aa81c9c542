import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    from aprilslam_b200 import synth
    from aprilslam_b200.detector import Detector
    n = int(sys.argv[2])
    frames = np.stack([synth.render(synth.grid_scene(1920, 1080, i, (10, 5))) for i in range(4)])
    t = torch.from_numpy(np.tile(frames, ((n + 3) // 4, 1, 1))[:n]).cuda()
    det = Detector("tag36h11", decimate=1.0, chunk_frames=n, pipeline_slots=1)
    det.set_profiling(True)
    best = 1e9
    for it in range(8):
        det.detect_batch(t)
        best = min(best, det.stage_ms()["image"])
    N = 1920 * 1080
    print(json.dumps({"frames": n, "minb": os.environ.get("AGPU_IMG_MINB"), "seg": os.environ.get("AGPU_SEG_TILES"), "image_ms": round(best, 4), "GBs": round(2 * N * n / (best / 1e3) / 1e9, 1)}))
else:
    for n in (64, 128):
        for minb in (3, 4, 5, 6):
            for seg in (8, 16):
                env = dict(os.environ, AGPU_SEG_TILES=str(seg), AGPU_IMG_MINB=str(minb))
                out = subprocess.run([sys.executable, __file__, "child", str(n)], env=env, capture_output=True, text=True)
                print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)
