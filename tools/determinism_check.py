#!/usr/bin/env python
"""Run-to-run determinism of the whole pipeline: the same batch (clean 1080p frames, augmented frames, noise frames)
through detect + pose `reps` times, with one and with three chunks in flight; every run must give BIT-IDENTICAL
detection and pose records.  Record emission order, cluster ids, dense component ids and work-list order are all
decided by atomics and differ from run to run -- the results may not (the quad fit sorts by a total order, reductions
that feed decisions run in a fixed order).  A race in one of the lock-free kernels would show up here as a flipped bit.
python tools/determinism_check.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aprilslam_b200 import synth
from aprilslam_b200.detector import Detector

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
K = synth.intrinsics(1920, 1080, 45.0)
rng = np.random.default_rng(3)
clean = [synth.render(synth.grid_scene(1920, 1080, 100 + i, (10, 5))) for i in range(24)]
aug = [synth.augment(synth.render(synth.grid_scene(1920, 1080, 300 + i, (10, 5), px_range=(60, 110))), 900 + i) for i in range(16)]
noise = [rng.integers(0, 256, (1080, 1920), dtype=np.uint8) for _ in range(2)] + \
        [np.kron(rng.integers(0, 2, (135, 240), dtype=np.uint8) * 200 + 25, np.ones((8, 8), np.uint8)).astype(np.uint8) for _ in range(2)]
frames = np.stack(clean + aug + noise)
bad = 0
for slots, chunk in ((1, 0), (3, 8), (3, 0)):
    det = Detector("tag36h11", decimate=1.0, pipeline_slots=slots, chunk_frames=chunk)
    ref = None
    for r in range(reps):
        d, p = det.detect_pose_batch(frames, K, None, 0.2, cap_per_frame=128)
        cur = (b"".join(np.asarray(x).tobytes() for x in d), b"".join(np.asarray(x).tobytes() for x in p))
        if ref is None:
            ref, ndet = cur, sum(len(x) for x in d)
        elif cur != ref:
            bad += 1
            print("DIFFERENT RESULT in run %d (slots %d, chunk %d)" % (r, slots, chunk))
    print("slots=%d chunk=%d: %d runs of %d frames, %d detections: %s" % (slots, chunk, reps, len(frames), ndet,
                                                                          "bit-identical" if bad == 0 else "MISMATCH"))
    det.close()
print("NON-DETERMINISTIC RUNS: %d" % bad)
