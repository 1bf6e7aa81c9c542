#!/usr/bin/env python
"""GPU vs oracle on augmented (blur / noise / lighting) frames of every family and on a 4K frame."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aprilslam_b200 import synth
from aprilslam_b200.detector import Detector
from oracle import binding as ob

bad = 0
def cmp(name, recs, ref):
    global bad
    ok = len(recs) == len(ref) and np.array_equal(recs["id"], ref["id"]) and np.array_equal(recs["hamming"], ref["hamming"]) \
        and np.array_equal(recs["family"], ref["family"])
    dc = float(np.abs(recs["p"] - ref["p"]).max()) if ok and len(ref) else 0.0
    dm = float(np.abs(recs["margin"] - ref["margin"]).max()) if ok and len(ref) else 0.0
    ok = ok and dc <= 0.05
    if not ok:
        bad += 1
        print("MISMATCH", name, "n=%d ref=%d" % (len(recs), len(ref)), "dc=%.3g" % dc)
        print("   gpu", recs["id"].tolist(), recs["hamming"].tolist())
        print("   ref", ref["id"].tolist(), ref["hamming"].tolist())
    return dc, dm

cases = [("tag36h11", (("tag36h11", range(587)),), 1.0, (1920, 1080), (10, 5)),
         ("tag36h11", (("tag36h11", range(587)),), 2.0, (1280, 720), (5, 2)),
         ("tag25h9 tagStandard41h12", (("tag25h9", range(35)), ("tagStandard41h12", range(5))), 1.0, (1920, 1080), (10, 5)),
         ("tag16h5", (("tag16h5", range(30)),), 2.0, (1280, 720), (6, 3))]
nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for fams, famspec, d, (W, H), grid in cases:
    g = Detector(fams, decimate=d)
    o = ob.OracleDetector(fams, decimate=d)
    frames = np.stack([synth.augment(synth.render(synth.grid_scene(W, H, 300 + s, grid, families=famspec,
                                                                     px_range=(60, 110))), 900 + s) for s in range(nframes)])
    t0 = time.time(); dets = g.detect_batch(frames, cap_per_frame=128); tg = time.time() - t0
    mdc = mdm = 0.0; nd = 0
    for b in range(nframes):
        ref = o.detect_records(frames[b])
        dc, dm = cmp("%s d=%g %dx%d frame %d" % (fams, d, W, H, b), dets[b], ref)
        mdc = max(mdc, dc); mdm = max(mdm, dm); nd += len(ref)
    print("[%s d=%g %dx%d] frames=%d dets=%d max|dcorner|=%.3g max|dmargin|=%.3g gpu %.1f ms" % (fams, d, W, H, nframes, nd, mdc, mdm, tg * 1e3))
    g.close()
# 4K (BASELINE configs[3]): ~200 tags
for d in (2.0, 1.0):
    sc = synth.grid_scene(3840, 2160, 77, (20, 10))
    img = synth.render(sc)
    g = Detector("tag36h11", decimate=d)
    recs = g.detect_batch(img, cap_per_frame=256)[0]
    ref = ob.OracleDetector("tag36h11", decimate=d).detect_records(img)
    dc, dm = cmp("4K d=%g" % d, recs, ref)
    print("[4K d=%g] dets=%d ref=%d tags=%d max|dcorner|=%.3g" % (d, len(recs), len(ref), len(sc.tags), dc))
    g.close()
# noise-only and heavy-noise frames
rng = np.random.default_rng(5)
g = Detector("tag36h11", decimate=1.0); o = ob.OracleDetector("tag36h11", decimate=1.0)
for k in range(4):
    im = rng.integers(0, 256, (720, 1280), dtype=np.uint8) if k < 2 else \
        (np.kron(rng.integers(0, 2, (90, 160), dtype=np.uint8) * 200 + 25, np.ones((8, 8), np.uint8))).astype(np.uint8)
    t0 = time.time(); recs = g.detect_batch(im, cap_per_frame=256)[0]; tg = time.time() - t0
    ref = o.detect_records(im)
    cmp("noise %d" % k, recs, ref)
    print("[noise %d] dets=%d ref=%d counters=%s gpu %.1f ms" % (k, len(recs), len(ref), g.counters(), tg * 1e3))
print("MISMATCHES:", bad)
sys.exit(1 if bad else 0)
