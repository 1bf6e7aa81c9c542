#!/usr/bin/env python
"""Achieved GPU-vs-oracle differences of every float quantity the parity tests bound (run on a B200 box).  The test
tolerances in tests/test_gpu_parity.py are set from this table (a few times the achieved maximum)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from aprilslam_b200 import synth
from aprilslam_b200.detector import Detector
from oracle import binding as ob

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = {
    "sim1000_41h12_d2": ("tagStandard41h12", 2.0), "sim640_41h12_d2": ("tagStandard41h12", 2.0),
    "sim640_36h11_d2": ("tag36h11", 2.0), "grid720_36h11_d2": ("tag36h11", 2.0),
    "grid1080_36h11_d1": ("tag36h11", 1.0), "grid1080_mixed_d1": ("tag25h9 tagStandard41h12", 1.0),
    "grid481_16h5_d1": ("tag16h5", 1.0),
}
acc = {k: 0.0 for k in ("quads", "quads_refined", "corners", "center", "margin", "H_rel", "tvec", "rot")}
n = {"frames": 0, "dets": 0}


def geodesic(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra @ Rb.T) - 1) / 2, -1, 1)))


detail = []


def one(fams, d, img, K=None, tag=""):
    g = Detector(fams, decimate=d, debug=True)
    if K is not None:
        dets, poses = g.detect_pose_batch(img, K, None, 0.2, cap_per_frame=256)
        recs, poses = dets[0], poses[0]
    else:
        recs, poses = g.detect_batch(img, cap_per_frame=256)[0], None
    ref, dbg = ob.OracleDetector(fams, decimate=d).detect_records(img, debug=True)
    assert recs["id"].tolist() == ref["id"].tolist(), "id mismatch"
    if len(dbg["quads"]):
        assert np.array_equal(g.debug_fetch("quad_keys"), dbg["quad_keys"]), "quad key sets differ"
        dq = np.abs(g.debug_fetch("quads") - dbg["quads"]).max(axis=1)
        dr = np.abs(g.debug_fetch("quads_refined") - dbg["quads_refined"]).max(axis=1)
        acc["quads"] = max(acc["quads"], float(dq.max()))
        acc["quads_refined"] = max(acc["quads_refined"], float(dr.max()))
        detail.append({"frame": tag, "quads": len(dq), "quads_not_bit_equal": int((dq > 0).sum()), "quads_over_1e-3": int((dq > 1e-3).sum()),
                       "max_quad": float(dq.max()), "refined_over_1e-3": int((dr > 1e-3).sum()), "max_refined": float(dr.max())})
    if len(ref):
        acc["corners"] = max(acc["corners"], float(np.abs(recs["p"] - ref["p"]).max()))
        acc["center"] = max(acc["center"], float(np.abs(recs["c"] - ref["c"]).max()))
        acc["margin"] = max(acc["margin"], float(np.abs(recs["margin"] - ref["margin"]).max()))
        acc["H_rel"] = max(acc["H_rel"], float((np.abs(recs["H"] - ref["H"]) / (np.abs(ref["H"]) + 1e-3)).max()))
    if poses is not None:
        for r, p in zip(recs, poses):
            ok, rv, tv, T = ob.reference_pose(r["p"], K, np.zeros((4, 1)), 0.2)
            acc["tvec"] = max(acc["tvec"], float(np.abs(p["tvec"] - tv.ravel()).max()))
            acc["rot"] = max(acc["rot"], geodesic(p["R"].reshape(3, 3), T[:3, :3]))
    n["frames"] += 1
    n["dets"] += len(ref)
    g.close()


det_gold = np.load(os.path.join(GOLD, "detect_golden.npz"))
for name, (fams, d) in CASES.items():
    one(fams, d, det_gold[name + "_frame"], tag=name)
for s in range(6):
    sc = synth.grid_scene(1920, 1080, s, (10, 5))
    one("tag36h11", 1.0, synth.render(sc), sc.K, tag="grid1080 seed %d" % s)
for s in range(4):
    famspec = (("tag25h9", range(35)), ("tagStandard41h12", range(5)))
    im = synth.augment(synth.render(synth.grid_scene(1920, 1080, 300 + s, (10, 5), families=famspec, px_range=(60, 110))), 900 + s)
    one("tag25h9 tagStandard41h12", 1.0, im, tag="augmented 1080p %d" % s)
for d in (1.0, 2.0):
    sc = synth.grid_scene(3840, 2160, 77, (20, 10))
    one("tag36h11", d, synth.render(sc), sc.K, tag="4K d=%g" % d)
for dd in detail:
    print(json.dumps(dd))
print(json.dumps({"achieved_max_abs_diff_gpu_vs_oracle": acc, "frames": n["frames"], "detections": n["dets"]}))
