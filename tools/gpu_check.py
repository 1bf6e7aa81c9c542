#!/usr/bin/env python
"""Stage-by-stage GPU-vs-oracle report (run on a B200 box: `python tools/gpu_check.py`).

Not a test (tests/ holds those); this prints everything in one go so that a single gpurun call
shows where the CUDA path and the oracle part ways.
"""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from aprilslam_b200 import synth
from aprilslam_b200.detector import Detector
from oracle import binding as ob

FAILS = []


def check(name, ok, info=""):
    print("[%s] %s %s" % ("ok" if ok else "FAIL", name, info), flush=True)
    if not ok:
        FAILS.append(name)


def stage_checks():
    rng = np.random.default_rng(0)
    for (W, H, d) in [(640, 480, 1), (640, 480, 2), (643, 481, 1), (1000, 1000, 2), (1001, 997, 4), (1920, 1080, 1),
                      (333, 77, 1), (64, 64, 2), (97, 131, 3)]:
        det = Detector("tag36h11", decimate=float(d))
        for kind in ("noise", "blocks", "scene"):
            if kind == "noise":
                im = rng.integers(0, 256, (H, W), dtype=np.uint8)
            elif kind == "blocks":
                im = (rng.integers(0, 2, (H // 7 + 1, W // 7 + 1), dtype=np.uint8) * 200 + 20)
                im = np.kron(im, np.ones((7, 7), np.uint8))[:H, :W].copy()
                im = (im + rng.integers(0, 6, (H, W), dtype=np.uint8)).astype(np.uint8)
            else:
                im = synth.render(synth.grid_scene(W, H, 3, (max(1, W // 200), max(1, H // 200)),
                                                   px_range=(40, 90)))
            q, t = det.stage_threshold(im)
            q_ref = im[::d, ::d]
            t_ref = ob.stage_threshold(np.ascontiguousarray(q_ref))
            check("threshold %dx%d d=%d %s" % (W, H, d, kind), np.array_equal(q, q_ref) and np.array_equal(t, t_ref),
                  "quad_im mism=%d thresh mism=%d" % ((q != q_ref).sum(), (t != t_ref).sum()))
            lab, sz = det.stage_labels(t_ref)
            lab_ref, sz_ref = ob.stage_labels(t_ref)
            check("labels %dx%d d=%d %s" % (W, H, d, kind), np.array_equal(lab, lab_ref) and np.array_equal(sz, sz_ref),
                  "label mism=%d size mism=%d" % ((lab != lab_ref).sum(), (sz != sz_ref).sum()))
        det.close()


def compare_dets(name, recs, ref, tol=0.05):
    ok = len(recs) == len(ref)
    info = "n=%d ref=%d" % (len(recs), len(ref))
    if ok and len(ref):
        ok &= np.array_equal(recs["id"], ref["id"]) and np.array_equal(recs["hamming"], ref["hamming"])
        ok &= np.array_equal(recs["family"], ref["family"])
        dc = np.abs(recs["p"] - ref["p"]).max() if ok else float("nan")
        dm = np.abs(recs["margin"] - ref["margin"]).max() if ok else float("nan")
        dH = np.abs(recs["H"] - ref["H"]).max() if ok else float("nan")
        info += " max|dcorner|=%.3g max|dmargin|=%.3g max|dH|=%.3g" % (dc, dm, dH)
        ok &= bool(dc <= tol)
    check(name, ok, info)
    if not ok and len(recs) and len(ref):
        print("   gpu ids:", recs["id"].tolist()[:60])
        print("   ref ids:", ref["id"].tolist()[:60])


def pipeline_checks():
    cases = [
        ("sim 1000x1000 41h12 d=2", synth.sim_settings_scene(1000, 1000), "tagStandard41h12", 2.0),
        ("sim 640x480 41h12 d=2", synth.sim_settings_scene(640, 480, cam_pos=(10, 2, 5)), "tagStandard41h12", 2.0),
        ("grid 1280x720 36h11 d=2", synth.grid_scene(1280, 720, 1, (5, 2), px_range=(60, 110)), "tag36h11", 2.0),
        ("grid 1920x1080 36h11 d=1", synth.grid_scene(1920, 1080, 0, (10, 5)), "tag36h11", 1.0),
        ("grid 1920x1080 36h11 d=2", synth.grid_scene(1920, 1080, 2, (10, 5)), "tag36h11", 2.0),
        ("grid 1920x1080 mixed d=1", synth.grid_scene(1920, 1080, 5, (10, 5), families=(
            ("tag25h9", range(35)), ("tagStandard41h12", range(5)))), "tag25h9 tagStandard41h12", 1.0),
        ("grid 643x481 16h5 d=1", synth.grid_scene(643, 481, 7, (3, 2), families=(("tag16h5", range(30)),),
                                                  px_range=(50, 90)), "tag16h5", 1.0),
    ]
    for name, sc, fams, d in cases:
        try:
            img = synth.render(sc)
            o = ob.OracleDetector(fams, decimate=d)
            ref, dbg = o.detect_records(img, debug=True)
            g = Detector(fams, decimate=d, debug=True)
            recs = g.detect_batch(img, cap_per_frame=256)[0]
            print("   counters:", g.counters(), "oracle: npoints=%d nclusters=%d nquads=%d ndet=%d" % (
                dbg["npoints"], len(dbg["cluster_keys"]), len(dbg["quads"]), len(ref)))
            th = g.debug_fetch("thresh")
            check(name + " thresh", np.array_equal(th, dbg["thresh"]), "mism=%d" % (th != dbg["thresh"]).sum())
            lab = g.debug_fetch("labels")
            check(name + " labels", np.array_equal(lab, dbg["labels"]), "mism=%d" % (lab != dbg["labels"]).sum())
            sz = g.debug_fetch("sizes")
            check(name + " sizes", np.array_equal(sz, dbg["sizes"]), "mism=%d" % (sz != dbg["sizes"]).sum())
            ck, cs = g.debug_fetch("cluster_keys"), g.debug_fetch("cluster_sizes")
            check(name + " clusters", np.array_equal(ck, dbg["cluster_keys"]) and np.array_equal(cs, dbg["cluster_sizes"]),
                  "n=%d ref=%d" % (len(ck), len(dbg["cluster_keys"])))
            qk, qq = g.debug_fetch("quad_keys"), g.debug_fetch("quads")
            okq = np.array_equal(qk, dbg["quad_keys"])
            dq = np.abs(qq - dbg["quads"]).max() if okq and len(qq) else (0.0 if okq else float("nan"))
            check(name + " quads", okq and dq < 1e-3, "n=%d ref=%d max|d|=%.3g" % (len(qk), len(dbg["quad_keys"]), dq))
            qr = g.debug_fetch("quads_refined")
            dr = np.abs(qr - dbg["quads_refined"]).max() if okq and len(qr) else float("nan")
            check(name + " refined", okq and dr < 1e-2, "max|d|=%.3g" % dr)
            compare_dets(name + " detections", recs, ref)
            g.close()
        except Exception:
            traceback.print_exc()
            FAILS.append(name + " (exception)")


def pose_checks():
    import cv2
    rng = np.random.default_rng(1)
    g = Detector("tag36h11")
    K = synth.intrinsics(1920, 1080, 45.0)
    for dist in (np.zeros(4), np.array([0.08, -0.15, 0.001, -0.002, 0.05])):
        corners, refs = [], []
        for i in range(400):
            sc = synth.grid_scene(1920, 1080, 100 + i, (1, 1), px_range=(40, 400), max_tilt_deg=65)
            R, t = synth.gt_pose(sc.tags[0])
            obj = np.array([[-.5, -.5, 0], [.5, -.5, 0], [.5, .5, 0], [-.5, .5, 0]]) * 0.2
            t = t * 0.2
            rv, _ = cv2.Rodrigues(R)
            pts, _ = cv2.projectPoints(obj, rv, t, K, dist)
            c = pts.reshape(4, 2) + rng.normal(0, 0.05, (4, 2))
            corners.append(c)
            refs.append(ob.reference_pose(c, K, dist.reshape(-1, 1), 0.2))
        poses = g.estimate_pose(np.array(corners), K, dist, 0.2)
        ang, dt = [], []
        for p, (ok, rvec, tvec, T) in zip(poses, refs):
            Rg = p["R"].reshape(3, 3)
            dR = Rg @ T[:3, :3].T
            ang.append(np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1)))
            dt.append(np.abs(p["tvec"] - tvec.ravel()).max())
        ang, dt = np.array(ang), np.array(dt)
        check("pose vs cv2.solvePnP dist=%s" % ("zero" if not dist.any() else "5-term"),
              bool((ang < 1e-4).all() and (dt < 1e-4).all() and poses["ok"].all()),
              "max ang=%.3g rad, max |dt|=%.3g, iters<=%d, n_bad=%d" % (ang.max(), dt.max(), poses["iters"].max(),
                                                                      int(((ang >= 1e-4) | (dt >= 1e-4)).sum())))
        p1 = g.estimate_pose(np.array(corners), K, dist, 0.2, method=1)
        ang1 = []
        for p, (ok, rvec, tvec, T) in zip(p1, refs):
            dR = p["R"].reshape(3, 3) @ T[:3, :3].T
            ang1.append(np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1)))
        print("   orthogonal iteration vs solvePnP: median ang=%.3g max=%.3g iters<=%d" % (
            np.median(ang1), np.max(ang1), p1["iters"].max()))
    g.close()


def timing():
    import torch
    frames = np.stack([synth.render(synth.grid_scene(1920, 1080, i, (10, 5))) for i in range(8)])
    t = torch.from_numpy(np.tile(frames, (8, 1, 1))).cuda()
    K = synth.intrinsics(1920, 1080, 45.0)
    for d in (1.0, 2.0):
        g = Detector("tag36h11", decimate=d)
        g.set_profiling(True)
        for it in range(3):
            torch.cuda.synchronize()
            t0 = time.time()
            dets, poses = g.detect_pose_batch(t, K, None, 0.2)
            dtm = time.time() - t0
        print("timing 1080p d=%g B=%d: %.2f ms total, %.1f frames/s, dets/frame=%.1f launches=%d" % (
            d, t.shape[0], dtm * 1e3, t.shape[0] / dtm, np.mean([len(x) for x in dets]), g.launch_count()))
        print("   stage ms:", {k: round(v, 3) for k, v in g.stage_ms().items()})
        print("   counters:", g.counters())
        g.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["stage", "pipeline", "pose", "timing"]
    for w in which:
        try:
            {"stage": stage_checks, "pipeline": pipeline_checks, "pose": pose_checks, "timing": timing}[w]()
        except Exception:
            traceback.print_exc()
            FAILS.append(w + " (exception)")
    print("FAILS:", len(FAILS), FAILS[:40])
    sys.exit(1 if FAILS else 0)
