#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own wrapper code.

Needs /root/reference (read-only) and therefore only runs in the build container; the fixtures it
writes are committed so that nothing under tests/ reads /root/reference at run time.

  pose_golden.npz    corners/K/dist/tag_size -> (retval, rvec, tvec, T) from the reference's
                     TagDetector.get_pose / transformation (src/detection/tag_detector.py:30-52),
                     i.e. cv2.solvePnP + cv2.Rodrigues exactly as the reference calls them.
  detect_golden.npz  frames -> detections through the reference's TagDetector.detect
                     (tag_detector.py:23-28: cv2.cvtColor(BGR2GRAY) -> detector.detect -> sorted by id).
                     The native detector behind `from apriltag import apriltag` (tag_detector.py:11) is
                     not available offline, so the oracle stands in for it here: these vectors pin the
                     wrapper behaviour and freeze the oracle's answers (regression), they are NOT
                     outputs of upstream's binary ("parity unpinned").
  bgr2gray_golden.npz  random BGR pixels -> cv2.cvtColor(BGR2GRAY) (the call at tag_detector.py:25).
"""
import os
import sys
import types
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from aprilslam_b200 import synth  # noqa: E402
from oracle.binding import OracleDetector  # noqa: E402


def import_reference_tag_detector(**oracle_kwargs):
    """Import the reference's tag_detector.py unmodified, with an `apriltag` module whose class is the oracle."""
    stub = types.ModuleType("apriltag")

    class apriltag(OracleDetector):  # noqa: N801
        def __init__(self, family, *a, **k):
            kw = dict(oracle_kwargs)
            kw.update(k)
            super().__init__(family, *a, **kw)

    stub.apriltag = apriltag
    sys.modules["apriltag"] = stub
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    sys.modules.pop("src.detection.tag_detector", None)   # re-import so that it binds THIS stub's class
    mod = importlib.import_module("src.detection.tag_detector")
    return mod


def make_pose():
    mod = import_reference_tag_detector()
    rng = np.random.default_rng(7)
    out = {}
    sets = [("sim", synth.intrinsics(1000, 1000, 45.0), np.zeros((4, 1)), 10.0, 150, (30, 400)),
            ("webcam", np.array([[612.3, 0, 318.7], [0, 609.9, 243.1], [0, 0, 1.0]]),
             np.array([[0.11], [-0.23], [0.0012], [-0.0021], [0.09]]), 0.06, 80, (25, 200))]
    import cv2
    for name, K, dist, size, n, pxr in sets:
        td = mod.TagDetector({"camera_matrix": K, "dist_coeffs": dist}, "tag36h11", size)
        W, H = int(2 * K[0, 2]), int(2 * K[1, 2])
        corners, rv, tv, Ts, oks = [], [], [], [], []
        for i in range(n):
            sc = synth.grid_scene(W, H, 500 + i, (1, 1), px_range=pxr, max_tilt_deg=65)
            R, t = synth.gt_pose(sc.tags[0])
            obj = np.array([[-.5, -.5, 0], [.5, -.5, 0], [.5, .5, 0], [-.5, .5, 0]]) * size
            rvec, _ = cv2.Rodrigues(R)
            pts, _ = cv2.projectPoints(obj, rvec, t * size, K, dist)
            c = pts.reshape(4, 2) + rng.normal(0, 0.08, (4, 2))
            retval, r, tt, T = td.get_pose({"lb-rb-rt-lt": c})
            corners.append(c); rv.append(r.ravel()); tv.append(tt.ravel()); Ts.append(T); oks.append(retval)
        out[name + "_K"] = K; out[name + "_dist"] = dist; out[name + "_size"] = np.float64(size)
        out[name + "_corners"] = np.array(corners); out[name + "_rvec"] = np.array(rv)
        out[name + "_tvec"] = np.array(tv); out[name + "_T"] = np.array(Ts); out[name + "_ok"] = np.array(oks)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pose_golden.npz"), **out)
    print("pose_golden:", {k: np.asarray(v).shape for k, v in out.items()})


DETECT_CASES = [
    # name, scene builder, family string, decimate
    ("sim1000_41h12_d2", lambda: synth.sim_settings_scene(1000, 1000), "tagStandard41h12", 2.0),
    ("sim640_41h12_d2", lambda: synth.sim_settings_scene(640, 480, cam_pos=(10.0, 2.0, 5.0)), "tagStandard41h12", 2.0),
    ("sim640_36h11_d2", lambda: synth.sim_settings_scene(640, 480, family="tag36h11"), "tag36h11", 2.0),
    ("grid720_36h11_d2", lambda: synth.grid_scene(1280, 720, 1, (5, 2), px_range=(60, 110)), "tag36h11", 2.0),
    ("grid1080_36h11_d1", lambda: synth.grid_scene(1920, 1080, 0, (10, 5)), "tag36h11", 1.0),
    ("grid1080_mixed_d1", lambda: synth.grid_scene(1920, 1080, 5, (10, 5), families=(
        ("tag25h9", range(35)), ("tagStandard41h12", range(5)))), "tag25h9 tagStandard41h12", 1.0),
    ("grid481_16h5_d1", lambda: synth.grid_scene(643, 481, 7, (3, 2), families=(("tag16h5", range(30)),),
                                               px_range=(50, 90)), "tag16h5", 1.0),
]


def make_detect():
    out = {}
    for name, build, fams, d in DETECT_CASES:
        mod = import_reference_tag_detector(decimate=d)
        sc = build()
        gray = synth.render(sc)
        bgr = np.repeat(gray[..., None], 3, axis=2)
        td = mod.TagDetector({"camera_matrix": sc.K, "dist_coeffs": np.zeros((4, 1))}, fams, 1.0)
        dets = td.detect(bgr)
        out[name + "_frame"] = gray
        out[name + "_K"] = sc.K
        out[name + "_id"] = np.array([x["id"] for x in dets], np.int32)
        out[name + "_hamming"] = np.array([x["hamming"] for x in dets], np.int32)
        out[name + "_margin"] = np.array([x["margin"] for x in dets], np.float32)
        out[name + "_center"] = np.array([x["center"] for x in dets], np.float64).reshape(-1, 2)
        out[name + "_corners"] = np.array([x["lb-rb-rt-lt"] for x in dets], np.float64).reshape(-1, 4, 2)
        gt_ids = np.array([t.tag_id for t in sc.tags], np.int32)
        gt_c = np.array([synth.gt_corners(sc, t) for t in sc.tags])
        gt_fam = np.array([t.family for t in sc.tags])
        out[name + "_gt_id"] = gt_ids; out[name + "_gt_corners"] = gt_c; out[name + "_gt_family"] = gt_fam
        print(name, "dets", len(dets), "tags", len(sc.tags), "crc", zlib.crc32(gray.tobytes()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "detect_golden.npz"), **out)


def make_gray():
    import cv2
    rng = np.random.default_rng(3)
    bgr = rng.integers(0, 256, (64, 257, 3), dtype=np.uint8)
    bgr[0, :4] = [[128, 0, 128], [0, 0, 0], [255, 255, 255], [1, 2, 3]]
    gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bgr2gray_golden.npz"), bgr=bgr, gray=gray)


if __name__ == "__main__":
    make_pose()
    make_detect()
    make_gray()
    print("sizes:", {f: os.path.getsize(os.path.join(ROOT, "tests", "golden", f)) for f in os.listdir(
        os.path.join(ROOT, "tests", "golden"))})
