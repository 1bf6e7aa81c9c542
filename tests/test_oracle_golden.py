"""CPU tests: the oracle against every pin that exists for this path.

The reference has no tests and its native detector is an un-vendored, unpinned dependency
("parity unpinned"), so the pins are: code-book known-answer words, fixtures produced by the
reference's own wrapper code (tools/make_golden.py), analytic ground truth of the reference's
renderer geometry, a cv2.aruco cross-check of ids, and the reference's committed run log / CSV.
"""
import os

import numpy as np
import pytest

from aprilslam_b200 import synth
from aprilslam_b200.families_data import FAMILIES
from oracle import binding as ob

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DET = np.load(os.path.join(GOLD, "detect_golden.npz"))
POSE = np.load(os.path.join(GOLD, "pose_golden.npz"))

CASES = {  # name -> (families, decimate)    (tools/make_golden.py DETECT_CASES)
    "sim1000_41h12_d2": ("tagStandard41h12", 2.0), "sim640_41h12_d2": ("tagStandard41h12", 2.0),
    "sim640_36h11_d2": ("tag36h11", 2.0), "grid720_36h11_d2": ("tag36h11", 2.0),
    "grid1080_36h11_d1": ("tag36h11", 1.0), "grid1080_mixed_d1": ("tag25h9 tagStandard41h12", 1.0),
    "grid481_16h5_d1": ("tag16h5", 1.0),
}


def test_codebook_known_answers():
    # SURVEY.md 8a family table: first code words published with the upstream families
    assert FAMILIES["tag36h11"]["codes"][:3] == [0xd7e00984b, 0xdda664ca7, 0xdc4a1c821]
    assert FAMILIES["tag25h9"]["codes"][:3] == [0x156f1f4, 0x1f28cd5, 0x16ce32c]
    assert FAMILIES["tag16h5"]["codes"][:3] == [0x27c8, 0x31b6, 0x3859]
    assert FAMILIES["tagStandard41h12"]["codes"] == [0x1bd8a64ad10, 0x1bdc4f3b2d5, 0x1bdff82b89a, 0x1be3a11be5f,
                                                      0x1be74a0c424]
    assert [len(FAMILIES[f]["codes"]) for f in ("tag36h11", "tag25h9", "tag16h5")] == [587, 35, 30]
    # 41h12 code words advance by a constant
    c = FAMILIES["tagStandard41h12"]["codes"]
    assert all(b - a == 982451653 for a, b in zip(c, c[1:]))


def test_codebook_min_hamming_distance():
    for name in ("tag36h11", "tag25h9", "tag16h5"):
        f = FAMILIES[name]
        codes = np.array(f["codes"], np.uint64)
        rots = [codes]
        for _ in range(3):
            rots.append(np.array([ob.lib().ao_rotate90(int(c), f["nbits"]) for c in rots[-1]], np.uint64))
        best = 64
        for r, rc in enumerate(rots):
            x = codes[:, None] ^ rc[None, :]
            pc = np.zeros(x.shape, np.int32)
            for b in range(f["nbits"]):
                pc += ((x >> np.uint64(b)) & np.uint64(1)).astype(np.int32)
            if r == 0:
                pc[np.arange(len(codes)), np.arange(len(codes))] = 64
            best = min(best, int(pc.min()))
        assert best >= f["h"], (name, best)


def test_rotate90_is_a_quarter_turn_of_the_bit_layout():
    for name, f in FAMILIES.items():
        nb, wb = f["nbits"], f["width_at_border"]
        code = f["codes"][0]
        cells = {(x, y): (code >> (nb - 1 - i)) & 1 for i, (x, y) in enumerate(zip(f["bit_x"], f["bit_y"]))}
        r = ob.lib().ao_rotate90(code, nb)
        rot = {(x, y): (r >> (nb - 1 - i)) & 1 for i, (x, y) in enumerate(zip(f["bit_x"], f["bit_y"]))}
        # the rotated word holds, at cell (x, y), the bit the original had one quadrant further along the
        # spiral, i.e. at (wb-1-y, x): a quarter turn of the whole bit layout
        assert all(rot[(x, y)] == cells[(wb - 1 - y, x)] for (x, y) in cells), name
        w = code
        for _ in range(4):
            w = ob.lib().ao_rotate90(w, nb)
        assert w == code


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_wrapper_fixture(name):
    """Frames -> detections as the reference's TagDetector.detect returned them (tag_detector.py:23-28)."""
    fams, d = CASES[name]
    recs = ob.OracleDetector(fams, decimate=d).detect_records(DET[name + "_frame"])
    assert recs["id"].tolist() == DET[name + "_id"].tolist()          # bit-exact, sorted by id
    assert recs["hamming"].tolist() == DET[name + "_hamming"].tolist()
    assert np.array_equal(recs["p"], DET[name + "_corners"])
    assert np.array_equal(recs["c"], DET[name + "_center"])
    assert np.array_equal(recs["margin"], DET[name + "_margin"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_against_analytic_ground_truth(name):
    """Ids and corners against the renderer geometry (renderer.py:91-96,188-251)."""
    fams, d = CASES[name]
    fam_list = fams.split()
    recs = ob.OracleDetector(fams, decimate=d).detect_records(DET[name + "_frame"])
    gt_id, gt_c, gt_f = DET[name + "_gt_id"], DET[name + "_gt_corners"], DET[name + "_gt_family"]
    Himg, Wimg = DET[name + "_frame"].shape
    def side(c):
        return float(np.linalg.norm(c[1] - c[0]))

    # a tag must be found when it is fully inside the frame (Standard families carry data 2 cells outside the
    # border) and its border is at least 30 working (decimated) pixels wide
    visible = [(int(i), str(f), c) for i, f, c in zip(gt_id, gt_f, gt_c)
               if side(c) >= 30 * d and c[:, 0].min() > 0.5 * side(c) and c[:, 0].max() < Wimg - 0.5 * side(c)
               and c[:, 1].min() > 0.5 * side(c) and c[:, 1].max() < Himg - 0.5 * side(c)]
    found = {(int(r["id"]), fam_list[int(r["family"])]): r for r in recs}
    assert len(found) == len(recs)
    for tid, fam, c in visible:
        assert (tid, fam) in found, (name, tid, fam)
        r = found[(tid, fam)]
        assert r["hamming"] == 0
        # aliased (non-antialiased) rasterisation moves an edge by up to half a pixel
        assert np.abs(r["p"] - c).max() < 0.75, (name, tid, np.abs(r["p"] - c).max())
    assert len(recs) >= len(visible) and len(visible) >= 1
    gt_keys = {(int(i), str(f)) for i, f in zip(gt_id, gt_f)}
    assert set(found) <= gt_keys          # no false positives


def test_reference_pose_restatement_equals_reference_outputs():
    """oracle.binding.reference_pose == the reference's TagDetector.get_pose (tag_detector.py:30-52)."""
    for s in ("sim", "webcam"):
        K, dist, size = POSE[s + "_K"], POSE[s + "_dist"], float(POSE[s + "_size"])
        for c, rv, tv, T, ok in zip(POSE[s + "_corners"], POSE[s + "_rvec"], POSE[s + "_tvec"], POSE[s + "_T"],
                                    POSE[s + "_ok"]):
            retval, rvec, tvec, TT = ob.reference_pose(c, K, dist, size)
            assert bool(retval) == bool(ok)
            assert np.allclose(rvec.ravel(), rv, atol=1e-9) and np.allclose(tvec.ravel(), tv, atol=1e-9)
            assert np.allclose(TT, T, atol=1e-9)


def test_soft_vectors_from_the_reference_run_log():
    """data/logs/simulation_runner.log:26-27 -- camera at the origin of config/sim_settings.json:
    'Tag ID 1 ... translation length = 76.34146240389457', 'Tag ID 2 ... = 45.47295093668058'
    (world frame = tag 0, src/core/slam_graph.py:45-49).  A different rasteriser (OpenGL) produced those
    frames, so agreement is soft: 1 % (the log itself is 0.24 % / 0.17 % off the analytic 76.1577 / 45.5522)."""
    sc = synth.sim_settings_scene(1000, 1000)
    img = synth.render(sc)
    recs = ob.OracleDetector("tagStandard41h12", decimate=2.0).detect_records(img)
    assert recs["id"].tolist() == [0, 1, 2]       # tags 3 and 4 are outside the 1000x1000 view from the origin
    T = {int(r["id"]): ob.reference_pose(r["p"], sc.K, np.zeros((4, 1)), 10.0)[3] for r in recs}
    for tid, logged, analytic in ((1, 76.34146240389457, 76.15773105863909), (2, 45.47295093668058, 45.55216789572032)):
        Wt = np.linalg.inv(T[0]) @ T[tid]
        length = float(np.linalg.norm(Wt[:3, 3]))
        assert abs(length - logged) / logged < 0.01
        assert abs(length - analytic) / analytic < 0.01


def test_soft_vector_from_the_reference_csv():
    """data/csv/slam_clustered_data.csv row 1: ground-truth camera pose (0, 0, 50) in the world (= tag 0) frame,
    i.e. the camera at the scene origin; the reference's estimate there was (-0.0040036, 0.0041643, 50.0195103)."""
    sc = synth.sim_settings_scene(1000, 1000, cam_pos=(0.0, 0.0, 0.0))
    img = synth.render(sc)
    recs = ob.OracleDetector("tagStandard41h12", decimate=2.0).detect_records(img)
    r0 = [r for r in recs if r["id"] == 0][0]
    T = ob.reference_pose(r0["p"], sc.K, np.zeros((4, 1)), 10.0)[3]
    cam_in_tag = np.linalg.inv(T)[:3, 3]          # camera position in the tag-0 frame (tag z points at the camera)
    est = np.array([-0.0040035578004714, 0.0041642559669223, 50.0195102906674])
    assert np.abs(np.abs(cam_in_tag) - np.array([0.0, 0.0, 50.0])).max() < 0.05
    assert np.abs(np.abs(cam_in_tag) - np.abs(est)).max() < 0.05


def test_aruco_cross_check_of_ids():
    """Independent detector (cv2.aruco, AprilTag 36h11 dictionary): same ids; corner order lb,rb,rt,lt ==
    aruco corners[[1,0,3,2]]."""
    import cv2
    img = DET["grid720_36h11_d2_frame"]
    recs = ob.OracleDetector("tag36h11", decimate=2.0).detect_records(img)
    d = cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_APRILTAG_36h11)
    prm = cv2.aruco.DetectorParameters()
    prm.cornerRefinementMethod = cv2.aruco.CORNER_REFINE_APRILTAG
    corners, ids, _ = cv2.aruco.ArucoDetector(d, prm).detectMarkers(img)
    assert ids is not None
    a = {int(i): c.reshape(4, 2)[[1, 0, 3, 2]] for i, c in zip(ids.ravel(), corners)}
    assert sorted(a) == recs["id"].tolist()
    for r in recs:
        assert np.abs(a[int(r["id"])] - r["p"]).max() < 1.0


def test_bgr2gray_formula_matches_cv2():
    g = np.load(os.path.join(GOLD, "bgr2gray_golden.npz"))
    bgr = g["bgr"].astype(np.uint32)
    mine = (bgr[..., 0] * 3735 + bgr[..., 1] * 19235 + bgr[..., 2] * 9798 + 16384) >> 15
    assert np.array_equal(mine.astype(np.uint8), g["gray"])
    assert g["gray"][0, 0] == 53   # glClearColor(0.5, 0, 0.5) background (renderer.py:206)


def test_oracle_edge_cases():
    o = ob.OracleDetector("tag36h11", decimate=2.0)
    assert len(o.detect_records(np.zeros((100, 100), np.uint8))) == 0       # verify_installation.py:45-51 smoke input
    assert len(o.detect_records(np.full((480, 640), 255, np.uint8))) == 0
    assert len(o.detect_records(np.zeros((5, 7), np.uint8))) == 0
    rng = np.random.default_rng(0)
    assert len(o.detect_records(rng.integers(0, 256, (240, 320), dtype=np.uint8))) == 0
    with pytest.raises(RuntimeError):
        o.detect_records(np.zeros((10, 10, 3), np.uint8))
    with pytest.raises(RuntimeError):
        ob.OracleDetector("tag99h1")


def test_threshold_leftover_pixels_follow_upstreams_fix_up_loop():
    """Upstream thresholds the pixels right of / below the last full 4x4 tile in a separate fix-up loop that uses the
    last full tile's (dilated) extrema and has NO low-contrast test: those pixels are 0 or 255, never 127."""
    flat = np.full((9, 10), 93, np.uint8)                       # tw = 2, th = 2: column 8-9 and row 8 are leftovers
    t = ob.stage_threshold(flat)
    assert (t[:8, :8] == 127).all()                             # full tiles: max - min < 5 -> unknown
    assert (t[:, 8:] == 0).all() and (t[8, :] == 0).all()       # leftovers: 93 > 93 + 0 is false -> 0, not 127
    im = np.full((9, 10), 93, np.uint8)
    im[0:4, 4:8] = [[10, 200, 10, 200]] * 4                     # tile (1, 0): min 10, max 200 -> threshold 105
    im[1, 8], im[2, 9], im[8, 5] = 150, 60, 250
    t = ob.stage_threshold(im)
    assert t[1, 8] == 255 and t[2, 9] == 0                      # right leftovers of tile row 0 use tile (1, 0): 150 > 105
    assert t[8, 5] == 255 and t[8, 1] == 0                      # bottom leftovers use tile row 1, whose dilation sees tile (1, 0)
    assert set(np.unique(t[:, 8:])) <= {0, 255} and set(np.unique(t[8])) <= {0, 255}


def test_connected_components_last_column_follows_upstreams_guards():
    """Columns 0 and w-1 never initiate a union, and upstream skips the up-right union whenever the upper neighbour has
    the pixel's value (do_unionfind_line2): a white pixel of the last column is reached only through the diagonal of
    (w-2, y+1), and only when (w-2, y) is not white."""
    t = np.full((5, 6), 255, np.uint8)
    lab, sz = ob.stage_labels(t)
    assert (lab[:, :5] == 0).all() and sz[0, 0] == 25           # columns 0..4: one component
    assert np.array_equal(lab[:, 5], np.arange(5) * 6 + 5)      # column 5: singletons (up == up-right everywhere)
    t[1, 4] = 0                                                 # now (4, 2)'s upper neighbour is black ...
    lab, sz = ob.stage_labels(t)
    assert lab[1, 5] == 0 and lab[2, 4] == 0                    # ... and its up-right union with (5, 1) happens
    assert lab[0, 5] == 5 and lab[2, 5] == 17 and sz[0, 0] == 25
    assert lab[1, 4] == 10 and sz[1, 4] == 1                    # the black pixel: alone (4-connectivity, no equal neighbour)
    b = np.zeros((4, 5), np.uint8)                              # black: left + up only
    lab, sz = ob.stage_labels(b)
    assert (lab[:, :4] == 0).all() and np.array_equal(lab[:, 4], np.arange(4) * 5 + 4)


def test_oracle_blur_and_threshold_stage_properties():
    rng = np.random.default_rng(1)
    im = rng.integers(0, 256, (61, 83), dtype=np.uint8)
    t = ob.stage_threshold(im)
    assert set(np.unique(t)) <= {0, 127, 255}
    flat = np.full((40, 40), 90, np.uint8)
    assert (ob.stage_threshold(flat) == 127).all()      # min_white_black_diff = 5
    b = ob.stage_blur(im, 0.8)
    assert b.shape == im.shape and not np.array_equal(b, im)
    assert np.array_equal(ob.stage_blur(im, 0.0), im)
    lab, sz = ob.stage_labels(t)
    assert sz.sum() == t.size and (lab <= np.arange(t.size, dtype=np.uint32).reshape(t.shape)).all()
