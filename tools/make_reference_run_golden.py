#!/usr/bin/env python
"""Generate tests/golden/reference_run.npz: everything the REFERENCE itself recorded about its hot path.

Needs /root/reference (read-only), so it only runs in the build container; the fixture is committed and the tests
(tests/test_reference_run.py) never read /root/reference.

What goes in (all of it data the reference holds, none of it produced by this repo's detector):

  textures   uint8 [5,354,354]  the reference's own tag images assets/tags/tag{0..4}.png (R = G = B; alpha is ignored by
                                the renderer's opaque writes, renderer.py:150-176), the textures its OpenGL renderer maps
                                on the tag quads with GL_LINEAR (renderer.py:172-173)
  traj_*     the committed run data/csv/slam_clustered_data.csv (570 rows) reduced to its camera trajectory: one entry
             per run of consecutive rows with the same ground-truth pose (89 entries, 75 distinct poses).  Per entry:
             GT_X/Y/Z (= camera position - tag 0 position, ground_truth.py:146-188), the logged estimate
             Est_X/Y/Z + roll/pitch/yaw (SLAM.my_pose, slam.py:36-63, as simulation_engine.py:240-300 logs it) and the
             logged number of graph nodes.  Rows inside one run are identical (checked here).
  log_*      data/logs/simulation_runner.log: every "Tag ID n (reference: 0): World transform translation length = v"
             line (slam_graph.py:45-49) as (line number, tag id, value).  The camera pose is not logged; that run was
             driven from the keyboard in steps of movement_speed * size_scale = 2 units (camera_controller.py:48,90-95),
             so the pose behind a (tag 1, tag 2) pair of lengths is found by a search over that lattice
             (--search, minutes); `log_pos` holds the lattice point per pair and `log_margin` how much worse the
             runner-up lattice point fits.  The pairs used by the tests are the unambiguous ones.
"""
import argparse
import itertools
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "reference_run.npz")

TAG0_POS = np.array([0.0, 0.0, -50.0])   # config/sim_settings.json:11-16


def load_textures():
    import cv2
    tex = []
    for i in range(5):
        im = cv2.imread(os.path.join(REF, "assets", "tags", "tag%d.png" % i), cv2.IMREAD_UNCHANGED)
        assert im.shape == (354, 354, 4) and np.array_equal(im[..., 0], im[..., 1]) and np.array_equal(im[..., 1], im[..., 2])
        tex.append(im[..., 0].copy())
    return np.stack(tex)


def load_trajectory():
    import pandas as pd
    df = pd.read_csv(os.path.join(REF, "data", "csv", "slam_clustered_data.csv"))
    gt = ["GT_X", "GT_Y", "GT_Z"]
    est = ["Est_X", "Est_Y", "Est_Z", "Est_Roll", "Est_Pitch", "Est_Yaw"]
    start = (df[gt].shift() != df[gt]).any(axis=1).to_numpy()
    run_id = np.cumsum(start) - 1
    # rows of one run carry identical estimates (the scene is static while the camera rests)
    for r in np.unique(run_id):
        blk = df.loc[run_id == r, est + ["Number of Nodes"]].to_numpy()
        assert np.all(blk == blk[0]), r
    u = df.loc[start].reset_index(drop=True)
    assert (u[["GT_Roll", "GT_Pitch", "GT_Yaw"]].to_numpy() == np.array([np.pi, -0.0, 0.0])).all()   # no camera rotation
    return dict(traj_gt=u[gt].to_numpy(np.float64), traj_est=u[est].to_numpy(np.float64),
                traj_nodes=u["Number of Nodes"].to_numpy(np.int32),
                traj_rows=np.flatnonzero(start).astype(np.int32),
                traj_frames=np.bincount(run_id).astype(np.int32))


def load_log():
    pat = re.compile(r"Tag ID (\d+) \(reference: 0\): World transform translation length = ([0-9.eE+-]+)")
    rows = []
    with open(os.path.join(REF, "data", "logs", "simulation_runner.log")) as f:
        for ln, line in enumerate(f, 1):
            m = pat.search(line)
            if m:
                rows.append((ln, int(m.group(1)), float(m.group(2))))
    return rows


def _lengths_at(args):
    cam, tex = args
    from aprilslam_b200 import synth
    from oracle import binding as ob
    sc = synth.sim_settings_scene(1000, 1000, cam_pos=cam, textures=tex)
    recs = ob.OracleDetector("tagStandard41h12", decimate=2.0).detect_records(synth.render(sc))
    T = {int(r["id"]): ob.reference_pose(r["p"], sc.K, np.zeros((4, 1)), 10.0)[3] for r in recs}
    out = [np.nan, np.nan]
    if 0 in T:
        for k, t in enumerate((1, 2)):
            if t in T:
                out[k] = float(np.linalg.norm((np.linalg.inv(T[0]) @ T[t])[:3, 3]))
    return out


def search_log_positions(tex, pairs):
    """Keyboard lattice (steps of 2) around the start pose: which lattice point reproduces a logged pair best."""
    from multiprocessing import Pool
    cams = list(itertools.product(range(-12, 13, 2), range(-8, 9, 2), range(-12, 13, 2)))
    with Pool(min(16, os.cpu_count() or 1)) as p:
        tab = np.array(p.map(_lengths_at, [(c, tex) for c in cams], chunksize=8))
    cams = np.array(cams, np.float64)
    pos, margin, fit = [], [], []
    for a, b in pairs:
        d = np.abs(tab[:, 0] - a) + np.abs(tab[:, 1] - b)
        d = np.where(np.isnan(d), np.inf, d)
        o = np.argsort(d)
        pos.append(cams[o[0]]); fit.append(d[o[0]]); margin.append(d[o[1]] - d[o[0]])
    return np.array(pos), np.array(fit), np.array(margin)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--search", action="store_true", help="redo the lattice search for the log lines (minutes)")
    args = ap.parse_args()
    tex = load_textures()
    out = dict(textures=tex, tag0_pos=TAG0_POS)
    out.update(load_trajectory())
    log = load_log()
    out["log_line"] = np.array([r[0] for r in log], np.int32)
    out["log_tag"] = np.array([r[1] for r in log], np.int32)
    out["log_len"] = np.array([r[2] for r in log], np.float64)
    # distinct (tag 1, tag 2) pairs logged in the same frame, in order of first appearance
    pairs, first_line = [], []
    for (l1, t1, v1), (l2, t2, v2) in zip(log, log[1:]):
        if t1 == 1 and t2 == 2 and l2 == l1 + 1 and (v1, v2) not in pairs:
            pairs.append((v1, v2)); first_line.append(l1)
    out["pair_len"] = np.array(pairs)
    out["pair_line"] = np.array(first_line, np.int32)
    if args.search or not os.path.exists(OUT) or "pair_pos" not in np.load(OUT):
        pos, fit, margin = search_log_positions(tex, pairs)
    else:
        old = np.load(OUT)
        pos, fit, margin = old["pair_pos"], old["pair_fit"], old["pair_margin"]
    out["pair_pos"], out["pair_fit"], out["pair_margin"] = pos, fit, margin
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(out["traj_gt"]), "trajectory entries,", len(log), "log lines,",
          len(pairs), "distinct pairs")
    for ln, (a, b), p, f, m in zip(first_line, pairs, pos, fit, margin):
        print("  log line %4d: %.5f / %.5f -> camera %s  (misfit %.4f, runner-up +%.4f)" % (ln, a, b, p.tolist(), f, m))


if __name__ == "__main__":
    main()
