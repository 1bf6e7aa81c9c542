"""Replay of the reference's recorded run through the CPU oracle chain -- TEST INFRASTRUCTURE ONLY.

tests/golden/reference_run.npz (tools/make_reference_run_golden.py) holds what the reference itself recorded: its tag
textures, the camera trajectory + logged `my_pose` estimates of data/csv/slam_clustered_data.csv and the logged
tag-to-tag distances of data/logs/simulation_runner.log.  Here the frames are re-rendered with the reference's textures
and geometry (aprilslam_b200.synth, renderer.py:91-96,188-251) and pushed through
detect (oracle) -> cv2.solvePnP (tag_detector.py:30-43) -> SLAMGraph / my_pose restatement (graph_oracle.py),
the same caller loop as simulation_engine.py:212-238.
"""
import os

import numpy as np

from aprilslam_b200 import synth
from . import binding as ob
from .graph_oracle import GraphOracle

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_run.npz")
TAG_SIZE = 10.0          # tag_size_inner * size_scale (simulation_engine.py:141, config/sim_settings.json:7-8)
FAMILY = "tagStandard41h12"


def load():
    return np.load(GOLD)


def camera_of(gt_xyz, tag0_pos):
    """GT_X/Y/Z = camera position - tag 0 position for the unrotated tag 0 (ground_truth.py:146-188)."""
    return np.asarray(gt_xyz, float) + np.asarray(tag0_pos, float)


def render_frame(gold, cam):
    sc = synth.sim_settings_scene(1000, 1000, cam_pos=tuple(cam), textures=gold["textures"])
    return sc, synth.render(sc)


def chain_oracle(gold, detect=None, pose=None):
    """-> list of dicts per trajectory entry: visible ids, nodes in the graph, my_pose (4x4 or None).
    detect(img) -> records with 'id' and 'p'; pose(corners, K) -> (ok, T); both default to the CPU oracle chain."""
    o = ob.OracleDetector(FAMILY, decimate=2.0)
    detect = detect or o.detect_records
    pose = pose or (lambda p, K: (lambda r: (r[0], r[3]))(ob.reference_pose(p, K, np.zeros((4, 1)), TAG_SIZE)))
    g = GraphOracle(4)
    out = []
    for gt in gold["traj_gt"]:
        sc, img = render_frame(gold, camera_of(gt, gold["tag0_pos"]))
        recs = detect(img)
        vis = [int(r["id"]) for r in recs]
        for r in recs:                                   # simulation_engine.py:222-226
            ok, T = pose(r["p"], sc.K)
            if ok:
                g.add_or_update(int(r["id"]), T, vis)    # slam.py:29-31
        mp = g.my_pose(vis)                              # simulation_engine.py:232
        out.append(dict(visible=vis, nodes=int(g.present.sum()), my_pose=None if mp is None else mp.copy()))
    return out


def world_lengths(gold, cam, detect=None):
    """Translation lengths of get_world(0, T_i) for tags 1 and 2 (slam_graph.py:44-49) from camera position cam."""
    o = ob.OracleDetector(FAMILY, decimate=2.0)
    sc, img = render_frame(gold, cam)
    recs = (detect or o.detect_records)(img)
    T = {int(r["id"]): ob.reference_pose(r["p"], sc.K, np.zeros((4, 1)), TAG_SIZE)[3] for r in recs}
    return [float(np.linalg.norm((np.linalg.inv(T[0]) @ T[t])[:3, 3])) if (0 in T and t in T) else float("nan")
            for t in (1, 2)]


def report_rows(gold, chain):
    """Per trajectory entry: (index, gt xyz, visible, nodes logged/ours, |ours - logged| xyz, |logged - gt| xyz)."""
    rows = []
    for k, (gt, est, nn, c) in enumerate(zip(gold["traj_gt"], gold["traj_est"], gold["traj_nodes"], chain)):
        d = None if c["my_pose"] is None else c["my_pose"][:3, 3] - est[:3]
        rows.append(dict(k=k, gt=gt.tolist(), visible=c["visible"], nodes_logged=int(nn), nodes=c["nodes"],
                         diff=None if d is None else d.tolist(), logged_err=(est[:3] - gt).tolist()))
    return rows
