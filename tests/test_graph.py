"""Tag graph + camera pose estimate: the consumer of the detect + pose path.

CPU: the numpy restatement (oracle/graph_oracle.py) against tests/golden/graph_golden.npz, which was produced by the
reference's own SLAMGraph / SLAM.my_pose code (tools/make_graph_golden.py).
GPU: k_graph_update behind agpu_graph_update against the same fixture and against the oracle on fresh inputs.
Tolerance: 1e-9 relative to the magnitude of the transforms (4x4 products / inverse in float64; numpy's BLAS / LAPACK
and the kernel only differ in rounding order).
"""
import os

import numpy as np
import pytest

from oracle.graph_oracle import run_streams

GOLD = os.path.join(os.path.dirname(__file__), "golden", "graph_golden.npz")
TOL = 1e-9


def _close(a, b, scale):
    return np.abs(np.asarray(a) - np.asarray(b)).max() <= TOL * max(1.0, scale)


def _check_state(st, g, s, scale):
    assert st["coordinate_id"] == int(g["coordinate_id"][s])
    assert np.array_equal(st["present"], g["present"][s].astype(bool))
    p = st["present"]
    for k in ("reference", "weight"):
        assert np.array_equal(st[k][p], g[k][s][p]), k
    for k in ("updated", "visible"):
        assert np.array_equal(st[k][p], g[k][s][p].astype(bool)), k
    assert _close(st["local"][p], g["local"][s][p], scale) and _close(st["world"][p], g["world"][s][p], scale)
    assert _close(st["estimated_pose"], g["estimated_pose"][s], scale)


def test_oracle_matches_reference_golden():
    g = np.load(GOLD)
    my_pose, valid, graphs = run_streams(g["ids"], g["ok"], g["T"], g["counts"], int(g["max_id"]))
    assert np.array_equal(valid, g["valid"].astype(bool))
    scale = np.abs(g["my_pose"]).max()
    assert _close(my_pose, g["my_pose"], scale)
    for s, go in enumerate(graphs):
        st = {k: getattr(go, k) for k in ("coordinate_id", "present", "reference", "weight", "updated", "visible", "local",
                                          "world", "estimated_pose")}
        _check_state(st, g, s, scale)
    # the fixture exercises every branch of add_or_update_node
    assert set(np.unique(g["weight"][g["present"] > 0])) >= {1, 2} and (g["updated"][g["present"] > 0] == 0).any()
    assert (g["valid"] == 0).any() and (g["ok"][g["counts"][..., None] > np.arange(g["ids"].shape[2])] == 0).any()


@pytest.mark.gpu
def test_gpu_graph_matches_reference_golden():
    from aprilslam_b200.detector import Detector
    from aprilslam_b200.slam_graph import SLAMGraphBatch, transforms_to_records
    g = np.load(GOLD)
    det = Detector("tag36h11")
    S, F, cap = g["ids"].shape
    gb = SLAMGraphBatch(det, S, int(g["max_id"]))
    dets, poses = transforms_to_records(g["ids"], g["T"], g["ok"])
    scale = np.abs(g["my_pose"]).max()
    # in two calls: the graph has to survive between calls
    h = F // 2
    mp1, v1 = gb.update(dets[:, :h], poses[:, :h], g["counts"][:, :h])
    mp2, v2 = gb.update(dets[:, h:], poses[:, h:], g["counts"][:, h:])
    mp, v = np.concatenate([mp1, mp2], 1), np.concatenate([v1, v2], 1)
    assert np.array_equal(v, g["valid"].astype(bool))
    assert _close(mp[v], g["my_pose"][v], scale)
    for s in range(S):
        _check_state(gb.state(s), g, s, scale)
    assert det.launch_count() == 1
    # reset -> same answers again in one call
    gb.reset()
    mp3, v3 = gb.update(dets, poses, g["counts"])
    assert np.array_equal(v3, v) and _close(mp3[v], mp[v], scale)
    nodes = gb.get_nodes(0)
    assert sorted(nodes) == list(np.nonzero(g["present"][0])[0]) and gb.get_coordinate_id(0) == int(g["coordinate_id"][0])
    gb.close()
    det.close()


@pytest.mark.gpu
def test_gpu_graph_after_detection_matches_oracle():
    """detect + pose on rendered frames of two 'cameras', piped straight into the graph; oracle on the same records."""
    from aprilslam_b200 import synth
    from aprilslam_b200.detector import Detector
    from aprilslam_b200.slam_graph import SLAMGraphBatch
    det = Detector("tag36h11", decimate=2.0)
    S, F, W, H = 2, 4, 640, 480
    K = synth.intrinsics(W, H, 45.0)
    # the reference's own 5-tag scene (config/sim_settings.json) seen by two cameras that slide sideways in opposite
    # directions: tags enter and leave the view, and for camera 1 lower ids arrive late (the world tag changes)
    cams = [[(x, 0.0, 0.0) for x in (0.0, 15.0, 30.0, 45.0)], [(x, 0.0, 0.0) for x in (60.0, 40.0, 20.0, 0.0)]]
    frames = np.stack([synth.render(synth.sim_settings_scene(W, H, cam_pos=cams[s][f], family="tag36h11"))
                       for s in range(S) for f in range(F)])
    dl, pl = det.detect_pose_batch(frames, K, None, 10.0)
    cap = max(1, max(len(d) for d in dl))
    from aprilslam_b200._lib import DET_DTYPE, POSE_DTYPE
    dets, poses = np.zeros((S, F, cap), DET_DTYPE), np.zeros((S, F, cap), POSE_DTYPE)
    counts = np.zeros((S, F), np.int32)
    for i, (d, p) in enumerate(zip(dl, pl)):
        s, f = divmod(i, F)
        dets[s, f, :len(d)], poses[s, f, :len(d)], counts[s, f] = d, p, len(d)
    assert counts.min() >= 1 and counts.max() >= 3
    gb = SLAMGraphBatch(det, S, 586)
    mp, v = gb.update(dets, poses, counts)
    T = np.zeros((S, F, cap, 4, 4))
    T[..., :3, :3] = poses["R"].reshape(S, F, cap, 3, 3)
    T[..., :3, 3] = poses["tvec"]
    T[..., 3, 3] = 1
    mp_ref, v_ref, graphs = run_streams(dets["id"], poses["ok"], T, counts, 586)
    assert v[:, 0].all() and np.array_equal(v, v_ref)
    assert _close(mp[v], mp_ref[v], np.abs(mp_ref).max())
    for s in range(S):
        st = gb.state(s)
        assert st["coordinate_id"] == graphs[s].coordinate_id and np.array_equal(st["present"], graphs[s].present)
        assert np.array_equal(st["weight"], graphs[s].weight)
    gb.close()
    det.close()


@pytest.mark.gpu
def test_closing_the_detector_closes_its_graphs_first_and_41h12_warns():
    from aprilslam_b200.detector import Detector
    from aprilslam_b200.slam_graph import SLAMGraphBatch
    with pytest.warns(RuntimeWarning, match="only 5 of upstream's 2115 code words"):
        det = Detector("tagStandard41h12", decimate=2.0)
    g = SLAMGraphBatch(det, nstreams=2, max_tag_id=4)
    det.close()                      # must not leave g with a dangling C handle
    assert g._g is None
    g.close()                        # idempotent
    del g
