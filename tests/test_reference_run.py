"""The hard pin of the detector oracle: everything the REFERENCE recorded about its own hot path.

tests/golden/reference_run.npz (tools/make_reference_run_golden.py) = the reference's tag textures
(assets/tags/tag0..4.png), the 89-entry camera trajectory with the logged SLAM.my_pose estimates of
data/csv/slam_clustered_data.csv, and the logged tag-to-tag distances of data/logs/simulation_runner.log -- outputs of
the reference running the REAL upstream apriltag detector + cv2.solvePnP + its SLAMGraph.  The frames are re-rendered
here with the reference's textures and renderer geometry and pushed through this repo's chain.

What agreement means: the reference's own estimate is off the ground truth by up to 1.7 units on these rows (pose
noise of a 100-pixel tag 120 units away); the chain reproduces the LOGGED value -- noise included -- to a few
hundredths, i.e. it tracks the reference's corner localisation to ~0.01 px.  With the code-book cell textures instead
of the reference's anti-aliased PNGs the depth estimate moves by 0.02 units on every row (the test below shows it), so
the comparison is sensitive at that level.
"""
import numpy as np
import pytest

from oracle import replay

TRAJ_TOL = 0.06       # units (scene: tags 50..120 units away); achieved: max 0.046, median 0.004 (profiles/r3a_*)
TRAJ_MEDIAN_TOL = 0.006
WALK_PAIRS = (26, 115, 117, 119, 121, 123, 125, 127)   # log lines of the tracked keyboard walk (0,0,0) -> +z ... -> +x


@pytest.fixture(scope="module")
def gold():
    return replay.load()


def test_fixture_holds_the_references_recorded_data(gold):
    assert gold["textures"].shape == (5, 354, 354) and gold["textures"].dtype == np.uint8
    assert len(gold["traj_gt"]) == 89 and len(np.unique(gold["traj_gt"], axis=0)) == 75
    assert int(gold["traj_frames"].sum()) == 570                      # every CSV row belongs to one entry
    # data/csv/slam_clustered_data.csv row 1 and data/logs/simulation_runner.log:26-27,149-150
    assert np.allclose(gold["traj_est"][0, :3], [-0.0040035578004714, 0.0041642559669223, 50.0195102906674], atol=0)
    ll = dict(zip(gold["log_line"].tolist(), gold["log_len"].tolist()))
    assert (ll[26], ll[27], ll[149], ll[150]) == (76.34146240389457, 45.47295093668058, 76.13008383126994,
                                                  45.62343209718027)
    # the textures decode to tagStandard41h12 ids 0..4 (9x9 cells of 39.33 px)
    from aprilslam_b200 import synth
    for i, t in enumerate(gold["textures"]):
        c = (np.arange(9) + 0.5) * 354 / 9
        cells = t[np.ix_(c.astype(int), c.astype(int))] > 127
        assert np.array_equal(cells, synth.tag_cells("tagStandard41h12", i).astype(bool))


def test_oracle_chain_reproduces_the_logged_trajectory(gold):
    """data/csv/slam_clustered_data.csv: camera pose -> logged my_pose, all 89 trajectory entries in order."""
    chain = replay.chain_oracle(gold)
    rows = replay.report_rows(gold, chain)
    assert [r["nodes"] for r in rows] == [r["nodes_logged"] for r in rows]        # 'Number of Nodes' column, every entry
    with0 = [r for r in rows if 0 in r["visible"]]
    assert len(with0) == 78
    d = np.abs(np.array([r["diff"] for r in with0]))
    assert d.max() <= TRAJ_TOL and np.median(d.max(axis=1)) <= TRAJ_MEDIAN_TOL, (d.max(), np.median(d.max(axis=1)))
    # the logged values are reproduced far better than they match the ground truth: where the reference itself is off
    # by more than 0.3 units the chain is off the LOG by less than a sixth of that
    noisy = [r for r in with0 if np.abs(r["logged_err"]).max() > 0.3]
    assert len(noisy) >= 20
    assert all(np.abs(r["diff"]).max() < np.abs(r["logged_err"]).max() / 6 for r in noisy)
    # entries 78..88: tag 0 has left the view and the estimate runs on world transforms frozen at entry 77
    # (slam_graph.py:50-54).  The logged error grows linearly with the distance moved since (0.6 -> 5.2 units), the
    # signature of a rotation error frozen into a world transform (one of the tilted tags' planar pose ambiguity at
    # the freeze frame); that branch is not reproducible from outside, only the structure is compared
    for r in rows[78:]:
        assert r["visible"] == [2, 3, 4] and r["diff"] is not None


def test_reference_textures_matter_at_the_level_compared(gold):
    """Sensitivity of the pin: the same row with code-book cell textures instead of the reference's PNGs."""
    from aprilslam_b200 import synth
    from oracle import binding as ob
    o = ob.OracleDetector("tagStandard41h12", decimate=2.0)
    est = gold["traj_est"][0, :3]
    z = {}
    for name, tex in (("png", gold["textures"]), ("cells", None)):
        sc = synth.sim_settings_scene(1000, 1000, cam_pos=(0.0, 0.0, 0.0), textures=tex)
        r0 = [r for r in o.detect_records(synth.render(sc)) if r["id"] == 0][0]
        T = ob.reference_pose(r0["p"], sc.K, np.zeros((4, 1)), replay.TAG_SIZE)[3]
        z[name] = np.linalg.inv(T)[2, 3]
    assert abs(z["png"] - est[2]) < 0.003 and abs(z["cells"] - est[2]) > 0.015, z


def test_oracle_chain_reproduces_the_logged_world_transform_lengths(gold):
    """data/logs/simulation_runner.log: 'Tag ID n (reference: 0): World transform translation length' (slam_graph.py:45-49).
    Line 26-27: camera at the start pose.  Lines 115-128: the keyboard walk +z ... +x in steps of 2 (poses found by the
    lattice search of tools/make_reference_run_golden.py; each is the next lattice point of the previous one).
    Line 149-150: the pose the run rested at for its last 101 frames."""
    line = gold["pair_line"].tolist()
    for ln in WALK_PAIRS + (149,):
        i = line.index(ln)
        got = replay.world_lengths(gold, gold["pair_pos"][i])
        want = gold["pair_len"][i]
        assert np.abs(np.array(got) - want).max() < 0.012, (ln, got, want)      # achieved: <= 0.0101 (1.3e-4 relative)
        assert (np.abs(np.array(got) - want) / want).max() < 2e-4
    walk = np.array([gold["pair_pos"][line.index(ln)] for ln in WALK_PAIRS])
    assert np.array_equal(np.abs(np.diff(walk, axis=0)).sum(axis=1), np.full(len(walk) - 1, 2.0))   # one key press each
