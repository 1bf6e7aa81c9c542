#!/usr/bin/env python
"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck), run on a B200 box:

  compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_run.py

Covers every kernel of the pipeline on inputs small enough for the sanitizer's 10-100x slowdown: ragged stage-level
shapes (threshold leftovers, last-column unions), a scene with tags at decimate 1 and 2 (all quad-fit tiers that a small
frame can reach, decode, reconcile, pose), noise / block frames (thousands of components: the lock-free union-find of
k_cc_local / k_cc_boundary, the id-width switch, the segmented radix sort), a BGR frame, blur, and the graph update.
Every result is compared with the CPU oracle, so a run under the sanitizer is also a parity run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from aprilslam_b200 import synth
from aprilslam_b200.detector import Detector
from oracle import binding as ob

rng = np.random.default_rng(3)
n_checked = 0


def same(recs, ref):
    global n_checked
    assert recs["id"].tolist() == ref["id"].tolist() and recs["hamming"].tolist() == ref["hamming"].tolist()
    if len(ref):
        assert np.abs(recs["p"] - ref["p"]).max() <= 0.05
    n_checked += 1


for (W, H, d) in [(333, 77, 1), (97, 131, 3), (163, 121, 2), (256, 96, 1)]:
    det = Detector("tag36h11", decimate=float(d))
    for im in (rng.integers(0, 256, (H, W), dtype=np.uint8),
               (np.kron(rng.integers(0, 2, (H // 5 + 1, W // 5 + 1), dtype=np.uint8) * 180 + 30, np.ones((5, 5), np.uint8))[:H, :W]
                + rng.integers(0, 7, (H, W), dtype=np.uint8)).astype(np.uint8)):
        im = np.ascontiguousarray(im)
        q, t = det.stage_threshold(im)
        t_ref = ob.stage_threshold(np.ascontiguousarray(im[::d, ::d]))
        assert np.array_equal(t, t_ref)
        lab, sz = det.stage_labels(t_ref)
        lab_ref, sz_ref = ob.stage_labels(t_ref)
        assert np.array_equal(lab, lab_ref) and np.array_equal(sz, sz_ref)
        n_checked += 1
    det.close()

for d in (1.0, 2.0):
    sc = synth.grid_scene(640, 480, 11, (3, 2), px_range=(50, 140))
    img = synth.render(sc)
    det = Detector("tag36h11", decimate=d, debug=True)
    dets, poses = det.detect_pose_batch(np.stack([img, img[::-1].copy(), img]), sc.K, None, 0.1)
    o = ob.OracleDetector("tag36h11", decimate=d)
    same(dets[0], o.detect_records(img))
    same(dets[1], o.detect_records(img[::-1].copy()))
    assert poses[0]["ok"].all()
    bgr = np.repeat(img[..., None], 3, axis=2)
    same(det.detect_batch(bgr, bgr=True)[0], o.detect_records(img))
    det.close()

det = Detector("tag36h11", decimate=1.0)
o = ob.OracleDetector("tag36h11", decimate=1.0)
clean = synth.render(synth.grid_scene(480, 360, 5, (2, 2), px_range=(50, 90)))
noise = rng.integers(0, 256, (360, 480), dtype=np.uint8)
blocks = np.kron(rng.integers(0, 2, (90, 120), dtype=np.uint8) * 200 + 25, np.ones((4, 4), np.uint8)).astype(np.uint8)
for im in (clean, noise, clean, blocks, clean):     # (few components -> 11-bit ids; blocks -> re-run with 16-bit ids)
    same(det.detect_batch(im, cap_per_frame=256)[0], o.detect_records(im))
det.close()

det = Detector("tag36h11", decimate=2.0, blur=0.8)
same(det.detect_batch(clean)[0], ob.OracleDetector("tag36h11", decimate=2.0, blur=0.8).detect_records(clean))
det.close()

from aprilslam_b200.slam_graph import SLAMGraphBatch
sc = synth.sim_settings_scene(640, 480)
det = Detector("tagStandard41h12", decimate=2.0)
dets, poses = det.detect_pose_batch(synth.render(sc), sc.K, None, 10.0, cap_per_frame=8)
g = SLAMGraphBatch(det, nstreams=1, max_tag_id=4)
mp = g.update_lists([dets[0]], [poses[0]])
assert mp[0] is not None
g.close()
det.close()
print("sanitize_run ok: %d comparisons with the oracle" % n_checked)
