#!/usr/bin/env python
"""Generate tests/golden/graph_golden.npz by running the REFERENCE's own graph code.

Imports /root/reference/src/core/slam_graph.py and src/core/slam.py UNMODIFIED (matplotlib / apriltag, which the
reference imports but this container lacks, are satisfied by empty stub modules) and drives, per frame, exactly what
the reference's caller does (src/simulation/simulation_engine.py:219-232):

    slam.visible_tags = [ids of all detections]                 # SLAM.detect, slam.py:21-25
    for every detection whose solvePnP succeeded:
        slam.graph.add_or_update_node(id, T, slam.visible_tags)  # SLAM.get_pose, slam.py:27-32
    slam.my_pose()                                               # slam.py:36-63

Inputs are seeded random streams (tag subsets that come and go, rigid camera<-tag transforms, failed poses, frames
without detections, a lower id appearing late so that the world tag changes).  Only runs in the build container;
the fixture is committed so that nothing under tests/ reads /root/reference at run time.
"""
import io
import logging
import os
import sys
import types
import contextlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def import_reference_slam():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.lines", "matplotlib.patches", "matplotlib.animation",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d", "apriltag"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []                                   # (so that sub-module imports resolve to the stubs)
            m.__getattr__ = lambda attr: object               # any `from x import Y` yields a placeholder
            sys.modules[name] = m
    sys.modules["apriltag"].apriltag = lambda *a, **k: None
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    return importlib.import_module("src.core.slam")


def rigid(rng, tscale=50.0):
    a = rng.normal(size=3)
    a /= np.linalg.norm(a)
    th = rng.uniform(-np.pi, np.pi)
    Kx = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = rng.normal(size=3) * tscale
    return T


def main():
    slam_mod = import_reference_slam()
    rng = np.random.default_rng(20261018)
    S, F, cap, max_id = 6, 40, 12, 30
    ids = np.zeros((S, F, cap), np.int32)
    ok = np.zeros((S, F, cap), np.uint8)
    T = np.zeros((S, F, cap, 4, 4))
    counts = np.zeros((S, F), np.int32)
    my_pose = np.zeros((S, F, 4, 4))
    valid = np.zeros((S, F), np.uint8)
    present = np.zeros((S, max_id + 1), np.uint8)
    reference = np.full((S, max_id + 1), -1, np.int32)
    weight = np.zeros((S, max_id + 1), np.int32)
    updated = np.zeros((S, max_id + 1), np.uint8)
    visible = np.zeros((S, max_id + 1), np.uint8)
    local = np.zeros((S, max_id + 1, 4, 4))
    world = np.zeros((S, max_id + 1, 4, 4))
    coord = np.zeros(S, np.int32)
    est = np.zeros((S, 4, 4))
    log = logging.getLogger("graph_golden")
    log.addHandler(logging.NullHandler())
    for s in range(S):
        slam = slam_mod.SLAM.__new__(slam_mod.SLAM)          # (the constructor wants a native detector; the graph code does not)
        slam.logger = log
        slam.graph = slam_mod.SLAMGraph(log)
        slam.visible_tags = []
        lo = int(rng.integers(3, 10))                         # ids below `lo` only appear in the second half
        for f in range(F):
            pool = np.arange(lo if f < F // 2 else 0, max_id + 1)
            n = 0 if rng.random() < 0.08 else int(rng.integers(1, cap + 1))
            sel = np.sort(rng.choice(pool, size=min(n, len(pool)), replace=False))
            if n >= 3 and rng.random() < 0.15:
                sel[1] = sel[0]                               # a duplicated id (two non-overlapping detections)
            counts[s, f] = len(sel)
            for i, tid in enumerate(sel):
                ids[s, f, i] = tid
                T[s, f, i] = rigid(rng)
                ok[s, f, i] = 0 if rng.random() < 0.07 else 1
            # ---- the reference's caller loop
            slam.visible_tags = [int(t) for t in sel]
            with contextlib.redirect_stdout(io.StringIO()):
                for i, tid in enumerate(sel):
                    if ok[s, f, i]:
                        slam.graph.add_or_update_node(int(tid), T[s, f, i].copy(), slam.visible_tags)
                mp = slam.my_pose()
            if mp is not None:
                my_pose[s, f] = mp
                valid[s, f] = 1
        coord[s] = slam.graph.get_coordinate_id()
        est[s] = slam.graph.get_estimated_pose()
        for tid, node in slam.graph.get_nodes().items():
            present[s, tid] = 1
            reference[s, tid] = node.reference
            weight[s, tid] = node.weight
            updated[s, tid] = node.updated
            visible[s, tid] = node.visible
            local[s, tid] = node.local
            world[s, tid] = node.world
    out = dict(ids=ids, ok=ok, T=T, counts=counts, my_pose=my_pose, valid=valid, present=present, reference=reference,
               weight=weight, updated=updated, visible=visible, local=local, world=world, coordinate_id=coord,
               estimated_pose=est, max_id=np.int32(max_id))
    path = os.path.join(ROOT, "tests", "golden", "graph_golden.npz")
    np.savez_compressed(path, **out)
    print("graph_golden:", {k: np.asarray(v).shape for k, v in out.items()})
    print("valid frames per stream:", valid.sum(1), "nodes per stream:", present.sum(1), "coordinate ids:", coord)


if __name__ == "__main__":
    main()
