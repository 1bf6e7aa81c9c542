"""CPU restatement of the reference's tag-graph update + camera pose estimate -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import this; the product path
(aprilslam_b200/slam_graph.py -> agpu_graph_update -> k_graph_update) never does.

Follows /root/reference/src/core/slam_graph.py:29-70 (add_or_update_node, find_world, get_world) and
/root/reference/src/core/slam.py:36-63 (my_pose), driven per frame the way the reference's caller does
(src/simulation/simulation_engine.py:219-232).  Pinned by tests/golden/graph_golden.npz, which
tools/make_graph_golden.py produced by running the reference's own classes.
"""
import numpy as np


class GraphOracle:
    """State of one camera stream as flat arrays over tag ids 0..max_id (the layout the CUDA kernel uses)."""

    def __init__(self, max_id: int):
        n = max_id + 1
        self.coordinate_id = -1                       # slam_graph.py:21
        self.estimated_pose = np.zeros((4, 4))        # slam_graph.py:22
        self.present = np.zeros(n, bool)
        self.reference = np.full(n, -1, np.int32)
        self.weight = np.zeros(n, np.int32)
        self.updated = np.zeros(n, bool)
        self.visible = np.zeros(n, bool)
        self.local = np.zeros((n, 4, 4))
        self.world = np.zeros((n, 4, 4))
        self.skipped = 0

    def _set(self, tid, local, world, reference, weight=1, updated=True):
        self.present[tid] = True
        self.local[tid], self.world[tid] = local, world
        self.reference[tid], self.weight[tid], self.updated[tid], self.visible[tid] = reference, weight, updated, False

    def add_or_update(self, tid: int, T: np.ndarray, visible_ids):
        Ti = np.linalg.inv(T)                                                    # slam_graph.py:24-27
        c = self.coordinate_id
        if c == -1 or c == tid or tid < c:                                       # :33-39 (update_world is a no-op, :72-76)
            self.coordinate_id = tid
            self._set(tid, Ti, np.eye(4), tid)
            return
        ref = min(visible_ids)                                                   # :41
        if ref == c:                                                             # :42-44
            self._set(tid, Ti, self.local[ref] @ T, c)
        elif self.present[tid] and self.reference[tid] == c:                     # :50-54
            self._set(tid, Ti, self.world[tid].copy(), c, int(self.weight[tid]), False)
        elif ref != tid and self.present[ref]:                                   # :55-57 + find_world :61-66
            world = self.world[ref] @ (self.local[ref] @ T)
            self._set(tid, Ti, world, int(self.reference[ref]), int(self.weight[ref]) + 1, bool(self.updated[ref]))
        else:                                                                    # :58-59
            self.skipped += 1

    def my_pose(self, visible_ids):
        if len(visible_ids) == 0:                                                # slam.py:38-39
            return None
        self.visible[:] = False                                                  # :45-46
        T_sum, count = np.zeros((4, 4)), 0.0
        for tid in visible_ids:                                                  # :48-55
            if self.present[tid]:
                self.visible[tid] = True
                T_sum += (self.world[tid] @ self.local[tid]) / self.weight[tid]
                count += 1 / self.weight[tid]
        if count == 0:                                                           # :57-58
            return None
        self.estimated_pose = T_sum / count                                      # :60-62
        return self.estimated_pose


def run_streams(ids, ok, T, counts, max_id):
    """ids/ok [S,F,cap], T [S,F,cap,4,4], counts [S,F] -> (my_pose [S,F,4,4], valid [S,F], [GraphOracle per stream])"""
    S, F = counts.shape
    my_pose, valid, graphs = np.zeros((S, F, 4, 4)), np.zeros((S, F), bool), []
    for s in range(S):
        g = GraphOracle(max_id)
        for f in range(F):
            n = int(counts[s, f])
            vis = [int(t) for t in ids[s, f, :n]]
            for i in range(n):
                if ok[s, f, i]:
                    g.add_or_update(vis[i], T[s, f, i], vis)
            mp = g.my_pose(vis)
            if mp is not None:
                my_pose[s, f], valid[s, f] = mp, True
        graphs.append(g)
    return my_pose, valid, graphs
