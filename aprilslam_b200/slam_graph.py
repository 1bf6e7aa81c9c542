"""Batched tag graph + camera pose estimate: the consumer of the detect + pose path, on the GPU.

Mirrors `SLAMGraph.add_or_update_node / find_world / get_world` (/root/reference/src/core/slam_graph.py:29-70) and
`SLAM.my_pose` (/root/reference/src/core/slam.py:36-63) for S independent camera streams at once; the reference's
per-frame caller loop (`simulation_engine.py:219-232`) is what one frame of `update` replays:

    detections = slam.detect(frame)            # visible_tags = ids of all detections
    for d in detections: slam.get_pose(d)      # graph.add_or_update_node(id, T, visible_tags) when solvePnP succeeded
    slam.my_pose()                             # weighted average of world @ local over the visible tags

There is no CPU implementation here: the work runs in `k_graph_update` behind `agpu_graph_update`.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from ._lib import DET_DTYPE, POSE_DTYPE


class Node:
    """Same fields as the reference's Node (slam_graph.py:5-12)."""

    def __init__(self, local, world, reference, weight=1, updated=True, visible=False):
        self.local, self.world, self.reference = local, world, reference
        self.weight, self.updated, self.visible = weight, updated, visible


class SLAMGraphBatch:
    """Tag graphs of `nstreams` cameras, resident on the detector's GPU."""

    def __init__(self, detector, nstreams: int = 1, max_tag_id: int = 586):
        self._det = detector
        self._L = detector._L
        self.nstreams, self.max_tag_id = int(nstreams), int(max_tag_id)
        h = C.c_void_p()
        detector._check(self._L.agpu_graph_create(detector._h, self.nstreams, self.max_tag_id, C.byref(h)))
        self._g = h
        import weakref
        detector._graphs.append(weakref.ref(self))     # Detector.close() closes its graphs before the handle goes away

    def close(self):
        if getattr(self, "_g", None):
            self._L.agpu_graph_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self._det._check(self._L.agpu_graph_reset(self._g))

    def update(self, dets: np.ndarray, poses: np.ndarray, counts: np.ndarray):
        """dets / poses: structured arrays [S, F, cap] (DET_DTYPE / POSE_DTYPE, e.g. Detector.detect_pose_records);
        counts [S, F].  -> (my_pose float64 [S, F, 4, 4], valid bool [S, F]); valid False = my_pose() returned None."""
        dets = np.ascontiguousarray(dets, DET_DTYPE)
        poses = np.ascontiguousarray(poses, POSE_DTYPE)
        counts = np.ascontiguousarray(counts, np.int32)
        if dets.ndim != 3 or dets.shape != poses.shape or dets.shape[0] != self.nstreams or counts.shape != dets.shape[:2]:
            raise ValueError("expected dets/poses [S, F, cap] and counts [S, F] with S = %d" % self.nstreams)
        S, F, cap = dets.shape
        my_pose = np.zeros((S, F, 4, 4), np.float64)
        valid = np.zeros((S, F), np.uint8)
        self._det._check(self._L.agpu_graph_update(self._g, F, dets.ctypes.data, poses.ctypes.data, counts.ctypes.data, cap,
                                                   my_pose.ctypes.data, valid.ctypes.data))
        return my_pose, valid.astype(bool)

    def update_lists(self, det_lists, pose_lists):
        """One frame per stream from per-stream record lists (what Detector.detect_pose_batch returns per frame)."""
        S = self.nstreams
        if len(det_lists) != S or len(pose_lists) != S:
            raise ValueError("one detection list and one pose list per stream expected")
        cap = max(1, max(len(d) for d in det_lists))
        dets = np.zeros((S, 1, cap), DET_DTYPE)
        poses = np.zeros((S, 1, cap), POSE_DTYPE)
        counts = np.zeros((S, 1), np.int32)
        for s in range(S):
            n = len(det_lists[s])
            dets[s, 0, :n] = det_lists[s]
            poses[s, 0, :n] = pose_lists[s]
            counts[s, 0] = n
        mp, ok = self.update(dets, poses, counts)
        return [mp[s, 0] if ok[s, 0] else None for s in range(S)]

    def state(self, stream: int = 0) -> dict:
        n = self.max_tag_id + 1
        coord, skipped = np.zeros(1, np.int32), np.zeros(1, np.int32)
        est = np.zeros((4, 4))
        present, updated, visible = (np.zeros(n, np.uint8) for _ in range(3))
        reference, weight = np.zeros(n, np.int32), np.zeros(n, np.int32)
        local, world = np.zeros((n, 4, 4)), np.zeros((n, 4, 4))
        self._det._check(self._L.agpu_graph_get(self._g, stream, coord.ctypes.data, est.ctypes.data, present.ctypes.data,
                                                reference.ctypes.data, weight.ctypes.data, updated.ctypes.data,
                                                visible.ctypes.data, local.ctypes.data, world.ctypes.data, skipped.ctypes.data))
        return {"coordinate_id": int(coord[0]), "estimated_pose": est, "present": present.astype(bool),
                "reference": reference, "weight": weight, "updated": updated.astype(bool), "visible": visible.astype(bool),
                "local": local, "world": world, "skipped": int(skipped[0])}

    # --- the reference's accessors (slam_graph.py:80-90), for stream 0 unless told otherwise
    def get_nodes(self, stream: int = 0) -> dict:
        st = self.state(stream)
        return {int(i): Node(st["local"][i], st["world"][i], int(st["reference"][i]), int(st["weight"][i]),
                             bool(st["updated"][i]), bool(st["visible"][i])) for i in np.nonzero(st["present"])[0]}

    def get_coordinate_id(self, stream: int = 0) -> int:
        return self.state(stream)["coordinate_id"]

    def get_estimated_pose(self, stream: int = 0) -> np.ndarray:
        return self.state(stream)["estimated_pose"]


def transforms_to_records(ids, T, ok: Optional[np.ndarray] = None):
    """Helper for callers that hold 4x4 transforms instead of pose records: -> (dets, poses) structured arrays."""
    ids = np.asarray(ids, np.int32)
    T = np.asarray(T, np.float64).reshape(ids.shape + (4, 4))
    dets = np.zeros(ids.shape, DET_DTYPE)
    poses = np.zeros(ids.shape, POSE_DTYPE)
    dets["id"] = ids
    poses["R"] = T[..., :3, :3].reshape(ids.shape + (9,))
    poses["tvec"] = T[..., :3, 3]
    poses["ok"] = 1 if ok is None else np.asarray(ok, np.int32)
    return dets, poses
