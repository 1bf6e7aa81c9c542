"""ctypes binding of oracle/libapriltag_oracle.so (test infrastructure; see apriltag_oracle.cpp).

The detector restated here is the native call behind
/root/reference/src/detection/tag_detector.py:18,26; the pose oracle is the reference's own
call, cv2.solvePnP + cv2.Rodrigues (tag_detector.py:30-52), restated in `reference_pose`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

from aprilslam_b200.families_data import FAMILIES  # generated code-book data (tools/gen_codebooks.py)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libapriltag_oracle.so")


class AoDetection(C.Structure):
    _fields_ = [("family", C.c_int32), ("id", C.c_int32), ("hamming", C.c_int32), ("margin", C.c_float),
                ("c", C.c_double * 2), ("p", (C.c_double * 2) * 4), ("H", C.c_double * 9)]


class AoDebug(C.Structure):
    _fields_ = [("quad_im", C.c_void_p), ("thresh", C.c_void_p), ("labels", C.c_void_p), ("sizes", C.c_void_p),
                ("wd", C.c_int), ("hd", C.c_int), ("npoints", C.c_int), ("nclusters", C.c_int),
                ("cluster_keys", C.c_void_p), ("cluster_sizes", C.c_void_p), ("cap_clusters", C.c_int),
                ("nquads", C.c_int), ("quads", C.c_void_p), ("quads_refined", C.c_void_p),
                ("quad_keys", C.c_void_p), ("cap_quads", C.c_int), ("noversize", C.c_int)]


DET_DTYPE = np.dtype([("family", "<i4"), ("id", "<i4"), ("hamming", "<i4"), ("margin", "<f4"),
                      ("c", "<f8", (2,)), ("p", "<f8", (4, 2)), ("H", "<f8", (9,))])
assert DET_DTYPE.itemsize == C.sizeof(AoDetection)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "apriltag_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ao_create.restype = C.c_void_p
        L.ao_create.argtypes = [C.c_float, C.c_float, C.c_int, C.c_double, C.c_int]
        L.ao_destroy.argtypes = [C.c_void_p]
        L.ao_add_family.argtypes = [C.c_void_p, C.c_char_p] + [C.c_int] * 6 + [C.c_void_p] * 3
        L.ao_detect.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.ao_detect_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_int, C.c_void_p]
        L.ao_stage_threshold.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ao_stage_labels.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ao_stage_blur.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float]
        L.ao_rotate90.restype = C.c_uint64
        L.ao_rotate90.argtypes = [C.c_uint64, C.c_int]
        _lib = L
    return _lib


class OracleDetector:
    """Same constructor keywords as upstream's Python wrapper: apriltag(family, threads, maxhamming, ...)."""

    def __init__(self, families: Sequence[str] | str = "tag36h11", threads: int = 1, maxhamming: int = 1,
                 decimate: float = 2.0, blur: float = 0.0, refine_edges: bool = True, debug: bool = False,
                 decode_sharpening: float = 0.25):
        if isinstance(families, str):
            families = families.split()
        self.families = list(families)
        self.threads = threads
        self.decimate = decimate
        L = lib()
        self._h = L.ao_create(decimate, blur, int(bool(refine_edges)), decode_sharpening, maxhamming)
        for name in self.families:
            if name not in FAMILIES:
                raise RuntimeError("Unrecognized tag family name: %s" % name)
            f = FAMILIES[name]
            codes = np.array(f["codes"], np.uint64)
            bx = np.array(f["bit_x"], np.int32)
            by = np.array(f["bit_y"], np.int32)
            L.ao_add_family(self._h, name.encode(), f["nbits"], f["h"], len(codes), f["width_at_border"],
                            f["total_width"], f["reversed_border"], codes.ctypes.data, bx.ctypes.data, by.ctypes.data)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ao_destroy(self._h)
            self._h = None

    def detect_records(self, gray: np.ndarray, debug: bool = False, cap: int = 1024):
        if gray.ndim != 2 or gray.dtype != np.uint8:
            raise RuntimeError("expected a 2-D uint8 image")
        gray = np.ascontiguousarray(gray)
        h, w = gray.shape
        out = np.zeros(cap, DET_DTYPE)
        dbg = None
        keep = {}
        if debug:
            f = max(1, int(self.decimate))
            wd, hd = 1 + (w - 1) // f, 1 + (h - 1) // f
            keep = dict(quad_im=np.zeros((hd, wd), np.uint8), thresh=np.zeros((hd, wd), np.uint8),
                        labels=np.zeros((hd, wd), np.uint32), sizes=np.zeros((hd, wd), np.uint32),
                        cluster_keys=np.zeros(1 << 20, np.uint64), cluster_sizes=np.zeros(1 << 20, np.int32),
                        quads=np.zeros((4096, 9), np.float32), quads_refined=np.zeros((4096, 8), np.float32),
                        quad_keys=np.zeros(4096, np.uint64))
            dbg = AoDebug()
            for k, v in keep.items():
                setattr(dbg, k, v.ctypes.data)
            dbg.cap_clusters = 1 << 20
            dbg.cap_quads = 4096
        n = lib().ao_detect(self._h, gray.ctypes.data, w, h, gray.strides[0], out.ctypes.data, cap,
                            C.byref(dbg) if dbg is not None else None)
        recs = out[:min(n, cap)].copy()
        if not debug:
            return recs
        nc, nq = dbg.nclusters, dbg.nquads
        keep["cluster_keys"] = keep["cluster_keys"][:nc].copy()
        keep["cluster_sizes"] = keep["cluster_sizes"][:nc].copy()
        for k in ("quads", "quads_refined", "quad_keys"):
            keep[k] = keep[k][:nq].copy()
        keep["npoints"] = dbg.npoints
        keep["noversize"] = dbg.noversize
        return recs, keep

    def detect(self, gray: np.ndarray):
        """Upstream pywrap result shape: tuple of dicts (tag_detector.py:26-27,32 index 'id', 'lb-rb-rt-lt')."""
        recs = self.detect_records(gray)
        return tuple({"hamming": int(r["hamming"]), "margin": float(r["margin"]), "id": int(r["id"]),
                      "center": r["c"].copy(), "lb-rb-rt-lt": r["p"].copy()} for r in recs)

    def detect_batch(self, frames: np.ndarray, nthreads: Optional[int] = None, cap: int = 256):
        assert frames.ndim == 3 and frames.dtype == np.uint8 and frames.flags.c_contiguous
        B, h, w = frames.shape
        out = np.zeros((B, cap), DET_DTYPE)
        counts = np.zeros(B, np.int32)
        lib().ao_detect_batch(self._h, frames.ctypes.data, B, w, h, w, nthreads or os.cpu_count() or 1,
                              out.ctypes.data, cap, counts.ctypes.data)
        return [out[b, :min(int(counts[b]), cap)].copy() for b in range(B)]


def stage_threshold(im: np.ndarray, min_diff: int = 5) -> np.ndarray:
    im = np.ascontiguousarray(im)
    out = np.zeros_like(im)
    lib().ao_stage_threshold(im.ctypes.data, im.shape[1], im.shape[0], min_diff, out.ctypes.data)
    return out


def stage_labels(thr: np.ndarray):
    thr = np.ascontiguousarray(thr)
    labels = np.zeros(thr.shape, np.uint32)
    sizes = np.zeros(thr.shape, np.uint32)
    lib().ao_stage_labels(thr.ctypes.data, thr.shape[1], thr.shape[0], labels.ctypes.data, sizes.ctypes.data)
    return labels, sizes


def stage_blur(im: np.ndarray, sigma: float) -> np.ndarray:
    out = np.ascontiguousarray(im).copy()
    lib().ao_stage_blur(out.ctypes.data, out.shape[1], out.shape[0], sigma)
    return out


def reference_pose(corners_lb_rb_rt_lt, K, dist, tag_size):
    """The reference's pose path, verbatim in behaviour: tag_detector.py:30-52."""
    import cv2
    corners = np.array(corners_lb_rb_rt_lt, dtype=np.float32)
    s = tag_size
    obj = np.array([[-s / 2, -s / 2, 0], [s / 2, -s / 2, 0], [s / 2, s / 2, 0], [-s / 2, s / 2, 0]], dtype=np.float32)
    retval, rvec, tvec = cv2.solvePnP(obj, corners, np.asarray(K, float), np.asarray(dist, float))
    R, _ = cv2.Rodrigues(rvec)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = tvec.flatten()
    return retval, rvec, tvec, T
