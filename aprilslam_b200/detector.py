"""Host-side mirror of the reference's detector interface, on top of the C ABI.

  * `apriltag`      -- drop-in for upstream's Python class that the reference constructs at
                       /root/reference/src/detection/tag_detector.py:18 and calls at :26
                       (same ctor keywords as apriltag_pywrap.c, same tuple-of-dicts result).
  * `Detector`      -- the batched entry: uint8 [B,H,W] (numpy or CUDA torch tensor) -> per-frame records.
  * `TagDetector`   -- same class name / methods / return shapes as the reference's
                       src/detection/tag_detector.py:14-52 (detect, get_pose, transformation), with
                       the gray conversion, detection and solvePnP running on the B200.

torch is optional and only used to recognise CUDA tensors (data_ptr / stream); all computation is in
libaprilgpu.so.  Nothing here falls back to a CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import sys
from collections.abc import Sequence as _SequenceABC
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import DET_DTYPE, POSE_DTYPE, STAGE_NAMES, AgpuConfig

KNOWN_FAMILIES = ("tag36h11", "tag25h9", "tag16h5", "tagStandard41h12")


class FrameLists(_SequenceABC):
    """Per-frame record arrays of one batch call: element b is a view `buf[b, :n[b]]`, created on demand (a batch of
    1024 frames would otherwise cost a millisecond of Python just to slice).  Behaves like a list of arrays."""
    __slots__ = ("_buf", "_n")

    def __init__(self, buf: np.ndarray, n: np.ndarray):
        self._buf, self._n = buf, n

    def __len__(self) -> int:
        return int(self._buf.shape[0])

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        i = int(i)
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError("frame index out of range")
        return self._buf[i, :int(self._n[i])]

    def counts(self) -> np.ndarray:
        """Number of records per frame."""
        return self._n

    def __repr__(self):
        return "FrameLists(%d frames, %d records)" % (len(self), int(self._n.sum()))


def _calibrate_pool_idle_refs():
    """Reference count `sys.getrefcount(pool[i])` reports for an array that only a pool list refers to, measured on this
    interpreter instead of assumed; None (no pooling: every call gets a fresh buffer) where reference counts do not exist
    or do not move when a second reference appears."""
    getref = getattr(sys, "getrefcount", None)
    if getref is None:
        return None
    pool = [np.empty((1, 1), np.uint8)]
    idle = getref(pool[0])
    extra = pool[0][0, :1]          # a view, as FrameLists hands out: must raise the count
    busy = getref(pool[0])
    del extra
    return idle if busy > idle and getref(pool[0]) == idle else None


_POOL_IDLE_REFS = _calibrate_pool_idle_refs()


def family_code_counts(family: str):
    """(code words shipped, code words of upstream's family table) -- they differ for tagStandard41h12 (ids 0..4 only)."""
    L = _lib.load()
    a, b = C.c_int(), C.c_int()
    if L.agpu_family_info(family.encode(), C.byref(a), C.byref(b)) != 0:
        return 0, 0
    return a.value, b.value


def _is_torch_cuda(x) -> bool:
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda") and bool(x.is_cuda)


class Detector:
    """Batched B200 detector.  One instance = one C handle = one GPU + one stream set (not thread-safe)."""

    def __init__(self, families: Sequence[str] | str = "tag36h11", threads: int = 1, maxhamming: int = 1,
                 decimate: float = 2.0, blur: float = 0.0, refine_edges: bool = True, debug: bool = False,
                 decode_sharpening: float = 0.25, device: int = 0, chunk_frames: int = 0, pipeline_slots: int = 0,
                 max_points_per_frame: int = 0, max_clusters_per_frame: int = 0, max_quads_per_frame: int = 0):
        if not isinstance(families, str):
            families = " ".join(families)
        self.families = families.replace(",", " ").split()
        self.decimate = float(decimate)
        self._L = _lib.load()
        self._pool = {}   # (B, cap, dtype) -> result buffers, see _result_buffer
        cfg = AgpuConfig()
        self._L.agpu_default_config(C.byref(cfg))
        self._fam_bytes = " ".join(self.families).encode()
        cfg.families = self._fam_bytes
        cfg.threads = int(threads)
        cfg.maxhamming = int(maxhamming)
        cfg.quad_decimate = float(decimate)
        cfg.quad_sigma = float(blur)
        cfg.refine_edges = int(bool(refine_edges))
        cfg.decode_sharpening = float(decode_sharpening)
        cfg.debug = int(bool(debug))
        cfg.device = int(device)
        cfg.chunk_frames = int(chunk_frames)
        cfg.pipeline_slots = int(pipeline_slots)
        cfg.max_points_per_frame = int(max_points_per_frame)
        cfg.max_clusters_per_frame = int(max_clusters_per_frame)
        cfg.max_quads_per_frame = int(max_quads_per_frame)
        h = C.c_void_p()
        rc = self._L.agpu_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise RuntimeError((self._L.agpu_last_error(None) or b"agpu_create failed").decode())
        self._h = h
        self.device = int(device)
        self._graphs = []     # weak references to the tag graphs that live on this handle (closed before it)
        for f in self.families:
            have, full = family_code_counts(f)
            if have < full:
                import warnings
                warnings.warn("%s: only %d of upstream's %d code words are available offline (ids 0..%d); tags with other "
                              "ids will not be reported" % (f, have, full, have - 1), RuntimeWarning, stacklevel=2)

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            for ref in getattr(self, "_graphs", []):      # a graph holds the C handle: close it first
                g = ref()
                if g is not None:
                    g.close()
            self._graphs = []
            self._L.agpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, allow_truncated: bool = False):
        if rc == 0 or (allow_truncated and rc == _lib.AGPU_E_TRUNCATED):
            return
        msg = (self._L.agpu_last_error(self._h) or b"").decode()
        raise RuntimeError("libaprilgpu error %d: %s" % (rc, msg))

    # -- input plumbing -----------------------------------------------------------------------
    def _frames(self, frames, channels: int):
        """-> (ptr, on_device, B, W, H, stride_bytes, stream_ptr, keepalive)"""
        if _is_torch_cuda(frames):
            import torch
            t = frames
            if t.dtype != torch.uint8:
                raise RuntimeError("expected uint8 frames")
            if t.dim() == (2 if channels == 1 else 3):
                t = t.unsqueeze(0)
            if t.dim() != (3 if channels == 1 else 4) or (channels == 3 and t.shape[-1] != 3):
                raise RuntimeError("expected frames of shape [B,H,W]" + (",3" if channels == 3 else ""))
            if not t.is_contiguous():
                t = t.contiguous()
            if t.device.index != self.device:
                raise RuntimeError("frames live on cuda:%s but the detector is bound to cuda:%d" % (t.device.index, self.device))
            B, H, W = int(t.shape[0]), int(t.shape[1]), int(t.shape[2])
            stream = torch.cuda.current_stream(t.device).cuda_stream
            return t.data_ptr(), 1, B, W, H, W * channels, stream, t
        a = np.asarray(frames)
        if a.dtype != np.uint8:
            raise RuntimeError("expected uint8 frames")
        if a.ndim == (2 if channels == 1 else 3):
            a = a[None]
        if a.ndim != (3 if channels == 1 else 4) or (channels == 3 and a.shape[-1] != 3):
            raise RuntimeError("expected frames of shape [B,H,W]" + (",3" if channels == 3 else ""))
        a = np.ascontiguousarray(a)
        B, H, W = a.shape[:3]
        return a.ctypes.data, 0, B, W, H, W * channels, None, a

    # -- result buffers -----------------------------------------------------------------------
    def _result_buffer(self, B: int, cap: int, dtype) -> np.ndarray:
        """[B, cap] record array for one call.  A fresh 10-20 MB array per call costs ~2 ms of page faults on a 1024-frame
        batch (10 % of the whole step), so buffers are pooled and one is handed out again once NOTHING outside the pool
        references it any more (the per-frame views a caller still holds keep their buffer alive and out of
        circulation -- results never change under a caller's feet)."""
        key = (B, cap, dtype.str if hasattr(dtype, "str") else str(dtype))
        pool = self._pool.setdefault(key, [])
        getref = getattr(sys, "getrefcount", None)
        if getref is not None and _POOL_IDLE_REFS is not None:
            for i in range(len(pool)):
                if getref(pool[i]) == _POOL_IDLE_REFS:     # nothing but the pool refers to it (count calibrated at import)
                    return pool[i]
        buf = np.empty((B, cap), dtype)
        if getref is not None and _POOL_IDLE_REFS is not None and len(pool) < 4:
            pool.append(buf)
        return buf

    # -- detection ----------------------------------------------------------------------------
    def detect_batch(self, frames, cap_per_frame: int = 64, bgr: bool = False) -> FrameLists:
        """frames: uint8 [B,H,W] (gray) or [B,H,W,3] (bgr=True) -> per-frame DET_DTYPE record arrays (a list-like)."""
        ch = 3 if bgr else 1
        ptr, on_dev, B, W, H, stride, stream, keep = self._frames(frames, ch)
        self._last_width = W
        out = self._result_buffer(B, cap_per_frame, DET_DTYPE)
        counts = np.zeros(B, np.int32)
        fn = self._L.agpu_detect_bgr if bgr else self._L.agpu_detect
        rc = fn(self._h, ptr, on_dev, B, W, H, stride, stream, out.ctypes.data, cap_per_frame, counts.ctypes.data)
        self._check(rc, allow_truncated=True)
        return FrameLists(out, np.minimum(counts, cap_per_frame))

    def detect_pose_batch(self, frames, camera_matrix, dist_coeffs, tag_size: float, cap_per_frame: int = 64,
                          bgr: bool = False) -> Tuple[FrameLists, FrameLists]:
        """Detection + per-tag pose in one pass: (detections, poses) per frame (POSE_DTYPE records)."""
        ch = 3 if bgr else 1
        ptr, on_dev, B, W, H, stride, stream, keep = self._frames(frames, ch)
        out = self._result_buffer(B, cap_per_frame, DET_DTYPE)
        poses = self._result_buffer(B, cap_per_frame, POSE_DTYPE)
        counts = np.zeros(B, np.int32)
        K = np.ascontiguousarray(np.asarray(camera_matrix, np.float64).reshape(3, 3))
        dist = np.ascontiguousarray(np.asarray(dist_coeffs if dist_coeffs is not None else [], np.float64).ravel())
        rc = self._L.agpu_detect_pose(self._h, ptr, on_dev, ch, B, W, H, stride, stream, K.ctypes.data,
                                      dist.ctypes.data if dist.size else None, int(dist.size), float(tag_size),
                                      out.ctypes.data, poses.ctypes.data, cap_per_frame, counts.ctypes.data)
        self._check(rc, allow_truncated=True)
        n = np.minimum(counts, cap_per_frame)
        return FrameLists(out, n), FrameLists(poses, n)

    def estimate_pose(self, corners, camera_matrix, dist_coeffs, tag_size: float, method: int = 0) -> np.ndarray:
        """corners [M,4,2] (lb, rb, rt, lt) -> POSE_DTYPE[M] (ok, rvec, tvec, R)."""
        c = np.ascontiguousarray(np.asarray(corners, np.float64).reshape(-1, 4, 2))
        M = c.shape[0]
        K = np.ascontiguousarray(np.asarray(camera_matrix, np.float64).reshape(3, 3))
        dist = np.ascontiguousarray(np.asarray(dist_coeffs if dist_coeffs is not None else [], np.float64).ravel())
        poses = np.zeros(M, POSE_DTYPE)
        rc = self._L.agpu_pose(self._h, c.ctypes.data, M, K.ctypes.data, dist.ctypes.data if dist.size else None,
                               int(dist.size), float(tag_size), int(method), poses.ctypes.data)
        self._check(rc)
        return poses

    # -- instrumentation ------------------------------------------------------------------------
    def set_profiling(self, on: bool = True):
        self._check(self._L.agpu_set_profiling(self._h, int(on)))

    def kernel_ms(self, kernel: str = "k_cc_local") -> float:
        """CUDA-event time of one kernel summed over the chunks of the last call (profiling on)."""
        ms = np.zeros(1, np.float32)
        self._check(self._L.agpu_get_kernel_ms(self._h, kernel.encode(), ms.ctypes.data))
        return float(ms[0])

    def kernel_table(self) -> dict:
        """{kernel name: (milliseconds, launches)} of the last call (profiling on): every launch has its own event pair."""
        n = int(self._L.agpu_get_kernel_table(self._h, None, 0))
        buf = C.create_string_buffer(max(n, 1))
        self._L.agpu_get_kernel_table(self._h, buf, n)
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, cnt = line.split("\t")
            out[name] = (float(ms), int(cnt))
        return out

    def stage_ms(self) -> dict:
        ms = np.zeros(len(STAGE_NAMES), np.float32)
        self._check(self._L.agpu_get_stage_ms(self._h, ms.ctypes.data))
        return dict(zip(STAGE_NAMES, (float(v) for v in ms)))

    def timeline(self) -> np.ndarray:
        """Per chunk of the last call (profiling on): [first frame, frames, slot, 10 stage boundary marks in ms]."""
        n = int(self._L.agpu_get_timeline(self._h, None, 0))
        buf = np.zeros(max(n, 1), np.float32)
        self._L.agpu_get_timeline(self._h, buf.ctypes.data, n)
        return buf[:n].reshape(-1, len(STAGE_NAMES) + 4)

    def launch_count(self) -> int:
        v = C.c_longlong()
        self._check(self._L.agpu_get_launch_count(self._h, C.byref(v)))
        return int(v.value)

    def counters(self) -> dict:
        c = np.zeros(8, np.int64)
        self._check(self._L.agpu_get_counters(self._h, c.ctypes.data))
        return dict(edge_points=int(c[0]), clusters=int(c[1]), quads=int(c[2]), raw_detections=int(c[3]),
                    oversize_clusters=int(c[4]), tier_clusters=(int(c[1] - c[5] - c[6] - c[7]), int(c[5]), int(c[6]), int(c[7])))

    def tier_stats(self) -> dict:
        """Clusters and edge-point records handed to the four quad-fit size tiers in the last call."""
        c = np.zeros(8, np.int64)
        self._check(self._L.agpu_get_tier_stats(self._h, c.ctypes.data))
        return dict(clusters=[int(v) for v in c[:4]], records=[int(v) for v in c[4:]])

    def debug_fetch(self, what: str, frame: int = 0) -> np.ndarray:
        wd, hd = C.c_int(), C.c_int()
        self._check(self._L.agpu_debug_dims(self._h, C.byref(wd), C.byref(hd)))
        dt = {"gray": np.uint8, "quad_im": np.uint8, "thresh": np.uint8, "labels": np.uint32, "sizes": np.uint32,
              "cluster_keys": np.uint64, "cluster_sizes": np.int32, "quads": np.float32, "quads_refined": np.float32,
              "quad_keys": np.uint64}[what]
        per = {"quads": 9, "quads_refined": 8}.get(what, 1)
        n = self._L.agpu_debug_fetch(self._h, what.encode(), frame, np.zeros(1, np.uint8).ctypes.data, 0)
        if n < 0:
            self._check(int(n))
        buf = np.zeros(max(int(n) * per, 1), dt)
        n2 = self._L.agpu_debug_fetch(self._h, what.encode(), frame, buf.ctypes.data, buf.nbytes)
        if n2 < 0:
            self._check(int(n2))
        buf = buf[:int(n2) * per]
        if what == "gray":
            return buf.reshape(-1, self._last_width) if getattr(self, "_last_width", 0) else buf
        if what in ("quad_im", "thresh", "labels", "sizes"):
            return buf.reshape(hd.value, wd.value)
        if per > 1:
            return buf.reshape(-1, per)
        return buf

    def stage_threshold(self, im: np.ndarray):
        """(decimated image, threshold image) through the pipeline's own image kernels."""
        im = np.ascontiguousarray(im, np.uint8)
        H, W = im.shape
        f = max(1, int(self.decimate))
        wd, hd = 1 + (W - 1) // f, 1 + (H - 1) // f
        q = np.zeros((hd, wd), np.uint8)
        t = np.zeros((hd, wd), np.uint8)
        self._check(self._L.agpu_stage_threshold(self._h, im.ctypes.data, W, H, q.ctypes.data, t.ctypes.data))
        return q, t

    def stage_labels(self, thresh: np.ndarray):
        thresh = np.ascontiguousarray(thresh, np.uint8)
        H, W = thresh.shape
        lab = np.zeros((H, W), np.uint32)
        sz = np.zeros((H, W), np.uint32)
        self._check(self._L.agpu_stage_labels(self._h, thresh.ctypes.data, W, H, lab.ctypes.data, sz.ctypes.data))
        return lab, sz


def records_to_dicts(recs: np.ndarray, families: Sequence[str]) -> tuple:
    """Upstream pywrap dict ('hamming','margin','id','center','lb-rb-rt-lt') plus the pip-API / north-star
    names ('tag_family','tag_id','decision_margin','corners','homography')."""
    out = []
    for r in recs:
        corners = np.array(r["p"], dtype=np.float64)
        d = {"hamming": int(r["hamming"]), "margin": float(r["margin"]), "id": int(r["id"]),
             "center": np.array(r["c"], dtype=np.float64), "lb-rb-rt-lt": corners,
             "tag_family": families[int(r["family"])], "tag_id": int(r["id"]),
             "decision_margin": float(r["margin"]), "corners": corners,
             "homography": np.array(r["H"], dtype=np.float64).reshape(3, 3)}
        out.append(d)
    return tuple(out)


class apriltag:  # noqa: N801  (upstream's class name; tag_detector.py:11 does `from apriltag import apriltag`)
    """Drop-in for upstream's Python wrapper class: apriltag(family, threads=1, maxhamming=1, decimate=2.0,
    blur=0.0, refine_edges=True, debug=False).detect(gray) -> tuple of dicts."""

    def __init__(self, family, threads=1, maxhamming=1, decimate=2.0, blur=0.0, refine_edges=True, debug=False,
                 device=0):
        fams = family.split() if isinstance(family, str) else list(family)
        for f in fams:
            if f not in KNOWN_FAMILIES:
                raise RuntimeError("Unrecognized tag family name: %s. Use e.g. \"tag36h11\"." % f)
        self.families = fams
        self._det = Detector(fams, threads=threads, maxhamming=maxhamming, decimate=decimate, blur=blur,
                             refine_edges=refine_edges, debug=False, device=device)

    def detect(self, image):
        if _is_torch_cuda(image):
            if image.dim() != 2:
                raise RuntimeError("detect() expects a 2-D uint8 image")
        else:
            image = np.asarray(image)
            if image.ndim != 2 or image.dtype != np.uint8:
                raise RuntimeError("detect() expects a 2-D uint8 numpy array")
        recs = self._det.detect_batch(image, cap_per_frame=256)[0]
        return records_to_dicts(recs, self.families)


class TagDetector:
    """Same interface as the reference's TagDetector (src/detection/tag_detector.py:14-52)."""

    def __init__(self, camera_params, tag_type="tagStandard41h12", tag_size=0.06, device=0, **detector_kwargs):
        self.detector = Detector(tag_type, device=device, **detector_kwargs)
        self.families = self.detector.families
        self.tag_size = tag_size
        self.camera_matrix = camera_params["camera_matrix"]
        self.dist_coeffs = camera_params["dist_coeffs"]

    def detect(self, image):
        """BGR frame -> detections sorted by id (tag_detector.py:23-28); gray conversion fused on the GPU."""
        image = image if _is_torch_cuda(image) else np.asarray(image)
        bgr = image.ndim == 3
        recs = self.detector.detect_batch(image, cap_per_frame=256, bgr=bgr)[0]
        return list(records_to_dicts(recs, self.families))  # already sorted by id

    def get_pose(self, detection):
        """-> (retval, rvec (3,1), tvec (3,1), T 4x4)  (tag_detector.py:30-43)"""
        p = self.detector.estimate_pose(np.asarray(detection["lb-rb-rt-lt"], np.float64)[None], self.camera_matrix,
                                        self.dist_coeffs, self.tag_size)[0]
        rvec = np.array(p["rvec"], np.float64).reshape(3, 1)
        tvec = np.array(p["tvec"], np.float64).reshape(3, 1)
        T = np.eye(4)
        T[:3, :3] = np.array(p["R"]).reshape(3, 3)
        T[:3, 3] = tvec.ravel()
        return bool(p["ok"]), rvec, tvec, T

    def transformation(self, rvec, tvec):
        """Rodrigues(rvec) and tvec -> 4x4 (tag_detector.py:45-52); tiny host math."""
        r = np.asarray(rvec, np.float64).ravel()
        th = float(np.linalg.norm(r))
        T = np.eye(4)
        if th > 0:
            k = r / th
            Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
            T[:3, :3] = np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * Kx
        T[:3, 3] = np.asarray(tvec, np.float64).ravel()
        return T
