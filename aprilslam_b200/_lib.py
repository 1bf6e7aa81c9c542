"""ctypes loader for libaprilgpu.so (the C ABI of include/aprilgpu.h).

There is no CPU fallback: if the library is missing or no B200 is usable, loading / creating a
detector raises.  `build()` compiles the library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AGPU_LIB") or os.path.join(_HERE, "libaprilgpu.so")   # (AGPU_LIB: an experimental build)
CSRC = os.path.join(_HERE, "csrc")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false",  # no FMA contraction: float decisions must match the CPU oracle bit for bit
              "--shared", "-Xcompiler", "-fPIC"]

AGPU_OK, AGPU_E_INVALID, AGPU_E_CUDA, AGPU_E_TRUNCATED, AGPU_E_WORKSPACE, AGPU_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5
STAGE_NAMES = ["h2d", "image", "cc", "edges", "sort", "quads", "decode", "reconcile_pose", "d2h"]


class AgpuConfig(C.Structure):
    _fields_ = [("families", C.c_char_p), ("threads", C.c_int), ("maxhamming", C.c_int),
                ("quad_decimate", C.c_float), ("quad_sigma", C.c_float), ("refine_edges", C.c_int),
                ("decode_sharpening", C.c_double), ("debug", C.c_int), ("device", C.c_int),
                ("chunk_frames", C.c_int), ("pipeline_slots", C.c_int), ("max_points_per_frame", C.c_int),
                ("max_clusters_per_frame", C.c_int), ("max_quads_per_frame", C.c_int)]


DET_DTYPE = np.dtype([("family", "<i4"), ("id", "<i4"), ("hamming", "<i4"), ("margin", "<f4"),
                      ("c", "<f8", (2,)), ("p", "<f8", (4, 2)), ("H", "<f8", (9,))])
POSE_DTYPE = np.dtype([("rvec", "<f8", (3,)), ("tvec", "<f8", (3,)), ("R", "<f8", (9,)), ("err", "<f8"),
                       ("ok", "<i4"), ("iters", "<i4")])
assert DET_DTYPE.itemsize == 168 and POSE_DTYPE.itemsize == 136


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libaprilgpu.so in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    srcs = sources() + [os.path.join(_HERE, "..", "include", "aprilgpu.h")]
    if not force and os.path.exists(LIB_PATH) and all(
            os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs if os.path.exists(s)):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB_PATH, os.path.join(CSRC, "aprilgpu.cu")]
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def load():
    """Load the shared library and declare every entry point of include/aprilgpu.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libaprilgpu.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                           "aprilslam_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    L.agpu_version.restype = ci
    L.agpu_family_info.argtypes = [C.c_char_p, vp, vp]
    L.agpu_default_config.argtypes = [C.POINTER(AgpuConfig)]
    L.agpu_default_config.restype = None
    L.agpu_create.argtypes = [C.POINTER(AgpuConfig), C.POINTER(vp)]
    L.agpu_destroy.argtypes = [vp]
    L.agpu_last_error.argtypes = [vp]
    L.agpu_last_error.restype = C.c_char_p
    L.agpu_detect.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp, vp, ci, vp]
    L.agpu_detect_bgr.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp, vp, ci, vp]
    L.agpu_detect_pose.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, vp, vp, vp, ci, cd, vp, vp, ci, vp]
    L.agpu_pose.argtypes = [vp, vp, ci, vp, vp, ci, cd, ci, vp]
    L.agpu_set_profiling.argtypes = [vp, ci]
    L.agpu_get_stage_ms.argtypes = [vp, vp]
    L.agpu_get_kernel_ms.argtypes = [vp, C.c_char_p, vp]
    L.agpu_get_kernel_table.argtypes = [vp, vp, ci]
    L.agpu_get_launch_count.argtypes = [vp, vp]
    L.agpu_get_counters.argtypes = [vp, vp]
    L.agpu_get_tier_stats.argtypes = [vp, vp]
    L.agpu_debug_fetch.argtypes = [vp, C.c_char_p, ci, vp, C.c_longlong]
    L.agpu_debug_fetch.restype = C.c_longlong
    L.agpu_debug_dims.argtypes = [vp, vp, vp]
    L.agpu_stage_threshold.argtypes = [vp, vp, ci, ci, vp, vp]
    L.agpu_stage_labels.argtypes = [vp, vp, ci, ci, vp, vp]
    L.agpu_render.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp, vp]
    L.agpu_get_timeline.argtypes = [vp, vp, ci]
    L.agpu_graph_create.argtypes = [vp, ci, ci, C.POINTER(vp)]
    L.agpu_graph_reset.argtypes = [vp]
    L.agpu_graph_destroy.argtypes = [vp]
    L.agpu_graph_update.argtypes = [vp, ci, vp, vp, vp, ci, vp, vp]
    L.agpu_graph_get.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    _lib = L
    return L


EXPORTS = ["agpu_version", "agpu_family_info", "agpu_default_config", "agpu_create", "agpu_destroy", "agpu_last_error", "agpu_detect",
           "agpu_detect_bgr", "agpu_detect_pose", "agpu_pose", "agpu_set_profiling", "agpu_get_stage_ms",
           "agpu_get_kernel_ms", "agpu_get_kernel_table", "agpu_get_timeline", "agpu_get_launch_count", "agpu_get_counters", "agpu_get_tier_stats", "agpu_debug_fetch", "agpu_debug_dims",
           "agpu_stage_threshold", "agpu_stage_labels", "agpu_render", "agpu_graph_create", "agpu_graph_reset",
           "agpu_graph_destroy", "agpu_graph_update", "agpu_graph_get"]
