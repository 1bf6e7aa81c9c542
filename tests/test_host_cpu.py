"""CPU tests of the host layer: the C-ABI library loads and exports what include/aprilgpu.h declares,
there is no CPU fallback, the reference-facing shim behaves like the reference's import, and the
frame-sharding plumbing works at world_size 2 (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    from aprilslam_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "aprilgpu.h")).read()
    declared = set(re.findall(r"\b(agpu_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"agpu_status"}
    L = _lib.load()
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert set(_lib.EXPORTS) == declared
    assert L.agpu_version() == 100


def test_struct_layouts_match_the_header():
    from aprilslam_b200 import _lib
    assert _lib.DET_DTYPE.itemsize == 4 * 4 + 8 * (2 + 8 + 9)      # agpu_detection
    assert _lib.POSE_DTYPE.itemsize == 8 * (3 + 3 + 9 + 1) + 8      # agpu_pose_t
    cfg = _lib.AgpuConfig()
    _lib.load().agpu_default_config(ctypes.byref(cfg))
    # defaults = what AprilSLAM gets from apriltag(tag_type) (tag_detector.py:18): decimate 2, refine, maxhamming 1
    assert (cfg.threads, cfg.maxhamming, cfg.quad_decimate, cfg.quad_sigma, cfg.refine_edges) == (1, 1, 2.0, 0.0, 1)
    assert cfg.decode_sharpening == 0.25


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from aprilslam_b200.detector import Detector, apriltag
    with pytest.raises(RuntimeError, match="CUDA"):
        Detector("tag36h11")
    with pytest.raises(RuntimeError, match="CUDA"):
        apriltag("tag36h11")


def test_shim_argument_errors_like_upstream():
    from aprilslam_b200.detector import apriltag
    with pytest.raises(RuntimeError, match="Unrecognized tag family"):
        apriltag("tag99h99")


def test_product_never_imports_the_oracle():
    """The product package must not reference oracle/ (a routed-through oracle voids parity claims)."""
    pkg = os.path.join(ROOT, "aprilslam_b200")
    bad = re.compile(r"(^\s*(import|from)\s+oracle\b)|(#include\s*[<\"][^>\"]*oracle)|(libapriltag_oracle)|(ao_detect)",
                     re.MULTILINE)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), (dirpath, f)
    code = ("import sys; sys.path.insert(0, %r); import aprilslam_b200, aprilslam_b200.detector, aprilslam_b200.shard, "
            "aprilslam_b200.apriltag; assert not any(m.startswith('oracle') for m in sys.modules), 'oracle imported'" % ROOT)
    subprocess.check_call([sys.executable, "-c", code])


def test_drop_in_module_name():
    """`from apriltag import apriltag` (tag_detector.py:11) resolves to the shim when the package dir is on sys.path."""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); from apriltag import apriltag, Detector; "
            "print(apriltag.__module__)" % (ROOT, os.path.join(ROOT, "aprilslam_b200")))
    out = subprocess.check_output([sys.executable, "-c", code], text=True)
    assert "aprilslam_b200.detector" in out


def test_shard_ranges_cover_the_batch():
    from aprilslam_b200.shard import shard_range
    for B in (0, 1, 7, 1024, 1031):
        for world in (1, 2, 3, 8):
            spans = [shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_pack_unpack_roundtrip():
    from aprilslam_b200._lib import DET_DTYPE
    from aprilslam_b200.shard import pack_lists, unpack_lists
    rng = np.random.default_rng(0)
    lists = []
    for n in (0, 3, 1, 0, 5):
        a = np.zeros(n, DET_DTYPE)
        a["id"] = rng.integers(0, 587, n)
        a["p"] = rng.normal(size=(n, 4, 2))
        lists.append(a)
    counts, flat = pack_lists(lists)
    back = unpack_lists(counts, flat)
    assert all(np.array_equal(x, y) for x, y in zip(lists, back))


_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from aprilslam_b200._lib import DET_DTYPE
from aprilslam_b200.shard import shard_range, gather_lists, env_rank
rank, local_rank, world = env_rank()
dist.init_process_group("gloo", rank=rank, world_size=world)
B = 11
lo, hi = shard_range(B, rank, world)
# a stand-in for Detector.detect_batch on this rank's shard: frame f holds f %% 4 detections with id = 100*f + k
lists = []
for f in range(lo, hi):
    a = np.zeros(f %% 4, DET_DTYPE); a["id"] = 100 * f + np.arange(f %% 4); a["c"][:, 0] = f
    lists.append(a)
out = gather_lists(lists, DET_DTYPE)
if rank == 0:
    assert len(out) == B, len(out)
    for f, a in enumerate(out):
        assert len(a) == f %% 4 and (a["id"] == 100 * f + np.arange(f %% 4)).all() and (a["c"][:, 0] == f).all()
    print("GATHER_OK")
else:
    assert out is None
dist.barrier(); dist.destroy_process_group()
"""


def test_world_size_2_gather_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER % {"root": ROOT})
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT="29611")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0]


def test_bench_reference_arm_prints_one_json_line():
    import json
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                                   "--warmup", "0", "--ref-frames", "4"], text=True, timeout=600)
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_result_buffers_are_recycled_only_when_unreferenced():
    """Detector._result_buffer hands a pooled [B, cap] array out again only when no FrameLists / view of it is alive."""
    from aprilslam_b200._lib import DET_DTYPE
    from aprilslam_b200.detector import Detector, FrameLists

    class PoolOnly(Detector):          # (no GPU here: only the host-side pool logic)
        def __init__(self):
            self._pool = {}

        def close(self):
            pass

    d = PoolOnly()
    a = d._result_buffer(4, 8, DET_DTYPE)
    ida = id(a)
    fl = FrameLists(a, np.array([1, 2, 0, 3], np.int32))
    del a
    b = d._result_buffer(4, 8, DET_DTYPE)
    assert id(b) != ida                                   # first buffer still referenced by `fl`
    view = fl[1]
    del fl
    c = d._result_buffer(4, 8, DET_DTYPE)
    assert id(c) != ida and len(view) == 2                # ... and by a per-frame view
    del view, b, c
    e = d._result_buffer(4, 8, DET_DTYPE)
    assert id(e) in {id(x) for x in d._pool[(4, 8, DET_DTYPE.str)]}   # recycled from the pool
    fl = FrameLists(e, np.array([1, 2, 0, 3], np.int32))
    assert len(fl) == 4 and [len(x) for x in fl] == [1, 2, 0, 3] and len(fl[-1]) == 3 and len(fl[1:3]) == 2
    with pytest.raises(IndexError):
        fl[4]


def test_bind_to_gpu_numa_degrades_without_nvml():
    """No GPU / NVML here: the helper reports 0 bound CPUs and leaves the process affinity untouched."""
    from aprilslam_b200.shard import bind_to_gpu_numa
    before = os.sched_getaffinity(0)
    n = bind_to_gpu_numa(0)
    assert isinstance(n, int) and n >= 0
    if n == 0:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)


def test_family_info_reports_the_truncated_41h12_codebook():
    """tagStandard41h12 (the reference's default tag_type) ships with ids 0..4 only; the C ABI and the Python layer say so."""
    from aprilslam_b200.detector import family_code_counts
    assert family_code_counts("tagStandard41h12") == (5, 2115)
    assert family_code_counts("tag36h11") == (587, 587)
    assert family_code_counts("tag25h9") == (35, 35) and family_code_counts("tag16h5") == (30, 30)
    assert family_code_counts("tag99h1") == (0, 0)


def test_result_pool_reference_count_is_calibrated():
    """Pooled result buffers are recycled when nothing but the pool refers to them; the idle reference count is measured on
    this interpreter (not a magic constant), and a live per-frame view keeps a buffer out of circulation."""
    import sys
    import numpy as np
    from aprilslam_b200 import detector
    assert detector._POOL_IDLE_REFS is not None
    pool = [np.empty((4, 4), np.uint8)]
    idle = sys.getrefcount(pool[0])          # (taken outside `assert`: pytest's rewriting would hold a reference of its own)
    view = pool[0][1, :2]
    busy = sys.getrefcount(pool[0])
    del view
    idle_again = sys.getrefcount(pool[0])
    assert idle == idle_again == detector._POOL_IDLE_REFS and busy > idle
