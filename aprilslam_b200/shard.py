"""Frame sharding across the GPUs of one box: frames are independent units, so the
batch index range is split across ranks, every rank runs its own Detector (one C handle, one stream
set) on its own GPU, and only the compact detection lists travel.  There is no data-path collective;
`gather_lists` (rank 0 collects the per-frame lists in frame order) is host-side plumbing over
torch.distributed (nccl or gloo) and is optional.
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import numpy as np


def shard_range(num_frames: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [start, stop) frame range of `rank`; ranges differ by at most one frame."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, rem = divmod(num_frames, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def env_rank() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1-process default)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def bind_to_gpu_numa(device_index: int) -> int:
    """Pin this process to the CPU cores next to GPU `device_index` (NVML's ideal CPU affinity), so that the pinned
    frame buffers it allocates afterwards live on the NUMA node the GPU's PCIe link hangs off -- with one process per
    GPU every rank then streams its frames over its own socket's memory controllers.  Returns the number of CPUs
    bound to (0: NVML unavailable or nothing to do; the process is left as it was)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * wi + b for wi, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def pack_lists(lists: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    """Per-frame record arrays -> (counts int32[B], concatenated records)."""
    counts = np.array([len(x) for x in lists], np.int32)
    if len(lists) and sum(counts):
        flat = np.concatenate([np.asarray(x) for x in lists])
    else:
        flat = np.zeros(0, lists[0].dtype if len(lists) else np.uint8)
    return counts, flat


def unpack_lists(counts: np.ndarray, flat: np.ndarray) -> List[np.ndarray]:
    out, o = [], 0
    for c in counts:
        out.append(flat[o:o + int(c)])
        o += int(c)
    return out


def gather_lists(lists: Sequence[np.ndarray], dtype: np.dtype, group=None) -> List[np.ndarray] | None:
    """Collect every rank's per-frame lists on rank 0, concatenated in rank (= frame) order.

    Works on any torch.distributed backend; payloads are raw bytes, sizes are exchanged first
    (ragged: ranks may hold different numbers of frames and detections).  Returns None on ranks != 0.
    """
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(lists)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    counts, flat = pack_lists(lists)
    payload = np.concatenate([counts.view(np.uint8), np.ascontiguousarray(flat).view(np.uint8).ravel()])
    meta = torch.tensor([len(counts), payload.size], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    maxlen = int(max(int(m[1]) for m in metas))
    buf = torch.zeros(max(maxlen, 1), dtype=torch.uint8, device=dev)
    buf[:payload.size] = torch.from_numpy(payload.copy()).to(dev)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)
    if rank != 0:
        return None
    out: List[np.ndarray] = []
    for m, b in zip(metas, bufs):
        nframes, nbytes = int(m[0]), int(m[1])
        raw = b[:nbytes].cpu().numpy()
        cnt = raw[:4 * nframes].view(np.int32)
        recs = raw[4 * nframes:].view(dtype)
        out.extend(unpack_lists(cnt, recs))
    return out
