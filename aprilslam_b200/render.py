"""GPU synthetic frame source: Scene objects (synth.py) -> uint8 [B,H,W] frames in HBM, bit-identical to
synth.render (the numpy restatement of the reference's OpenGL renderer, renderer.py:180-274).  Frame source for
benchmarks and tests; not part of the detector."""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np

from . import synth

TAG_DTYPE = np.dtype([("Gi", "<f8", (9,)), ("cells", "<u8", (2,)), ("total_width", "<i4"), ("ppc", "<i4"),
                      ("x0", "<i4"), ("x1", "<i4"), ("y0", "<i4"), ("y1", "<i4")])
assert TAG_DTYPE.itemsize == 112


def scene_records(scene: synth.Scene) -> np.ndarray:
    """Per-tag render records with the same bounding boxes synth.render uses."""
    W, H = scene.width, scene.height
    recs = synth.tag_records(scene)
    out = np.zeros(len(scene.tags), TAG_DTYPE)
    for i, (tag, rec) in enumerate(zip(scene.tags, recs)):
        Gi = rec[:9].reshape(3, 3)
        G = np.linalg.inv(Gi)
        cs = np.array([[0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1.0]]) @ G.T
        if np.any(cs[:, 2] <= 1e-9):
            x0, x1, y0, y1 = 0, W, 0, H
        else:
            px, py = cs[:, 0] / cs[:, 2], cs[:, 1] / cs[:, 2]
            x0, x1 = max(0, int(math.floor(px.min())) - 1), min(W, int(math.ceil(px.max())) + 2)
            y0, y1 = max(0, int(math.floor(py.min())) - 1), min(H, int(math.ceil(py.max())) + 2)
        cells = synth.tag_cells(tag.family, tag.tag_id)
        bits = [0, 0]
        for k, v in enumerate(cells.ravel()):
            if v:
                bits[k >> 6] |= 1 << (k & 63)
        out[i]["Gi"] = Gi.reshape(-1)
        out[i]["cells"] = bits
        out[i]["total_width"] = cells.shape[0]
        out[i]["ppc"] = tag.ppc
        out[i]["x0"], out[i]["x1"], out[i]["y0"], out[i]["y1"] = x0, max(x0, x1), y0, max(y0, y1)
    return out


def render_batch(detector, scenes: Sequence[synth.Scene]):
    """-> torch.uint8 [B,H,W] CUDA tensor on the detector's device."""
    import torch
    B = len(scenes)
    W, H = scenes[0].width, scenes[0].height
    assert all(s.width == W and s.height == H for s in scenes)
    recs = [scene_records(s) for s in scenes]
    offsets = np.zeros(B + 1, np.int32)
    offsets[1:] = np.cumsum([len(r) for r in recs])
    tags = np.concatenate(recs) if offsets[-1] else np.zeros(1, TAG_DTYPE)
    bg = np.array([s.background for s in scenes], np.uint8)
    dev = torch.device("cuda", detector.device)
    frames = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = detector._L.agpu_render(detector._h, tags.ctypes.data, offsets.ctypes.data, bg.ctypes.data, B, W, H,
                                 frames.data_ptr(), stream)
    detector._check(rc)
    return frames
