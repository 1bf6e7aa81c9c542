"""Synthetic frame source: a numpy restatement of the reference's OpenGL renderer geometry.

The reference renders its test scenes with fixed-function OpenGL
(/root/reference/src/simulation/renderer.py:91-96 projection, :188-195 view matrix,
:222-251 one textured quad per tag, :253-274 glReadPixels + flipud + RGB->BGR).  There is no
GL here, so the same geometry is evaluated per pixel:

  * pinhole fx = fy = 0.5*H/tan(fov_y/2), cx = W/2, cy = H/2
    (simulation_engine.py:124-126), pixel (c, r) sampled at (c+0.5, r+0.5);
  * OpenGL eye space (-Z forward, +Y up); model = T(pos) Rz(rot[2]) Ry(rot[1]) Rx(rot[0])
    (renderer.py:232-237); view = Rz(-roll) Rx(-pitch) Ry(-yaw) T(-cam) (:192-195);
  * the whole tag image (quiet zone included) is mapped on a quad of half size
    tag_size_outer/2, texture (0,0) at the bottom-left vertex, PNG upright (:164, :243-249);
  * GL_LINEAR magnification / minification without mipmaps, GL_REPEAT wrap (:172-173);
  * depth test on, aliased polygon edges, background glClearColor(0.5, 0, 0.5) -> gray 53.

A tag texture is "virtual": a cell grid (total_width x total_width bits) rendered at
``ppc`` texels per cell, so that no image atlas is needed (the CUDA renderer in
csrc/render.cuh evaluates the same function).

This module is the frame generator for tests and benchmarks; it is not part of the detector.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .families_data import FAMILIES

BACKGROUND_GRAY = 53  # BGR (128, 0, 128) through cv2.cvtColor(BGR2GRAY), SURVEY.md section 8c
DEFAULT_PPC = 40      # texels per tag cell (reference PNGs: 354 px / 9 cells = 39.3)


# --------------------------------------------------------------------------------------
# tag cell grids
# --------------------------------------------------------------------------------------
def tag_cells(family: str, tag_id: int) -> np.ndarray:
    """total_width x total_width array of {0,1} (1 = white), row 0 = top of the upright tag."""
    f = FAMILIES[family]
    wb, tw, nbits = f["width_at_border"], f["total_width"], f["nbits"]
    off = (tw - wb) // 2
    code = f["codes"][tag_id]
    rev = f["reversed_border"]
    g = np.zeros((tw, tw), np.uint8)
    for y in range(-off, wb + off):
        for x in range(-off, wb + off):
            ring = min(x, y, wb - 1 - x, wb - 1 - y)  # 0 = border ring, <0 outside, >0 inside
            if ring == 0:
                v = 1 if rev else 0
            elif ring == -1:
                v = 0 if rev else 1
            else:
                v = 0
            g[y + off, x + off] = v
    for i in range(nbits):
        bit = (code >> (nbits - 1 - i)) & 1
        g[f["bit_y"][i] + off, f["bit_x"][i] + off] = bit
    return g


def tag_image(family: str, tag_id: int, ppc: int = DEFAULT_PPC) -> np.ndarray:
    """Nearest-neighbour up-scaled gray image of the tag (0 / 255)."""
    g = tag_cells(family, tag_id)
    return np.kron(g, np.ones((ppc, ppc), np.uint8)) * 255


# --------------------------------------------------------------------------------------
# geometry helpers (OpenGL conventions, degrees)
# --------------------------------------------------------------------------------------
def _rx(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], float)


def _ry(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], float)


def _rz(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], float)


def model_matrix(position, rotation_deg) -> np.ndarray:
    """T(pos) Rz(rot[2]) Ry(rot[1]) Rx(rot[0]) -- renderer.py:232-237."""
    r = np.radians(rotation_deg)
    M = np.eye(4)
    M[:3, :3] = _rz(r[2]) @ _ry(r[1]) @ _rx(r[0])
    M[:3, 3] = position
    return M


def view_matrix(cam_pos, cam_rot_deg=(0.0, 0.0, 0.0)) -> np.ndarray:
    """Rz(-roll) Rx(-pitch) Ry(-yaw) T(-cam), camera_rotation = [pitch, yaw, roll] -- renderer.py:188-195."""
    p, y, r = np.radians(cam_rot_deg)
    V = np.eye(4)
    V[:3, :3] = _rz(-r) @ _rx(-p) @ _ry(-y)
    T = np.eye(4)
    T[:3, 3] = -np.asarray(cam_pos, float)
    return V @ T


def intrinsics(width: int, height: int, fov_y_deg: float) -> np.ndarray:
    """simulation_engine.py:124-132."""
    f = 0.5 * height / math.tan(0.5 * math.radians(fov_y_deg))
    return np.array([[f, 0, 0.5 * width], [0, f, 0.5 * height], [0, 0, 1.0]])


FLIP = np.diag([1.0, -1.0, -1.0])  # GL eye axes -> OpenCV camera axes (ground_truth.py:70-88)


# --------------------------------------------------------------------------------------
# scene description
# --------------------------------------------------------------------------------------
@dataclass
class SceneTag:
    family: str
    tag_id: int
    eye_from_tag: np.ndarray  # 4x4, GL eye space <- tag plane coords (u right, v up, w out)
    half_outer: float         # half size of the textured quad (tag_size_outer / 2)
    half_inner: float         # half size of the border square (tag_size_inner / 2)
    ppc: int = DEFAULT_PPC
    texture: Optional[np.ndarray] = None   # gray uint8 image (row 0 = top) instead of the virtual cell texture,
                                           # e.g. the reference's own assets/tags/tagN.png (renderer.py:150-176)


@dataclass
class Scene:
    width: int
    height: int
    K: np.ndarray
    tags: List[SceneTag] = field(default_factory=list)
    background: int = BACKGROUND_GRAY


def tag_records(scene: Scene) -> np.ndarray:
    """Per tag: 9 (pixel -> (s*q, t*q, q) texture homography) + 3 (1/depth plane) doubles.

    (s, t) in [0,1]^2 over the textured quad with t = 0 at the bottom edge; the same record
    layout feeds the CUDA renderer.
    """
    fx, fy, cx, cy = scene.K[0, 0], scene.K[1, 1], scene.K[0, 2], scene.K[1, 2]
    out = np.zeros((len(scene.tags), 12))
    for i, t in enumerate(scene.tags):
        M = t.eye_from_tag
        h = t.half_outer
        # plane point P(u,v) = M[:, 0] u + M[:, 1] v + M[:, 3]; (s,t) = (u + h, v + h) / 2h
        A = np.stack([M[:3, 0] * 2 * h, M[:3, 1] * 2 * h, M[:3, 3] - h * (M[:3, 0] + M[:3, 1])], axis=1)
        # image homogeneous coords: x = cx + fx X/(-Z), y = cy - fy Y/(-Z)
        P = np.array([[fx, 0, -cx], [0, -fy, -cy], [0, 0, -1.0]])
        G = P @ A                      # (s, t, 1) -> (x*w, y*w, w), w = -Z = depth
        Gi = np.linalg.inv(G)
        out[i, :9] = Gi.reshape(-1)
        # for a pixel (x, y): Gi @ (x, y, 1) = (s, t, 1) / w  => q = 1/depth
        out[i, 9:12] = Gi[2]
    return out


def gt_corners(scene: Scene, tag: SceneTag) -> np.ndarray:
    """Image positions of the border-square corners in the detector's lb, rb, rt, lt order."""
    K = scene.K
    pts = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], float) * tag.half_inner
    out = np.zeros((4, 2))
    for i, (u, v) in enumerate(pts):
        p = tag.eye_from_tag @ np.array([u, v, 0, 1.0])
        out[i, 0] = K[0, 2] + K[0, 0] * p[0] / (-p[2])
        out[i, 1] = K[1, 2] - K[1, 1] * p[1] / (-p[2])
    return out


def gt_pose(tag: SceneTag) -> Tuple[np.ndarray, np.ndarray]:
    """Camera<-tag (R, t) in OpenCV axes: diag(1,-1,-1) applied to the GL eye transform."""
    return FLIP @ tag.eye_from_tag[:3, :3], FLIP @ tag.eye_from_tag[:3, 3]


# --------------------------------------------------------------------------------------
# rasteriser
# --------------------------------------------------------------------------------------
def _sample_virtual(cells: np.ndarray, ppc: int, s: np.ndarray, t: np.ndarray) -> np.ndarray:
    """GL_LINEAR + GL_REPEAT lookup in the (total_width*ppc)^2 virtual texture; t=0 is the bottom."""
    tw = cells.shape[0]
    n = tw * ppc
    xt = s * n - 0.5
    yt = (1.0 - t) * n - 0.5
    x0 = np.floor(xt)
    y0 = np.floor(yt)
    ax = xt - x0
    ay = yt - y0
    x0 = x0.astype(np.int64)
    y0 = y0.astype(np.int64)
    x0m, x1m = np.mod(x0, n) // ppc, np.mod(x0 + 1, n) // ppc
    y0m, y1m = np.mod(y0, n) // ppc, np.mod(y0 + 1, n) // ppc
    c = cells.astype(np.float64) * 255.0
    v = ((1 - ax) * (1 - ay) * c[y0m, x0m] + ax * (1 - ay) * c[y0m, x1m]
         + (1 - ax) * ay * c[y1m, x0m] + ax * ay * c[y1m, x1m])
    return v


def _sample_image(tex: np.ndarray, s: np.ndarray, t: np.ndarray) -> np.ndarray:
    """GL_LINEAR + GL_REPEAT lookup in an image texture (texel centres at +0.5; t = 0 is the bottom row, the way
    renderer.py:164 uploads the PNG bottom-up)."""
    h, w = tex.shape
    xt = s * w - 0.5
    yt = (1.0 - t) * h - 0.5
    x0 = np.floor(xt)
    y0 = np.floor(yt)
    ax = xt - x0
    ay = yt - y0
    x0 = x0.astype(np.int64)
    y0 = y0.astype(np.int64)
    x0m, x1m = np.mod(x0, w), np.mod(x0 + 1, w)
    y0m, y1m = np.mod(y0, h), np.mod(y0 + 1, h)
    c = tex.astype(np.float64)
    return ((1 - ax) * (1 - ay) * c[y0m, x0m] + ax * (1 - ay) * c[y0m, x1m]
            + (1 - ax) * ay * c[y1m, x0m] + ax * ay * c[y1m, x1m])


def render(scene: Scene) -> np.ndarray:
    """Gray uint8 [H, W] frame."""
    W, H = scene.width, scene.height
    img = np.full((H, W), scene.background, np.uint8)
    zbuf = np.zeros((H, W))  # stores 1/depth; larger = nearer
    recs = tag_records(scene)
    for tag, rec in zip(scene.tags, recs):
        Gi = rec[:9].reshape(3, 3)
        G = np.linalg.inv(Gi)
        cs = np.array([[0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1.0]]) @ G.T
        if np.any(cs[:, 2] <= 1e-9):
            x0, x1, y0, y1 = 0, W, 0, H
        else:
            px, py = cs[:, 0] / cs[:, 2], cs[:, 1] / cs[:, 2]
            x0, x1 = max(0, int(math.floor(px.min())) - 1), min(W, int(math.ceil(px.max())) + 2)
            y0, y1 = max(0, int(math.floor(py.min())) - 1), min(H, int(math.ceil(py.max())) + 2)
        if x0 >= x1 or y0 >= y1:
            continue
        xs = np.arange(x0, x1) + 0.5
        ys = np.arange(y0, y1) + 0.5
        X, Y = np.meshgrid(xs, ys)
        a = Gi[0, 0] * X + Gi[0, 1] * Y + Gi[0, 2]
        b = Gi[1, 0] * X + Gi[1, 1] * Y + Gi[1, 2]
        q = Gi[2, 0] * X + Gi[2, 1] * Y + Gi[2, 2]
        ok = q > 0
        qq = np.where(ok, q, 1.0)
        s = a / qq
        t = b / qq
        ok &= (s >= 0) & (s <= 1) & (t >= 0) & (t <= 1)
        zb = zbuf[y0:y1, x0:x1]
        ok &= q > zb
        if not ok.any():
            continue
        if tag.texture is not None:
            v = _sample_image(tag.texture, np.where(ok, s, 0.5), np.where(ok, t, 0.5))
        else:
            cells = tag_cells(tag.family, tag.tag_id)
            v = _sample_virtual(cells, tag.ppc, np.where(ok, s, 0.5), np.where(ok, t, 0.5))
        sub = img[y0:y1, x0:x1]
        sub[ok] = np.floor(v[ok] + 0.5).astype(np.uint8)
        zb[ok] = q[ok]
    return img


def gray_to_bgr(gray: np.ndarray, background: int = BACKGROUND_GRAY) -> np.ndarray:
    """BGR frame whose cv2.cvtColor(BGR2GRAY) is `gray`; pure-background pixels are GL purple."""
    bgr = np.repeat(gray[..., None], 3, axis=2)
    return bgr


# --------------------------------------------------------------------------------------
# the reference's own scene (config/sim_settings.json)
# --------------------------------------------------------------------------------------
SIM_SETTINGS_TAGS = [  # /root/reference/config/sim_settings.json:11-42
    (0, (0, 0, -50), (0, 0, 0)),
    (1, (-30, 0, -120), (0, 0, 0)),
    (2, (25, 15, -85), (0, 0, 0)),
    (3, (55, -10, -75), (0, 20, 10)),
    (4, (80, 5, -65), (0, 20, 0)),
]


def sim_settings_scene(width=1000, height=1000, cam_pos=(0.0, 0.0, 0.0), cam_rot=(0.0, 0.0, 0.0),
                       fov_y=45.0, family="tagStandard41h12", size_scale=2.0,
                       tag_size_inner=5.0, tag_size_outer=9.0, textures=None) -> Scene:
    """The shipped 5-tag scene (sizes scaled by size_scale as config_manager.py:147 does).  textures: optional
    gray images indexed by tag id (the reference's assets/tags/tag{0..4}.png) instead of code-book cell textures."""
    sc = Scene(width, height, intrinsics(width, height, fov_y))
    V = view_matrix(cam_pos, cam_rot)
    wb = FAMILIES[family]["width_at_border"]
    tw = FAMILIES[family]["total_width"]
    inner = tag_size_inner * size_scale
    outer = tag_size_outer * size_scale if family == "tagStandard41h12" else inner * tw / wb
    for tid, pos, rot in SIM_SETTINGS_TAGS:
        sc.tags.append(SceneTag(family, tid, V @ model_matrix(pos, rot), outer / 2, inner / 2,
                                texture=None if textures is None else textures[tid]))
    return sc


# --------------------------------------------------------------------------------------
# benchmark scenes (BASELINE.json configs 2-5): jittered grid of tags
# --------------------------------------------------------------------------------------
def _axis_angle(axis, ang):
    axis = np.asarray(axis, float)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)


def grid_scene(width: int, height: int, frame_index: int, grid: Tuple[int, int],
               families: Sequence[Tuple[str, Sequence[int]]] = (("tag36h11", range(587)),),
               px_range: Tuple[float, float] = (60.0, 110.0), max_tilt_deg: float = 50.0,
               fov_y: float = 45.0, seed_base: int = 1000) -> Scene:
    """One frame of the benchmark workload: grid[0] x grid[1] tags, jittered, rotated and tilted.

    Seeded by ``seed_base + frame_index``.  The border width in pixels is drawn
    from px_range and clipped so that a tag stays inside ~1.05 grid cells.
    """
    rng = np.random.default_rng(seed_base + frame_index)
    K = intrinsics(width, height, fov_y)
    sc = Scene(width, height, K, background=int(rng.integers(30, 201)))
    gx, gy = grid
    cw, ch = width / gx, height / gy
    pools = {}
    for fam, ids in families:
        ids = np.array(list(ids))
        rng.shuffle(ids)
        pools[fam] = list(ids)
    fams = [f for f, _ in families]
    for j in range(gy):
        for i in range(gx):
            fam = fams[int(rng.integers(len(fams)))]
            if not pools[fam]:
                continue
            tid = int(pools[fam].pop())
            f = FAMILIES[fam]
            wb, tw = f["width_at_border"], f["total_width"]
            # data outside the border (Standard families) needs no extra quiet zone here
            lim = 0.72 * min(cw, ch) * wb / tw
            s_px = min(float(rng.uniform(*px_range)), lim)
            cxp = (i + 0.5) * cw + float(rng.uniform(-0.08, 0.08)) * cw
            cyp = (j + 0.5) * ch + float(rng.uniform(-0.08, 0.08)) * ch
            inner = 1.0
            Z = K[0, 0] * inner / s_px
            X = (cxp - K[0, 2]) / K[0, 0] * Z
            Y = -(cyp - K[1, 2]) / K[1, 1] * Z
            inplane = float(rng.uniform(-math.pi, math.pi))
            phi = float(rng.uniform(0, 2 * math.pi))
            tilt = math.radians(float(rng.uniform(-max_tilt_deg, max_tilt_deg)))
            R = _axis_angle((math.cos(phi), math.sin(phi), 0.0), tilt) @ _rz(inplane)
            M = np.eye(4)
            M[:3, :3] = R
            M[:3, 3] = (X, Y, -Z)
            sc.tags.append(SceneTag(fam, tid, M, 0.5 * inner * tw / wb, 0.5 * inner))
    return sc


# --------------------------------------------------------------------------------------
# C5 augmentations (builder-defined, seeded): the reference's randomised
# config only perturbs tag poses, so blur / noise / lighting are defined here
# --------------------------------------------------------------------------------------
def augment(img: np.ndarray, seed: int) -> np.ndarray:
    """Gaussian blur sigma~U(0,1.5) px, gain U(0.6,1.2), linear illumination ramp +-30, additive N(0, U(0,8))."""
    import cv2
    rng = np.random.default_rng(seed)
    sigma = float(rng.uniform(0.0, 1.5))
    out = img.astype(np.float32)
    if sigma > 0.05:
        out = cv2.GaussianBlur(out, (0, 0), sigma)
    gain = float(rng.uniform(0.6, 1.2))
    H, W = img.shape
    ang = float(rng.uniform(0, 2 * math.pi))
    amp = float(rng.uniform(0, 30))
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    ramp = amp * ((xx / W - 0.5) * math.cos(ang) + (yy / H - 0.5) * math.sin(ang)) * 2
    out = out * gain + ramp
    sn = float(rng.uniform(0, 8))
    out = out + rng.normal(0, sn, img.shape).astype(np.float32)
    return np.clip(np.floor(out + 0.5), 0, 255).astype(np.uint8)
