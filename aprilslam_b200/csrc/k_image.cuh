// k_image.cuh -- image stages: (BGR->gray) + quad_decimate + Gaussian blur + adaptive threshold.
//
// Restates upstream stages U1-U3, i.e. the first part of the
// native call at /root/reference/src/detection/tag_detector.py:26, plus cv2.cvtColor(BGR2GRAY)
// of tag_detector.py:25.
//
// Fast path (k_decimate_threshold): HBM-bound streaming kernel with zero shared memory.  One warp
// walks DOWN a vertical strip of the image.  Every lane owns one 128-bit column chunk (16 source
// bytes per row = 4 / 2 / 1 threshold tiles for decimate 1 / 2 / 4), so a warp reads 512
// contiguous bytes per source row (fully coalesced LDG.128).  Tile min/max come from packed-byte
// SIMD (__vminu4/__vmaxu4), the 3x3 tile dilation takes its horizontal neighbours with warp
// shuffles and its vertical neighbours from a sliding register window (previous / current / next
// tile row), so each pixel is read from HBM once and written once.  Lanes 0 and 31 are halo lanes
// (their tiles are needed by the dilation of lanes 1 and 30 but are written by the neighbouring
// strip).
#pragma once
#include "common.cuh"

template <int F>
__device__ __forceinline__ void unpack16(const uint4 v, uint32_t (&w)[4 / F]) {
    if (F == 1) {
        w[0] = v.x; w[1 % (4 / F)] = v.y; w[2 % (4 / F)] = v.z; w[3 % (4 / F)] = v.w;
    } else if (F == 2) {
        w[0] = __byte_perm(v.x, v.y, 0x6420);
        w[1 % (4 / F)] = __byte_perm(v.z, v.w, 0x6420);
    } else {
        uint32_t a = __byte_perm(v.x, v.y, 0x0040), b = __byte_perm(v.z, v.w, 0x0040);
        w[0] = __byte_perm(a, b, 0x5410);
    }
}

__device__ __forceinline__ uint32_t bytes_min(uint32_t w) {
    uint32_t a = __vminu4(w, w >> 16);
    a = __vminu4(a, a >> 8);
    return a & 0xffu;
}
__device__ __forceinline__ uint32_t bytes_max(uint32_t w) {
    uint32_t a = __vmaxu4(w, w >> 16);
    a = __vmaxu4(a, a >> 8);
    return a & 0xffu;
}

template <int TPL>
__device__ __forceinline__ uint32_t nb_left(uint32_t own, uint32_t from_up) {
    // byte j <- tile j-1 (byte 0 from the lane on the left)
    uint32_t r = (own << 8) | ((from_up >> (8 * (TPL - 1))) & 0xffu);
    if (TPL < 4) r &= (1u << (8 * (TPL & 3))) - 1u;
    return r;
}
template <int TPL>
__device__ __forceinline__ uint32_t nb_right(uint32_t own, uint32_t from_down) {
    // byte j <- tile j+1 (last byte from the lane on the right)
    uint32_t lowm = (TPL == 1) ? 0u : ((1u << (8 * ((TPL - 1) & 3))) - 1u);
    return ((own >> 8) & lowm) | ((from_down & 0xffu) << (8 * (TPL - 1)));
}

// threshold of one word (4 pixels) against tile extrema (mn, mx)
__device__ __forceinline__ uint32_t thresh_word(uint32_t px, uint32_t mn, uint32_t mx, int min_diff) {
    int diff = (int)mx - (int)mn;
    if (diff < min_diff) return 0x7f7f7f7fu;
    uint32_t thr = mn + (uint32_t)(diff >> 1);
    return __vcmpgtu4(px, thr * 0x01010101u);  // 0xff where v > thr
}

template <int F>
__global__ void __launch_bounds__(256)
k_decimate_threshold(const uint8_t* __restrict__ src, int W, int H, size_t src_stride, size_t src_frame_stride,
                     uint8_t* __restrict__ quad_im, uint8_t* __restrict__ thresh, Geom g, int nstrips, int nsegs,
                     int seg_tiles, int nframes, int min_diff, int vec_ok) {
    constexpr int TPL = 4 / F;  // tiles (4-pixel words) per lane per row
    const int warp = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    const int per_frame = nstrips * nsegs;
    if (warp >= nframes * per_frame) return;
    const int frame = warp / per_frame;
    const int rem = warp - frame * per_frame;
    const int seg = rem / nstrips, strip = rem - seg * nstrips;
    const uint8_t* fsrc = src + (size_t)frame * src_frame_stride;
    uint8_t* fth = thresh + (size_t)frame * g.plane;
    uint8_t* fq = quad_im ? quad_im + (size_t)frame * g.plane : nullptr;
    const int tw = g.wd >> 2, th = g.hd >> 2;
    const int tile0 = (strip * 30 + lane - 1) * TPL;
    const int px0 = tile0 * 4;
    const long sc0 = (long)px0 * F;
    const bool is_out = lane >= 1 && lane <= 30 && px0 < g.wd;
    const bool vec = vec_ok && sc0 >= 0 && sc0 + 16 <= W;

    auto load_row = [&](int gy, uint32_t(&w)[TPL]) {
#pragma unroll
        for (int j = 0; j < TPL; j++) w[j] = 0;
        if (gy >= g.hd) return;
        const uint8_t* row = fsrc + (size_t)gy * F * src_stride;
        if (vec) {
            uint4 v = __ldg(reinterpret_cast<const uint4*>(row + sc0));
            unpack16<F>(v, w);
        } else {
#pragma unroll
            for (int k = 0; k < TPL * 4; k++) {
                long sx = sc0 + (long)k * F;
                if (sx >= 0 && sx < W) w[k >> 2] |= (uint32_t)row[sx] << (8 * (k & 3));
            }
        }
    };
    // packed per-tile extrema of tile row T (neutral for tiles outside the tile grid), horizontally dilated
    auto tile_extrema = [&](int T, const uint32_t(&px)[4][TPL], uint32_t& hmn, uint32_t& hmx) {
        uint32_t mn = 0, mx = 0;
#pragma unroll
        for (int j = 0; j < TPL; j++) {
            uint32_t a = __vminu4(__vminu4(px[0][j], px[1][j]), __vminu4(px[2][j], px[3][j]));
            uint32_t b = __vmaxu4(__vmaxu4(px[0][j], px[1][j]), __vmaxu4(px[2][j], px[3][j]));
            uint32_t tmn = bytes_min(a), tmx = bytes_max(b);
            int tx = tile0 + j;
            if (tx < 0 || tx >= tw || T < 0 || T >= th) { tmn = 255u; tmx = 0u; }
            mn |= tmn << (8 * j);
            mx |= tmx << (8 * j);
        }
        uint32_t mn_up = __shfl_up_sync(FULL_MASK, mn, 1), mn_dn = __shfl_down_sync(FULL_MASK, mn, 1);
        uint32_t mx_up = __shfl_up_sync(FULL_MASK, mx, 1), mx_dn = __shfl_down_sync(FULL_MASK, mx, 1);
        if (lane == 0) { mn_up = 0xffffffffu; mx_up = 0u; }
        if (lane == 31) { mn_dn = 0xffffffffu; mx_dn = 0u; }
        hmn = __vminu4(__vminu4(nb_left<TPL>(mn, mn_up), mn), nb_right<TPL>(mn, mn_dn));
        hmx = __vmaxu4(__vmaxu4(nb_left<TPL>(mx, mx_up), mx), nb_right<TPL>(mx, mx_dn));
        if (TPL < 4) {  // keep unused bytes neutral
            hmn |= ~((1u << (8 * (TPL & 3))) - 1u);
            hmx &= (1u << (8 * (TPL & 3))) - 1u;
        }
    };
    auto store_row = [&](int gy, const uint32_t(&px)[TPL], uint32_t dmn, uint32_t dmx, uint32_t lmn, uint32_t lmx) {
        uint32_t o[TPL];
#pragma unroll
        for (int j = 0; j < TPL; j++) {
            int tx = tile0 + j;
            uint32_t mn = (dmn >> (8 * j)) & 0xffu, mx = (dmx >> (8 * j)) & 0xffu;
            if (tx >= tw) {  // right leftover columns use the last tile column
                mn = (lmn >> (8 * j)) & 0xffu;
                mx = (lmx >> (8 * j)) & 0xffu;
            }
            o[j] = thresh_word(px[j], mn, mx, min_diff);
        }
        if (!is_out || gy >= g.hd) return;
        size_t off = (size_t)gy * g.wp + px0;
        if (TPL == 4) {
            *reinterpret_cast<uint4*>(fth + off) = make_uint4(o[0], o[1 % TPL], o[2 % TPL], o[3 % TPL]);
            if (fq) *reinterpret_cast<uint4*>(fq + off) = make_uint4(px[0], px[1 % TPL], px[2 % TPL], px[3 % TPL]);
        } else if (TPL == 2) {
            *reinterpret_cast<uint2*>(fth + off) = make_uint2(o[0], o[1 % TPL]);
            if (fq) *reinterpret_cast<uint2*>(fq + off) = make_uint2(px[0], px[1 % TPL]);
        } else {
            *reinterpret_cast<uint32_t*>(fth + off) = o[0];
            if (fq) *reinterpret_cast<uint32_t*>(fq + off) = px[0];
        }
    };

    const int T0 = seg * seg_tiles;
    const int T1 = min(T0 + seg_tiles, th);
    if (T0 >= th) return;

    uint32_t cur[4][TPL], nxt[4][TPL];
    uint32_t prevMn = 0xffffffffu, prevMx = 0u, curMn, curMx;
    if (T0 > 0) {
#pragma unroll
        for (int r = 0; r < 4; r++) load_row((T0 - 1) * 4 + r, cur[r]);
        tile_extrema(T0 - 1, cur, prevMn, prevMx);
    }
#pragma unroll
    for (int r = 0; r < 4; r++) load_row(T0 * 4 + r, cur[r]);
    tile_extrema(T0, cur, curMn, curMx);

    for (int T = T0; T < T1; T++) {
        uint32_t nextMn = 0xffffffffu, nextMx = 0u;
        if (T + 1 < th) {
#pragma unroll
            for (int r = 0; r < 4; r++) load_row((T + 1) * 4 + r, nxt[r]);
        } else {
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int j = 0; j < TPL; j++) nxt[r][j] = 0;
        }
        tile_extrema(T + 1, nxt, nextMn, nextMx);  // neutral when T+1 == th
        uint32_t dmn = __vminu4(__vminu4(prevMn, curMn), nextMn);
        uint32_t dmx = __vmaxu4(__vmaxu4(prevMx, curMx), nextMx);
        uint32_t lmn = nb_left<TPL>(dmn, __shfl_up_sync(FULL_MASK, dmn, 1));
        uint32_t lmx = nb_left<TPL>(dmx, __shfl_up_sync(FULL_MASK, dmx, 1));
#pragma unroll
        for (int r = 0; r < 4; r++) store_row(T * 4 + r, cur[r], dmn, dmx, lmn, lmx);
        if (T == th - 1 && (g.hd & 3)) {  // bottom leftover rows use the last tile row
            for (int gy = th * 4; gy < g.hd; gy++) {
                uint32_t w[TPL];
                load_row(gy, w);
                store_row(gy, w, dmn, dmx, lmn, lmx);
            }
        }
        prevMn = curMn; prevMx = curMx;
        curMn = nextMn; curMx = nextMx;
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int j = 0; j < TPL; j++) cur[r][j] = nxt[r][j];
    }
}

// cv2.cvtColor(BGR2GRAY) of this OpenCV build: (B*3735 + G*19235 + R*9798 + 16384) >> 15
__device__ __forceinline__ uint32_t bgr2gray(uint32_t b, uint32_t g, uint32_t r) {
    return (b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15;
}

// Generic front end: any integer decimation factor, 1 or 3 channels -> pitched quad_im.
// 4 output pixels per thread (one 32-bit store).
__global__ void __launch_bounds__(256)
k_pack(const uint8_t* __restrict__ src, int W, int H, size_t src_stride, size_t src_frame_stride, int channels,
       int F, uint8_t* __restrict__ quad_im, Geom g, int nframes) {
    const int words_per_row = g.wp >> 2;
    size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t total = (size_t)nframes * g.hd * words_per_row;
    if (idx >= total) return;
    int wx = (int)(idx % words_per_row);
    size_t t = idx / words_per_row;
    int gy = (int)(t % g.hd);
    int frame = (int)(t / g.hd);
    const uint8_t* row = src + (size_t)frame * src_frame_stride + (size_t)gy * F * src_stride;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int gx = wx * 4 + k;
        if (gx < g.wd) {
            size_t sx = (size_t)gx * F * channels;
            uint32_t v = channels == 1 ? row[sx] : bgr2gray(row[sx], row[sx + 1], row[sx + 2]);
            out |= v << (8 * k);
        }
    }
    *reinterpret_cast<uint32_t*>(quad_im + (size_t)frame * g.plane + (size_t)gy * g.wp + wx * 4) = out;
}

// Separable integer Gaussian (upstream image_u8_gaussian_blur / convolve).
// dir 0: along rows, dir 1: along columns.  Positions outside [ksz/2, sz-ksz+ksz/2) are copied.
struct BlurKernel {
    int ksz;
    uint8_t k[64];
};
__global__ void __launch_bounds__(256)
k_blur_pass(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, Geom g, int nframes, BlurKernel bk, int dir) {
    size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t total = (size_t)nframes * g.hd * g.wd;
    if (idx >= total) return;
    int x = (int)(idx % g.wd);
    size_t t = idx / g.wd;
    int y = (int)(t % g.hd);
    int frame = (int)(t / g.hd);
    const uint8_t* fi = in + (size_t)frame * g.plane;
    int sz = dir == 0 ? g.wd : g.hd;
    int p = dir == 0 ? x : y;
    int half = bk.ksz / 2;
    uint32_t v;
    if (p >= half && p < sz - bk.ksz + half) {
        uint32_t acc = 0;
        for (int j = 0; j < bk.ksz; j++) {
            int q = p - half + j;
            acc += (uint32_t)bk.k[j] * (dir == 0 ? fi[(size_t)y * g.wp + q] : fi[(size_t)q * g.wp + x]);
        }
        v = acc >> 8;
    } else {
        v = fi[(size_t)y * g.wp + x];
    }
    out[(size_t)frame * g.plane + (size_t)y * g.wp + x] = (uint8_t)v;
}

// quad_sigma < 0: v = clamp(2*orig - blurred)
__global__ void __launch_bounds__(256)
k_unsharp(const uint8_t* __restrict__ orig, uint8_t* __restrict__ blurred_inout, Geom g, int nframes) {
    size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t total = (size_t)nframes * g.plane;
    if (idx >= total) return;
    int v = 2 * (int)orig[idx] - (int)blurred_inout[idx];
    blurred_inout[idx] = (uint8_t)min(255, max(0, v));
}
