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
#include <cuda_fp16.h>
#include "common.cuh"
#include "k_cc.cuh"   // tile-major mask layout (cc_tile_index)

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

// Byte arithmetic on the packed-half pipe: a byte b becomes the half 0x6400 | b (= 1024 + b, exact), two per
// register, so min / max / compare of pixel pairs are single HMNMX2 / HSET2 instructions (the byte-SIMD
// __vminu4 family is emulated with ~10 LOP3/PRMT each on sm_100).
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 px_lo(uint32_t w) { return u2h(__byte_perm(w, 0x64646464u, 0x4140)); }
__device__ __forceinline__ __half2 px_hi(uint32_t w) { return u2h(__byte_perm(w, 0x64646464u, 0x4342)); }
#define H2_BIG 0x7bff7bffu   // 65504: neutral element of min
#define H2_ZERO 0x00000000u  // 0 < 1024: neutral element of max

// cv2.cvtColor(BGR2GRAY) of this OpenCV build: (B*3735 + G*19235 + R*9798 + 16384) >> 15
__device__ __forceinline__ uint32_t bgr2gray(uint32_t b, uint32_t g, uint32_t r) {
    return (b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15;
}

// Four interleaved BGR pixels (12 bytes in w0 w1 w2) -> four gray bytes.  A pixel's three bytes are gathered into one
// register with PRMT and go through two 16-bit x 8-bit dot products (DP2A): lower byte pair x (3735, 19235), upper
// byte pair x (9798, 0) -- the fourth byte is multiplied by zero, so whatever PRMT leaves there does not matter.
__device__ __forceinline__ uint32_t bgr4_to_gray4(uint32_t w0, uint32_t w1, uint32_t w2) {
    const uint32_t cBG = 3735u | (19235u << 16), cR = 9798u;
    const uint32_t p1 = __byte_perm(w0, w1, 0x6543), p2 = __byte_perm(w1, w2, 0x5432), p3 = w2 >> 8;
    const uint32_t a0 = __dp2a_hi(cR, w0, __dp2a_lo(cBG, w0, 16384u)) << 1;   // gray in byte 2
    const uint32_t a1 = __dp2a_hi(cR, p1, __dp2a_lo(cBG, p1, 16384u)) << 1;
    const uint32_t a2 = __dp2a_hi(cR, p2, __dp2a_lo(cBG, p2, 16384u)) << 1;
    const uint32_t a3 = __dp2a_hi(cR, p3, __dp2a_lo(cBG, p3, 16384u)) << 1;
    return __byte_perm(__byte_perm(a0, a1, 0x0062), __byte_perm(a2, a3, 0x0062), 0x5410);
}

// MASKS (decimate 1 only): instead of the threshold BYTES the kernel writes what the connected-components pass
// actually consumes -- the tile-major bit masks (2 bits per pixel: white / black, neither = 127) of k_cc.cuh.  A lane's
// 16 pixels are half a mask row; lanes pair up (1,2), (3,4), ... (a strip of 30 lanes starts on a 32-pixel tile
// boundary), swap two rows' worth of bits with one shuffle each, and every lane stores ONE 16-byte word (two mask
// rows) per tile row instead of four: the threshold image never reaches HBM (N/4 bytes written instead of N, and
// k_cc_local reads N/4 instead of N).
// CH == 3 (the reference's frames are BGR, tag_detector.py:25): the source is interleaved BGR; a lane
// reads 48 bytes (three LDG.128) per source row, converts its 16 pixels with cv2's fixed-point formula, writes them to
// the full-resolution gray plane `gray_out` (refine_edges and decode sample it; at decimate 1 it is also the quad image)
// and carries on with the gray words exactly as in the one-channel case: gray conversion, decimation and threshold in
// ONE pass over the frame.  With decimation the source rows between two sampled rows are converted as well.
template <int F, int MINB, bool MASKS = false, int CH = 1>
__global__ void __launch_bounds__(128, MINB)
k_decimate_threshold(const uint8_t* __restrict__ src, int W, int H, size_t src_stride, size_t src_frame_stride,
                     uint8_t* __restrict__ quad_im, uint8_t* __restrict__ thresh, Geom g, int nstrips, int nsegs,
                     int seg_tiles, int nframes, int min_diff, int vec_ok, uint2* __restrict__ masks = nullptr,
                     uint8_t* __restrict__ gray_out = nullptr, size_t gray_pitch = 0, size_t gray_frame = 0) {
    static_assert(!MASKS || F == 1, "mask output needs 16 pixels per lane");
    static_assert(CH == 1 || (CH == 3 && !MASKS), "BGR input writes threshold bytes");
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
    const bool lane_in = sc0 + 16 > 0 && sc0 < W;           // the lane's column chunk overlaps the frame
    const bool vec = vec_ok && sc0 >= 0 && sc0 + 16 <= W;   // ... and is a whole aligned 16-byte chunk
    // mask output: this lane's half of the 32-pixel mask row, its partner, and the tile column of the pair
    const bool m_low = (lane & 1) != 0;                         // lanes 1, 3, ...: columns 0-15 of the tile
    const int m_partner = (m_low ? lane + 1 : lane - 1) & 31;
    const int m_px0 = m_low ? px0 : px0 - 16;
    const bool m_out = MASKS && lane >= 1 && lane <= 30 && m_px0 < g.wd;
    uint32_t m_valid = 0;                                       // the lane's columns that exist in the image
    if (MASKS && px0 >= 0 && px0 < g.wd) m_valid = g.wd - px0 >= 16 ? 0xffffu : ((1u << (g.wd - px0)) - 1u);
    const int m_tiles_x = cc_tiles_x(g), m_tiles_y = cc_tiles_y(g);
    auto store_mask_rows = [&](int gy0, const uint32_t(&v)[4]) {   // v[r]: white bits | black bits << 16 of row gy0 + r
        // the low lane keeps rows 0-1 and takes the partner's, the high lane keeps rows 2-3
        const uint32_t ra = __shfl_sync(FULL_MASK, m_low ? v[2] : v[0], m_partner);
        const uint32_t rb = __shfl_sync(FULL_MASK, m_low ? v[3] : v[1], m_partner);
        if (!m_out) return;
        const uint32_t lo_a = m_low ? v[0] : ra, hi_a = m_low ? ra : v[2];
        const uint32_t lo_b = m_low ? v[1] : rb, hi_b = m_low ? rb : v[3];
        const size_t tile = ((size_t)frame * m_tiles_y + (gy0 >> 5)) * m_tiles_x + (m_px0 >> 5);
        uint2* dst = masks + tile * 32 + (gy0 & 31) + (m_low ? 0 : 2);
        *reinterpret_cast<uint4*>(dst) = make_uint4(__byte_perm(lo_a, hi_a, 0x5410), __byte_perm(lo_a, hi_a, 0x7632),
                                                    __byte_perm(lo_b, hi_b, 0x5410), __byte_perm(lo_b, hi_b, 0x7632));
    };

    uint8_t* fgray = CH == 3 ? gray_out + (size_t)frame * gray_frame : nullptr;
    const bool gray_lane = lane >= 1 && lane <= 30;          // (halo lanes' columns are stored by the neighbouring strip)
    auto load_row = [&](int gy, uint32_t(&w)[TPL], bool own = false) {   // own: this segment stores the row's gray pixels
#pragma unroll
        for (int j = 0; j < TPL; j++) w[j] = 0;
        if (gy >= g.hd || !lane_in) return;
        if (CH == 3) {
            // 16 source pixels of source row sy -> 16 gray bytes (stored to the gray plane when this segment owns the row)
            auto convert_row = [&](int sy, bool store) -> uint4 {
                const uint8_t* row = fsrc + (size_t)sy * src_stride;
                if (vec) {
                    const uint4* p = reinterpret_cast<const uint4*>(row + sc0 * 3);
                    const uint4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2);
                    const uint4 gq = make_uint4(bgr4_to_gray4(v0.x, v0.y, v0.z), bgr4_to_gray4(v0.w, v1.x, v1.y),
                                                bgr4_to_gray4(v1.z, v1.w, v2.x), bgr4_to_gray4(v2.y, v2.z, v2.w));
                    if (store) *reinterpret_cast<uint4*>(fgray + (size_t)sy * gray_pitch + sc0) = gq;
                    return gq;
                }
                uint32_t gw[4] = {0u, 0u, 0u, 0u};
#pragma unroll 1
                for (int k = 0; k < 16; k++) {
                    const long sx = sc0 + k;
                    if (sx < 0 || sx >= W) continue;
                    const uint32_t v = bgr2gray(row[sx * 3], row[sx * 3 + 1], row[sx * 3 + 2]);
                    gw[k >> 2] |= v << (8 * (k & 3));
                    if (store) fgray[(size_t)sy * gray_pitch + sx] = (uint8_t)v;
                }
                return make_uint4(gw[0], gw[1], gw[2], gw[3]);
            };
            const bool store = own && gray_lane;
            unpack16<F>(convert_row(gy * F, store), w);
            if (F > 1 && store) {   // the source rows between two sampled rows only feed the gray plane
#pragma unroll 1
                for (int sub = 1; sub < F; sub++)
                    if (gy * F + sub < H) (void)convert_row(gy * F + sub, true);
            }
            return;
        }
        const uint8_t* row = fsrc + (size_t)gy * F * src_stride;
        if (vec) {
            uint4 v = __ldg(reinterpret_cast<const uint4*>(row + sc0));
            unpack16<F>(v, w);
        } else {
            for (int k = 0; k < TPL * 4; k++) {
                long sx = sc0 + (long)k * F;
                if (sx >= 0 && sx < W) w[k >> 2] |= (uint32_t)row[sx] << (8 * (k & 3));
            }
        }
    };
    const int own_T0 = seg * seg_tiles, own_T1 = min(own_T0 + seg_tiles, th);
    auto load_tile_row = [&](int T, uint32_t(&px)[4][TPL]) {
        if (T >= 0 && T < th) {
#pragma unroll
            for (int r = 0; r < 4; r++) load_row(T * 4 + r, px[r], T >= own_T0 && T < own_T1);
        } else {
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int j = 0; j < TPL; j++) px[r][j] = 0;
        }
    };
    // per-tile extrema of tile row T as replicated half2 (neutral outside the tile grid), horizontally dilated
    auto tile_extrema = [&](int T, const uint32_t(&px)[4][TPL], uint32_t(&hmn)[TPL], uint32_t(&hmx)[TPL]) {
        uint32_t mn[TPL], mx[TPL];
        const bool rowv = T >= 0 && T < th;
#pragma unroll
        for (int j = 0; j < TPL; j++) {
            __half2 a = px_lo(px[0][j]), b = px_hi(px[0][j]);
            __half2 lo = __hmin2(a, b), hi = __hmax2(a, b);
#pragma unroll
            for (int r = 1; r < 4; r++) {
                a = px_lo(px[r][j]);
                b = px_hi(px[r][j]);
                lo = __hmin2(lo, __hmin2(a, b));
                hi = __hmax2(hi, __hmax2(a, b));
            }
            lo = __hmin2(lo, u2h(__byte_perm(h2u(lo), 0, 0x1032)));
            hi = __hmax2(hi, u2h(__byte_perm(h2u(hi), 0, 0x1032)));
            const int tx = tile0 + j;
            const bool v = rowv && tx >= 0 && tx < tw;
            mn[j] = v ? h2u(lo) : H2_BIG;
            mx[j] = v ? h2u(hi) : H2_ZERO;
        }
        uint32_t mn_l = __shfl_up_sync(FULL_MASK, mn[TPL - 1], 1), mn_r = __shfl_down_sync(FULL_MASK, mn[0], 1);
        uint32_t mx_l = __shfl_up_sync(FULL_MASK, mx[TPL - 1], 1), mx_r = __shfl_down_sync(FULL_MASK, mx[0], 1);
        if (lane == 0) { mn_l = H2_BIG; mx_l = H2_ZERO; }
        if (lane == 31) { mn_r = H2_BIG; mx_r = H2_ZERO; }
#pragma unroll
        for (int j = 0; j < TPL; j++) {
            const uint32_t l_mn = j == 0 ? mn_l : mn[(j + TPL - 1) % TPL], r_mn = j == TPL - 1 ? mn_r : mn[(j + 1) % TPL];
            const uint32_t l_mx = j == 0 ? mx_l : mx[(j + TPL - 1) % TPL], r_mx = j == TPL - 1 ? mx_r : mx[(j + 1) % TPL];
            hmn[j] = h2u(__hmin2(__hmin2(u2h(l_mn), u2h(mn[j])), u2h(r_mn)));
            hmx[j] = h2u(__hmax2(__hmax2(u2h(l_mx), u2h(mx[j])), u2h(r_mx)));
        }
    };
    // threshold + store the four rows of tile row T (or one leftover row) against dilated extrema d
    // (leftover pixels -- right of the last full tile column, or the rows below the last full tile row (`bottom`) --
    // are thresholded against the last full tile's extrema and never become 127: upstream's fix-up loop has no
    // low-contrast test)
    auto store_rows = [&](int gy0, int nrows, const uint32_t(&px)[4][TPL], const uint32_t(&dmn)[TPL],
                          const uint32_t(&dmx)[TPL], bool bottom) {
        // the tile column left of this lane's first tile: right leftover pixels (x >= 4*tw) use tile tw-1
        const uint32_t lmn = __shfl_up_sync(FULL_MASK, dmn[TPL - 1], 1), lmx = __shfl_up_sync(FULL_MASK, dmx[TPL - 1], 1);
        uint32_t thr2[TPL];
        bool flat[TPL];
#pragma unroll
        for (int j = 0; j < TPL; j++) {
            const int tx = tile0 + j;
            uint32_t a = dmn[j], b = dmx[j];
            if (tx >= tw) { a = j == 0 ? lmn : dmn[(j + TPL - 1) % TPL]; b = j == 0 ? lmx : dmx[(j + TPL - 1) % TPL]; }
            const int imn = a & 0xff, imx = b & 0xff;
            const int diff = imx - imn;
            flat[j] = !bottom && tx < tw && diff < min_diff;
            const uint32_t thr = (uint32_t)(imn + (diff >> 1));
            thr2[j] = 0x64006400u | thr | (thr << 16);
        }
        if (MASKS) {
            uint32_t nf = 0;   // pixels of non-flat tiles (flat ones are 127: neither white nor black)
#pragma unroll
            for (int j = 0; j < TPL; j++) nf |= flat[j] ? 0u : (0xfu << (4 * j));
            nf &= m_valid;
            uint32_t v[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                uint32_t acc = 0;   // pixel 4j+k above the threshold -> bit 4j+k (after folding the two halves)
#pragma unroll
                for (int j = 0; j < TPL; j++) {
                    const uint32_t m0 = __hgt2_mask(px_lo(px[r][j]), u2h(thr2[j]));
                    const uint32_t m1 = __hgt2_mask(px_hi(px[r][j]), u2h(thr2[j]));
                    acc |= ((m0 & 0x00020001u) | (m1 & 0x00080004u)) << (4 * j);
                }
                const uint32_t w16 = (acc | (acc >> 16)) & nf;
                v[r] = (r < nrows && gy0 + r < g.hd) ? (w16 | ((~w16 & nf) << 16)) : 0u;
            }
            store_mask_rows(gy0, v);
            return;
        }
        if (!is_out) return;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int gy = gy0 + r;
            if (r >= nrows || gy >= g.hd) break;
            uint32_t o[TPL];
#pragma unroll
            for (int j = 0; j < TPL; j++) {
                const uint32_t m0 = __hgt2_mask(px_lo(px[r][j]), u2h(thr2[j]));
                const uint32_t m1 = __hgt2_mask(px_hi(px[r][j]), u2h(thr2[j]));
                o[j] = flat[j] ? 0x7f7f7f7fu : __byte_perm(m0, m1, 0x6420);
            }
            const size_t off = (size_t)gy * g.wp + px0;
            if (TPL == 4) {
                __stcs(reinterpret_cast<uint4*>(fth + off), make_uint4(o[0], o[1 % TPL], o[2 % TPL], o[3 % TPL]));
                if (fq) *reinterpret_cast<uint4*>(fq + off) = make_uint4(px[r][0], px[r][1 % TPL], px[r][2 % TPL], px[r][3 % TPL]);
            } else if (TPL == 2) {
                *reinterpret_cast<uint2*>(fth + off) = make_uint2(o[0], o[1 % TPL]);
                if (fq) *reinterpret_cast<uint2*>(fq + off) = make_uint2(px[r][0], px[r][1 % TPL]);
            } else {
                *reinterpret_cast<uint32_t*>(fth + off) = o[0];
                if (fq) *reinterpret_cast<uint32_t*>(fq + off) = px[r][0];
            }
        }
    };

    const int T0 = seg * seg_tiles;
    const int T1 = min(T0 + seg_tiles, th);
    if (T0 >= th) return;

    // software pipeline, two tile rows deep: while tile row T is thresholded the loads of T+2 are in flight
    uint32_t cur[4][TPL], nxt[4][TPL], nn[4][TPL];
    uint32_t prevMn[TPL], prevMx[TPL], curMn[TPL], curMx[TPL], nextMn[TPL], nextMx[TPL];
    load_tile_row(T0 - 1, nn);
    load_tile_row(T0, cur);
    load_tile_row(T0 + 1, nxt);
    tile_extrema(T0 - 1, nn, prevMn, prevMx);
    tile_extrema(T0, cur, curMn, curMx);
    for (int T = T0; T < T1; T++) {
        load_tile_row(T + 2 < T1 + 1 ? T + 2 : th, nn);     // nothing beyond this segment's lower halo row
        tile_extrema(T + 1, nxt, nextMn, nextMx);
        uint32_t dmn[TPL], dmx[TPL];
#pragma unroll
        for (int j = 0; j < TPL; j++) {
            dmn[j] = h2u(__hmin2(__hmin2(u2h(prevMn[j]), u2h(curMn[j])), u2h(nextMn[j])));
            dmx[j] = h2u(__hmax2(__hmax2(u2h(prevMx[j]), u2h(curMx[j])), u2h(nextMx[j])));
        }
        store_rows(T * 4, 4, cur, dmn, dmx, false);
        if (T == th - 1 && (g.hd & 3)) {  // bottom leftover rows use the last tile row
            uint32_t lr[4][TPL];
#pragma unroll
            for (int r = 0; r < 4; r++) load_row(th * 4 + r, lr[r], true);
            store_rows(th * 4, g.hd - th * 4, lr, dmn, dmx, true);
        }
        if (MASKS && T == th - 1) {   // mask rows of the last tile row that lie below the image: empty
            const uint32_t zero[4] = {0u, 0u, 0u, 0u};
            for (int Tz = th + ((g.hd & 3) ? 1 : 0); Tz < m_tiles_y * 8; Tz++) store_mask_rows(Tz * 4, zero);
        }
#pragma unroll
        for (int j = 0; j < TPL; j++) {
            prevMn[j] = curMn[j]; prevMx[j] = curMx[j];
            curMn[j] = nextMn[j]; curMx[j] = nextMx[j];
        }
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int j = 0; j < TPL; j++) { cur[r][j] = nxt[r][j]; nxt[r][j] = nn[r][j]; }
    }
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

// U1 + U2 fused: decimate and blur in ONE pass (upstream image_u8_decimate + image_u8_gaussian_blur, SURVEY.md A.3 / A.4;
// `quad_sigma` of the detector, /root/reference/src/detection/tag_detector.py:18).  Separable integer Gaussian, rows then
// columns; positions outside [ksz/2, sz - ksz + ksz/2) are copied, as upstream's convolution does; sigma < 0 sharpens,
// v = clamp(2 * decimated - blurred).
//
// One CTA produces a 128 x 32 tile of the decimated image.  (1) The tile and its halo of ksz/2 pixels come in from HBM
// once -- 128-bit loads, decimation by 2 as a byte permute of two such loads -- into a shared-memory tile; (2) the row
// pass runs on that tile with DP4A (four taps per instruction on an unaligned four-byte window cut out of two words by
// a funnel shift) into a second shared tile; (3) the column pass reads words of four pixels and accumulates even and
// odd bytes as two 16-bit lanes of one register (the taps sum to at most 255, so a lane cannot overflow); (4) one
// 32-bit store per four pixels.  HBM traffic: F*F*N_d (what the decimation touches) in, N_d out -- the three-kernel
// version moved 7 N_d.
struct BlurKernel {
    int ksz;
    uint8_t k[64];
};
#define BL_TW 128
#define BL_TH 32
__host__ __device__ inline int bl_halo_cols(int half) { return (half + 15) & ~15; }            // halo rounded up to whole 16-byte chunks
__host__ __device__ inline int bl_pitch0(int half) { return BL_TW + 2 * bl_halo_cols(half) + 16; }
__host__ __device__ inline size_t bl_smem_bytes(int ksz) {
    const int half = ksz / 2, rows = BL_TH + 2 * half;
    return (size_t)rows * bl_pitch0(half) + (size_t)rows * BL_TW + 64;
}

template <int F>   // 1, 2: vector loads; 0: any factor (`Fdyn`), byte gathers
__global__ void __launch_bounds__(256)
k_decimate_blur(const uint8_t* __restrict__ src, size_t src_stride, size_t src_frame_stride, int Fdyn, int vec_ok,
                uint8_t* __restrict__ quad_im, Geom g, const __grid_constant__ BlurKernel bk, int sharpen) {
    extern __shared__ __align__(16) unsigned char bl_smem[];
    const int ksz = bk.ksz, half = ksz / 2, hal = bl_halo_cols(half);
    const int rows0 = BL_TH + 2 * half, pitch0 = bl_pitch0(half);
    uint8_t* S0 = bl_smem;                                  // [rows0][pitch0] decimated input, halo included
    uint8_t* S1 = bl_smem + (size_t)rows0 * pitch0;         // [rows0][BL_TW]  after the row pass
    uint32_t* KW = reinterpret_cast<uint32_t*>(S1 + (size_t)rows0 * BL_TW);   // [16] taps, four per word, zero padded
    const int frame = blockIdx.z;
    const int tx0 = blockIdx.x * BL_TW, ty0 = blockIdx.y * BL_TH;
    const int fac = F ? F : Fdyn;
    const uint8_t* fs = src + (size_t)frame * src_frame_stride;
    if (threadIdx.x < 16) {
        uint32_t wv = 0;
        for (int i = 0; i < 4; i++) {
            const int j = threadIdx.x * 4 + i;
            wv |= (uint32_t)(j < ksz ? bk.k[j] : 0) << (8 * i);
        }
        KW[threadIdx.x] = wv;
    }
    // ---- (1) load + decimate
    const int chunks = (BL_TW + 2 * hal) / 16;
    for (int i = threadIdx.x; i < rows0 * chunks; i += blockDim.x) {
        const int r = i / chunks, ch = i - r * chunks;
        const int gy = ty0 - half + r, gx = tx0 - hal + ch * 16;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (gy >= 0 && gy < g.hd && gx + 15 >= 0 && gx < g.wd) {
            const uint8_t* row = fs + (size_t)gy * fac * src_stride;
            if (F == 1 && vec_ok && gx >= 0 && gx + 16 <= g.wd) {
                v = __ldg(reinterpret_cast<const uint4*>(row + gx));
            } else if (F == 2 && vec_ok && gx >= 0 && gx + 16 < g.wd) {   // (< : an odd source width ends one byte early)
                const uint4 a = __ldg(reinterpret_cast<const uint4*>(row + 2 * gx));
                const uint4 b = __ldg(reinterpret_cast<const uint4*>(row + 2 * gx + 16));
                v.x = __byte_perm(a.x, a.y, 0x6420); v.y = __byte_perm(a.z, a.w, 0x6420);   // the even bytes
                v.z = __byte_perm(b.x, b.y, 0x6420); v.w = __byte_perm(b.z, b.w, 0x6420);
            } else {
                uint32_t wd4[4] = {0u, 0u, 0u, 0u};
                for (int k = 0; k < 16; k++) {
                    const int x = gx + k;
                    if (x >= 0 && x < g.wd) wd4[k >> 2] |= (uint32_t)row[(size_t)x * fac] << (8 * (k & 3));
                }
                v = make_uint4(wd4[0], wd4[1], wd4[2], wd4[3]);
            }
        }
        *reinterpret_cast<uint4*>(S0 + (size_t)r * pitch0 + ch * 16) = v;
    }
    __syncthreads();
    // ---- (2) row pass: four pixels (one word) per thread and step
    const int ngroups = (ksz + 3) >> 2;
    for (int i = threadIdx.x; i < rows0 * (BL_TW / 4); i += blockDim.x) {
        const int r = i / (BL_TW / 4), wq = i - r * (BL_TW / 4);
        const uint8_t* row = S0 + (size_t)r * pitch0;
        const int c0 = hal + 4 * wq;                         // column of the first of the four pixels inside S0
        uint32_t out = *reinterpret_cast<const uint32_t*>(row + c0);   // (the copy; pixels inside the range are replaced)
        const int gx0 = tx0 + 4 * wq;
        if (gx0 + 3 >= half && gx0 < g.wd - ksz + half) {
            const int a0 = c0 - half;                        // first byte of pixel 0's window
            const uint32_t* wrow = reinterpret_cast<const uint32_t*>(row) + (a0 >> 2);
            const int sh0 = a0 & 3;
            uint32_t acc[4] = {0u, 0u, 0u, 0u};
            uint32_t w0 = wrow[0], w1 = wrow[1];
            for (int jg = 0; jg < ngroups; jg++) {
                const uint32_t w2 = wrow[jg + 2];
                const uint32_t kw = KW[jg];
#pragma unroll
                for (int px = 0; px < 4; px++) {
                    const int sh = sh0 + px;                 // 0 .. 6
                    const uint32_t win = sh < 4 ? __funnelshift_r(w0, w1, 8 * sh) : __funnelshift_r(w1, w2, 8 * (sh - 4));
                    acc[px] = __dp4a(win, kw, acc[px]);
                }
                w0 = w1; w1 = w2;
            }
#pragma unroll
            for (int px = 0; px < 4; px++) {
                const int gx = gx0 + px;
                if (gx >= half && gx < g.wd - ksz + half) out = (out & ~(0xffu << (8 * px))) | (((acc[px] >> 8) & 0xffu) << (8 * px));
            }
        }
        *reinterpret_cast<uint32_t*>(S1 + (size_t)r * BL_TW + 4 * wq) = out;
    }
    __syncthreads();
    // ---- (3) column pass + (4) store
    uint8_t* fo = quad_im + (size_t)frame * g.plane;
    for (int i = threadIdx.x; i < BL_TH * (BL_TW / 4); i += blockDim.x) {
        const int r = i / (BL_TW / 4), wq = i - r * (BL_TW / 4);
        const int gy = ty0 + r, gx0 = tx0 + 4 * wq;
        if (gy >= g.hd || gx0 >= g.wp) continue;
        uint32_t out = *reinterpret_cast<const uint32_t*>(S1 + (size_t)(half + r) * BL_TW + 4 * wq);
        if (gy >= half && gy < g.hd - ksz + half) {
            uint32_t ae = 0u, ao = 0u;
            for (int j = 0; j < ksz; j++) {
                const uint32_t wv = *reinterpret_cast<const uint32_t*>(S1 + (size_t)(r + j) * BL_TW + 4 * wq);
                const uint32_t kj = (KW[j >> 2] >> (8 * (j & 3))) & 0xffu;
                ae += kj * (wv & 0x00ff00ffu);
                ao += kj * ((wv >> 8) & 0x00ff00ffu);
            }
            out = ((ae >> 8) & 0x00ff00ffu) | (ao & 0xff00ff00u);
        }
        if (sharpen) {
            const uint32_t orig = *reinterpret_cast<const uint32_t*>(S0 + (size_t)(half + r) * pitch0 + hal + 4 * wq);
            uint32_t sw = 0u;
#pragma unroll
            for (int px = 0; px < 4; px++) {
                const int v = 2 * (int)((orig >> (8 * px)) & 0xffu) - (int)((out >> (8 * px)) & 0xffu);
                sw |= (uint32_t)min(255, max(0, v)) << (8 * px);
            }
            out = sw;
        }
        // (bytes right of the image inside the row pitch stay zero, as k_pack leaves them)
        if (gx0 + 4 > g.wd) out &= gx0 >= g.wd ? 0u : (0xffffffffu >> (8 * (gx0 + 4 - g.wd)));
        *reinterpret_cast<uint32_t*>(fo + (size_t)gy * g.wp + gx0) = out;
    }
}

// U1 + U2, the common case (3, 5 or 7 taps -- quad_sigma up to 2 --, decimation 1 or 2, 16-byte aligned rows): NO shared
// memory.  One warp walks down a strip of 512 decimated pixels, lane = one 16-pixel chunk per row (one LDG.128, two for
// decimation 2 with the even bytes picked by a byte permute).  Row pass: the neighbour words come from the neighbour
// lanes by shuffle, every output pixel is one DP4A (two for 5 / 7 taps) on a four-byte window cut out of two words by a
// funnel shift.  Column pass: the row-pass results of the last KSZ rows stay in registers, split into even / odd bytes as
// 16-bit lanes, so a tap is two IMADs per four pixels; the row loop is unrolled KSZ times so the ring of rows is indexed
// statically.  One STG.128 per 16 pixels.  Borders as upstream (positions outside [ksz/2, sz - ksz + ksz/2) are copies).
#define BLS_ROWS 32      // output rows per warp (+ KSZ - 1 halo rows it recomputes)
template <int F, int KSZ, bool SHARPEN>
__global__ void __launch_bounds__(128)   // (occupancy bounds were tried: 6 / 4 / 3 CTAs per SM spill and lose 1 .. 15 %)
k_decimate_blur_strip(const uint8_t* __restrict__ src, size_t src_stride, size_t src_frame_stride,
                      uint8_t* __restrict__ quad_im, Geom g, const __grid_constant__ BlurKernel bk, int nstrips, int nsegs, int nframes) {
    constexpr int HALF = KSZ / 2, NG = (KSZ + 3) / 4;
    const int lane = threadIdx.x & 31;
    const long long unit = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (unit >= (long long)nstrips * nsegs * nframes) return;
    const int strip = (int)(unit % nstrips), seg = (int)((unit / nstrips) % nsegs), frame = (int)(unit / ((long long)nstrips * nsegs));
    const int x = strip * 512 + lane * 16;               // first decimated pixel of this lane's chunk
    const int y0 = seg * BLS_ROWS;
    const uint8_t* fs = src + (size_t)frame * src_frame_stride;
    uint8_t* fo = quad_im + (size_t)frame * g.plane;
    uint32_t kw[NG], kt[KSZ];
#pragma unroll
    for (int j = 0; j < KSZ; j++) kt[j] = bk.k[j];
#pragma unroll
    for (int gI = 0; gI < NG; gI++) {
        kw[gI] = 0;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (gI * 4 + i < KSZ) kw[gI] |= kt[gI * 4 + i] << (8 * i);
    }
    // per word: which of its four pixels are inside the horizontal convolution range / inside the image at all
    uint32_t cmask[4], vmask[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        cmask[j] = vmask[j] = 0;
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const int gx = x + 4 * j + p;
            if (gx >= HALF && gx < g.wd - KSZ + HALF) cmask[j] |= 0xffu << (8 * p);
            if (gx < g.wd) vmask[j] |= 0xffu << (8 * p);
        }
    }
    const bool in_pitch = x < g.wp;                      // (the pitch is a multiple of 16: a chunk is inside or outside)
    const bool have_l = x >= 4, have_r = x + 16 < g.wp;  // a word left / right of the chunk exists
    // one row of decimated pixels: the lane's four words + the word before and after them
    auto fetch_row = [&](int gy, uint4& v, uint32_t& wl, uint32_t& wr) {
        v = make_uint4(0u, 0u, 0u, 0u);
        wl = 0u; wr = 0u;
        if (gy >= 0 && gy < g.hd && in_pitch) {
            const uint8_t* row = fs + (size_t)gy * F * src_stride;
            if (F == 1) {
                v = __ldg(reinterpret_cast<const uint4*>(row + x));
                if (lane == 0 && have_l) wl = __ldg(reinterpret_cast<const uint32_t*>(row + x - 4));
                if (lane == 31 && have_r) wr = __ldg(reinterpret_cast<const uint32_t*>(row + x + 16));
            } else if ((size_t)(2 * x + 40) <= src_stride) {
                const uint4 a = __ldg(reinterpret_cast<const uint4*>(row + 2 * x));
                const uint4 b = __ldg(reinterpret_cast<const uint4*>(row + 2 * x + 16));
                v.x = __byte_perm(a.x, a.y, 0x6420); v.y = __byte_perm(a.z, a.w, 0x6420);
                v.z = __byte_perm(b.x, b.y, 0x6420); v.w = __byte_perm(b.z, b.w, 0x6420);
                if (lane == 0 && have_l) {
                    const uint2 c = __ldg(reinterpret_cast<const uint2*>(row + 2 * x - 8));
                    wl = __byte_perm(c.x, c.y, 0x6420);
                }
                if (lane == 31 && have_r) {
                    const uint2 c = __ldg(reinterpret_cast<const uint2*>(row + 2 * x + 32));
                    wr = __byte_perm(c.x, c.y, 0x6420);
                }
            } else {   // the chunk at the end of the source row: byte by byte, inside the image only
                uint32_t w6[6] = {0u, 0u, 0u, 0u, 0u, 0u};
                for (int k = -4; k < 20; k++) {
                    const int gx = x + k;
                    if (gx >= 0 && gx < g.wd) w6[(k + 4) >> 2] |= (uint32_t)row[(size_t)gx * 2] << (8 * ((k + 4) & 3));
                }
                v = make_uint4(w6[1], w6[2], w6[3], w6[4]);
                wl = w6[0]; wr = w6[5];
            }
        }
    };
    auto assemble_row = [&](const uint4& v, uint32_t wl, uint32_t wr, uint32_t (&W)[7]) {
        const uint32_t nl = __shfl_up_sync(FULL_MASK, v.w, 1), nr = __shfl_down_sync(FULL_MASK, v.x, 1);
        W[0] = lane == 0 ? wl : nl;
        W[1] = v.x; W[2] = v.y; W[3] = v.z; W[4] = v.w;
        W[5] = lane == 31 ? wr : nr;
        W[6] = 0u;
    };
    uint32_t Re[KSZ][4], Ro[KSZ][4];      // ring of row-pass results, even / odd bytes as 16-bit lanes
    uint32_t Og[SHARPEN ? KSZ : 1][4];    // ring of the decimated pixels themselves (sharpen only)
#pragma unroll
    for (int s = 0; s < KSZ; s++)
#pragma unroll
        for (int j = 0; j < 4; j++) Re[s][j] = Ro[s][j] = 0;
    // rows y0 - HALF .. y0 + BLS_ROWS - 1 + HALF come in; row `gy` completes the window of output row gy - HALF
    const int ylast = min(y0 + BLS_ROWS, g.hd) - 1 + HALF;
#pragma unroll 1
    for (int yb = y0 - HALF; yb <= ylast; yb += KSZ) {
#pragma unroll
        for (int s = 0; s < KSZ; s++) {
            const int gy = yb + s;
            if (gy > ylast) break;
            uint32_t W[7];
            {
                uint4 cv;
                uint32_t cwl, cwr;
                fetch_row(gy, cv, cwl, cwr);
                assemble_row(cv, cwl, cwr, W);
            }
            // ---- row pass
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t acc[4];
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    acc[p] = 0;
#pragma unroll
                    for (int gI = 0; gI < NG; gI++) {
                        const int off = p - HALF + 4 * gI;                  // byte offset of the window from the word's start
                        const int q = off >= 0 ? off / 4 : -((3 - off) / 4);  // floor(off / 4)
                        const int sh = off - 4 * q;                          // 0 .. 3
                        const uint32_t lo = W[1 + j + q], hi = W[2 + j + q];
                        const uint32_t win = sh == 0 ? lo : __funnelshift_r(lo, hi, 8 * sh);
                        acc[p] = __dp4a(win, kw[gI], acc[p]);
                    }
                }
                const uint32_t conv = ((acc[0] >> 8) & 0xffu) | (acc[1] & 0xff00u) | ((acc[2] << 8) & 0xff0000u) | ((acc[3] << 16) & 0xff000000u);
                const uint32_t r = (conv & cmask[j]) | (W[1 + j] & ~cmask[j]);
                Re[s][j] = r & 0x00ff00ffu;
                Ro[s][j] = (r >> 8) & 0x00ff00ffu;
                if (SHARPEN) Og[SHARPEN ? s : 0][j] = W[1 + j];
            }
            // ---- column pass for output row yo = gy - HALF (its window is the ring, oldest row first: slot s + 1)
            const int yo = gy - HALF;
            if (yo >= y0 && yo < g.hd && in_pitch) {
                uint4 o;
                uint32_t ow[4];
                const bool vconv = yo >= HALF && yo < g.hd - KSZ + HALF;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int sc = (s + KSZ - HALF) % KSZ;                   // slot of row yo itself
                    uint32_t word = Re[sc][j] | (Ro[sc][j] << 8);
                    if (vconv) {
                        uint32_t ae = 0u, ao = 0u;
#pragma unroll
                        for (int t = 0; t < KSZ; t++) {
                            const int st = (s + 1 + t) % KSZ;                // tap t sits on row gy - (KSZ - 1) + t
                            ae += kt[t] * Re[st][j];
                            ao += kt[t] * Ro[st][j];
                        }
                        word = ((ae >> 8) & 0x00ff00ffu) | (ao & 0xff00ff00u);
                    }
                    if (SHARPEN) {
                        const uint32_t og = Og[SHARPEN ? sc : 0][j];   // the decimated pixels of row yo itself
                        const uint32_t oe = og & 0x00ff00ffu, oo = (og >> 8) & 0x00ff00ffu;
                        const uint32_t be = word & 0x00ff00ffu, bo = (word >> 8) & 0x00ff00ffu;
                        // per 16-bit lane: clamp(2 * orig - blurred, 0, 255)
                        const uint32_t de = __vminu2(__vmaxs2(__vsub2(__vadd2(oe, oe), be), 0u), 0x00ff00ffu);
                        const uint32_t dn = __vminu2(__vmaxs2(__vsub2(__vadd2(oo, oo), bo), 0u), 0x00ff00ffu);
                        word = de | (dn << 8);
                    }
                    ow[j] = word & vmask[j];
                }
                o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                *reinterpret_cast<uint4*>(fo + (size_t)yo * g.wp + x) = o;
            }
        }
    }
}
