// k_cc.cuh -- connected components of the threshold image (upstream stage U4, SURVEY.md A.6;
// part of the native call at /root/reference/src/detection/tag_detector.py:26).
//
// Partition restated: pixels with value 127 are singletons.  A pixel (x, y) with 1 <= x <= w-2
// links to its equal-valued neighbours left (x-1, y) and, for y >= 1, up (x, y-1); white (255)
// pixels additionally link up-left and up-right.  Columns 0 and w-1 never initiate links.
//
// Three kernels (block-local union-find + boundary merge + canonical relabel):
//   k_cc_local    32x64-pixel tile per CTA, warp-ballot row runs + union-find in shared memory (atomicMin
//                 hooks: the root of a set is always its smallest pixel id), labels written as global
//                 pixel ids; sizes[] zeroed at tile-local roots.
//   k_cc_boundary one thread per tile-border pixel, lock-free unions across tile edges in global memory.
//   k_cc_finalize pointer jumping to the global root (= smallest pixel id of the component, so the
//                 labelling is canonical by construction) + component sizes, aggregated per tile in
//                 shared memory before the global atomics.
#pragma once
#include "common.cuh"

// tile of the local pass: one warp row = 32 pixels, 64 rows, 8 warps x 8 rows
#define CC_TW 32
#define CC_TH 64
#define CC_THREADS 256
#define CC_ROWS_PER_WARP 8
// tile of the finalize pass (8-pixel runs per thread)
#define CCF_TW 64
#define CCF_TH 32
#define CC_RUN 8

__device__ __forceinline__ uint32_t sfind(volatile uint32_t* L, uint32_t a) {
    uint32_t p = L[a];
    while (p != a) {
        a = p;
        p = L[a];
    }
    return a;
}

__device__ __forceinline__ void sunion(uint32_t* L, uint32_t a, uint32_t b) {
    for (;;) {
        a = sfind(L, a);
        b = sfind(L, b);
        if (a == b) return;
        if (a < b) { uint32_t t = a; a = b; b = t; }
        uint32_t old = atomicMin(&L[a], b);  // hook the larger root under the smaller
        if (old == a) return;
        a = old;  // a was hooked elsewhere meanwhile; keep uniting its new parent with b
    }
}

__device__ __forceinline__ uint32_t gfind(const uint32_t* L, uint32_t a) {
    uint32_t p = __ldcg(&L[a]);
    while (p != a) {
        a = p;
        p = __ldcg(&L[a]);
    }
    return a;
}

__device__ __forceinline__ void gunion(uint32_t* L, uint32_t a, uint32_t b) {
    for (;;) {
        a = gfind(L, a);
        b = gfind(L, b);
        if (a == b) return;
        if (a < b) { uint32_t t = a; a = b; b = t; }
        uint32_t old = atomicMin(&L[a], b);
        if (old == a) return;
        a = old;
    }
}

// Local pass.  Lane = column: a warp ballot finds the horizontal runs of a row (a pixel's first label is the
// start of its run -- no atomics), then only the first column of every vertical / diagonal contact between
// runs of adjacent rows issues a union, so the number of shared-memory atomics follows the number of runs,
// not the number of pixels.
__global__ void __launch_bounds__(CC_THREADS)
k_cc_local(const uint8_t* __restrict__ thresh, uint32_t* __restrict__ labels, uint32_t* __restrict__ sizes, Geom g,
           int tiles_x, int tiles_y) {
    __shared__ uint8_t sv[CC_TH][CC_TW];
    __shared__ uint32_t L[CC_TH * CC_TW];
    const int frame = blockIdx.z;
    const int x0 = blockIdx.x * CC_TW, y0 = blockIdx.y * CC_TH;
    const uint8_t* ft = thresh + (size_t)frame * g.plane;
    uint32_t* fl = labels + (size_t)frame * g.plane;
    uint32_t* fs = sizes + (size_t)frame * g.plane;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int x = x0 + lane;
    const bool col_init = x >= 1 && x <= g.wd - 2;   // columns 0 and w-1 never initiate links

    uint32_t v[CC_ROWS_PER_WARP], starts[CC_ROWS_PER_WARP];
#pragma unroll
    for (int k = 0; k < CC_ROWS_PER_WARP; k++) {
        const int ry = w * CC_ROWS_PER_WARP + k, gy = y0 + ry;
        uint32_t c = 127;
        if (gy < g.hd && x < g.wd) c = ft[(size_t)gy * g.wp + x];
        const uint32_t cl = __shfl_up_sync(FULL_MASK, c, 1);
        const bool link = lane > 0 && c != 127 && c == cl && col_init;
        const uint32_t st = __ballot_sync(FULL_MASK, !link);
        const int rs = 31 - __clz(st & (0xffffffffu >> (31 - lane)));
        L[ry * CC_TW + lane] = (uint32_t)(ry * CC_TW + rs);
        sv[ry][lane] = (uint8_t)c;
        v[k] = c;
        starts[k] = st;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CC_ROWS_PER_WARP; k++) {
        const int ry = w * CC_ROWS_PER_WARP + k, gy = y0 + ry;
        const uint32_t c = v[k];
        if (ry == 0 || gy < 1 || gy >= g.hd) continue;           // warp-uniform
        const uint32_t up = sv[ry - 1][lane];
        const uint32_t upl = __shfl_up_sync(FULL_MASK, up, 1), upr = __shfl_down_sync(FULL_MASK, up, 1);
        const uint32_t cl = __shfl_up_sync(FULL_MASK, c, 1);
        if (c == 127 || !col_init) continue;
        const uint32_t li = (uint32_t)(ry * CC_TW + lane);
        if (up == c) {
            // implied when the left pixel is chained to this one, to the pixel above it, and that one to `up`
            const bool implied = lane > 0 && cl == c && upl == c && x - 1 >= 1;
            if (!implied) sunion(L, li, li - CC_TW);
        }
        if (c == 255) {
            // up == c implies the diagonals through (x, y-1)'s left link and the left link initiated by
            // (x+1, y-1); the latter only exists when x+1 <= w-2
            if (lane > 0 && upl == 255 && up != 255) sunion(L, li, li - CC_TW - 1);
            if (lane < 31 && upr == 255 && (up != 255 || x + 1 > g.wd - 2)) sunion(L, li, li - CC_TW + 1);
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CC_ROWS_PER_WARP; k++) {
        const int ry = w * CC_ROWS_PER_WARP + k, gy = y0 + ry;
        const uint32_t st = starts[k];
        const int rs = 31 - __clz(st & (0xffffffffu >> (31 - lane)));
        const uint32_t li = (uint32_t)(ry * CC_TW + lane);
        uint32_t r = 0;
        if ((st >> lane) & 1u) r = sfind(L, li);
        r = __shfl_sync(FULL_MASK, r, rs);
        if (gy < g.hd && x < g.wp) {
            const uint32_t gid = (uint32_t)((y0 + (int)(r / CC_TW)) * g.wp + x0 + (int)(r % CC_TW));
            fl[(size_t)gy * g.wp + x] = gid;
            if (r == li && v[k] != 127) fs[gid] = 0;   // size counters start at zero at every tile-local root
        }
    }
}

// One thread per tile-border pixel: top row (32) + left column (64) + right column (64).
__global__ void __launch_bounds__(160)
k_cc_boundary(const uint8_t* __restrict__ thresh, uint32_t* __restrict__ labels, Geom g, int tiles_x, int tiles_y) {
    const int frame = blockIdx.z;
    const int x0 = blockIdx.x * CC_TW, y0 = blockIdx.y * CC_TH;
    const uint8_t* ft = thresh + (size_t)frame * g.plane;
    uint32_t* fl = labels + (size_t)frame * g.plane;
    const int t = threadIdx.x;
    int x, y;
    if (t < CC_TW) { x = x0 + t; y = y0; }
    else if (t < CC_TW + CC_TH) { x = x0; y = y0 + (t - CC_TW); if (t == CC_TW) return; }   // corner: top row's job
    else { x = x0 + CC_TW - 1; y = y0 + (t - CC_TW - CC_TH); if (t == CC_TW + CC_TH) return; }
    if (x < 1 || x > g.wd - 2 || y >= g.hd) return;
    const uint8_t c = ft[(size_t)y * g.wp + x];
    if (c == 127) return;
    const uint32_t id = (uint32_t)(y * g.wp + x);
    const bool left_edge = (x == x0), top_edge = (y == y0), right_edge = (x == x0 + CC_TW - 1);
    if (left_edge && ft[(size_t)y * g.wp + x - 1] == c) gunion(fl, id, id - 1);
    if (y >= 1) {
        const uint8_t* up = ft + (size_t)(y - 1) * g.wp;
        if (top_edge && up[x] == c) gunion(fl, id, id - g.wp);
        if (c == 255) {
            if ((top_edge || left_edge) && up[x - 1] == c) gunion(fl, id, id - g.wp - 1);
            if ((top_edge || right_edge) && up[x + 1] == c) gunion(fl, id, id - g.wp + 1);
        }
    }
}

#define CC_HASH 128
__global__ void __launch_bounds__(CC_THREADS)
k_cc_finalize(const uint8_t* __restrict__ thresh, uint32_t* __restrict__ labels, uint32_t* __restrict__ sizes, Geom g) {
    __shared__ uint32_t hkey[CC_HASH];
    __shared__ uint32_t hcnt[CC_HASH];
    const int frame = blockIdx.z;
    const int x0 = blockIdx.x * CCF_TW, y0 = blockIdx.y * CCF_TH;
    const uint8_t* ft = thresh + (size_t)frame * g.plane;
    uint32_t* fl = labels + (size_t)frame * g.plane;
    uint32_t* fs = sizes + (size_t)frame * g.plane;
    const int t = threadIdx.x;
    if (t < CC_HASH) { hkey[t] = 0xffffffffu; hcnt[t] = 0; }
    __syncthreads();
    const int ry = t >> 3, rx = (t & 7) * CC_RUN;
    const int gy = y0 + ry, gx = x0 + rx;
    if (gy < g.hd && gx < g.wp) {
        uint2 raw = *reinterpret_cast<const uint2*>(ft + (size_t)gy * g.wp + gx);
        uint4* lp = reinterpret_cast<uint4*>(fl + (size_t)gy * g.wp + gx);
        uint4 a = __ldcg(lp), b = __ldcg(lp + 1);
        uint32_t lab[CC_RUN] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t root[CC_RUN];
        uint32_t prev_lab = 0xffffffffu, prev_root = 0, run_root = 0xffffffffu, run_cnt = 0;
        bool changed = false;
#pragma unroll
        for (int k = 0; k < CC_RUN; k++) {
            uint32_t v = ((k < 4 ? raw.x : raw.y) >> (8 * (k & 3))) & 0xff;
            uint32_t r;
            if (lab[k] == prev_lab) r = prev_root;
            else r = gfind(fl, lab[k]);
            prev_lab = lab[k];
            prev_root = r;
            root[k] = r;
            changed |= (r != lab[k]);
            bool counted = (v != 127) && (gx + k < g.wd);
            if (counted) {
                if (r == run_root) run_cnt++;
                else {
                    if (run_cnt) {
                        // flush the previous run into the tile hash (linear probing; overflow -> global)
                        uint32_t hsh = (run_root * 2654435761u) >> 25;
                        bool done = false;
                        for (int probe = 0; probe < 8 && !done; probe++) {
                            uint32_t s = (hsh + probe) & (CC_HASH - 1);
                            uint32_t old = atomicCAS(&hkey[s], 0xffffffffu, run_root);
                            if (old == 0xffffffffu || old == run_root) { atomicAdd(&hcnt[s], run_cnt); done = true; }
                        }
                        if (!done) atomicAdd(&fs[run_root], run_cnt);
                    }
                    run_root = r;
                    run_cnt = 1;
                }
            }
        }
        if (run_cnt) {
            uint32_t hsh = (run_root * 2654435761u) >> 25;
            bool done = false;
            for (int probe = 0; probe < 8 && !done; probe++) {
                uint32_t s = (hsh + probe) & (CC_HASH - 1);
                uint32_t old = atomicCAS(&hkey[s], 0xffffffffu, run_root);
                if (old == 0xffffffffu || old == run_root) { atomicAdd(&hcnt[s], run_cnt); done = true; }
            }
            if (!done) atomicAdd(&fs[run_root], run_cnt);
        }
        if (changed) {
            lp[0] = make_uint4(root[0], root[1], root[2], root[3]);
            lp[1] = make_uint4(root[4], root[5], root[6], root[7]);
        }
    }
    __syncthreads();
    if (t < CC_HASH && hkey[t] != 0xffffffffu) atomicAdd(&fs[hkey[t]], hcnt[t]);
}
