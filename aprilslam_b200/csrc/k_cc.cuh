// k_cc.cuh -- connected components of the threshold image (upstream stage U4, SURVEY.md A.6;
// part of the native call at /root/reference/src/detection/tag_detector.py:26).
//
// Partition restated: pixels with value 127 are singletons.  A pixel (x, y) with 1 <= x <= w-2
// links to its equal-valued neighbours left (x-1, y) and, for y >= 1, up (x, y-1); white (255)
// pixels additionally link up-left and up-right.  Columns 0 and w-1 never initiate links.
//
// Kernels (block-local union-find + boundary merge + sizes [+ canonical relabel]):
//   k_cc_local    32x32-pixel tile per warp, bit-parallel row runs + union-find over runs in shared memory
//                 (atomicMin hooks: the root of a set is always its smallest pixel id); writes labels as
//                 global pixel ids, the pixel count of every tile-local root, and the list of those roots.
//   k_cc_boundary one thread per tile-border pixel, lock-free unions across tile edges in global memory.
//   k_cc_sizes    folds the counts of the tile-local roots into the final roots.
//   k_cc_flatten  pointer jumping to the global root (= smallest pixel id of the component, so the
//                 labelling is canonical by construction); only needed for the stage dumps, the
//                 pipeline resolves representatives on the fly.
#pragma once
#include "common.cuh"

// tile of the local pass: 32 x 32 pixels per warp (lane = ROW), 4 tiles side by side per CTA
#define CC_TW 32
#define CC_TH 32
#define CC_WARPS 4
#define CC_THREADS (CC_WARPS * 32)
#define CC_SUBLISTS 16   // root sub-lists per frame (tile row mod 16): spreads the append atomics over 16 counters
#define CC_PITCH 33   // run-start slots of row r live at r*33 + c (16-bit entries)
// tile of the flatten pass (8-pixel runs per thread)
#define CCF_TW 64
#define CCF_TH 32
#define CC_RUN 8

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
        uint32_t old = atomicMin(&L[a], b);  // hook the larger root under the smaller
        if (old == a) return;
        a = old;  // a was hooked elsewhere meanwhile; keep uniting its new parent with b
    }
}

// A run is identified inside its tile by the local pixel id of its first pixel, row*32 + col (< 1024, 16 bits):
// the whole union-find state of a tile is 2 x 2 KB of shared memory, so an SM holds ~50 tiles at a time.
__device__ __forceinline__ uint32_t cc_slot(uint32_t id) { return id + (id >> 5); }

// atomicMin on a 16-bit shared-memory entry (CAS on the containing word); returns the previous value
__device__ __forceinline__ uint32_t atomic_min_u16(uint16_t* base, uint32_t slot, uint32_t val) {
    uint32_t* word = reinterpret_cast<uint32_t*>(base) + (slot >> 1);
    const int sh = (slot & 1) * 16;
    uint32_t old = *reinterpret_cast<volatile uint32_t*>(word);
    for (;;) {
        const uint32_t cur = (old >> sh) & 0xffffu;
        if (cur <= val) return cur;
        const uint32_t assumed = old;
        old = atomicCAS(word, assumed, (assumed & ~(0xffffu << sh)) | (val << sh));
        if (old == assumed) return cur;
    }
}

__device__ __forceinline__ uint32_t sfind(volatile uint16_t* L, uint32_t a) {
    uint32_t p = L[cc_slot(a)];
    while (p != a) {
        a = p;
        p = L[cc_slot(a)];
    }
    return a;
}

__device__ __forceinline__ void sunion(uint16_t* L, uint32_t a, uint32_t b) {
    for (;;) {
        a = sfind(L, a);
        b = sfind(L, b);
        if (a == b) return;
        if (a < b) { uint32_t t = a; a = b; b = t; }
        uint32_t old = atomic_min_u16(L, cc_slot(a), b);
        if (old == a) return;
        a = old;
    }
}

// 4 pixels (one word) -> 4 mask bits
__device__ __forceinline__ uint32_t gather4(uint32_t t /* bytes of 0/1 */) { return (t * 0x01020408u) >> 24; }

// bits s..e of the run that starts at s, given the row's continuation bits
__device__ __forceinline__ uint32_t run_mask(uint32_t cont, int s) {
    const uint32_t t = s == 31 ? 0u : (cont >> (s + 1));
    const int e = s + __ffs(~t) - 1;            // trailing ones of t
    const uint32_t hi = e >= 31 ? 0xffffffffu : ((2u << e) - 1u);
    return hi & ~((1u << s) - 1u);
}

// Local pass, bit-parallel.  A lane owns one ROW of the 32x32 tile as two bit masks (white / black); the runs of
// the row come from shifts and ANDs of the lane's own word, the contacts with the row above from the masks of
// the lane above (one shuffle), and only runs -- not pixels -- take part in the shared-memory union-find.
// Besides the labels the pass leaves, for every tile-local root: its pixel count in sizes[] and its id in the
// frame's root list (k_cc_sizes folds the counts into the final roots after the boundary merges).
__global__ void __launch_bounds__(CC_THREADS)
k_cc_local(const uint8_t* __restrict__ thresh, uint32_t* __restrict__ labels, uint32_t* __restrict__ sizes,
           uint32_t* __restrict__ roots, int* __restrict__ nroots, Geom g, size_t sub_stride) {
    __shared__ __align__(16) uint16_t sL[CC_WARPS][CC_TH * CC_PITCH + 8];   // parent links, then pixel counters
    __shared__ __align__(16) uint16_t sX[CC_WARPS][CC_TH * CC_PITCH + 8];   // root of every run
    const int frame = blockIdx.z;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x0 = (blockIdx.x * CC_WARPS + w) * CC_TW, y0 = blockIdx.y * CC_TH;
    if (x0 >= g.wd) return;   // (no block-level synchronisation anywhere below)
    const int y = y0 + lane;
    const uint8_t* ft = thresh + (size_t)frame * g.plane;
    uint32_t* fl = labels + (size_t)frame * g.plane;
    uint32_t* fs = sizes + (size_t)frame * g.plane;
    uint16_t* L = sL[w];
    uint16_t* X = sX[w];
    const bool second = x0 + 32 <= g.wp;   // the row pitch is a multiple of 16, not of 32

    // ---- row masks
    uint32_t Wm = 0, Bm = 0;
    if (y < g.hd) {
        const uint4* rp = reinterpret_cast<const uint4*>(ft + (size_t)y * g.wp + x0);
        const uint4 a = __ldg(rp);
        uint4 b = make_uint4(0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu);
        if (second) b = __ldg(rp + 1);
        const uint32_t wd8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            Wm |= gather4((wd8[k] >> 7) & 0x01010101u) << (4 * k);   // 255: bit 7 set
            Bm |= gather4(~wd8[k] & 0x01010101u) << (4 * k);          // 0: bit 0 clear (127 and 255 have it set)
        }
    }
    const int ncols = min(32, g.wd - x0);
    const uint32_t V = ncols >= 32 ? 0xffffffffu : ((1u << ncols) - 1u);
    Wm &= V;
    Bm &= V;
    // initiator columns: 1 <= x <= w-2
    uint32_t I = V;
    if (x0 == 0) I &= ~1u;
    if (g.wd - 1 - x0 < 32) I &= ~(1u << (g.wd - 1 - x0));
    const uint32_t cw = Wm & (Wm << 1) & I, cb = Bm & (Bm << 1) & I;   // bit x: x continues the run of x-1
    const uint32_t Sw = Wm & ~cw, Sb = Bm & ~cb;                       // run starts
    const uint32_t S = Sw | Sb;
    const uint32_t rid0 = (uint32_t)(lane * 32);
    for (uint32_t m = S; m; m &= m - 1) {
        const uint32_t id = rid0 + __ffs(m) - 1;
        L[cc_slot(id)] = (uint16_t)id;
    }
    __syncwarp();

    // ---- contacts with the row above (lane 0's upper row belongs to another tile: k_cc_boundary)
    {
        uint32_t Wu = __shfl_up_sync(FULL_MASK, Wm, 1), Bu = __shfl_up_sync(FULL_MASK, Bm, 1);
        uint32_t cwu = __shfl_up_sync(FULL_MASK, cw, 1), cbu = __shfl_up_sync(FULL_MASK, cb, 1);
        uint32_t Swu = __shfl_up_sync(FULL_MASK, Sw, 1), Sbu = __shfl_up_sync(FULL_MASK, Sb, 1);
        if (lane == 0) { Wu = Bu = 0; }
        // white: 8-connected (up-left, up, up-right), black: 4-connected (up)
        for (uint32_t m = Sw; m; m &= m - 1) {
            const int s = __ffs(m) - 1;
            const uint32_t rI = run_mask(cw, s) & I;
            uint32_t touched = ((rI << 1) | rI | (rI >> 1)) & Wu;
            while (touched) {
                const int x = __ffs(touched) - 1;
                const int su = 31 - __clz(Swu & (0xffffffffu >> (31 - x)));
                sunion(L, rid0 + s, rid0 - 32 + su);
                touched &= ~run_mask(cwu, su);
            }
        }
        for (uint32_t m = Sb; m; m &= m - 1) {
            const int s = __ffs(m) - 1;
            uint32_t touched = run_mask(cb, s) & I & Bu;
            while (touched) {
                const int x = __ffs(touched) - 1;
                const int su = 31 - __clz(Sbu & (0xffffffffu >> (31 - x)));
                sunion(L, rid0 + s, rid0 - 32 + su);
                touched &= ~run_mask(cbu, su);
            }
        }
    }
    __syncwarp();

    // ---- pointer jumping: rows hook to the row above concurrently, which leaves chains as deep as the tile is
    //      tall; every round halves them (L[x] = L[L[x]] only ever moves an entry closer to its root)
    for (int round = 0; round < 10; round++) {
        bool changed = false;
        for (uint32_t m = S; m; m &= m - 1) {
            const uint32_t slot = cc_slot(rid0 + __ffs(m) - 1);
            const uint32_t par = L[slot];
            const uint32_t gpar = L[cc_slot(par)];
            if (gpar != par) { L[slot] = (uint16_t)gpar; changed = true; }
        }
        __syncwarp();
        if (!__any_sync(FULL_MASK, changed)) break;
    }
    // ---- root of every run, then pixel counts per tile-local root (L is reused as the counter array: two 16-bit
    //      counters per word, a tile holds 1024 pixels so a carry can never reach the upper counter)
    for (uint32_t m = S; m; m &= m - 1) {
        const uint32_t id = rid0 + __ffs(m) - 1;
        X[cc_slot(id)] = (uint16_t)sfind(L, id);
    }
    __syncwarp();
    int nroot = 0;
    for (uint32_t m = S; m; m &= m - 1) {
        const uint32_t id = rid0 + __ffs(m) - 1;
        if (X[cc_slot(id)] == id) { L[cc_slot(id)] = 0; nroot++; }
    }
    __syncwarp();
    for (uint32_t m = S; m; m &= m - 1) {
        const int s = __ffs(m) - 1;
        const uint32_t cont = ((Sw >> s) & 1u) ? cw : cb;
        const uint32_t slot = cc_slot(X[cc_slot(rid0 + s)]);
        atomicAdd(reinterpret_cast<uint32_t*>(L) + (slot >> 1), (uint32_t)__popc(run_mask(cont, s)) << ((slot & 1) * 16));
    }
    __syncwarp();
    // append the tile's roots to the frame's root list (one atomic per tile) and publish their counts
    {
        int incl = nroot;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int n = __shfl_up_sync(FULL_MASK, incl, off);
            if (lane >= off) incl += n;
        }
        const int total = __shfl_sync(FULL_MASK, incl, 31);
        const int sub = blockIdx.y & (CC_SUBLISTS - 1);
        int base = 0;
        if (lane == 31 && total) base = atomicAdd(&nroots[frame * CC_SUBLISTS + sub], total);
        base = __shfl_sync(FULL_MASK, base, 31);
        int o = base + incl - nroot;
        uint32_t* fr = roots + ((size_t)frame * CC_SUBLISTS + sub) * sub_stride;
        for (uint32_t m = S; m; m &= m - 1) {
            const int s = __ffs(m) - 1;
            if (X[cc_slot(rid0 + s)] == rid0 + s) {
                const uint32_t gid = (uint32_t)(y * g.wp + x0 + s);   // a root is a run of this very row
                fs[gid] = L[cc_slot(rid0 + s)];
                fr[o++] = gid;
            }
        }
    }
    // ---- labels: every pixel of a run gets the global id of the run's root, everything else its own id
    if (y < g.hd) {
        const uint32_t own0 = (uint32_t)(y * g.wp + x0);
        const uint32_t fg = Wm | Bm;
        uint32_t cur = 0;
        uint4* dst = reinterpret_cast<uint4*>(fl + (size_t)y * g.wp + x0);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            uint32_t o4[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int c = q * 4 + k;
                if ((S >> c) & 1u) {
                    const uint32_t r = X[cc_slot(rid0 + c)];
                    cur = (uint32_t)((y0 + (int)(r >> 5)) * g.wp + x0 + (int)(r & 31));
                }
                o4[k] = ((fg >> c) & 1u) ? cur : own0 + c;
            }
            if (q < 4 || second) dst[q] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        }
    }
}

// One thread per tile-border pixel (top row, left column, right column of every 32x32 tile): lock-free unions
// across tile borders.  As in the local pass only the first pixel of every contact issues a union.
#define CCB_ITEMS (CC_TW + 2 * CC_TH)
__global__ void __launch_bounds__(256)
k_cc_boundary(const uint8_t* __restrict__ thresh, uint32_t* __restrict__ labels, Geom g, int tiles_x, int tiles_y) {
    const int frame = blockIdx.z;
    const uint8_t* ft = thresh + (size_t)frame * g.plane;
    uint32_t* fl = labels + (size_t)frame * g.plane;
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)tiles_x * tiles_y * CCB_ITEMS;
    if (item >= total) return;
    const int tile = (int)(item / CCB_ITEMS), t = (int)(item - (long long)tile * CCB_ITEMS);
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int x0 = tx * CC_TW, y0 = ty * CC_TH;
    int x, y;
    if (t < CC_TW) { x = x0 + t; y = y0; }
    else if (t < CC_TW + CC_TH) { x = x0; y = y0 + (t - CC_TW); if (t == CC_TW) return; }          // corner: top row's job
    else { x = x0 + CC_TW - 1; y = y0 + (t - CC_TW - CC_TH); if (t == CC_TW + CC_TH) return; }
    if (x < 1 || x > g.wd - 2 || y >= g.hd) return;
    const uint8_t* row = ft + (size_t)y * g.wp;
    const uint8_t c = row[x];
    if (c == 127) return;
    const uint32_t id = (uint32_t)(y * g.wp + x);
    const bool left_edge = (x == x0), top_edge = (y == y0), right_edge = (x == x0 + CC_TW - 1);
    const uint8_t* up = y >= 1 ? row - g.wp : nullptr;
    if (left_edge && row[x - 1] == c) {
        // implied when the row above carries the same contact and both pixels hang on it
        const bool implied = !top_edge && y >= 1 && up[x] == c && up[x - 1] == c && x - 1 >= 1;
        if (!implied) gunion(fl, id, id - 1);
    }
    if (y >= 1) {
        const bool upsame = up[x] == c;
        if (top_edge && upsame) {
            const bool implied = x - 1 >= 1 && row[x - 1] == c && up[x - 1] == c;
            if (!implied) gunion(fl, id, id - g.wp);
        }
        if (c == 255) {
            if ((top_edge || left_edge) && up[x - 1] == c && !upsame) gunion(fl, id, id - g.wp - 1);
            if ((top_edge || right_edge) && up[x + 1] == c && (!upsame || x + 1 > g.wd - 2)) gunion(fl, id, id - g.wp + 1);
        }
    }
}

// Fold the pixel counts of the tile-local roots into the final roots (after the boundary merges) and point every
// tile-local root straight at its final root: afterwards the representative of ANY pixel is labels[labels[id]]
// (pixel -> tile-local root -> final root), two loads instead of a chain walk.
__global__ void __launch_bounds__(256)
k_cc_sizes(uint32_t* __restrict__ labels, uint32_t* __restrict__ sizes, const uint32_t* __restrict__ roots,
           const int* __restrict__ nroots, Geom g, size_t sub_stride) {
    const int frame = blockIdx.y / CC_SUBLISTS;
    const int n = nroots[blockIdx.y];
    uint32_t* fl = labels + (size_t)frame * g.plane;
    uint32_t* fs = sizes + (size_t)frame * g.plane;
    const uint32_t* fr = roots + (size_t)blockIdx.y * sub_stride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t a = fr[i];
        const uint32_t r = gfind(fl, a);
        if (r != a) {
            atomicAdd(&fs[r], fs[a]);
            fl[a] = r;   // (monotone: still an ancestor for any concurrent walk)
        }
    }
}

// Dense ids for the components that can carry an edge point (final roots of >= 25 pixels): the edge-cluster key
// becomes a pair of 16-bit ids instead of a pair of 21..23-bit pixel ids, which halves the radix-sort record
// and its number of passes.  dense[rep] is defined at every final root (0xffffffff: component smaller than 25
// pixels, upstream's edge-point filter); dense2rep maps back.
#define AGPU_MAX_DENSE 65536
__global__ void __launch_bounds__(256)
k_cc_dense(const uint32_t* __restrict__ labels, const uint32_t* __restrict__ sizes, const uint32_t* __restrict__ roots,
           const int* __restrict__ nroots, uint32_t* __restrict__ dense, uint32_t* __restrict__ dense2rep,
           int* __restrict__ ndense, Geom g, size_t sub_stride) {
    const int frame = blockIdx.y / CC_SUBLISTS;
    const int n = nroots[blockIdx.y];
    const uint32_t* fl = labels + (size_t)frame * g.plane;
    const uint32_t* fs = sizes + (size_t)frame * g.plane;
    const uint32_t* fr = roots + (size_t)blockIdx.y * sub_stride;
    uint32_t* fd = dense + (size_t)frame * g.plane;
    uint32_t* f2 = dense2rep + (size_t)frame * AGPU_MAX_DENSE;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t a = fr[i];
        if (__ldcg(&fl[a]) != a) continue;                             // not a final root
        if (__ldcg(&fs[a]) < 25u) { fd[a] = 0xffffffffu; continue; }   // too small to carry edge points
        const int d = atomicAdd(&ndense[frame], 1);
        if (d < AGPU_MAX_DENSE) {
            fd[a] = (uint32_t)d;
            f2[d] = a;
        }
    }
}

// Pointer jumping to the global root (= smallest pixel id of the component): the canonical labelling.  The
// pipeline itself resolves representatives on the fly (k_edges); this pass runs for the stage dumps.
__global__ void __launch_bounds__(256)
k_cc_flatten(uint32_t* __restrict__ labels, Geom g) {
    const int frame = blockIdx.z;
    const int x0 = blockIdx.x * CCF_TW, y0 = blockIdx.y * CCF_TH;
    uint32_t* fl = labels + (size_t)frame * g.plane;
    const int t = threadIdx.x;
    const int ry = t >> 3, rx = (t & 7) * CC_RUN;
    const int gy = y0 + ry, gx = x0 + rx;
    if (gy < g.hd && gx < g.wp) {
        uint4* lp = reinterpret_cast<uint4*>(fl + (size_t)gy * g.wp + gx);
        uint4 a = __ldcg(lp), b = __ldcg(lp + 1);
        uint32_t lab[CC_RUN] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t root[CC_RUN];
        uint32_t prev_lab = 0xffffffffu, prev_root = 0;
        bool changed = false;
#pragma unroll
        for (int k = 0; k < CC_RUN; k++) {
            uint32_t r = lab[k] == prev_lab ? prev_root : gfind(fl, lab[k]);
            prev_lab = lab[k];
            prev_root = r;
            root[k] = r;
            changed |= (r != lab[k]);
        }
        if (changed) {
            lp[0] = make_uint4(root[0], root[1], root[2], root[3]);
            lp[1] = make_uint4(root[4], root[5], root[6], root[7]);
        }
    }
}
