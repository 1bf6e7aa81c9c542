// k_cc.cuh -- connected components of the threshold image (upstream stage U4, SURVEY.md A.6;
// part of the native call at /root/reference/src/detection/tag_detector.py:26).
//
// Partition restated: pixels with value 127 are singletons.  A pixel (x, y) with 1 <= x <= w-2
// links to its equal-valued neighbours left (x-1, y) and, for y >= 1, up (x, y-1); white (255)
// pixels additionally link up-left and -- unless the upper neighbour is white as well (upstream's
// guard; only at x = w-2 is that link not implied by the others) -- up-right.  Columns 0 and w-1
// never initiate links.
//
// Data layout (per frame, TILE-MAJOR: tile t = ty * tiles_x + tx covers 32x32 pixels):
//   masks  uint2 [tiles][32]     row r of the tile as two bit masks {white, black} (bit c = column c); everything
//                                downstream of the local pass (boundary merge, edge points) works on these
//                                2 bits per pixel instead of the threshold bytes
//   l16    u16   [tiles][1024]   SPARSE: defined only at the first pixel of every run; the tile-local root of the
//                                run as its ORDINAL among the tile's roots.  A pixel finds its run start with two
//                                bit operations on its row mask, so no per-pixel label image is ever written.
//   root tables (CcRoots)        COMPACT: the tile-local roots of a frame are numbered by their position in the
//                                frame's root lists (16 sub-lists by tile row); a root's HANDLE = sub-list * capacity
//                                + position, tile_base[tile] = handle of the tile's first root.  Union-find links,
//                                pixel counts, dense ids, the root's pixel id and the smallest pixel id of its set
//                                are arrays over handles: a few thousand entries per frame that stay in L2, instead
//                                of sparse touches all over three plane-sized u32 arrays.
//
// Kernels (block-local union-find + boundary merge + sizes [+ canonical relabel]):
//   k_cc_local    32x32-pixel tile per warp, bit-parallel row runs + union-find over runs in shared memory
//                 (atomicMin hooks: the root of a set is always its smallest pixel id); writes the masks, the
//                 run-start labels, the pixel count of every tile-local root, and the list of those roots.
//   k_cc_boundary one warp per tile, bit-parallel contact tests along the tile's top / left / right border,
//                 lock-free unions between the tile-local roots in global memory.
//   k_cc_sizes    folds the counts of the tile-local roots into the final roots.
//   k_cc_canonical per-pixel canonical labels (smallest pixel id of the component); stage dumps only, the
//                 pipeline resolves representatives on the fly.
#pragma once
#include "common.cuh"

// tile of the local pass: 32 x 32 pixels per warp (lane = ROW), CC_WARPS tiles side by side per CTA
#define CC_TW 32
#define CC_TH 32
#ifndef CC_WARPS
#define CC_WARPS 2   // (1: 461, 2: 450, 4: 463, 8: 510 us for the CC stage of 128 frames -- a flat tile's warp leaves at once, and a
#endif               //  small CTA gives its slot back sooner)
#define CC_THREADS (CC_WARPS * 32)
#define CC_SUBLISTS 16   // root sub-lists per frame (tile row mod 16): spreads the append atomics over 16 counters
#define CC_PITCH 33   // run-start slots of row r live at r*33 + c (16-bit entries)

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

__host__ __device__ __forceinline__ int cc_tiles_x(const Geom& g) { return (g.wd + CC_TW - 1) / CC_TW; }
__host__ __device__ __forceinline__ int cc_tiles_y(const Geom& g) { return (g.hd + CC_TH - 1) / CC_TH; }
__host__ __device__ __forceinline__ size_t cc_tile_index(const Geom& g, int frame, int tx, int ty) {
    return ((size_t)frame * cc_tiles_y(g) + ty) * cc_tiles_x(g) + tx;
}

// initiator columns of the tile column block that starts at x0: 1 <= x <= wd-2
__device__ __forceinline__ uint32_t cc_initiators(int x0, int wd) {
    const int ncols = min(32, wd - x0);
    uint32_t I = ncols >= 32 ? 0xffffffffu : (ncols <= 0 ? 0u : ((1u << ncols) - 1u));
    if (x0 == 0) I &= ~1u;
    if (wd - 1 - x0 < 32 && wd - 1 - x0 >= 0) I &= ~(1u << (wd - 1 - x0));
    return I;
}

// first column of the run that contains column c (M: the row mask of the pixel's own colour, bit c set)
__device__ __forceinline__ int cc_run_start(uint32_t M, uint32_t I, int c) {
    const uint32_t S = M & ~(M & (M << 1) & I);
    return 31 - __clz(S & (0xffffffffu >> (31 - c)));
}

// Root tables of a chunk.  All per-frame arrays have `cap` = CC_SUBLISTS * sub_cap entries per frame; a handle is an
// index into a frame's arrays.  When a frame has more tile-local roots than a sub-list holds the surplus tiles get no
// handles (tile_base = CC_NO_HANDLE), every consumer treats their pixels as "no component", and the host re-runs the
// chunk with larger lists.
#define CC_NO_HANDLE 0xffffffffu
struct CcRoots {
    uint32_t* links;      // [nframes][cap] union-find parent (a handle); root of a set = its smallest handle
    uint32_t* sizes;      // [nframes][cap] pixel count: of the tile-local root, after k_cc_sizes of the whole set at its root
    uint32_t* rootpix;    // [nframes][cap] raster pixel id (y * wp + x) of the tile-local root
    uint32_t* minpix;     // [nframes][cap] smallest root pixel id of the set, at its final root = the component's canonical label
    uint32_t* dense;      // [nframes][cap] dense component id at final roots (0xffffffff: fewer than 25 pixels)
    uint32_t* tile_base;  // [nframes][tiles] handle of the tile's first root
    int* nroots;          // [nframes][CC_SUBLISTS] roots appended per sub-list (may exceed sub_cap)
    int sub_cap;          // capacity of one sub-list
    int ntiles;           // tiles per frame
    __host__ __device__ size_t cap() const { return (size_t)CC_SUBLISTS * sub_cap; }
};

// handle of the tile-local root of the foreground pixel (column c, row r) of tile `tile` (frame-relative tile index)
__device__ __forceinline__ uint32_t cc_pixel_root(const uint16_t* __restrict__ l16, const uint32_t* __restrict__ tile_base,
                                                  size_t tile, int r, int c, uint32_t M, uint32_t I) {
    const int s = cc_run_start(M, I, c);
    const uint32_t base = __ldg(&tile_base[tile]);
    const uint32_t ord = l16[tile * 1024 + r * 32 + s];
    return base == CC_NO_HANDLE ? CC_NO_HANDLE : base + ord;
}

// Local pass, bit-parallel and RUN-BALANCED.  The row masks of the 32x32 tile are built with lane = row (two bit
// masks, white / black); the runs of every row come from shifts and ANDs of the lane's own word.  Everything after
// that works on the tile's RUN LIST (one 16-bit entry per run, in shared memory) dealt round-robin to the lanes:
// a tile that crosses a tag has a few rows with many runs and many rows with one, so "lane = row" loops idle half
// of the warp, while "lane = run j, j + 32, ..." keeps every lane busy in every phase (contacts with the row above,
// pointer jumping, pixel counts, label stores).  Only runs -- not pixels -- take part in the shared-memory union-find.
// Besides the labels the pass leaves, for every tile-local root: its pixel count in sizes[] and its id in the
// frame's root list (k_cc_sizes folds the counts into the final roots after the boundary merges).
// FROM_MASKS: the tile-major masks were already written by the threshold kernel (decimate 1); otherwise they are
// built here from the threshold bytes and stored.
#define CC_RUN_COLOUR 0x400u   // run list entry: local pixel id of the first pixel (10 bits) | colour (1 = black)
template <bool FROM_MASKS>
__global__ void __launch_bounds__(CC_THREADS)
k_cc_local(const uint8_t* __restrict__ thresh, uint2* __restrict__ masks, uint16_t* __restrict__ l16, CcRoots rt, Geom g) {
    __shared__ __align__(16) uint16_t sL[CC_WARPS][CC_TH * CC_PITCH + 8];   // parent links; at roots later the pixel counters
    __shared__ __align__(16) uint16_t sR[CC_WARPS][CC_TW * CC_TH];          // run list
    __shared__ uint2 sM[CC_WARPS][CC_TH];                                    // row masks {white, black}
    const int frame = blockIdx.z;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x0 = (blockIdx.x * CC_WARPS + w) * CC_TW, y0 = blockIdx.y * CC_TH;
    if (x0 >= g.wd) return;   // (no block-level synchronisation anywhere below)
    const int y = y0 + lane;
    uint16_t* L = sL[w];
    uint16_t* R = sR[w];
    uint2* Ms = sM[w];
    const size_t tile = cc_tile_index(g, frame, blockIdx.x * CC_WARPS + w, blockIdx.y);

    // ---- row masks (lane = row)
    uint32_t Wm = 0, Bm = 0;
    if (FROM_MASKS) {
        const uint2 m = __ldg(&masks[tile * 32 + lane]);
        Wm = m.x; Bm = m.y;
    } else {
        const uint8_t* ft = thresh + (size_t)frame * g.plane;
        const bool second = x0 + 32 <= g.wp;   // the row pitch is a multiple of 16, not of 32
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
        masks[tile * 32 + lane] = make_uint2(Wm, Bm);
    }
    if (!__any_sync(FULL_MASK, (Wm | Bm) != 0u)) {          // nothing but 127-pixels: no runs, no labels, no roots
        if (lane == 0) rt.tile_base[(size_t)frame * rt.ntiles + (size_t)blockIdx.y * cc_tiles_x(g) + blockIdx.x * CC_WARPS + w] = CC_NO_HANDLE;
        return;
    }
    const uint32_t I = cc_initiators(x0, g.wd);   // initiator columns: 1 <= x <= w-2

    // ---- run list: the runs of row r occupy the entries [off(r), off(r) + popc(S_r)), white and black in column order
    int nrun;
    {
        const uint32_t cw = Wm & (Wm << 1) & I, cb = Bm & (Bm << 1) & I;   // bit x: x continues the run of x-1
        const uint32_t Sw = Wm & ~cw, Sb = Bm & ~cb;                       // run starts
        Ms[lane] = make_uint2(Wm, Bm);
        const int mine = __popc(Sw | Sb);
        int incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int n = __shfl_up_sync(FULL_MASK, incl, off);
            if (lane >= off) incl += n;
        }
        nrun = __shfl_sync(FULL_MASK, incl, 31);
        int o = incl - mine;
        const uint32_t rid0 = (uint32_t)(lane * 32);
        for (uint32_t m = Sw | Sb; m; m &= m - 1) {
            const int c = __ffs(m) - 1;
            const uint32_t id = rid0 + c;
            R[o++] = (uint16_t)(id | (((Sb >> c) & 1u) ? CC_RUN_COLOUR : 0u));
            L[cc_slot(id)] = (uint16_t)id;
        }
    }
    __syncwarp();

    // ---- contacts with the row above (row 0's upper row belongs to another tile: k_cc_boundary); lane = run.
    //      White: 8-connected (up-left, up, up-right), black: 4-connected (up).  Upstream skips the up-right link when the
    //      upper neighbour is white too; inside the image that link is implied (up + the upper row's own run), but at
    //      x = wd-2 it is not (column wd-1 never continues a run), and the partition has to be upstream's.
    //      Returns the mask of upper-row pixels the run touches (0 for row 0) and the upper row's masks.
    auto contacts = [&](uint32_t e, uint32_t& cu, uint32_t& Su) -> uint32_t {
        const uint32_t id = e & 1023u;
        const int r = id >> 5, s = id & 31;
        if (r == 0) return 0u;
        const bool black = (e & CC_RUN_COLOUR) != 0u;
        const uint2 mr = Ms[r], mu = Ms[r - 1];
        const uint32_t M = black ? mr.y : mr.x, Mu = black ? mu.y : mu.x;
        cu = Mu & (Mu << 1) & I;
        Su = Mu & ~cu;
        const uint32_t rI = run_mask(M & (M << 1) & I, s) & I;
        return (black ? rI : (((rI & ~Mu) << 1) | rI | (rI >> 1))) & Mu;
    };
    // Part 1: every run links to the FIRST upper run it touches -- a plain store into its own entry, no find, no atomic
    // (nobody else writes that entry here).  A link climbs exactly one row, so the chains are at most 31 long.  Runs that
    // touch further upper runs (the places where two branches of a component meet) are remembered for part 3.
    uint32_t morebits = 0u;
    for (int j = lane, k = 0; j < nrun; j += 32, k++) {
        const uint32_t e = R[j];
        uint32_t cu, Su;
        const uint32_t touched = contacts(e, cu, Su);
        if (touched) {
            const uint32_t id = e & 1023u;
            const int x = __ffs(touched) - 1;
            const int su = 31 - __clz(Su & (0xffffffffu >> (31 - x)));
            L[cc_slot(id)] = (uint16_t)(id - (id & 31u) - 32 + su);
            if (touched & ~run_mask(cu, su)) morebits |= 1u << k;
        }
    }
    __syncwarp();
    // Part 2: pointer jumping -- every round halves the chains (L[x] = L[L[x]] only ever moves an entry closer to its
    // root), five rounds at most
    for (int round = 0; round < 5; round++) {
        bool changed = false;
        for (int j = lane; j < nrun; j += 32) {
            const uint32_t slot = cc_slot(R[j] & 1023u);
            const uint32_t par = L[slot];
            const uint32_t gpar = L[cc_slot(par)];
            if (gpar != par) { L[slot] = (uint16_t)gpar; changed = true; }
        }
        __syncwarp();
        if (!__any_sync(FULL_MASK, changed)) break;
    }
    // Part 3: the remaining contacts are real unions between (now flat) trees: lock-free, larger root under the smaller
    if (__any_sync(FULL_MASK, morebits != 0u)) {
        for (int j = lane, k = 0; j < nrun; j += 32, k++) {
            if (!((morebits >> k) & 1u)) continue;
            const uint32_t e = R[j];
            const uint32_t id = e & 1023u;
            uint32_t cu, Su;
            uint32_t touched = contacts(e, cu, Su);
            bool first = true;
            while (touched) {
                const int x = __ffs(touched) - 1;
                const int su = 31 - __clz(Su & (0xffffffffu >> (31 - x)));
                if (!first) sunion(L, id, id - (id & 31u) - 32 + su);
                first = false;
                touched &= ~run_mask(cu, su);
            }
        }
        __syncwarp();
    }
    // ---- root of every run; the entry of a ROOT then becomes its pixel counter (two 16-bit entries per word; id + count
    //      <= 2047, so a carry can never leave a counter), the entries of the other runs keep their root.
    //      A lane remembers which of its runs are roots in a bit mask (run j = lane + 32 k -> bit k).
    uint32_t rootbits = 0u;
    int nroot = 0;
    for (int j = lane, k = 0; j < nrun; j += 32, k++) {
        const uint32_t id = R[j] & 1023u;
        const uint32_t root = sfind(L, id);
        if (root == id) { rootbits |= 1u << k; nroot++; }
        else L[cc_slot(id)] = (uint16_t)root;
    }
    __syncwarp();
    // (a root's entry holds its own id, < 1024: the counts are added on top of it and the id is taken off again below)
    for (int j = lane, k = 0; j < nrun; j += 32, k++) {
        const uint32_t e = R[j];
        const uint32_t id = e & 1023u;
        const int r = id >> 5, s = id & 31;
        const uint2 mr = Ms[r];
        const uint32_t M = (e & CC_RUN_COLOUR) ? mr.y : mr.x;
        const uint32_t root = ((rootbits >> k) & 1u) ? id : (uint32_t)L[cc_slot(id)];
        const uint32_t slot = cc_slot(root);
        atomicAdd(reinterpret_cast<uint32_t*>(L) + (slot >> 1),
                  (uint32_t)__popc(run_mask(M & (M << 1) & I, s)) << ((slot & 1) * 16));
    }
    __syncwarp();
    // ---- the tile's roots get consecutive handles in the frame's root tables (one atomic per tile); their counter
    //      entries then become their ORDINAL inside the tile, which is what the run-start labels store
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
        if (lane == 31) base = atomicAdd(&rt.nroots[frame * CC_SUBLISTS + sub], total);
        base = __shfl_sync(FULL_MASK, base, 31);
        const bool fits = base + total <= rt.sub_cap;
        const uint32_t handle0 = (uint32_t)(sub * rt.sub_cap + base);
        const size_t ftile = (size_t)blockIdx.y * cc_tiles_x(g) + blockIdx.x * CC_WARPS + w;
        if (lane == 0) rt.tile_base[(size_t)frame * rt.ntiles + ftile] = fits ? handle0 : CC_NO_HANDLE;
        int o = incl - nroot;
        const size_t fo = (size_t)frame * rt.cap();
        for (uint32_t m = rootbits; m; m &= m - 1) {
            const int k = __ffs(m) - 1;
            const uint32_t id = R[lane + 32 * k] & 1023u;
            const uint32_t gid = (uint32_t)((y0 + (int)(id >> 5)) * g.wp + x0 + (int)(id & 31u));
            if (fits) {
                const size_t hnd = fo + handle0 + o;
                rt.sizes[hnd] = (uint32_t)L[cc_slot(id)] - id;
                rt.links[hnd] = handle0 + o;
                rt.rootpix[hnd] = gid;
                rt.minpix[hnd] = gid;
            }
            L[cc_slot(id)] = (uint16_t)o;
            o++;
        }
    }
    __syncwarp();
    uint16_t* tl = l16 + tile * 1024;
    for (int j = lane, k = 0; j < nrun; j += 32, k++) {
        const uint32_t id = R[j] & 1023u;
        const uint32_t root = ((rootbits >> k) & 1u) ? id : (uint32_t)L[cc_slot(id)];
        tl[id] = L[cc_slot(root)];   // run-start label: the ordinal of the run's tile-local root (sparse, 2 bytes per run)
    }
}

// Boundary merge: one warp per tile.  All contact tests are bit operations on the row masks of the tile and of its
// left / right / upper neighbours; only real, non-implied contacts reach the lock-free union in global memory, and
// the unions run between TILE-LOCAL ROOTS (found through the run-start labels), never between pixels.
// As in the local pass only initiator pixels (1 <= x <= wd-2) issue links: left and up for both colours, up-left
// and up-right for white.  Top row: lane = column; left / right column: lane = row.
__device__ __forceinline__ uint2 cc_ld_mask(const uint2* __restrict__ fm, const Geom& g, int tx, int ty, int r) {
    if (tx < 0 || ty < 0 || tx >= cc_tiles_x(g) || ty >= cc_tiles_y(g)) return make_uint2(0u, 0u);
    return __ldg(&fm[((size_t)ty * cc_tiles_x(g) + tx) * 32 + r]);
}

// union of the sets of a and b whose parents pa = L[a], pb = L[b] were already loaded: both chains are walked together
// (the two loads of a step in flight at once), larger root hooked under the smaller
__device__ __forceinline__ void gunion_p(uint32_t* L, uint32_t a, uint32_t pa, uint32_t b, uint32_t pb) {
    for (;;) {
        while (pa != a || pb != b) {
            a = pa; b = pb;
            pa = __ldcg(&L[a]); pb = __ldcg(&L[b]);
        }
        if (a == b) return;
        if (a < b) { const uint32_t t = a; a = b; b = t; }
        const uint32_t old = atomicMin(&L[a], b);   // hook the larger root under the smaller
        if (old == a) return;
        a = old;                                    // a was hooked elsewhere meanwhile: keep uniting its new parent with b
        pa = __ldcg(&L[a]); pb = __ldcg(&L[b]);
    }
}

// A union request of the boundary merge: run-start offsets (row * 32 + first column of the run) of the two pixels inside
// their tiles, and which neighbour tile the second one lies in.
#define CCB_REL_LEFT 0u
#define CCB_REL_RIGHT 1u
#define CCB_REL_UP 2u
#define CCB_REL_UPLEFT 3u
#define CCB_REL_UPRIGHT 4u
#define CCB_QUEUE 192   // 32 lanes x (3 top-row + 2 left-column + 1 right-column requests)

// TWO PHASES.  Collect: the contact tests (bit operations on row masks in registers) push one 32-bit request per real,
// non-implied contact into the warp's queue.  Execute: the requests are dealt one per lane, so all of a tile's unions run
// side by side and every dependent step -- run-start labels + tile bases, parents, hook -- is ONE memory round trip for
// the whole tile (the contacts used to be resolved where they were found, three divergent sections with five dependent
// round trips each).
template <int CCB_WARPS>
__global__ void __launch_bounds__(CCB_WARPS * 32)
k_cc_boundary(const uint2* __restrict__ masks, const uint16_t* __restrict__ l16, CcRoots rt, Geom g) {
    __shared__ uint32_t sreq[CCB_WARPS][CCB_QUEUE];
    __shared__ int scount[CCB_WARPS];
    const int frame = blockIdx.z;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int tiles_x = cc_tiles_x(g), tiles_y = cc_tiles_y(g);
    const int t = blockIdx.x * CCB_WARPS + w;
    if (t >= tiles_x * tiles_y) return;   // (no block-level synchronisation anywhere below)
    const int ty = t / tiles_x, tx = t - ty * tiles_x;
    const int x0 = tx * CC_TW, y0 = ty * CC_TH;
    const uint2* fm = masks + (size_t)frame * tiles_x * tiles_y * 32;
    const uint2 M = fm[(size_t)t * 32 + lane];
    if (!__any_sync(FULL_MASK, (M.x | M.y) != 0u)) return;
    const uint2 ML = cc_ld_mask(fm, g, tx - 1, ty, lane), MR = cc_ld_mask(fm, g, tx + 1, ty, lane);
    const uint2 U = cc_ld_mask(fm, g, tx, ty - 1, 31), UL = cc_ld_mask(fm, g, tx - 1, ty - 1, 31),
                UR = cc_ld_mask(fm, g, tx + 1, ty - 1, 31);
    const uint32_t I = cc_initiators(x0, g.wd), IL = cc_initiators(x0 - 32, g.wd), IR = cc_initiators(x0 + 32, g.wd);
    if (lane == 0) scount[w] = 0;
    __syncwarp();
    uint32_t* Q = sreq[w];
    auto push = [&](uint32_t me_off, uint32_t rel, uint32_t other_off) {
        Q[atomicAdd(&scount[w], 1)] = me_off | (other_off << 10) | (rel << 20);
    };

    // ---- top row of the tile (y = y0): lane = column
    {
        const uint2 R0 = make_uint2(__shfl_sync(FULL_MASK, M.x, 0), __shfl_sync(FULL_MASK, M.y, 0));
        const uint2 L0 = make_uint2(__shfl_sync(FULL_MASK, ML.x, 0), __shfl_sync(FULL_MASK, ML.y, 0));
        const int c = lane, x = x0 + c;
        if ((I >> c) & 1u) {
            const int col = (R0.y >> c) & 1u;   // 0: white, 1: black
            const uint32_t P0 = col ? R0.y : R0.x;
            if ((P0 >> c) & 1u) {
                const uint32_t PL0 = col ? L0.y : L0.x, PU = col ? U.y : U.x, PUL = col ? UL.y : UL.x, PUR = col ? UR.y : UR.x;
                const bool row_l = c > 0 ? ((P0 >> (c - 1)) & 1u) : (PL0 >> 31);
                const bool up_c = (PU >> c) & 1u;
                const bool up_l = c > 0 ? ((PU >> (c - 1)) & 1u) : (PUL >> 31);
                const bool up_r = c < 31 ? ((PU >> (c + 1)) & 1u) : (PUR & 1u);
                const bool do_left = c == 0 && row_l;
                const bool do_up = up_c && !(x - 1 >= 1 && row_l && up_l);
                const bool do_ul = col == 0 && up_l && !up_c;
                const bool do_ur = col == 0 && up_r && !up_c;   // (upstream: no up-right link when up is white)
                if (do_left || do_up || do_ul || do_ur) {
                    const uint32_t me = (uint32_t)cc_run_start(P0, I, c);
                    if (do_left) push(me, CCB_REL_LEFT, (uint32_t)cc_run_start(PL0, IL, 31));
                    if (do_up) push(me, CCB_REL_UP, 31u * 32u + (uint32_t)cc_run_start(PU, I, c));
                    if (do_ul) {
                        if (c > 0) push(me, CCB_REL_UP, 31u * 32u + (uint32_t)cc_run_start(PU, I, c - 1));
                        else push(me, CCB_REL_UPLEFT, 31u * 32u + (uint32_t)cc_run_start(PUL, IL, 31));
                    }
                    if (do_ur) {
                        if (c < 31) push(me, CCB_REL_UP, 31u * 32u + (uint32_t)cc_run_start(PU, I, c + 1));
                        else push(me, CCB_REL_UPRIGHT, 31u * 32u + (uint32_t)cc_run_start(PUR, IR, 0));
                    }
                }
            }
        }
    }
    // ---- left and right column (rows 1..31; the corners belong to the top row): lane = row
    {
        const uint2 Mu = make_uint2(__shfl_up_sync(FULL_MASK, M.x, 1), __shfl_up_sync(FULL_MASK, M.y, 1));
        const uint2 MLu = make_uint2(__shfl_up_sync(FULL_MASK, ML.x, 1), __shfl_up_sync(FULL_MASK, ML.y, 1));
        const uint32_t MRu_w = __shfl_up_sync(FULL_MASK, MR.x, 1);
        if (lane != 0 && y0 + lane < g.hd) {
            const uint32_t r32 = (uint32_t)lane * 32u;
            if (I & 1u) {   // x = x0 is an initiator (x0 >= 1: there is a left tile)
                const int col = M.y & 1u;
                const uint32_t P = col ? M.y : M.x;
                if (P & 1u) {
                    const uint32_t PL = col ? ML.y : ML.x, Pu = col ? Mu.y : Mu.x, PLu = col ? MLu.y : MLu.x;
                    const bool left = PL >> 31, up = Pu & 1u, upleft = PLu >> 31;
                    const bool do_left = left && !(up && upleft);
                    const bool do_ul = col == 0 && upleft && !up;
                    if (do_left) push(r32, CCB_REL_LEFT, r32 + (uint32_t)cc_run_start(PL, IL, 31));
                    if (do_ul) push(r32, CCB_REL_LEFT, r32 - 32u + (uint32_t)cc_run_start(PLu, IL, 31));
                }
            }
            if ((I >> 31) & 1u) {   // x = x0 + 31 is an initiator: white up-right contact into the right tile
                if ((M.x >> 31) && (MRu_w & 1u) && !(Mu.x >> 31))
                    push(r32 + (uint32_t)cc_run_start(M.x, I, 31), CCB_REL_RIGHT, r32 - 32u);   // (column 0 starts its run)
            }
        }
    }
    __syncwarp();
    const int total = scount[w];
    if (total == 0) return;
    // ---- execute
    const uint16_t* f16 = l16 + (size_t)frame * tiles_x * tiles_y * 1024;
    uint32_t* fl = rt.links + (size_t)frame * rt.cap();
    const uint32_t* tb = rt.tile_base + (size_t)frame * rt.ntiles;
    const uint32_t base_me = __ldg(&tb[t]);
    for (int q = lane; q < total; q += 32) {
        const uint32_t rq = Q[q];
        const uint32_t rel = rq >> 20;
        const int to = t + (rel == CCB_REL_LEFT ? -1 : (rel == CCB_REL_RIGHT ? 1 : (rel == CCB_REL_UP ? -tiles_x :
                           (rel == CCB_REL_UPLEFT ? -tiles_x - 1 : -tiles_x + 1))));
        const uint32_t base_o = __ldg(&tb[to]);
        const uint32_t ord_me = f16[(size_t)t * 1024 + (rq & 1023u)];
        const uint32_t ord_o = f16[(size_t)to * 1024 + ((rq >> 10) & 1023u)];
        // (a tile without handles -- root list overflow, the chunk is re-run -- takes no part in any union)
        if (base_me == CC_NO_HANDLE || base_o == CC_NO_HANDLE) continue;
        const uint32_t a = base_me + ord_me, b = base_o + ord_o;
        const uint32_t pa = __ldcg(&fl[a]), pb = __ldcg(&fl[b]);
        gunion_p(fl, a, pa, b, pb);
    }
}

// Fold the pixel counts (and the smallest root pixel id) of the tile-local roots into the final roots (after the boundary
// merges) and point every tile-local root straight at its final root: afterwards the final root of ANY pixel is
// links[handle], one load.
__global__ void __launch_bounds__(256)
k_cc_sizes(CcRoots rt) {
    const int frame = blockIdx.y / CC_SUBLISTS, sub = blockIdx.y % CC_SUBLISTS;
    const int n = min(rt.nroots[blockIdx.y], rt.sub_cap);
    const size_t fo = (size_t)frame * rt.cap();
    uint32_t* fl = rt.links + fo;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t a = (uint32_t)(sub * rt.sub_cap + i);
        const uint32_t r = gfind(fl, a);
        if (r != a) {
            atomicAdd(&rt.sizes[fo + r], rt.sizes[fo + a]);
            atomicMin(&rt.minpix[fo + r], rt.rootpix[fo + a]);
            fl[a] = r;   // (monotone: still an ancestor for any concurrent walk)
        }
    }
}

// Dense ids for the components that can carry an edge point (final roots of >= 25 pixels): the edge-cluster key
// becomes a pair of 16-bit ids.  dense[] is defined at every final root (0xffffffff: component smaller than 25
// pixels, upstream's edge-point filter); dense2rep maps back to the component's canonical pixel id.
#define AGPU_MAX_DENSE 65535   // ids 0 .. 65534 (0xffff marks "no id" in 16-bit tables)
__global__ void __launch_bounds__(256)
k_cc_dense(CcRoots rt, uint32_t* __restrict__ dense2rep, int* __restrict__ ndense) {
    const int frame = blockIdx.y / CC_SUBLISTS, sub = blockIdx.y % CC_SUBLISTS;
    const int n = min(rt.nroots[blockIdx.y], rt.sub_cap);
    const size_t fo = (size_t)frame * rt.cap();
    uint32_t* f2 = dense2rep + (size_t)frame * AGPU_MAX_DENSE;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const size_t a = fo + (size_t)(sub * rt.sub_cap + i);
        if (__ldcg(&rt.links[a]) != (uint32_t)(sub * rt.sub_cap + i)) continue;        // not a final root
        if (__ldcg(&rt.sizes[a]) < 25u) { rt.dense[a] = 0xffffffffu; continue; }          // too small to carry edge points
        const int d = atomicAdd(&ndense[frame], 1);
        if (d < AGPU_MAX_DENSE) {
            rt.dense[a] = (uint32_t)d;
            f2[d] = __ldcg(&rt.minpix[a]);
        } else {
            rt.dense[a] = 0xffffffffu;
        }
    }
}

// Canonical per-pixel labels (smallest pixel id of the component; 127-pixels are singletons) and, at every
// representative pixel, the component's size -- for the stage dumps.  The pipeline itself never materialises a label
// image: k_edges resolves components on the fly.
__global__ void __launch_bounds__(256)
k_cc_canonical(const uint2* __restrict__ masks, const uint16_t* __restrict__ l16, CcRoots rt, uint32_t* __restrict__ out,
               uint32_t* __restrict__ out_sizes, Geom g) {
    const int frame = blockIdx.z;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= g.wd || y >= g.hd) return;
    const int tiles_x = cc_tiles_x(g), tiles_y = cc_tiles_y(g);
    const int tx = x >> 5, ty = y >> 5, c = x & 31, r = y & 31;
    const size_t t = (size_t)ty * tiles_x + tx;
    const uint2 M = masks[((size_t)frame * tiles_x * tiles_y + t) * 32 + r];
    const uint32_t id = (uint32_t)(y * g.wp + x);
    uint32_t lab = id, size = 0u;
    const uint32_t P = ((M.x >> c) & 1u) ? M.x : (((M.y >> c) & 1u) ? M.y : 0u);
    if (P) {
        const uint32_t hnd = cc_pixel_root(l16 + (size_t)frame * tiles_x * tiles_y * 1024, rt.tile_base + (size_t)frame * rt.ntiles,
                                           t, r, c, P, cc_initiators(tx * 32, g.wd));
        if (hnd != CC_NO_HANDLE) {
            const size_t fo = (size_t)frame * rt.cap();
            const uint32_t root = gfind(rt.links + fo, hnd);
            lab = rt.minpix[fo + root];
            if (lab == id) size = rt.sizes[fo + root];
        }
    }
    out[(size_t)frame * g.plane + id] = lab;
    out_sizes[(size_t)frame * g.plane + id] = size;
}
