// k_cluster.cuh -- gradient-edge clusters (upstream stage U5; part of the native call
// at /root/reference/src/detection/tag_detector.py:26).
//
//   k_edges          one warp per 32x32 tile on the bit masks: emits up to four edge points per pixel as ONE
//                    64-bit record (pair key << 32) | packed point, the key being the unordered pair of the
//                    two components' dense ids (k_cc_dense); records go to the frame's own segment of the
//                    point list (warp-aggregated atomics).
//                    The pair key is renamed to a per-frame CLUSTER ID on the fly (lock-free pair table) and the
//                    record carries that id; the number of records per id is counted as they are emitted.
//   k_cluster_refs   exclusive prefix of the per-id counts = start of every cluster in the sorted order -> cluster
//                    work lists by size tier (no pass over the records).
//   k_sort_scatter   hand-written SEGMENTED one-sweep radix (counting) sort whose single digit is the cluster id:
//                    every frame's segment is sorted independently, grid = (blocks, frames); a warp claims the
//                    slots of all its records of one cluster with one atomic.
#pragma once
#include "common.cuh"
#include "k_cc.cuh"

// Cluster ids.  A cluster is the set of edge points between one PAIR of components; its sort key used to be the pair of
// dense component ids (2 x 11 or 2 x 16 bits: two or three radix passes).  The pairs that actually occur are few --
// a few hundred per frame, a few thousand under sensor noise -- so k_edges renames them on the fly: a per-frame
// open-addressing table maps the 32-bit pair key to a CLUSTER ID handed out in order of first appearance, and the record
// carries that id.  The sort then has ONE digit -- the cluster id itself: the per-id record counts (taken while the
// records are emitted) are the digit histogram AND the table of cluster sizes, their exclusive prefix gives the digit
// offsets AND the cluster starts (k_cluster_refs), and one scatter sweep puts every record in place (k_sort_scatter).
// The order of the records inside a cluster is left to the atomics: the quad fit sorts them by a total order anyway.
struct PairTable {
    unsigned long long* slots;   // [nframes][nslots] (pair key << 32) | cluster id; all ones = empty
    uint32_t* keys;              // [nframes][cap_keys] cluster id -> pair key (max dense id << 16 | min dense id)
    int* ncl;                    // [nframes] cluster ids handed out (may exceed cap_keys: the host re-runs the chunk)
    uint32_t* count;             // [nframes][cap_keys] records per cluster id (k_edges); then the fill cursor of the scatter
    uint32_t* start;             // [nframes][cap_keys] first slot of the cluster in the sorted segment (k_cluster_refs)
    int nslots;                  // power of two
    int cap_keys;
};
#define PT_EMPTY 0xffffffffffffffffull

__device__ __forceinline__ uint32_t pair_hash(uint32_t k) {
    k ^= k >> 15; k *= 0x2c1b3c6du; k ^= k >> 12; k *= 0x297a2d39u; k ^= k >> 15;
    return k;
}

// cluster id of pair key `key` in frame `frame` (inserted when new).  Lock-free: the 64-bit slot holds key and id, so a
// reader never sees a key without its id; an id drawn for an insertion that loses its race is simply never used (an
// empty cluster).  Returns 0xffffffff when the table is full (the host sees ncl > cap and re-runs the chunk).
__device__ __forceinline__ uint32_t pair_cluster_id(const PairTable& pt, int frame, uint32_t key) {
    unsigned long long* T = pt.slots + (size_t)frame * pt.nslots;
    const uint32_t mask = (uint32_t)pt.nslots - 1u;
    uint32_t slot = pair_hash(key) & mask;
    uint32_t fresh = 0xffffffffu;
    for (int probes = 0; probes < pt.nslots; probes++) {
        unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(T + slot);
        if (e == PT_EMPTY) {
            if (fresh == 0xffffffffu) {
                fresh = (uint32_t)atomicAdd(&pt.ncl[frame], 1);
                if (fresh >= (uint32_t)pt.cap_keys) return 0xffffffffu;
            }
            e = atomicCAS(T + slot, PT_EMPTY, ((unsigned long long)key << 32) | fresh);
            if (e == PT_EMPTY) {
                pt.keys[(size_t)frame * pt.cap_keys + fresh] = key;
                return fresh;
            }
        }
        if ((uint32_t)(e >> 32) == key) return (uint32_t)e;
        slot = (slot + 1) & mask;
    }
    return 0xffffffffu;
}

// One warp per 32x32 tile, lane = row, everything on the tile-major bit masks of k_cc_local (2 bits per pixel).
// An edge between v0 and v1 (one white, one black) in direction d is one AND of the row's white mask with the
// neighbour row's shifted black mask (and vice versa): a lane tests its 32 pixels x 4 directions with ~20 bit
// operations, and a tile without any edge leaves after three coalesced 256-byte loads.  The (pixel, direction)
// candidates of a warp are compacted through shared memory and then handled ONE PER LANE (run-start label +
// representative + dense-id look-ups, ballot compaction, one atomic per 32 candidates), so the expensive part is
// not serialised inside the few rows that sit on an edge.
#define EDGE_CAND_PER_PASS 1024   // a tile with more candidates is handled in four passes of 8 rows (8 x 32 x 4)
#ifndef EDGE_MINB
#define EDGE_MINB 24   // 40 registers, 48 warps per SM: the kernel waits on dependent global look-ups (46 registers / 40 warps
#endif                 // unbounded: +6 % time; 32 registers spill and gain nothing)
template <int EDGE_WARPS>
__global__ void __launch_bounds__(EDGE_WARPS * 32, EDGE_WARPS == 2 ? EDGE_MINB : 1)
k_edges(const uint2* __restrict__ masks, const uint16_t* __restrict__ l16, CcRoots rt, Geom g,
        unsigned long long* __restrict__ recs, int* __restrict__ npts, int* __restrict__ ndups, int cap, PairTable pt) {
    __shared__ uint16_t scand[EDGE_WARPS][EDGE_CAND_PER_PASS];
    __shared__ uint16_t sdense[EDGE_WARPS][1024];   // dense component id of every run of the tile, at its start pixel
    __shared__ unsigned long long scache[EDGE_WARPS][8];   // the warp's last pair keys and their cluster ids (a tile sees a handful)
    __shared__ uint16_t shalo[EDGE_WARPS][100];            // dense ids of the neighbour-tile pixels the tile's candidates point at
    // grid = (frames, x blocks, tile rows): consecutive CTAs belong to DIFFERENT frames, so the per-frame append
    // counters are not hammered by every resident warp at once
    const int frame = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int tiles_x = cc_tiles_x(g), tiles_y = cc_tiles_y(g);
    const int tx = blockIdx.y * EDGE_WARPS + w, ty = blockIdx.z;
    if (tx >= tiles_x) return;   // (no block-level synchronisation anywhere below)
    const size_t t = (size_t)ty * tiles_x + tx;
    const uint2* fm = masks + (size_t)frame * tiles_x * tiles_y * 32;
    const uint16_t* f16 = l16 + (size_t)frame * tiles_x * tiles_y * 1024;
    const uint32_t* fl = rt.links + (size_t)frame * rt.cap();      // tile-local root -> final root (flattened by k_cc_sizes)
    const uint32_t* fd = rt.dense + (size_t)frame * rt.cap();      // final root -> dense component id
    const uint32_t* tb = rt.tile_base + (size_t)frame * rt.ntiles;
    const uint2 M = fm[t * 32 + lane];
    if (!__any_sync(FULL_MASK, (M.x | M.y) != 0u)) return;
    const int x0 = tx * 32, y0 = ty * 32;
    const uint2 ML = cc_ld_mask(fm, g, tx - 1, ty, lane), MR = cc_ld_mask(fm, g, tx + 1, ty, lane);
    uint2 N = make_uint2(__shfl_down_sync(FULL_MASK, M.x, 1), __shfl_down_sync(FULL_MASK, M.y, 1));
    uint2 NL = make_uint2(__shfl_down_sync(FULL_MASK, ML.x, 1), __shfl_down_sync(FULL_MASK, ML.y, 1));
    uint2 NR = make_uint2(__shfl_down_sync(FULL_MASK, MR.x, 1), __shfl_down_sync(FULL_MASK, MR.y, 1));
    if (lane == 31) {
        N = cc_ld_mask(fm, g, tx, ty + 1, 0);
        NL = cc_ld_mask(fm, g, tx - 1, ty + 1, 0);
        NR = cc_ld_mask(fm, g, tx + 1, ty + 1, 0);
    }
    const uint32_t Ix = cc_initiators(x0, g.wd);
    uint32_t m[4] = {0u, 0u, 0u, 0u};
    uint32_t dup2 = 0u;   // direction-2 candidates that absorb the identical point of direction 3 at x-1
    if (y0 + lane <= g.hd - 2) {
        const uint32_t W = M.x, B = M.y;
        const uint32_t rW = (W >> 1) | (MR.x << 31), rB = (B >> 1) | (MR.y << 31);           // (x+1, y)
        const uint32_t dlW = (N.x << 1) | (NL.x >> 31), dlB = (N.y << 1) | (NL.y >> 31);     // (x-1, y+1)
        const uint32_t drW = (N.x >> 1) | (NR.x << 31), drB = (N.y >> 1) | (NR.y << 31);     // (x+1, y+1)
        m[0] = ((W & rB) | (B & rW)) & Ix;
        m[1] = ((W & N.y) | (B & N.x)) & Ix;                                                 // (x,   y+1)
        m[2] = ((W & dlB) | (B & dlW)) & Ix;
        m[3] = ((W & drB) | (B & drW)) & Ix;
        // Duplicates.  (x, y, dir 2) and (x-1, y, dir 3) are the two diagonals of one 2x2 block and give the same
        // point (2x-1, 2y+1).  When both exist the four pixels are two vertical or two horizontal same-coloured
        // pairs, each pair linked by the component rules (all of x-1, x are initiators), so both candidates carry
        // the SAME pair of components: the duplicate upstream removes after its slope sort.  Keep dir 2, flag it.
        const uint32_t c3_left = ((ML.x >> 31) & N.y & 1u) | ((ML.y >> 31) & N.x & 1u);                 // dir 3 at (x0-1, y)
        const uint32_t c2_right = x0 + 32 <= g.wd - 2 ? ((MR.x & (N.y >> 31)) | (MR.y & (N.x >> 31))) & 1u : 0u;   // dir 2 at (x0+32, y)
        dup2 = m[2] & ((m[3] << 1) | c3_left);
        m[3] &= ~((m[2] >> 1) | (c2_right << 31));
    }
    const int cnt_all = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
    int total_all = cnt_all;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) total_all += __shfl_xor_sync(FULL_MASK, total_all, off);
    if (total_all == 0) return;

    // ---- halo: a candidate on the tile's right / left / bottom border has its second pixel in a neighbour tile (a tenth
    //      of the candidates).  Their dense ids -- mask -> run start -> root -> final root -> dense id, four dependent
    //      loads -- are fetched HERE for all of them at once (lane = row for the right and left neighbour's border column,
    //      lane = column for the lower neighbour's first row, lanes 0 / 31 for the two corner pixels) and parked in a
    //      98-entry table, so the candidate batches below never wait on global memory for a pixel.
    //      Table: [0, 32) right tile (column 0, row i), [32, 64) left tile (column 31, row i), [64, 96) lower tile (row 0,
    //      column i), 96 lower-left pixel (31, 0), 97 lower-right pixel (0, 0).
    uint32_t hh[4];          // handles (CC_NO_HANDLE: not needed)
    {
        const uint32_t m3u = __shfl_up_sync(FULL_MASK, m[3], 1), m2u = __shfl_up_sync(FULL_MASK, m[2], 1);
        const uint32_t b1 = __shfl_sync(FULL_MASK, m[1], 31), b2 = __shfl_sync(FULL_MASK, m[2], 31), b3 = __shfl_sync(FULL_MASK, m[3], 31);
        const uint2 Nb = make_uint2(__shfl_sync(FULL_MASK, N.x, 31), __shfl_sync(FULL_MASK, N.y, 31));
        const uint2 NLb = make_uint2(__shfl_sync(FULL_MASK, NL.x, 31), __shfl_sync(FULL_MASK, NL.y, 31));
        const uint2 NRb = make_uint2(__shfl_sync(FULL_MASK, NR.x, 31), __shfl_sync(FULL_MASK, NR.y, 31));
        const uint32_t IL = cc_initiators(x0 - 32, g.wd);
        const bool needR = ((m[0] >> 31) | (lane > 0 ? m3u >> 31 : 0u)) & 1u;
        const bool needL = lane > 0 && (m2u & 1u);
        const bool needB = ((b1 | (b2 >> 1) | (b3 << 1)) >> lane) & 1u;
        const bool needC = lane == 0 ? (b2 & 1u) : (lane == 31 ? (b3 >> 31) : 0u);
        const size_t tR = t + 1, tL = t - 1, tB = t + tiles_x;
        const size_t tC = lane == 0 ? tB - 1 : tB + 1;
        const uint32_t PL = (ML.x >> 31) ? ML.x : ML.y, PB = ((Nb.x >> lane) & 1u) ? Nb.x : Nb.y;
        const uint32_t PC = lane == 0 ? ((NLb.x >> 31) ? NLb.x : NLb.y) : NRb.x /* unused for column 0 */;
        uint32_t base[4], ord[4];
        base[0] = needR ? __ldg(&tb[tR]) : CC_NO_HANDLE;
        base[1] = needL ? __ldg(&tb[tL]) : CC_NO_HANDLE;
        base[2] = needB ? __ldg(&tb[tB]) : CC_NO_HANDLE;
        base[3] = needC ? __ldg(&tb[tC]) : CC_NO_HANDLE;
        ord[0] = needR ? (uint32_t)f16[tR * 1024 + lane * 32] : 0u;                                   // (column 0 starts its run)
        ord[1] = needL ? (uint32_t)f16[tL * 1024 + lane * 32 + cc_run_start(PL, IL, 31)] : 0u;
        ord[2] = needB ? (uint32_t)f16[tB * 1024 + cc_run_start(PB, Ix, lane)] : 0u;
        ord[3] = needC ? (uint32_t)f16[tC * 1024 + (lane == 0 ? cc_run_start(PC, IL, 31) : 0)] : 0u;
#pragma unroll
        for (int u = 0; u < 4; u++) hh[u] = base[u] == CC_NO_HANDLE ? CC_NO_HANDLE : base[u] + ord[u];
#pragma unroll
        for (int u = 0; u < 4; u++) hh[u] = hh[u] == CC_NO_HANDLE ? CC_NO_HANDLE : fl[hh[u]];   // -> final roots (loads in flight together)
    }

    // ---- dense id of every run of this tile: run start -> ordinal of its tile-local root -> final root -> dense id, in
    //      three sweeps with ONE dependent memory round trip each: (1) every lane fetches the ordinals of its row's runs,
    //      four loads in flight, and parks them in the table; (2) the tile's roots -- a handful -- are resolved side by
    //      side, one per lane, into a little table that borrows the candidate buffer; (3) ordinals -> dense ids in place.
    {
        const uint32_t Sw = M.x & ~(M.x & (M.x << 1) & Ix), Sb = M.y & ~(M.y & (M.y << 1) & Ix);
        const uint32_t Sall = Sw | Sb;
        const uint16_t* tl = f16 + t * 1024 + lane * 32;
        const uint32_t tbase = __ldg(&tb[t]);       // handle of the tile's first root (CC_NO_HANDLE: root lists overflowed)
        uint16_t* row = &sdense[w][lane * 32];
        uint32_t S = Sall;
        int maxo = -1;
        while (__any_sync(FULL_MASK, S != 0u)) {
            int c[4];
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                c[u] = S ? __ffs(S) - 1 : -1;
                S &= S - 1;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = c[u] >= 0 ? (uint32_t)tl[c[u]] : 0u;
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (c[u] >= 0) { row[c[u]] = (uint16_t)v[u]; maxo = max(maxo, (int)v[u]); }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) maxo = max(maxo, __shfl_xor_sync(FULL_MASK, maxo, off));
        uint16_t* rootd = scand[w];                 // (free until the candidates are listed; a tile has at most 512 roots)
        for (int o = lane; o <= maxo; o += 32)
            rootd[o] = (uint16_t)(tbase != CC_NO_HANDLE ? min(fd[fl[tbase + o]], 0xffffu) : 0xffffu);
        __syncwarp();
        for (S = Sall; S; S &= S - 1) {
            const int cc = __ffs(S) - 1;
            row[cc] = rootd[row[cc]];
        }
    }
    {
        uint32_t hd[4];
#pragma unroll
        for (int u = 0; u < 4; u++) hd[u] = hh[u] == CC_NO_HANDLE ? 0xffffu : min(fd[hh[u]], 0xffffu);
        shalo[w][lane] = (uint16_t)hd[0];
        shalo[w][32 + lane] = (uint16_t)hd[1];
        shalo[w][64 + lane] = (uint16_t)hd[2];
        if (lane == 0) shalo[w][96] = (uint16_t)hd[3];
        if (lane == 31) shalo[w][97] = (uint16_t)hd[3];
    }
    __syncwarp();

    unsigned long long* fk = recs + (size_t)frame * cap;
    if (lane < 8) scache[w][lane] = PT_EMPTY;
    __syncwarp();
    // The slot reservation of a batch (an atomic with a return value) is not waited for: the batch's records stay in
    // registers and are stored when the NEXT batch has done its look-ups, by which time the reply is there.
    bool pend = false, pend_ok = false;
    int pend_base = 0;
    uint32_t pend_okm = 0u;
    unsigned long long pend_rec = 0ull;
    auto flush = [&]() {
        if (!pend) return;
        const int base = __shfl_sync(FULL_MASK, pend_base, 0);
        const int p = base + __popc(pend_okm & ((1u << lane) - 1u));
        if (pend_ok && p < cap) fk[p] = pend_rec;
        pend = false;
    };
    const int npass = total_all <= EDGE_CAND_PER_PASS ? 1 : 4;
    const int rsh = npass == 1 ? 5 : 3;   // rows per pass = 1 << rsh
#pragma unroll 1
    for (int pass = 0; pass < npass; pass++) {
        const int cnt = (lane >> rsh) == pass ? cnt_all : 0;
        int incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int n = __shfl_up_sync(FULL_MASK, incl, off);
            if (lane >= off) incl += n;
        }
        const int total = __shfl_sync(FULL_MASK, incl, 31);
        if (total == 0) continue;
        if (cnt) {
            int o = incl - cnt;
#pragma unroll
            for (int d = 0; d < 4; d++)
                for (uint32_t mm = m[d]; mm; mm &= mm - 1) {
                    const int c = __ffs(mm) - 1;
                    const uint32_t pos = (M.y >> c) & 1u;   // v1 > v0  <=>  v0 == 0 (black)
                    uint32_t e = lane | (c << 5) | (d << 10) | (pos << 12);
                    if (d == 2 && ((dup2 >> c) & 1u))       // merged: + polarity of the dir-3 point, whose v0 is (x-1, y)
                        e |= (1u << 13) | ((c > 0 ? (M.y >> (c - 1)) & 1u : ML.y >> 31) << 14);
                    scand[w][o++] = (uint16_t)e;
                }
        }
        __syncwarp();
        for (int b = 0; b < total; b += 32) {
            const bool have = b + lane < total;
            const uint32_t cd = have ? scand[w][b + lane] : 0u;
            const int r = cd & 31, c = (cd >> 5) & 31, d = (cd >> 10) & 3, pos = (cd >> 12) & 1;
            const int dx = (d == 0 || d == 3) ? 1 : (d == 2 ? -1 : 0), dy = d == 0 ? 0 : 1;
            int qc = c + dx, qr = r + dy;
            const int qtx = tx + (qc < 0 ? -1 : (qc > 31 ? 1 : 0)), qty = ty + (qr > 31 ? 1 : 0);
            qc &= 31;
            qr &= 31;
            // row masks of p and q inside the tile: shuffles (every lane takes part)
            const uint32_t Wp = __shfl_sync(FULL_MASK, M.x, r), Bp = __shfl_sync(FULL_MASK, M.y, r);
            const uint32_t Wq = __shfl_sync(FULL_MASK, M.x, qr), Bq = __shfl_sync(FULL_MASK, M.y, qr);
            uint32_t d0 = 0xffffu, d1 = 0xffffu;
            if (have) {
                d0 = sdense[w][r * 32 + cc_run_start(pos ? Bp : Wp, Ix, c)];
                if (qtx == tx && qty == ty) {
                    d1 = sdense[w][qr * 32 + cc_run_start(pos ? Wq : Bq, Ix, qc)];
                } else {   // neighbour tile: the halo table
                    d1 = shalo[w][qty != ty ? (qtx == tx ? 64 + qc : (qtx < tx ? 96 : 97)) : (qtx > tx ? qr : 32 + qr)];
                }
            }
            bool ok = have && d0 != 0xffffu && d1 != 0xffffu;   // both components have >= 25 pixels
            uint32_t okm = __ballot_sync(FULL_MASK, ok);
            if (okm == 0) continue;
            // pair key -> cluster id: one look-up per DISTINCT key of the batch (usually one or two), through the warp's
            // little cache first
            const uint32_t key = ok ? ((max(d0, d1) << 16) | min(d0, d1)) : 0xffffffffu;
            const uint32_t peers = __match_any_sync(FULL_MASK, key);
            const int leader = __ffs(peers) - 1;
            uint32_t cid = 0xffffffffu;
            if (ok && lane == leader) {
                const uint32_t cs = pair_hash(key) >> 29;
                const unsigned long long ce = scache[w][cs];
                if ((uint32_t)(ce >> 32) == key) cid = (uint32_t)ce;
                else {
                    cid = pair_cluster_id(pt, frame, key);
                    if (cid != 0xffffffffu) scache[w][cs] = ((unsigned long long)key << 32) | cid;
                }
            }
            if (ok && lane == leader && cid != 0xffffffffu)      // the digit histogram of the sort, taken at the source
                atomicAdd(&pt.count[(size_t)frame * pt.cap_keys + cid], (uint32_t)__popc(peers));
            cid = __shfl_sync(FULL_MASK, cid, leader);
            ok = ok && cid != 0xffffffffu;                       // (table full: the host re-runs the chunk with a larger one)
            okm = __ballot_sync(FULL_MASK, ok);
            if (okm == 0) continue;
            unsigned long long rec = 0;
            if (ok) {
                const int merged = (cd >> 13) & 1;
                const int kind = merged ? (8 | pos | (((cd >> 14) & 1) << 1)) : (d | (pos << 2));
                rec = ((unsigned long long)cid << 32) | pack_point(2 * (x0 + c) + dx, 2 * (y0 + r) + dy, kind);
            }
            const uint32_t dupm = __ballot_sync(FULL_MASK, ok && ((cd >> 13) & 1u));
            // (a slot reservation per TILE instead of per batch -- sentinel records in the holes -- was measured: -3 % on
            // clean frames, +70 % under sensor noise, where most candidates touch a component of < 25 pixels)
            flush();
            if (lane == 0) {
                pend_base = atomicAdd(&npts[frame], __popc(okm));
                if (dupm) atomicAdd(&ndups[frame], __popc(dupm));   // (raw point count = npts + ndups)
            }
            pend = true; pend_ok = ok; pend_okm = okm; pend_rec = rec;
        }
        __syncwarp();
    }
    flush();
}

// ---- segmented one-sweep sort by cluster id ---------------------------------------------------
// Records are single 64-bit words, (cluster id << 32) | point.
#define RS_THREADS 256
#define RS_ITEMS 8
#define RS_TILE (RS_THREADS * RS_ITEMS)  // records per scatter block (and the granularity of the list capacity)

// cluster work lists, by size tier (the tier decides how much shared memory the fitting warp gets)
#define AGPU_NTIERS 5
// counter block shared by the host and the kernels (ints): clusters per tier, clusters over upstream's size limit, all
// heads (debug), quads, per-tier work cursors, records per tier
enum { CNT_TIER0 = 0, CNT_OVERSIZE = 5, CNT_HEADS = 6, CNT_NQUADS = 7, CNT_CURSOR0 = 8, CNT_TIER_RECS0 = 16, CNT_FIXED = 24 };
struct ClusterLists {
    ClusterRef* list[AGPU_NTIERS];
    int cap[AGPU_NTIERS];    // largest cluster size of the tier
    int* counters;           // the CNT_* block
    int cap_list;
    ClusterRef* dbg_heads;   // all run heads (debug only, may be null)
    int cap_dbg;
};

// Cluster sizes -> cluster starts and work lists.  One CTA per frame: exclusive prefix of the per-id record counts (in id
// order, 8 ids per thread and round), start[id] for the scatter, one ClusterRef per non-empty id into its size tier; the
// counts are zeroed on the way and serve as the scatter's fill cursors afterwards.
// (`cap`: when a frame emitted more records than its segment holds the counts exceed the segment; clusters that would
// reach past it are not listed and the scatter drops what does not fit -- the host re-runs such a chunk with a larger
// capacity anyway.)
__global__ void __launch_bounds__(256)
k_cluster_refs(PairTable pt, Geom g, int min_size, ClusterLists cl, int cap) {
    __shared__ int wsum[8];
    __shared__ int s_carry;
    const int frame = blockIdx.x;
    const int n = min(pt.ncl[frame], pt.cap_keys);
    const int max_cluster = 3 * (2 * g.wd + 2 * g.hd);
    uint32_t* cnt = pt.count + (size_t)frame * pt.cap_keys;
    uint32_t* st = pt.start + (size_t)frame * pt.cap_keys;
    constexpr int PER = 8;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 256 * PER) {
        int sz[PER], sum = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int c = base + threadIdx.x * PER + k;
            sz[k] = c < n ? (int)cnt[c] : 0;
            sum += sz[k];
        }
        int incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(FULL_MASK, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        int start = s_carry + incl - sum;
        for (int ww = 0; ww < w; ww++) start += wsum[ww];
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int c = base + threadIdx.x * PER + k;
            const int size = sz[k];
            if (c < n) { st[c] = (uint32_t)start; cnt[c] = 0u; }
            // tier of this cluster (-1: none; 4: over upstream's size limit).  The work lists are appended to by every
            // frame's CTA at once: ONE atomic per warp and tier claims the slots of all its clusters.
            int tier = -1;
            const bool live = size > 0 && start + size <= cap;
            if (live && size >= min_size) {
                // upstream drops clusters of more than 3(2w+2h) RAW points; a record stands for one or two raw points,
                // so a cluster with more RECORDS than that is certainly over the limit (the exact raw count of the
                // others is taken by the fitting group, which has to read the records anyway)
                if (size > max_cluster) tier = AGPU_NTIERS;
                else {
#pragma unroll
                    for (int t = AGPU_NTIERS - 1; t >= 0; t--)
                        if (size <= cl.cap[t]) tier = t;
                }
            }
            ClusterRef ref;
            ref.frame = frame; ref.start = start; ref.size = size; ref.pad = 0;
            if (cl.dbg_heads) {
                const uint32_t m = __ballot_sync(FULL_MASK, live);
                int b0 = 0;
                if (lane == 0 && m) b0 = atomicAdd(&cl.counters[CNT_HEADS], __popc(m));
                b0 = __shfl_sync(FULL_MASK, b0, 0) + __popc(m & ((1u << lane) - 1u));
                if (live && b0 < cl.cap_dbg) cl.dbg_heads[b0] = ref;
            }
#pragma unroll
            for (int t = 0; t <= AGPU_NTIERS; t++) {
                const uint32_t m = __ballot_sync(FULL_MASK, tier == t);
                if (m == 0) continue;
                if (t == AGPU_NTIERS) {
                    if (lane == 0) atomicAdd(&cl.counters[CNT_OVERSIZE], __popc(m));
                    continue;
                }
                int recs_t = tier == t ? size : 0;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) recs_t += __shfl_xor_sync(FULL_MASK, recs_t, off);
                int s0 = 0;
                if (lane == 0) {
                    s0 = atomicAdd(&cl.counters[CNT_TIER0 + t], __popc(m));
                    atomicAdd(&cl.counters[CNT_TIER_RECS0 + t], recs_t);   // records handed to this tier (instrumentation)
                }
                s0 = __shfl_sync(FULL_MASK, s0, 0) + __popc(m & ((1u << lane) - 1u));
                if (tier == t && s0 < cl.cap_list) cl.list[t][s0] = ref;
            }
            start += size;
        }
        __syncthreads();
        if (threadIdx.x == 255) s_carry = start;   // (the last thread's running start = end of this round)
        __syncthreads();
    }
}

// The scatter sweep: record i of the frame's emission order goes to start[id] + (a slot of cluster id claimed with the
// fill cursor).  The records of a warp mostly share their cluster (neighbouring image tiles), so one atomic per distinct id
// and warp round claims the slots of all of them.
__global__ void __launch_bounds__(RS_THREADS)
k_sort_scatter(const unsigned long long* __restrict__ recs_in, unsigned long long* __restrict__ recs_out,
               const int* __restrict__ npts, int cap, PairTable pt) {
    const int frame = blockIdx.y;
    const int n = min(npts[frame], cap);
    const int base = blockIdx.x * RS_TILE;
    if (base >= n) return;
    const int lane = threadIdx.x & 31;
    const size_t seg = (size_t)frame * cap;
    uint32_t* fill = pt.count + (size_t)frame * pt.cap_keys;
    const uint32_t* st = pt.start + (size_t)frame * pt.cap_keys;
    unsigned long long rec[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const int i = base + r * RS_THREADS + threadIdx.x;
        rec[r] = i < n ? __ldg(recs_in + seg + i) : 0xffffffffffffffffull;
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const uint32_t cid = (uint32_t)(rec[r] >> 32);
        const bool valid = cid != 0xffffffffu;
        const uint32_t peers = __match_any_sync(FULL_MASK, cid);
        const int leader = __ffs(peers) - 1;
        uint32_t slot = 0;
        if (valid && lane == leader) slot = atomicAdd(&fill[cid], (uint32_t)__popc(peers));
        slot = __shfl_sync(FULL_MASK, slot, leader);
        const uint32_t pos = valid ? __ldg(&st[cid]) + slot + __popc(peers & ((1u << lane) - 1u)) : 0u;
        if (valid && pos < (uint32_t)cap) recs_out[seg + pos] = rec[r];
    }
}
