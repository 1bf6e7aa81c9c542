// k_cluster.cuh -- gradient-edge clusters (upstream stage U5; part of the native call
// at /root/reference/src/detection/tag_detector.py:26).
//
//   k_edges          four pixels per thread: emits up to four edge points per pixel as ONE 64-bit record
//                    (pair key << 32) | packed point, the key being the unordered pair of the two
//                    components' dense 16-bit ids (k_cc_dense); records go to the frame's own segment
//                    of the point list (warp-aggregated atomics).
//   k_sort_hist / k_sort_scan / k_sort_scatter
//                    hand-written SEGMENTED least-significant-digit radix sort (8-bit digits, stable):
//                    every frame's segment is sorted independently, grid = (blocks, frames).
//   k_cluster_heads  run heads of equal keys -> cluster work lists (binary search for the run end).
#pragma once
#include "common.cuh"
#include "k_cc.cuh"

// Four pixels (one 32-bit word of the threshold image) per thread.  An edge between v0 and v1 means
// v0 ^ v1 == 0xff (values are 0 / 127 / 255), tested for all four pixels and one direction at a time with
// three word operations; the vast majority of warps see no edge and leave after five word loads.  The
// (pixel, direction) candidates of a warp are compacted through shared memory and then handled ONE PER LANE
// (label + size look-ups, ballot compaction, one atomic per 32 candidates), so the expensive part is not
// serialised inside the few threads that sit on an edge.
#define EDGE_WORDS 4                                  // 16 pixels (one 128-bit load) per thread
#define EDGE_CAND_PER_WARP (32 * EDGE_WORDS * 16)     // lanes x pixels x directions
__global__ void __launch_bounds__(256)
k_edges(const uint8_t* __restrict__ thresh, const uint32_t* __restrict__ labels, const uint32_t* __restrict__ sizes,
        const uint32_t* __restrict__ dense, Geom g, unsigned long long* __restrict__ recs, int* __restrict__ npts, int cap,
        int id_bits) {
    __shared__ uint16_t scand[8][EDGE_CAND_PER_WARP];
    // grid = (frames, x blocks, y blocks): consecutive CTAs belong to DIFFERENT frames, so the per-frame append
    // counters are not hammered by every resident warp at once
    const int frame = blockIdx.x;
    const uint8_t* ft = thresh + (size_t)frame * g.plane;
    const uint32_t* fl = labels + (size_t)frame * g.plane;
    const uint32_t* fs = sizes + (size_t)frame * g.plane;
    const uint32_t* fd = dense + (size_t)frame * g.plane;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int kx = (blockIdx.y * 32 + lane) * EDGE_WORDS;   // first word of this thread in the row
    const int y = blockIdx.z * 8 + w;
    const int wpr = g.wp >> 2;                              // words per row (a multiple of 4)
    const int x0 = kx * 4;

    uint32_t m[EDGE_WORDS][4];
    uint32_t cur[EDGE_WORDS];
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < EDGE_WORDS; j++) {
        cur[j] = 0;
#pragma unroll
        for (int d = 0; d < 4; d++) m[j][d] = 0;
    }
    if (y <= g.hd - 2 && kx < wpr && x0 <= g.wd - 2) {
        const uint32_t* r0 = reinterpret_cast<const uint32_t*>(ft + (size_t)y * g.wp);
        const uint32_t* r1 = r0 + wpr;
        const uint32_t none = 0x7f7f7f7fu;
        const uint4 c4 = *reinterpret_cast<const uint4*>(r0 + kx), d4 = *reinterpret_cast<const uint4*>(r1 + kx);
        const uint32_t c[6] = {0, c4.x, c4.y, c4.z, c4.w, kx + 4 < wpr ? r0[kx + 4] : none};
        const uint32_t e[6] = {kx > 0 ? r1[kx - 1] : none, d4.x, d4.y, d4.z, d4.w, kx + 4 < wpr ? r1[kx + 4] : none};
#pragma unroll
        for (int j = 0; j < EDGE_WORDS; j++) {
            cur[j] = c[j + 1];
            uint32_t nb[4];
            nb[0] = __funnelshift_r(c[j + 1], c[j + 2], 8);   // (x+1, y)
            nb[1] = e[j + 1];                                 // (x,   y+1)
            nb[2] = __funnelshift_l(e[j], e[j + 1], 8);       // (x-1, y+1)
            nb[3] = __funnelshift_r(e[j + 1], e[j + 2], 8);   // (x+1, y+1)
            uint32_t xm = 0;                                  // initiators: 1 <= x <= w-2
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int x = x0 + 4 * j + i;
                if (x >= 1 && x <= g.wd - 2) xm |= 1u << (8 * i);
            }
#pragma unroll
            for (int d = 0; d < 4; d++) {
                const uint32_t t = cur[j] ^ nb[d];
                m[j][d] = (t >> 7) & t & xm;
                cnt += __popc(m[j][d]);
            }
        }
    }
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int n = __shfl_up_sync(FULL_MASK, incl, off);
        if (lane >= off) incl += n;
    }
    const int total = __shfl_sync(FULL_MASK, incl, 31);
    if (total == 0) return;
    if (cnt) {
        int o = incl - cnt;
#pragma unroll
        for (int j = 0; j < EDGE_WORDS; j++) {
            if (!(m[j][0] | m[j][1] | m[j][2] | m[j][3])) continue;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t pos = ((cur[j] >> (8 * i)) & 0xff) == 0 ? 1u : 0u;   // v1 > v0  <=>  v0 == 0
#pragma unroll
                for (int d = 0; d < 4; d++)
                    if ((m[j][d] >> (8 * i)) & 1u)
                        scand[w][o++] = (uint16_t)((lane * 16 + j * 4 + i) | (d << 9) | (pos << 11));
            }
        }
    }
    __syncwarp();
    unsigned long long* fk = recs + (size_t)frame * cap;
    const int xbase = blockIdx.y * (32 * EDGE_WORDS * 4);
    for (int b = 0; b < total; b += 32) {
        bool ok = false;
        unsigned long long rec = 0;
        const bool have = b + lane < total;
        int x = 0, d = 0, pos = 0, dx = 0, dy = 0;
        uint32_t l0 = 0xffffffffu, l1 = 0xfffffffeu;
        if (have) {
            const uint32_t c = scand[w][b + lane];
            x = xbase + (int)(c & 511); d = (c >> 9) & 3; pos = (c >> 11) & 1;
            dx = (d == 0 || d == 3) ? 1 : (d == 2 ? -1 : 0); dy = d == 0 ? 0 : 1;
            const size_t id = (size_t)y * g.wp + x;
            // pixel -> tile-local root (k_cc_local)
            l0 = fl[id];
            l1 = fl[id + (size_t)dy * g.wp + dx];
        }
        // tile-local root -> final root -> dense id (k_cc_sizes / k_cc_dense).  The candidates of a warp sit on a
        // handful of components, so one lane per distinct tile-local root does the two dependent loads.
        uint32_t d0 = 0xffffffffu, d1 = 0xffffffffu;
        {
            const uint32_t p0 = __match_any_sync(FULL_MASK, l0), p1 = __match_any_sync(FULL_MASK, l1);
            const int ld0 = __ffs(p0) - 1, ld1 = __ffs(p1) - 1;
            if (have && lane == ld0) d0 = fd[fl[l0]];
            if (have && lane == ld1) d1 = fd[fl[l1]];
            d0 = __shfl_sync(FULL_MASK, d0, ld0);
            d1 = __shfl_sync(FULL_MASK, d1, ld1);
        }
        if (have && d0 != 0xffffffffu && d1 != 0xffffffffu) {   // both components have >= 25 pixels
            ok = true;
            // 2*id_bits key bits: as few sort passes as needed.  Ids that do not fit (the host re-runs such a chunk
            // with wider ids) are clamped so that nothing downstream indexes out of range in the meantime.
            const uint32_t idmax = (1u << id_bits) - 1u;
            d0 = min(d0, idmax);
            d1 = min(d1, idmax);
            const uint32_t key = (max(d0, d1) << id_bits) | min(d0, d1);
            rec = ((unsigned long long)key << 32) | pack_point(2 * x + dx, 2 * y + dy, d, pos);
        }
        const uint32_t okm = __ballot_sync(FULL_MASK, ok);
        if (okm == 0) continue;
        int base = 0;
        if (lane == 0) base = atomicAdd(&npts[frame], __popc(okm));
        base = __shfl_sync(FULL_MASK, base, 0);
        const int p = base + __popc(okm & ((1u << lane) - 1u));
        if (ok && p < cap) fk[p] = rec;
    }
}

// ---- segmented LSD radix sort ---------------------------------------------------------------
// Records are single 64-bit words, (pair key << 32) | point; only the 32 key bits are sorted: three passes of
// 11-bit digits (2048 bins).
#define RS_THREADS 512
#define RS_ITEMS 16
#define RS_TILE (RS_THREADS * RS_ITEMS)  // 8192 records per block: 1 byte of histogram traffic per record
#define RS_BITS 11
#define RS_RADIX (1 << RS_BITS)
#define RS_SCAN_PARTS 8
#define RS_SCATTER_SMEM ((RS_THREADS / 32) * RS_RADIX * 2 + RS_RADIX * 4)

// per frame: hist[block][digit] (block-major, frame stride RS_RADIX * nblk_max)
__global__ void __launch_bounds__(RS_THREADS)
k_sort_hist(const unsigned long long* __restrict__ recs, const int* __restrict__ npts, int cap, int shift,
            uint32_t* __restrict__ hist, int nblk_max) {
    __shared__ uint32_t h[RS_RADIX];
    const int frame = blockIdx.y, b = blockIdx.x;
    const int n = min(npts[frame], cap);
    const int nblk = (n + RS_TILE - 1) / RS_TILE;
    if (b >= nblk) return;
    for (int i = threadIdx.x; i < RS_RADIX; i += RS_THREADS) h[i] = 0;
    __syncthreads();
    const unsigned long long* fk = recs + (size_t)frame * cap;
    const int base = b * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        int i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(fk[i] >> shift) & (RS_RADIX - 1)], 1u);
    }
    __syncthreads();
    uint32_t* out = hist + ((size_t)frame * nblk_max + b) * RS_RADIX;
    for (int i = threadIdx.x; i < RS_RADIX; i += RS_THREADS) out[i] = h[i];
}

// grid (frames, RS_SCAN_PARTS): thread = one digit; exclusive prefix over the blocks (in place) and the digit's
// total; the prefix over the digit totals is taken by every scatter block itself
__global__ void __launch_bounds__(RS_RADIX / RS_SCAN_PARTS)
k_sort_scan(const int* __restrict__ npts, int cap, uint32_t* __restrict__ hist, uint32_t* __restrict__ digit_total,
            int nblk_max) {
    const int frame = blockIdx.x;
    const int n = min(npts[frame], cap);
    const int nblk = (n + RS_TILE - 1) / RS_TILE;
    const int d = blockIdx.y * blockDim.x + threadIdx.x;
    uint32_t* fh = hist + (size_t)frame * nblk_max * RS_RADIX + d;
    uint32_t run = 0;
    for (int b0 = 0; b0 < nblk; b0 += 8) {   // batches of 8 independent loads, then the 8 dependent stores
        uint32_t c[8];
#pragma unroll
        for (int k = 0; k < 8; k++) c[k] = b0 + k < nblk ? fh[(size_t)(b0 + k) * RS_RADIX] : 0u;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (b0 + k < nblk) {
                fh[(size_t)(b0 + k) * RS_RADIX] = run;
                run += c[k];
            }
    }
    digit_total[(size_t)frame * RS_RADIX + d] = run;
}

__global__ void __launch_bounds__(RS_THREADS)
k_sort_scatter(const unsigned long long* __restrict__ recs_in, unsigned long long* __restrict__ recs_out,
               const int* __restrict__ npts, int cap, int shift, const uint32_t* __restrict__ hist,
               const uint32_t* __restrict__ digit_total, int nblk_max) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    uint16_t (*wcnt)[RS_RADIX] = reinterpret_cast<uint16_t (*)[RS_RADIX]>(rs_smem);   // per-warp digit counts (<= 512), then warp prefixes
    uint32_t* dbase = reinterpret_cast<uint32_t*>(rs_smem + (RS_THREADS / 32) * RS_RADIX * 2);   // first output slot of digit d
    __shared__ uint32_t wtot[RS_THREADS / 32];
    const int frame = blockIdx.y, b = blockIdx.x;
    const int n = min(npts[frame], cap);
    const int nblk = (n + RS_TILE - 1) / RS_TILE;
    if (b >= nblk) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(&wcnt[0][0]);
        for (int i = threadIdx.x; i < (RS_THREADS / 32) * RS_RADIX / 2; i += RS_THREADS) z[i] = 0;
    }
    static_assert(RS_RADIX / RS_THREADS == 4, "digit ownership below assumes 4 digits per thread");
    const size_t seg = (size_t)frame * cap;
    const int base = b * RS_TILE + w * (32 * RS_ITEMS);
    unsigned long long rec[RS_ITEMS];
    uint32_t rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const int i = base + r * 32 + lane;
        rec[r] = i < n ? recs_in[seg + i] : 0xffffffffffffffffull;
    }
    // exclusive prefix of the frame's digit totals: thread t owns digits 4t .. 4t+3
    {
        constexpr int DPT = RS_RADIX / RS_THREADS;
        const uint32_t* dt = digit_total + (size_t)frame * RS_RADIX + threadIdx.x * DPT;
        const uint4 a = *reinterpret_cast<const uint4*>(dt);
        const uint32_t v[DPT] = {a.x, a.y, a.z, a.w};
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < DPT; k++) sum += v[k];
        uint32_t incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t t = __shfl_up_sync(FULL_MASK, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) wtot[w] = incl;
        __syncthreads();
        uint32_t run = incl - sum;
#pragma unroll
        for (int ww = 0; ww < RS_THREADS / 32; ww++)
            if (ww < w) run += wtot[ww];
#pragma unroll
        for (int k = 0; k < DPT; k++) {
            dbase[threadIdx.x * DPT + k] = run;
            run += v[k];
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const int i = base + r * 32 + lane;
        const bool valid = i < n;
        const uint32_t d = valid ? ((uint32_t)(rec[r] >> shift) & (RS_RADIX - 1)) : 0xffffffffu;
        const uint32_t peers = __match_any_sync(FULL_MASK, d);
        const int leader = __ffs(peers) - 1;
        uint32_t before = 0;
        if (valid && lane == leader) {
            before = wcnt[w][d];
            wcnt[w][d] = (uint16_t)(before + __popc(peers));
        }
        before = __shfl_sync(FULL_MASK, before, leader);
        rank[r] = before + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    {   // per-warp counts -> exclusive warp prefixes; add the block's offset inside each digit
        const uint32_t* bh = hist + ((size_t)frame * nblk_max + b) * RS_RADIX;
        for (int d = threadIdx.x; d < RS_RADIX; d += RS_THREADS) {
            uint32_t run = 0;
#pragma unroll
            for (int ww = 0; ww < RS_THREADS / 32; ww++) {
                uint32_t c = wcnt[ww][d];
                wcnt[ww][d] = (uint16_t)run;
                run += c;
            }
            dbase[d] += bh[d];
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const int i = base + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t)(rec[r] >> shift) & (RS_RADIX - 1);
            recs_out[seg + dbase[d] + wcnt[w][d] + rank[r]] = rec[r];
        }
    }
}

// cluster work lists, by size tier (the tier decides how much shared memory the fitting warp gets)
#define AGPU_NTIERS 4
struct ClusterLists {
    ClusterRef* list[AGPU_NTIERS];
    int cap[AGPU_NTIERS];    // largest cluster size of the tier
    int* counters;           // [0..3] tier counts, [4] oversize (skipped), [5] all heads (debug), [8..11] tier work cursors
    int cap_list;
    ClusterRef* dbg_heads;   // all run heads (debug only, may be null)
    int cap_dbg;
};

__global__ void __launch_bounds__(256)
k_cluster_heads(const unsigned long long* __restrict__ recs, const int* __restrict__ npts, int cap, Geom g,
                int min_size, ClusterLists cl) {
    const int frame = blockIdx.y;
    const int n = min(npts[frame], cap);
    const unsigned long long* fk = recs + (size_t)frame * cap;
    const int max_cluster = 3 * (2 * g.wd + 2 * g.hd);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t k = (uint32_t)(fk[i] >> 32);
        if (i > 0 && (uint32_t)(fk[i - 1] >> 32) == k) continue;
        int lo = i + 1, hi = n;  // first index in (i, n] whose key differs
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if ((uint32_t)(fk[mid] >> 32) == k) lo = mid + 1; else hi = mid;
        }
        const int size = lo - i;
        ClusterRef ref;
        ref.frame = frame; ref.start = i; ref.size = size; ref.pad = 0;
        if (cl.dbg_heads) {
            int s = atomicAdd(&cl.counters[5], 1);
            if (s < cl.cap_dbg) cl.dbg_heads[s] = ref;
        }
        if (size < min_size || size > max_cluster) continue;
        bool placed = false;
#pragma unroll
        for (int t = 0; t < AGPU_NTIERS; t++)
            if (!placed && size <= cl.cap[t]) {
                int s = atomicAdd(&cl.counters[t], 1);
                if (s < cl.cap_list) cl.list[t][s] = ref;
                placed = true;
            }
        if (!placed) atomicAdd(&cl.counters[4], 1);
    }
}
