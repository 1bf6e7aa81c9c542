// k_quad.cuh -- quad fitting, one warp per cluster (upstream stage U6 fit_quad /
// quad_segment_maxima / fit_line; part of the native call at
// /root/reference/src/detection/tag_detector.py:26).
//
// Per cluster: records staged once in shared memory -> bounding box + border polarity (exact integer sums) -> slope
// keys -> stable radix sort on the slope bits (ping-pong shared memory / L2 scratch) + tie fix on (y, x) -> prefix
// line-fit moments (terms in parallel, running sums in upstream's sequential order on six lanes; stored in the group's
// own L2-resident scratch) -> windowed line-fit error, 7-tap smoothing, local
// maxima, top-10 selection -> exhaustive 4-corner search over a pre-computed table of pairwise line fits (dealt to
// the threads by triples) -> corners + gates on one warp.  (The only duplicate points upstream produces are merged at
// emission by k_edges, so there is no de-duplication pass.)
#pragma once
#include "common.cuh"
#include "k_cc.cuh"

struct QuadFitArgs {
    const unsigned long long* recs;   // sorted records (cluster id << 32 | packed point), per-frame segments of `cap`
    const uint32_t* dense2rep;        // [nframes][AGPU_MAX_DENSE] dense component id -> representative pixel id
    const uint32_t* pair_keys;        // [nframes][cap_keys] cluster id (upper half of a record) -> pair of dense ids
    int cap_keys;
    int cap;
    const uint8_t* quad_im;     // decimated gray image (may alias the source frames)
    size_t q_pitch, q_frame;    // bytes per row / per frame of quad_im
    Geom g;
    double* scratch;            // this tier's per-GROUP scratch, 7 * scratch_pts doubles each: [scratch_pts][6] prefix moments
    int scratch_pts;            // (+ ping-pong partner of the in-cluster sort) and [scratch_pts] smoothed line-fit errors.
                                // A persistent group re-uses ITS region for every cluster it fits, so the lines stay in L2
                                // and are overwritten there instead of streaming through HBM once per cluster
    const ClusterRef* list;
    const int* list_count;
    int* cursor;                // next unclaimed cluster of the list (dynamic work distribution)
    int list_cap;
    QuadRec* quads;
    int* nquads;
    int cap_quads;              // per chunk
    int* per_frame_quads;       // [nframes] count per frame (limit check / debug)
    int* oversize;              // clusters over upstream's raw-point limit 3(2w+2h) (counted, not fitted -- as upstream)
    unsigned long long* gsort;  // last tier only: per-CTA sort buffer in global memory for clusters that do not fit the
    size_t gsort_stride;        // shared-memory one (frames so large that 3(2w+2h) records exceed it); may be null
};

struct LineFit {
    double Ex, Ey, nx, ny, err, mse;
};

__device__ __forceinline__ void ld_lfp(const double* lf, int i, double (&m)[6]) {
    const double2* p = reinterpret_cast<const double2*>(lf + (size_t)i * 6);
    double2 a = __ldcg(p), b = __ldcg(p + 1), c = __ldcg(p + 2);
    m[0] = a.x; m[1] = a.y; m[2] = b.x; m[3] = b.y; m[4] = c.x; m[5] = c.y;
}

// same arithmetic as the oracle's fit_line (apriltag_oracle.cpp), moments by prefix differences
// (__noinline__ and the `unroll 1` pragmas below: a warp walks through most of this kernel ONCE per cluster, so the
// kernel's instruction footprint -- not its instruction count -- is what the SM's instruction caches see; ncu showed
// no_instruction as the top stall.  Every kilobyte of code less is measurable: -19 % SASS gave -10 % time.)
__device__ __noinline__ void fit_line_dev(const double* lf, int sz, int i0, int i1, LineFit& out, bool want_params) {
    double M[6], T[6];
    int N;
    if (i0 < i1) {
        N = i1 - i0 + 1;
        ld_lfp(lf, i1, M);
        if (i0 > 0) {
            ld_lfp(lf, i0 - 1, T);
#pragma unroll
            for (int k = 0; k < 6; k++) M[k] -= T[k];
        }
    } else {
        double A[6];
        ld_lfp(lf, sz - 1, A);
        ld_lfp(lf, i0 - 1, T);
        ld_lfp(lf, i1, M);
#pragma unroll
        for (int k = 0; k < 6; k++) M[k] = (A[k] - T[k]) + M[k];
        N = sz - i0 + i1 + 1;
    }
    const double Mx = M[0], My = M[1], Mxx = M[2], Mxy = M[3], Myy = M[4], W = M[5];
    double Ex = Mx / W, Ey = My / W;
    double Cxx = Mxx / W - Ex * Ex;
    double Cxy = Mxy / W - Ex * Ey;
    double Cyy = Myy / W - Ey * Ey;
    float disc = sqrtf((float)((Cxx - Cyy) * (Cxx - Cyy) + 4 * Cxy * Cxy));
    double eig_small = 0.5 * (Cxx + Cyy - disc);
    if (want_params) {
        out.Ex = Ex;
        out.Ey = Ey;
        double eig = 0.5 * (Cxx + Cyy + disc);
        double nx1 = Cxx - eig, ny1 = Cxy, M1 = nx1 * nx1 + ny1 * ny1;
        double nx2 = Cxy, ny2 = Cyy - eig, M2 = nx2 * nx2 + ny2 * ny2;
        double nx, ny, MM;
        if (M1 > M2) { nx = nx1; ny = ny1; MM = M1; } else { nx = nx2; ny = ny2; MM = M2; }
        double length = sqrtf((float)MM);
        if (fabs(length) < 1e-12) { out.nx = 0; out.ny = 0; }
        else { out.nx = nx / length; out.ny = ny / length; }
    }
    out.err = N * eig_small;
    out.mse = eig_small;
}

#define QF_PTAB_DOUBLES (100 * 6)

// the C(9,3) = 84 triples m0 < m1 < m2 <= 8 in colexicographic order (by m2, then m1, then m0), packed m0 | m1 << 4 | m2 << 8
struct QfTrips { unsigned short v[84]; };
constexpr QfTrips qf_make_trips() {
    QfTrips t{};
    int k = 0;
    for (int c = 2; c <= 8; c++)
        for (int b = 1; b < c; b++)
            for (int a = 0; a < b; a++) t.v[k++] = (unsigned short)(a | (b << 4) | (c << 8));
    return t;
}
__constant__ QfTrips c_qf_trips = qf_make_trips();

// A cluster is fitted by a GROUP of NW warps.  NW == 1: the group is a warp (several clusters per CTA, warp
// synchronisation); NW > 1: the group is the whole CTA (one cluster at a time, __syncthreads), which puts
// NW times more warps on every shared-memory sort buffer -- the buffer, not the thread count, is what limits
// how many clusters an SM can hold, so the latency-bound phases get NW times more warps to hide behind.
template <int NW>
struct QGroup {
    static constexpr int T = NW * 32;
    int tid, lane, w;
    int* si;        // [NW + 4] ints of scratch
    double* sd;     // [NW * 8] doubles of scratch
    __device__ __forceinline__ void sync() const {
        if (NW == 1) __syncwarp(); else __syncthreads();
    }
    // position of this thread's kept element among the kept elements of the group (thread order) and their number
    __device__ __forceinline__ int compact_pos(bool keep, int& total) const {
        const uint32_t m = __ballot_sync(FULL_MASK, keep);
        const int below = __popc(m & ((1u << lane) - 1u));
        if (NW == 1) { total = __popc(m); return below; }
        if (lane == 0) si[w] = __popc(m);
        __syncthreads();
        int off = 0, tot = 0;
#pragma unroll
        for (int i = 0; i < NW; i++) { const int c = si[i]; if (i < w) off += c; tot += c; }
        __syncthreads();
        total = tot;
        return off + below;
    }
    __device__ __forceinline__ int reduce_min(int v) const {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = min(v, __shfl_xor_sync(FULL_MASK, v, off));
        if (NW == 1) return v;
        if (lane == 0) si[w] = v;
        __syncthreads();
        int r = si[0];
#pragma unroll
        for (int i = 1; i < NW; i++) r = min(r, si[i]);
        __syncthreads();
        return r;
    }
    __device__ __forceinline__ int reduce_max(int v) const { return -reduce_min(-v); }
    // (max of the 64-bit keys, how many threads-worth of `cnt` carry it): cnt is only summed where v == max
    __device__ __forceinline__ unsigned long long reduce_max_count(unsigned long long v, int cnt, int& total) const {
        unsigned long long m = v;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(FULL_MASK, m, off);
            m = o > m ? o : m;
        }
        if (NW > 1) {
            unsigned long long* sl = reinterpret_cast<unsigned long long*>(sd);
            if (lane == 0) sl[w] = m;
            __syncthreads();
#pragma unroll
            for (int i = 0; i < NW; i++) m = sl[i] > m ? sl[i] : m;
            __syncthreads();
        }
        total = (int)reduce_sum(v == m ? (long long)cnt : 0ll);
        return m;
    }
    __device__ __forceinline__ long long reduce_sum(long long v) const {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL_MASK, v, off);
        if (NW == 1) return v;
        long long* sl = reinterpret_cast<long long*>(sd);
        if (lane == 0) sl[w] = v;
        __syncthreads();
        long long r = 0;
#pragma unroll
        for (int i = 0; i < NW; i++) r += sl[i];
        __syncthreads();
        return r;
    }
};

// Stable LSD radix sort of n 64-bit keys by their upper 32 bits (the slope), 8-bit digits, four passes that
// ping-pong between the shared-memory buffer and a global scratch (ends in shared memory).  Warp w owns the
// contiguous chunk [w*cw, (w+1)*cw) of the source; ranks come from __match_any_sync, so the order of equal
// digits is preserved.  cnt: NW*256 16-bit counters.  Roughly a third of the instructions of a bitonic
// network for the cluster sizes that matter (1000-2000 points).
template <int NW>
__device__ void group_radix_sort_hi32(const QGroup<NW>& G, unsigned long long* sbuf, unsigned long long* gbuf, int n,
                                      uint16_t* cnt) {
    constexpr int T = QGroup<NW>::T;
    const int lane = G.lane, w = G.w, tid = G.tid;
    const int cw = ((n + NW - 1) / NW + 31) & ~31;          // chunk per warp, whole rounds of 32
    const int lo = w * cw, hi = min(lo + cw, n);
#pragma unroll 1
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 32 + 8 * pass;
        const bool from_smem = (pass & 1) == 0;
#pragma unroll 1
        for (int i = tid; i < NW * 256; i += T) cnt[i] = 0;
        G.sync();
        // (the key of the NEXT round is loaded before the current one is ranked: on the passes that read the global
        // scratch the L2 round trip overlaps the match / counter chain instead of heading it)
        auto ld_key = [&](int i) -> unsigned long long { return i < hi ? (from_smem ? sbuf[i] : __ldcg(gbuf + i)) : 0ull; };
        // sweep 1: per-warp digit counts
        unsigned long long knext = ld_key(lo + lane);
#pragma unroll 1
        for (int base = lo; base < hi; base += 32) {
            const int i = base + lane;
            const bool valid = i < hi;
            const unsigned long long k = knext;
            knext = ld_key(i + 32);
            const uint32_t d = valid ? (uint32_t)(k >> shift) & 255u : 0xffffffffu;
            const uint32_t peers = __match_any_sync(FULL_MASK, d);
            if (valid && lane == __ffs(peers) - 1) cnt[w * 256 + d] += (uint16_t)__popc(peers);
            __syncwarp();
        }
        G.sync();
        // exclusive prefix over (digit, warp): thread t owns digits t*DPT .. t*DPT+DPT-1
        {
            constexpr int DPT = 256 / T > 0 ? 256 / T : 1;   // T <= 256 here
            uint32_t tot[DPT];
            uint32_t sum = 0;
#pragma unroll
            for (int q = 0; q < DPT; q++) {
                const int d = tid * DPT + q;
                uint32_t run = 0;
#pragma unroll
                for (int ww = 0; ww < NW; ww++) {
                    const uint32_t c = cnt[ww * 256 + d];
                    cnt[ww * 256 + d] = (uint16_t)run;      // offset of warp ww inside digit d
                    run += c;
                }
                tot[q] = run;
                sum += run;
            }
            uint32_t incl = sum;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL_MASK, incl, off);
                if (lane >= off) incl += t;
            }
            uint32_t excl = incl - sum;
            if (NW > 1) {
                if (lane == 31) G.si[w] = (int)incl;
                __syncthreads();
#pragma unroll
                for (int ww = 0; ww < NW; ww++)
                    if (ww < w) excl += (uint32_t)G.si[ww];
                __syncthreads();
            }
#pragma unroll
            for (int q = 0; q < DPT; q++) {
                const int d = tid * DPT + q;
#pragma unroll
                for (int ww = 0; ww < NW; ww++) cnt[ww * 256 + d] = (uint16_t)(cnt[ww * 256 + d] + excl);
                excl += tot[q];
            }
        }
        G.sync();
        // sweep 2: scatter
        knext = ld_key(lo + lane);
#pragma unroll 1
        for (int base = lo; base < hi; base += 32) {
            const int i = base + lane;
            const bool valid = i < hi;
            const unsigned long long k = knext;
            knext = ld_key(i + 32);
            const uint32_t d = valid ? (uint32_t)(k >> shift) & 255u : 0xffffffffu;
            const uint32_t peers = __match_any_sync(FULL_MASK, d);
            const int leader = __ffs(peers) - 1;
            uint32_t start = 0;
            if (valid && lane == leader) {
                start = cnt[w * 256 + d];
                cnt[w * 256 + d] = (uint16_t)(start + __popc(peers));
            }
            start = __shfl_sync(FULL_MASK, start, leader);
            if (valid) {
                const uint32_t pos = start + __popc(peers & ((1u << lane) - 1u));
                if (from_smem) __stcg(gbuf + pos, k); else sbuf[pos] = k;
            }
            __syncwarp();
        }
        __threadfence_block();
        G.sync();
    }
}

// After the slope sort: order the (rare) runs of equal slope by (y, x), i.e. finish the sort on the full 64-bit key
// with odd-even transposition rounds until nothing moves (equal points -- duplicates -- never move).
template <int NW>
__device__ void group_fix_ties(const QGroup<NW>& G, unsigned long long* s, int n) {
    constexpr int T = QGroup<NW>::T;
    for (;;) {
        bool moved = false;
#pragma unroll
        for (int phase = 0; phase < 2; phase++) {
#pragma unroll 1
            for (int i = 2 * G.tid + phase; i + 1 < n; i += 2 * T) {
                const unsigned long long a = s[i], b = s[i + 1];
                if (a > b) { s[i] = b; s[i + 1] = a; moved = true; }
            }
            G.sync();
        }
        int total;
        (void)G.compact_pos(moved, total);
        if (total == 0) break;
    }
}

// Bucket sort of the n 64-bit keys (slope bits << 32 | y << 16 | x, all distinct) in `sbuf`: ONE counting pass instead
// of four radix passes.  The bucket of a key is a monotone function of its slope float f -- quadrant k of the slope
// key (f lives near -65536, 0, 65536, 131072) and r / (1 + r) of the remainder r = tan of the angle inside the
// quadrant, i.e. roughly the polar angle of the point, which is what spreads the points of a contour evenly --, so the
// buckets are in key order; inside a bucket (one or two keys on average) one thread finishes the order on the full
// 64-bit key by insertion.  Counts with shared-memory atomics (two 16-bit counters per word), scatter through the L2
// scratch.  Returns false -- nothing moved -- when some bucket holds more than QF_BUCKET_LIMIT keys (degenerate
// contours); the caller then falls back to the radix sort.  The result is the same total order either way.
#define QF_BUCKET_LIMIT 24
__device__ __forceinline__ int qf_bucket(uint32_t orderable, int per_quadrant) {
    const uint32_t bits = (orderable & 0x80000000u) ? (orderable ^ 0x80000000u) : ~orderable;
    const float f = __uint_as_float(bits);
    if (!(f == f)) return 4 * per_quadrant - 1;                      // NaN keys carry the largest bit pattern
    const int k = f < 0.0f ? 0 : (f < 65536.0f ? 1 : (f < 131072.0f ? 2 : 3));
    const float r = fmaxf(f - (float)(k - 1) * 65536.0f, 0.0f);      // (f below -65536 cannot occur: clamp)
    const float h = r >= 1e30f ? 1.0f : r / (1.0f + r);              // IEEE division: monotone in r
    return k * per_quadrant + min(per_quadrant - 1, (int)(h * (float)per_quadrant));
}

template <int NW>
__device__ bool group_bucket_sort(const QGroup<NW>& G, unsigned long long* sbuf, unsigned long long* gbuf, int n,
                                  uint32_t* cnt /* NB / 2 words */) {
    constexpr int T = QGroup<NW>::T;
    constexpr int PQ = NW == 1 ? 128 : 256, NB = 4 * PQ, NWORDS = NB / 2;
    constexpr int WPT = NWORDS / T;                                  // counter words per thread: 8 (NW 1, 2), 4, 2
    static_assert(WPT >= 1 && WPT * T == NWORDS, "bucket ownership");
    const int tid = G.tid, lane = G.lane;
#pragma unroll 1
    for (int i = tid; i < NWORDS; i += T) cnt[i] = 0;
    G.sync();
#pragma unroll 1
    for (int i = tid; i < n; i += T) {
        const int b = qf_bucket((uint32_t)(sbuf[i] >> 32), PQ);
        atomicAdd(&cnt[b >> 1], 1u << ((b & 1) * 16));
    }
    G.sync();
    // thread t owns the buckets [2 * WPT * t, 2 * WPT * (t + 1)): largest bucket, then exclusive prefix -> start offsets
    uint32_t w[WPT];
    int mx = 0, sum = 0;
#pragma unroll
    for (int q = 0; q < WPT; q++) {
        w[q] = cnt[tid * WPT + q];
        const int lo = w[q] & 0xffffu, hi = w[q] >> 16;
        mx = max(mx, max(lo, hi));
        sum += lo + hi;
    }
    if (G.reduce_max(mx) > QF_BUCKET_LIMIT) return false;            // (uniform)
    int incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(FULL_MASK, incl, off);
        if (lane >= off) incl += t;
    }
    int excl = incl - sum;
    if (NW > 1) {
        if (lane == 31) G.si[G.w] = incl;
        __syncthreads();
#pragma unroll
        for (int ww = 0; ww < NW; ww++)
            if (ww < G.w) excl += G.si[ww];
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < WPT; q++) {
        const int lo = w[q] & 0xffffu, hi = w[q] >> 16;
        cnt[tid * WPT + q] = (uint32_t)excl | ((uint32_t)(excl + lo) << 16);
        excl += lo + hi;
    }
    G.sync();
    // scatter through the scratch (a bucket's keys arrive in any order; the insertion below fixes it), copy back
#pragma unroll 1
    for (int i = tid; i < n; i += T) {
        const unsigned long long k = sbuf[i];
        const int b = qf_bucket((uint32_t)(k >> 32), PQ);
        const uint32_t old = atomicAdd(&cnt[b >> 1], 1u << ((b & 1) * 16));
        __stcg(gbuf + ((old >> ((b & 1) * 16)) & 0xffffu), k);
    }
    __threadfence_block();
    G.sync();
#pragma unroll 1
    for (int i = tid; i < n; i += T) sbuf[i] = __ldcg(gbuf + i);
    G.sync();
    // the counters now hold the END of every bucket: thread t finishes the buckets t, t + T, ...
#pragma unroll 1
    for (int b = tid; b < NB; b += T) {
        const int e = (cnt[b >> 1] >> ((b & 1) * 16)) & 0xffff;
        const int s0 = b == 0 ? 0 : (int)((cnt[(b - 1) >> 1] >> (((b - 1) & 1) * 16)) & 0xffff);
#pragma unroll 1
        for (int i = s0 + 1; i < e; i++) {
            const unsigned long long k = sbuf[i];
            int j = i - 1;
            while (j >= s0 && sbuf[j] > k) { sbuf[j + 1] = sbuf[j]; j--; }
            sbuf[j + 1] = k;
        }
    }
    G.sync();
    return true;
}

// Tiny clusters (sensor noise produces thousands per frame): rank sort.  Every thread counts the keys below its own -- the
// keys are distinct, so the count is the final position -- with broadcast reads of the shared buffer; no counters, no
// scratch traffic, a few hundred instructions for the 30..60 keys such a cluster has.
#define QF_RANK_SORT_MAX 64
template <int NW>
__device__ void group_rank_sort(const QGroup<NW>& G, unsigned long long* sbuf, unsigned long long* tmp, int n) {
    constexpr int T = QGroup<NW>::T;
#pragma unroll 1
    for (int i = G.tid; i < n; i += T) {
        const unsigned long long k = sbuf[i];
        int r = 0;
#pragma unroll 4
        for (int j = 0; j < n; j++) r += sbuf[j] < k;
        tmp[r] = k;
    }
    G.sync();
#pragma unroll 1
    for (int i = G.tid; i < n; i += T) sbuf[i] = tmp[i];
    G.sync();
}

// Returns true (uniformly over the group) and fills q when the cluster yields a quad.
template <int NW>
__device__ bool fit_cluster_group(const QGroup<NW>& G, const QuadFitArgs& a, const DevParams& P, const ClusterRef ref,
                                  unsigned long long* sbuf, double* ptab, int* sidx, uint16_t* scnt, double* stage, double* lf,
                                  long long* red, QuadRec& q) {
    constexpr int T = QGroup<NW>::T;
    const int tid = G.tid, lane = G.lane;
    const size_t seg = (size_t)ref.frame * a.cap + ref.start;
    const unsigned long long* pv = a.recs + seg;
    int sz = ref.size;

    // ---- bounding box and polarity sums
    int xmin = 1 << 30, xmax = -1, ymin = 1 << 30, ymax = -1;
    long long Sxgx = 0, Sygy = 0, Sgx = 0, Sgy = 0;
    int nmerged = 0;
    // the cluster's records come in from L2 / HBM exactly once, four loads in flight per thread, and are parked in the
    // sort buffer: the passes below (thread t always touches the entries t, t + T, ...) read shared memory
#pragma unroll 1
    for (int i0 = tid; i0 < sz; i0 += 4 * T) {
        uint32_t v4[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v4[u] = i0 + u * T < sz ? (uint32_t)__ldg(pv + i0 + u * T) : 0u;
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (i0 + u * T < sz) sbuf[i0 + u * T] = v4[u];
    }
#pragma unroll 1
    for (int i = tid; i < sz; i += T) {
        uint32_t v = (uint32_t)sbuf[i];
        int px = v & 0x3fff, py = (v >> 14) & 0x3fff;
        const int kind = v >> 28;
        int gx, gy;
        if (kind < 8) {
            const int dir = kind & 3, s = (kind & 4) ? 255 : -255;
            const int dx = (dir == 0 || dir == 3) ? 1 : (dir == 2 ? -1 : 0);
            const int dy = dir == 0 ? 0 : 1;
            gx = dx * s; gy = dy * s;
        } else {   // two raw points at this (x, y): direction 2 (-1, 1) and direction 3 (1, 1)
            const int s2 = (kind & 1) ? 255 : -255, s3 = (kind & 2) ? 255 : -255;
            gx = s3 - s2; gy = s2 + s3;
            nmerged++;
        }
        xmin = min(xmin, px); xmax = max(xmax, px);
        ymin = min(ymin, py); ymax = max(ymax, py);
        Sxgx += (long long)px * gx; Sgx += gx;
        Sygy += (long long)py * gy; Sgy += gy;
    }
    // all nine group reductions at once: REDUX inside the warps (the 64-bit sums as three 20-bit slices of the biased
    // value, so no lane sum can wrap), one shared-memory exchange and ONE barrier between the warps
    {
        xmin = __reduce_min_sync(FULL_MASK, xmin); xmax = __reduce_max_sync(FULL_MASK, xmax);
        ymin = __reduce_min_sync(FULL_MASK, ymin); ymax = __reduce_max_sync(FULL_MASK, ymax);
        nmerged = __reduce_add_sync(FULL_MASK, nmerged);
        auto warp_sum64 = [](long long v) -> long long {   // |v| < 2^41 (98280 records x 16380 x 510 < 2^40)
            const unsigned long long u = (unsigned long long)(v + (1ll << 41));
            const unsigned long long c0 = __reduce_add_sync(FULL_MASK, (unsigned)(u & 0xfffffu));
            const unsigned long long c1 = __reduce_add_sync(FULL_MASK, (unsigned)((u >> 20) & 0xfffffu));
            const unsigned long long c2 = __reduce_add_sync(FULL_MASK, (unsigned)(u >> 40));
            return (long long)(c0 + (c1 << 20) + (c2 << 40)) - (32ll << 41);
        };
        Sxgx = warp_sum64(Sxgx); Sygy = warp_sum64(Sygy); Sgx = warp_sum64(Sgx); Sgy = warp_sum64(Sgy);
        if (NW > 1) {
            long long* sr = red;   // [NW][10]
            if (lane == 0) {
                long long* o = sr + G.w * 10;
                o[0] = xmin; o[1] = xmax; o[2] = ymin; o[3] = ymax; o[4] = nmerged; o[5] = Sxgx; o[6] = Sygy; o[7] = Sgx; o[8] = Sgy;
            }
            __syncthreads();
            xmin = (int)sr[0]; xmax = (int)sr[1]; ymin = (int)sr[2]; ymax = (int)sr[3]; nmerged = (int)sr[4];
            Sxgx = sr[5]; Sygy = sr[6]; Sgx = sr[7]; Sgy = sr[8];
#pragma unroll
            for (int ww = 1; ww < NW; ww++) {
                const long long* o = sr + ww * 10;
                xmin = min(xmin, (int)o[0]); xmax = max(xmax, (int)o[1]); ymin = min(ymin, (int)o[2]); ymax = max(ymax, (int)o[3]);
                nmerged += (int)o[4]; Sxgx += o[5]; Sygy += o[6]; Sgx += o[7]; Sgy += o[8];
            }
        }
    }
    // upstream's size limit counts the raw points (duplicates included)
    if (sz + nmerged > 3 * (2 * a.g.wd + 2 * a.g.hd)) {
        if (tid == 0) atomicAdd(a.oversize, 1);
        return false;
    }
    if ((xmax - xmin) * (ymax - ymin) < P.min_tag_width) return false;
    const float cx = (xmin + xmax) * 0.5f + 0.05118f;
    const float cy = (ymin + ymax) * 0.5f - 0.028581f;
    const double dot = ((double)Sxgx - (double)cx * (double)Sgx) + ((double)Sygy - (double)cy * (double)Sgy);
    const int reversed = dot < 0;
    if (!P.reversed_border && reversed) return false;
    if (!P.normal_border && !reversed) return false;

    // ---- slope keys
#pragma unroll 1
    for (int i = tid; i < sz; i += T) {
        unsigned long long key = ~0ull;
        {
            uint32_t v = (uint32_t)sbuf[i];
            int px = v & 0x3fff, py = (v >> 14) & 0x3fff;
            float dx = (float)px - cx, dy = (float)py - cy;
            float qd;
            if (dy > 0) qd = (dx > 0) ? 65536.0f : 131072.0f;
            else qd = (dx > 0) ? 0.0f : -65536.0f;
            if (dy < 0) { dy = -dy; dx = -dx; }
            if (dx < 0) { float t = dx; dx = dy; dy = -t; }
            float slope = qd + dy / dx;
            key = ((unsigned long long)float_orderable(slope) << 32) | (uint32_t)((py << 16) | px);
        }
        sbuf[i] = key;
    }
    G.sync();
    // (the bucket counters / the rank sort's second buffer borrow the prefix-moment stage: 1 KB / 2 KB of the 1.6 KB / 3+ KB
    // it has)
    if (sz <= QF_RANK_SORT_MAX) group_rank_sort<NW>(G, sbuf, reinterpret_cast<unsigned long long*>(stage), sz);
    else if (sz > (NW == 1 ? 512 : 4096) ||
        !group_bucket_sort<NW>(G, sbuf, reinterpret_cast<unsigned long long*>(lf), sz, reinterpret_cast<uint32_t*>(stage))) {
        group_radix_sort_hi32<NW>(G, sbuf, reinterpret_cast<unsigned long long*>(lf), sz, scnt);
        group_fix_ties<NW>(G, sbuf, sz);
    }

    // (no duplicate points to remove here: the only duplicates upstream produces were merged at emission, k_edges)
    if (sz < 24) return false;

    // ---- gradient weights: squared gradient magnitude of the decimated image at every point, gathered in a
    //      fully parallel pass (independent loads) and parked in the upper half of the sort slot (the slope bits
    //      are dead after the sort), so that the scan below never waits on global memory
    {
        const uint8_t* im = a.quad_im + (size_t)ref.frame * a.q_frame;
#pragma unroll 2
        for (int i = tid; i < sz; i += T) {
            const uint32_t xy = (uint32_t)sbuf[i];
            const int px = xy & 0xffff, py = xy >> 16;
            const double x = px * .5 + 0.5, y = py * .5 + 0.5;
            const int ix = (int)x, iy = (int)y;
            uint32_t g2 = 0;   // W = sqrt(g2) + 1 = 1 outside the interior, as upstream
            if (ix > 0 && ix + 1 < a.g.wd && iy > 0 && iy + 1 < a.g.hd) {
                int grad_x = (int)im[(size_t)iy * a.q_pitch + ix + 1] - (int)im[(size_t)iy * a.q_pitch + ix - 1];
                int grad_y = (int)im[(size_t)(iy + 1) * a.q_pitch + ix] - (int)im[(size_t)(iy - 1) * a.q_pitch + ix];
                g2 = (uint32_t)(grad_x * grad_x + grad_y * grad_y);
            }
            sbuf[i] = ((unsigned long long)g2 << 32) | xy;
        }
    }
    G.sync();
    // ---- prefix moments (inclusive).  Upstream accumulates them sequentially, P[i] = P[i-1] + t[i], and every line fit
    //      below is a DIFFERENCE of two of these sums, so their rounding is part of the result: a parallel scan adds in
    //      another order and the last bits -- sometimes a decision -- come out differently.  So: T points at a time, every
    //      thread computes the six terms of its point into a shared-memory stage (moment-major, padded: conflict-free
    //      both ways), six lanes of warp 0 -- one per moment -- run upstream's sequential sums over the stage in place
    //      (LDS / DADD / STS per point, the DADD chain is the only dependency), and every thread writes its point's
    //      six sums to the group's scratch.  Bit-identical to the sequential loop, and cheaper in issue slots than a
    //      6 x double warp scan.
    {
        constexpr int SP = T + 2;   // stage pitch in doubles
        double carry = 0;           // (lanes 0..5 of warp 0: running sum of "their" moment)
#pragma unroll 1
        for (int base = 0; base < sz; base += T) {
            const int i = base + tid;
            double t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0;
            if (i < sz) {
                const unsigned long long e = sbuf[i];
                const uint32_t xy = (uint32_t)e;
                const int px = xy & 0xffff, py = xy >> 16;
                const double x = px * .5 + 0.5, y = py * .5 + 0.5;
                const double W = sqrt((double)(uint32_t)(e >> 32)) + 1;
                t0 = W * x; t1 = W * y; t2 = W * x * x; t3 = W * x * y; t4 = W * y * y; t5 = W;
            }
            stage[0 * SP + tid] = t0; stage[1 * SP + tid] = t1; stage[2 * SP + tid] = t2;
            stage[3 * SP + tid] = t3; stage[4 * SP + tid] = t4; stage[5 * SP + tid] = t5;
            G.sync();
            if (G.w == 0 && lane < 6) {
                double* p = stage + lane * SP;
#pragma unroll 1
                for (int j = 0; j < T; j += 8) {   // (entries past the cluster's end hold zeros; 128-bit shared-memory accesses)
                    double2 v[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) v[u] = *reinterpret_cast<const double2*>(p + j + 2 * u);
#pragma unroll
                    for (int u = 0; u < 4; u++) { carry += v[u].x; v[u].x = carry; carry += v[u].y; v[u].y = carry; }
#pragma unroll
                    for (int u = 0; u < 4; u++) *reinterpret_cast<double2*>(p + j + 2 * u) = v[u];
                }
            }
            G.sync();
            if (i < sz) {
                double2* p = reinterpret_cast<double2*>(lf + (size_t)i * 6);
                __stcg(p, make_double2(stage[0 * SP + tid], stage[1 * SP + tid]));
                __stcg(p + 1, make_double2(stage[2 * SP + tid], stage[3 * SP + tid]));
                __stcg(p + 2, make_double2(stage[4 * SP + tid], stage[5 * SP + tid]));
            }
        }
    }
    __threadfence_block();
    G.sync();

    // ---- windowed line-fit error -> sbuf (as doubles), smoothed -> scratch
    const int ksz = min(20, sz / 12);
    if (ksz < 2) return false;
    double* sraw = reinterpret_cast<double*>(sbuf);
#pragma unroll 1
    for (int i = tid; i < sz; i += T) {
        LineFit f;
        fit_line_dev(lf, sz, i - ksz < 0 ? i - ksz + sz : i - ksz, i + ksz >= sz ? i + ksz - sz : i + ksz, f, false);   // (index % sz without the division)
        sraw[i] = f.err;
    }
    G.sync();
    double* es = lf + (size_t)a.scratch_pts * 6;
#pragma unroll 1
    for (int i = tid; i < sz; i += T) {
        double acc = 0;
#pragma unroll
        for (int j = 0; j < 7; j++) {
            int q = i + j - 3;
            q = q < 0 ? q + sz : (q >= sz ? q - sz : q);
            acc += sraw[q] * P.smooth_f[j];
        }
        __stcg(es + i, acc);
    }
    __threadfence_block();
    G.sync();

    // ---- local maxima, in index order
    const int half = max((sz + 1) >> 1, 64);                // strict local maxima are never adjacent: at most sz/2 of them
    unsigned long long* mvals = sbuf;                       // [half]
    int* midx = reinterpret_cast<int*>(sbuf + half);        // [half] ints: 12 * half <= 8 * sz bytes in all
    int nmax = 0;
#pragma unroll 1
    for (int base = 0; base < sz; base += T) {
        const int i = base + tid;
        bool ismax = false;
        double e = 0;
        if (i < sz) {
            e = __ldcg(es + i);
            const double en = __ldcg(es + (i + 1 == sz ? 0 : i + 1)), ep = __ldcg(es + (i == 0 ? sz - 1 : i - 1));
            ismax = e > en && e > ep;
        }
        int total;
        const int pos = G.compact_pos(ismax, total);
        if (ismax) {
            midx[nmax + pos] = i;
            mvals[nmax + pos] = double_orderable(e);
        }
        nmax += total;
    }
    G.sync();
    if (nmax < 4) return false;
    if (nmax > P.max_nmaxima) {
        // threshold = the (max_nmaxima+1)-th largest value, multiplicities counted (what upstream reads off a
        // descending sort): walk down the distinct values, at most max_nmaxima+1 group reductions
        unsigned long long thr = 0ull, bound = ~0ull;
        int taken = 0;
#pragma unroll 1
        for (int it = 0; it <= P.max_nmaxima; it++) {
            unsigned long long best = 0ull;   // largest value strictly below `bound` among this thread's elements
#pragma unroll 1
            for (int i = tid; i < nmax; i += T) {
                const unsigned long long v = mvals[i];
                if (v < bound && v > best) best = v;
            }
            int mine = 0;                     // multiplicity of `best` among this thread's elements
#pragma unroll 1
            for (int i = tid; i < nmax; i += T) mine += (mvals[i] == best);
            int cnt;
            const unsigned long long m = G.reduce_max_count(best, mine, cnt);
            if (taken + cnt > P.max_nmaxima) { thr = m; break; }
            taken += cnt;
            bound = m;
        }
        int outn = 0;
#pragma unroll 1
        for (int base = 0; base < nmax; base += T) {
            const int i = base + tid;
            int id = 0;
            bool keep = false;
            if (i < nmax) {
                id = midx[i];
                keep = double_orderable(__ldcg(es + id)) > thr;
            }
            G.sync();
            int total;
            const int pos = G.compact_pos(keep, total);
            if (keep) midx[outn + pos] = id;
            outn += total;
            G.sync();
        }
        nmax = outn;
        if (nmax < 4) return false;   // (ties at the threshold can drop below 4: no 4-subset exists)
    }
    if (tid < nmax) sidx[tid] = midx[tid];
    G.sync();

    // ---- table of pairwise line fits between maxima (the table may alias sbuf: everything in it is dead now)
#pragma unroll 1
    for (int t = tid; t < nmax * nmax; t += T) {
        int ia = t / nmax, ib = t - ia * nmax;
        if (ia == ib) continue;
        LineFit f;
        fit_line_dev(lf, sz, sidx[ia], sidx[ib], f, true);
        double* e = ptab + (ia * 10 + ib) * 6;
        e[0] = f.Ex; e[1] = f.Ey; e[2] = f.nx; e[3] = f.ny; e[4] = f.err; e[5] = f.mse;
    }
    G.sync();

    // ---- exhaustive search over 4-subsets (first minimum in lexicographic order wins).  The work is dealt to the
    //      threads by (m0, m1, m2) TRIPLES -- the inner loop of a triple is at most seven candidates long, so the threads
    //      finish together -- taken from a table in colexicographic order (the triples for n maxima are a prefix of those
    //      for n + 1), so a thread goes straight to ITS triples instead of counting through all of them.  Which pairs of
    //      maxima are joined by an acceptable line (mse <= max) is a 100-bit mask in four registers (one ballot per 32
    //      pairs): the inner loop tests bits and touches the table only for the subsets that survive.  The tie-break
    //      key is the subset itself packed most-significant-first, which orders exactly like upstream's nested loops.
    double best = HUGE_VALF;
    int best_c = 0x7fffffff, best_pack = 0;
    {
        const double max_mse = P.max_line_fit_mse, max_dot = P.cos_critical_rad;
        uint32_t okw[4];
#pragma unroll
        for (int wd = 0; wd < 4; wd++) {
            const int i = wd * 32 + lane, ia = i / 10, ib = i - ia * 10;
            const bool ok = i < 100 && ia < nmax && ib < nmax && ia != ib && !(ptab[i * 6 + 5] > max_mse);
            okw[wd] = __ballot_sync(FULL_MASK, ok);
        }
        const unsigned long long ok_lo = okw[0] | ((unsigned long long)okw[1] << 32), ok_hi = okw[2] | ((unsigned long long)okw[3] << 32);
        auto pair_ok = [&](int a_, int b_) -> bool {
            const int i = a_ * 10 + b_;
            return ((i < 64 ? ok_lo >> i : ok_hi >> (i - 64)) & 1ull) != 0ull;
        };
        const int n1 = nmax - 1, ntrip = n1 * (n1 - 1) * (n1 - 2) / 6;   // triples with m2 <= nmax - 2
#pragma unroll 1
        for (int idx = tid; idx < ntrip; idx += T) {
            const int tr = c_qf_trips.v[idx];
            const int m0 = tr & 15, m1 = (tr >> 4) & 15, m2 = tr >> 8;
            if (!pair_ok(m0, m1) || !pair_ok(m1, m2)) continue;
            const double* e01 = ptab + (m0 * 10 + m1) * 6;
            const double* e12 = ptab + (m1 * 10 + m2) * 6;
            const double d = e01[2] * e12[2] + e01[3] * e12[3];
            if (fabs(d) > max_dot) continue;
            const double e012 = e01[4] + e12[4];
#pragma unroll 1
            for (int m3 = m2 + 1; m3 < nmax; m3++) {
                if (!pair_ok(m2, m3) || !pair_ok(m3, m0)) continue;
                const double err = e012 + ptab[(m2 * 10 + m3) * 6 + 4] + ptab[(m3 * 10 + m0) * 6 + 4];
                const int lex = (m0 << 12) | (m1 << 8) | (m2 << 4) | m3;
                if (err < best || (err == best && lex < best_c)) {
                    best = err;
                    best_c = lex;
                    best_pack = m0 | (m1 << 4) | (m2 << 8) | (m3 << 12);
                }
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ob = __shfl_xor_sync(FULL_MASK, best, off);
        int oc = __shfl_xor_sync(FULL_MASK, best_c, off);
        int op = __shfl_xor_sync(FULL_MASK, best_pack, off);
        if (ob < best || (ob == best && oc < best_c)) { best = ob; best_c = oc; best_pack = op; }
    }
    if (NW > 1) {
        if (lane == 0) { G.sd[G.w * 8] = best; G.si[G.w] = best_c; G.si[NW + 4 + G.w] = best_pack; }
        __syncthreads();
        best = G.sd[0]; best_c = G.si[0]; best_pack = G.si[NW + 4];
#pragma unroll
        for (int ww = 1; ww < NW; ww++) {
            const double ob = G.sd[ww * 8];
            const int oc = G.si[ww], op = G.si[NW + 4 + ww];
            if (ob < best || (ob == best && oc < best_c)) { best = ob; best_c = oc; best_pack = op; }
        }
        __syncthreads();
    }
    if (best_c == 0x7fffffff) return false;
    if (!(best / sz < P.max_line_fit_mse)) return false;
    // the rest is straight-line code that runs once per cluster (corners, area, angles): one warp is enough, the
    // others would only fetch the same instructions again (the caller lets thread 0 write the quad)
    if (NW > 1 && G.w != 0) return false;

    // ---- corners and gates (every thread computes the same values)
    int mi[4] = {best_pack & 15, (best_pack >> 4) & 15, (best_pack >> 8) & 15, (best_pack >> 12) & 15};
    double lines[4][4];
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
        const double* e = ptab + (mi[i] * 10 + mi[(i + 1) & 3]) * 6;
        if (e[5] > P.max_line_fit_mse) return false;
        lines[i][0] = e[0]; lines[i][1] = e[1]; lines[i][2] = e[2]; lines[i][3] = e[3];
    }
    float p[4][2];
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
        double A00 = lines[i][3], A01 = -lines[(i + 1) & 3][3];
        double A10 = -lines[i][2], A11 = lines[(i + 1) & 3][2];
        double B0 = -lines[i][0] + lines[(i + 1) & 3][0];
        double B1 = -lines[i][1] + lines[(i + 1) & 3][1];
        double det = A00 * A11 - A10 * A01;
        if (fabs(det) < 0.001) return false;
        double W00 = A11 / det, W01 = -A01 / det;
        double L0 = W00 * B0 + W01 * B1;
        p[i][0] = (float)(lines[i][0] + L0 * A00);
        p[i][1] = (float)(lines[i][1] + L0 * A10);
    }
    {
        double area = 0, length[3], pp;
#pragma unroll 1
        for (int i = 0; i < 3; i++) {
            int ia = i, ib = (i + 1) % 3;
            double ddx = p[ib][0] - p[ia][0], ddy = p[ib][1] - p[ia][1];
            length[i] = sqrt(ddx * ddx + ddy * ddy);
        }
        pp = (length[0] + length[1] + length[2]) / 2;
        area += sqrt(pp * (pp - length[0]) * (pp - length[1]) * (pp - length[2]));
        const int idxs[4] = {2, 3, 0, 2};
#pragma unroll 1
        for (int i = 0; i < 3; i++) {
            int ia = idxs[i], ib = idxs[i + 1];
            double ddx = p[ib][0] - p[ia][0], ddy = p[ib][1] - p[ia][1];
            length[i] = sqrt(ddx * ddx + ddy * ddy);
        }
        pp = (length[0] + length[1] + length[2]) / 2;
        area += sqrt(pp * (pp - length[0]) * (pp - length[1]) * (pp - length[2]));
        if (area < 0.95 * P.min_tag_width * P.min_tag_width) return false;
    }
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
        int i0 = i, i1 = (i + 1) & 3, i2 = (i + 2) & 3;
        double dx1 = p[i1][0] - p[i0][0], dy1 = p[i1][1] - p[i0][1];
        double dx2 = p[i2][0] - p[i1][0], dy2 = p[i2][1] - p[i1][1];
        double cos_dtheta = (dx1 * dx2 + dy1 * dy2) / sqrt((dx1 * dx1 + dy1 * dy1) * (dx2 * dx2 + dy2 * dy2));
        if ((cos_dtheta > P.cos_critical_rad || cos_dtheta < -P.cos_critical_rad) || dx1 * dy2 < dy1 * dx2) return false;
    }
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
        if (P.quad_decimate > 1) {
            q.p[i][0] = (p[i][0] - 0.5f) * P.quad_decimate + 0.5f;
            q.p[i][1] = (p[i][1] - 0.5f) * P.quad_decimate + 0.5f;
        } else {
            q.p[i][0] = p[i][0];
            q.p[i][1] = p[i][1];
        }
    }
    q.frame = ref.frame;
    q.reversed_border = reversed;
    {
        const uint32_t ck = a.pair_keys[(size_t)ref.frame * a.cap_keys + (uint32_t)(a.recs[seg] >> 32)];
        const uint32_t* f2 = a.dense2rep + (size_t)ref.frame * AGPU_MAX_DENSE;
        const uint32_t ra = f2[ck >> 16], rb = f2[ck & 0xffffu];
        q.key = ((unsigned long long)max(ra, rb) << 32) | min(ra, rb);
    }
    return true;
}

// Dynamic shared memory per GROUP: `wcap` u64 (sort buffer; from 1024 keys on it also hosts the 600-double pair
// table once the sort is over) + 16 ints, plus a separate pair table for the small tier.  QF_SMEM_CAP: the largest
// sort buffer one CTA can have (227 KB of shared memory per SM).
#define QF_SMEM_CAP 26000
__host__ __device__ inline size_t qf_stage_bytes(int nw) { return (size_t)6 * (nw * 32 + 2) * 8; }   // prefix-moment stage
__host__ __device__ inline size_t qf_smem_per_group(int wcap, int nw) {
    return (size_t)wcap * 8 + 64 + (size_t)nw * 512 + qf_stage_bytes(nw) + (wcap >= 1024 ? 0 : QF_PTAB_DOUBLES * 8);
}

// Persistent groups: group i takes clusters i, i + ngroups, ...   NW == 1: blockDim/32 warp groups per CTA;
// NW > 1: the CTA (NW warps) is the group.
#ifndef QF_MINB2
#define QF_MINB2 16
#endif
#ifndef QF_MINB4
#define QF_MINB4 8
#endif
template <int NW>
__global__ void __launch_bounds__(NW == 1 ? 256 : NW * 32, NW == 1 ? 4 : (NW == 2 ? QF_MINB2 : (NW == 4 ? QF_MINB4 : 2)))   // <= 64 registers
k_fit_quads(QuadFitArgs a, DevParams P, int wcap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_i[2 * NW + 8];
    __shared__ double s_d[NW * 8];
    __shared__ int s_ok;
    __shared__ long long s_red[NW * 10];   // the bounding-box / polarity reductions of the cluster in hand
    QGroup<NW> G;
    G.lane = threadIdx.x & 31;
    G.w = NW == 1 ? 0 : (threadIdx.x >> 5);
    G.tid = NW == 1 ? G.lane : threadIdx.x;
    G.si = s_i;
    G.sd = s_d;
    const int gi = NW == 1 ? (threadIdx.x >> 5) : 0;
    unsigned char* base = smem_raw + qf_smem_per_group(wcap, NW) * gi;
    unsigned long long* sbuf = reinterpret_cast<unsigned long long*>(base);
    int* sidx = reinterpret_cast<int*>(base + (size_t)wcap * 8);
    uint16_t* scnt = reinterpret_cast<uint16_t*>(base + (size_t)wcap * 8 + 64);
    double* stage = reinterpret_cast<double*>(base + (size_t)wcap * 8 + 64 + (size_t)NW * 512);
    double* ptab = wcap >= 1024 ? reinterpret_cast<double*>(base)
                                : reinterpret_cast<double*>(base + (size_t)wcap * 8 + 64 + (size_t)NW * 512 + qf_stage_bytes(NW));
    const int n = min(*a.list_count, a.list_cap);
    // clusters differ by two orders of magnitude in cost: groups claim them one at a time from a shared cursor
    for (;;) {
        int ci = 0;
        if (NW == 1) {
            if (G.lane == 0) ci = atomicAdd(a.cursor, 1);
            ci = __shfl_sync(FULL_MASK, ci, 0);
        } else {
            if (threadIdx.x == 0) s_ok = atomicAdd(a.cursor, 1);
            __syncthreads();
            ci = s_ok;
            __syncthreads();
        }
        if (ci >= n) break;
        const ClusterRef ref = a.list[ci];
        QuadRec q;
        unsigned long long* sb = sbuf;
        if (NW == 8 && ref.size > wcap) sb = a.gsort + (size_t)blockIdx.x * a.gsort_stride;   // (the pair table stays in shared memory)
        double* lf = a.scratch + (size_t)(NW == 1 ? blockIdx.x * 8 + gi : blockIdx.x) * a.scratch_pts * 7;
        const bool ok = fit_cluster_group<NW>(G, a, P, ref, sb, ptab, sidx, scnt, stage, lf, s_red, q);
        if (ok && G.tid == 0) {
            int s = atomicAdd(a.nquads, 1);
            atomicAdd(&a.per_frame_quads[ref.frame], 1);
            if (s < a.cap_quads) a.quads[s] = q;
        }
        G.sync();
    }
}
