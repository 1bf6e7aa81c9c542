// aprilgpu.cu -- host side of libaprilgpu.so: the C ABI declared in include/aprilgpu.h, workspace
// management and the per-chunk kernel pipeline.  sm_100a only; no CPU fallback of any stage.
//
// Pipeline per chunk of frames (all on one stream):
//   [H2D] -> image (k_pack / k_decimate_blur / k_decimate_threshold) -> CC (k_cc_local, k_cc_boundary,
//   k_cc_finalize) -> k_edges (records carry cluster ids) -> k_cluster_refs -> k_sort_scatter (one-sweep segmented
//   sort by cluster id) -> k_fit_quads (4 size tiers) -> k_decode_quads -> k_reconcile [-> k_pose] -> D2H
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/aprilgpu.h"
#include "common.cuh"
#include "k_image.cuh"
#include "k_cc.cuh"
#include "k_cluster.cuh"
#include "k_quad.cuh"
#include "k_decode.cuh"
#include "k_pose.cuh"
#include "k_render.cuh"
#include "k_graph.cuh"

#include "families_data.inc"

static_assert(sizeof(DetRec) == sizeof(agpu_detection), "DetRec layout");
static_assert(sizeof(PoseRec) == sizeof(agpu_pose_t), "PoseRec layout");

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct HostBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMallocHost(&p, need);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// cluster size tiers of the quad-fit kernel (largest cluster per tier; 1 / 2 / 4 / 8 warps per cluster: tier 0 packs
// eight one-warp groups into a CTA, the others use one CTA per cluster).  The last tier takes everything up to
// upstream's own size limit, 3(2w+2h) points (set per call from the frame geometry), so no cluster upstream would
// fit is ever skipped.
constexpr int TIER_CAP_DEFAULT[AGPU_NTIERS] = {256, 1024, 2048, 6144, 0};
// (warps per cluster: 1 -- eight one-warp groups per CTA --, 2, 4, 4, 8.  The fourth tier (2049..6144 records: three CTAs
// of four warps per SM) exists for noisy frames, whose background contours would otherwise all queue up behind the
// one-CTA-per-SM last tier.)

}  // namespace

// One pipeline slot: the workspaces, streams and pinned result buffers of ONE chunk in flight.  Several
// slots run concurrently (each on its own stream) so that the latency-bound stages of one chunk (quad
// fitting, decode, reconcile, pose, the radix-sort scans) overlap the bandwidth-bound stages of another.
struct Slot {
    cudaStream_t stream = nullptr;     // dense, bandwidth-bound front half: image, CC, edge points, radix sort
    cudaStream_t tail = nullptr;       // latency-bound back half (quad fit, decode, reconcile, pose, D2H): HIGH priority
    cudaStream_t aux[AGPU_NTIERS - 1] = {nullptr, nullptr, nullptr, nullptr};   // side streams: quad-fit tiers run concurrently
    cudaEvent_t ev_mid = nullptr, ev_fork = nullptr, ev_join[AGPU_NTIERS - 1] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> events;   // stage timing
    struct KEv { const char* name; cudaEvent_t a, b; };
    std::vector<KEv> kev;              // profiling only: one event pair around EVERY kernel launch of the chunk
    size_t kev_next = 0;
    DevBuf d_in, d_gray, d_quad_im, d_thresh, d_masks, d_l16, d_canon, d_canon_sizes, d_roottab, d_tilebase, d_dense2rep;
    DevBuf d_recs[2], d_qscratch, d_gsort, d_pairslots, d_pairkeys, d_paircount, d_pairstart;
    DevBuf d_counters;  // CNT_FIXED ints + per-frame: npts[chunk], frame_quads[chunk], ndets[chunk], out_counts[chunk], ndense[chunk], nroots[16*chunk], ndups[chunk], ncl[chunk]
    DevBuf d_clusters[AGPU_NTIERS], d_dbg_heads, d_quads, d_refined, d_dets, d_out, d_poses;
    HostBuf h_out, h_counts, h_poses;
    bool pending = false, from_masks = false;
    int b0 = 0, n = 0, sorted = 0;
    // CUDA graph of the whole per-chunk pipeline for the synchronous small-batch path (a webcam loop calls detect once
    // per frame: ~30 launches and a dozen event calls per call otherwise).  Captured the second time a call with the
    // same key arrives; any change of geometry, capacities, pose parameters or buffer addresses re-captures.
    struct GraphKey {
        int W, H, stride, channels, n, cap, maxcl, maxq, cap_out, cap_keys, roots_cap, pose;
        double K[9], dist[8], tag_size;
        int ndist;
        unsigned long long buffers;   // hash of every buffer address the kernels were given
    };
    GraphKey graph_key, seen_key;
    bool have_seen = false, graph_launched = false;
    cudaGraphExec_t graph_exec = nullptr;
    cudaEvent_t ev_end = nullptr;
    long long graph_launches = 0;
    int graph_sorted = 0;
    bool graph_from_masks = false;

    void release() {
        DevBuf* bufs[] = {&d_in, &d_gray, &d_quad_im, &d_thresh, &d_masks, &d_l16, &d_canon, &d_canon_sizes, &d_roottab,
                          &d_tilebase, &d_dense2rep, &d_recs[0], &d_recs[1], &d_qscratch, &d_gsort, &d_pairslots, &d_pairkeys, &d_paircount, &d_pairstart, &d_counters,
                          &d_clusters[0], &d_clusters[1], &d_clusters[2], &d_clusters[3], &d_clusters[4], &d_dbg_heads, &d_quads,
                          &d_refined, &d_dets, &d_out, &d_poses};
        for (DevBuf* bb : bufs) bb->release();
        h_out.release(); h_counts.release(); h_poses.release();
        for (cudaEvent_t e : events) cudaEventDestroy(e);
        events.clear();
        for (KEv& k : kev) { cudaEventDestroy(k.a); cudaEventDestroy(k.b); }
        kev.clear();
        for (int t = 0; t < AGPU_NTIERS - 1; t++) {
            if (aux[t]) cudaStreamDestroy(aux[t]);
            if (ev_join[t]) cudaEventDestroy(ev_join[t]);
            aux[t] = nullptr; ev_join[t] = nullptr;
        }
        if (graph_exec) cudaGraphExecDestroy(graph_exec);
        graph_exec = nullptr;
        if (ev_end) cudaEventDestroy(ev_end);
        ev_end = nullptr;
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_mid) cudaEventDestroy(ev_mid);
        if (tail) cudaStreamDestroy(tail);
        if (stream) cudaStreamDestroy(stream);
        ev_fork = nullptr; ev_mid = nullptr; stream = nullptr; tail = nullptr;
    }
};

struct agpu_handle {
    agpu_config cfg;
    std::string families_str;
    std::vector<const FamilyDef*> fams;
    DevParams prm;
    int device = 0;
    int num_sms = 148;
    std::string err;
    bool profiling = false;
    float stage_ms[AGPU_NUM_STAGES];
    std::map<std::string, std::pair<float, int>> kernel_ms;   // profiling: per kernel {ms, launches} of the last call
    long long launches = 0;
    long long counters[8];
    long long tier_stats[8];   // [0..3] clusters, [4..7] records handed to the quad-fit tiers in the last call

    // scheduling knobs (defaults below; AGPU_PRIO / AGPU_TIER_CTAS / AGPU_DECODE_CTAS override them for experiments)
    struct Tune {
        int prio = 1;                                  // back half of a chunk on high-priority streams
        int tier_ctas[AGPU_NTIERS] = {3, 16, 8, 3, 1};    // persistent quad-fit CTAs per SM, by size tier
        int tier_cap[AGPU_NTIERS] = {TIER_CAP_DEFAULT[0], TIER_CAP_DEFAULT[1], TIER_CAP_DEFAULT[2], TIER_CAP_DEFAULT[3],
                                     TIER_CAP_DEFAULT[4]};
        int decode_ctas = 4;                           // persistent decode CTAs (of 4 warps) per SM
        int edge_warps = 2, boundary_warps = 8;        // tiles (warps) per CTA of k_edges / k_cc_boundary
        int masks = 0;                                 // 1 (AGPU_MASKS=1), decimate 1: the threshold kernel writes the CC bit masks instead of
                                                       // threshold bytes -- 0.75 N less HBM traffic each way, but both kernels are issue-bound:
                                                       // measured +-0 on the whole pipeline, so the simpler byte image stays the default
        int tail_threads = 32;                         // CTA size of decode / reconcile / pose: small CTAs find room on SMs
                                                       // that the streaming kernels of the next chunk keep full
        int graph = 1;                                 // CUDA graph for single-chunk calls of up to graph_max_frames frames
        int graph_max_frames = 8;
        int seg_tiles = 0, img_minb = 4, slots = 0;    // AGPU_SEG_TILES / AGPU_IMG_MINB / AGPU_SLOTS (0 = default), read once
        int blur_strip = 1;                            // AGPU_BLUR_STRIP=0: the shared-memory blur kernel for every kernel size (tests)
    } tune;
    DevBuf d_fams, d_codes, d_pose_in, d_pose_out;
    std::vector<Slot> slots;
    cudaEvent_t ev_user = nullptr;
    cudaEvent_t ev_t0 = nullptr;          // profiling: start of the call, origin of the timeline
    std::vector<float> timeline;          // profiling: per finished chunk {b0, n, slot, marks[AGPU_NUM_STAGES + 1]} in ms since ev_t0

    // growable per-frame list capacities (0 = not chosen yet)
    int cap_points = 0, cap_clusters = 0, cap_quads = 0;
    int max_dense_seen = -1, max_clusters_seen = -1;   // most components / clusters seen in one frame so far (statistics)
    int cap_keys = 0;   // cluster-id capacity per frame (grows like the other work lists)
    int cap_roots = 0;  // tile-local roots per sub-list (likewise)
    int roots_hint = 0; // longest root sub-list of the last finished chunk (sizes the grids of k_cc_sizes / k_cc_dense)

    // state of the last finished chunk (debug fetch)
    Geom geom;
    int last_slot = 0, last_chunk = 0, last_cap = 0, last_cap_keys = 0;
    bool have_last = false, last_from_masks = false, last_bgr = false;

    void set_err(const std::string& s) { err = s; }
};

// Tag graphs of S independent camera streams, resident in HBM between agpu_graph_update calls.
struct agpu_graph {
    agpu_handle* h = nullptr;
    int S = 0, nid = 0;
    DevBuf d_coord, d_est, d_present, d_updated, d_visible, d_reference, d_weight, d_local, d_world, d_skipped;
    DevBuf d_dets, d_poses, d_counts, d_my_pose, d_valid;
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            h->set_err(std::string(#call) + ": " + cudaGetErrorString(e__));                      \
            return AGPU_E_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

#define LAUNCH_CHECK(name)                                                                         \
    do {                                                                                           \
        h->launches++;                                                                             \
        cudaError_t e__ = cudaGetLastError();                                                      \
        if (e__ != cudaSuccess) {                                                                  \
            h->set_err(std::string("launch ") + name + ": " + cudaGetErrorString(e__));           \
            return AGPU_E_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

const FamilyDef* find_family(const std::string& name) {
    for (const FamilyDef& f : k_builtin_families)
        if (name == f.name) return &f;
    return nullptr;
}

int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

Geom make_geom(int W, int H, int f) {
    Geom g;
    g.W = W; g.H = H;
    g.wd = 1 + (W - 1) / f;
    g.hd = 1 + (H - 1) / f;
    g.wp = (g.wd + 15) & ~15;
    g.plane = (size_t)g.wp * g.hd;
    return g;
}

int gaussian_kernel_host(float sigma, uint8_t* k) {
    int ksz = (int)(4 * sigma);
    if ((ksz & 1) == 0) ksz++;
    if (ksz <= 1) return 0;
    if (ksz > 63) ksz = 63;
    double dk[64], acc = 0;
    for (int i = 0; i < ksz; i++) {
        int x = -ksz / 2 + i;
        dk[i] = std::exp(-.5 * (x / (double)sigma) * (x / (double)sigma));
        acc += dk[i];
    }
    for (int i = 0; i < ksz; i++) k[i] = (uint8_t)(dk[i] / acc * 255.0);
    return ksz;
}

int init_slot(agpu_handle* h, Slot& s) {
    if (s.stream) return AGPU_OK;
    // The back half of a chunk is a handful of persistent, latency-bound kernels; the front half of the NEXT chunks
    // are huge grids of short CTAs.  Giving the back half the higher stream priority lets its CTAs take the SM
    // slots that the streaming kernels free all the time, so both kinds of work share every SM instead of
    // alternating kernel by kernel.
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    if (!h->tune.prio) prio_hi = prio_lo;
    CK(cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, prio_lo));
    CK(cudaStreamCreateWithPriority(&s.tail, cudaStreamNonBlocking, prio_hi));
    for (int t = 0; t < AGPU_NTIERS - 1; t++) {
        CK(cudaStreamCreateWithPriority(&s.aux[t], cudaStreamNonBlocking, prio_hi));
        CK(cudaEventCreateWithFlags(&s.ev_join[t], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s.ev_mid, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s.ev_end, cudaEventDisableTiming));
    return AGPU_OK;
}

// Profiling only: CUDA events right before and after one kernel launch on the stream it is launched on.  With one
// chunk in flight (and the quad-fit tiers serialised, see launch_chunk) the interval is that kernel alone.
struct KScope {
    agpu_handle* h;
    Slot& s;
    cudaStream_t st;
    Slot::KEv* e = nullptr;
    KScope(agpu_handle* hh, Slot& ss, const char* name, cudaStream_t stream) : h(hh), s(ss), st(stream) {
        if (!h->profiling) return;
        if (s.kev_next >= s.kev.size()) {
            Slot::KEv k{name, nullptr, nullptr};
            if (cudaEventCreate(&k.a) != cudaSuccess || cudaEventCreate(&k.b) != cudaSuccess) return;
            s.kev.push_back(k);
        }
        e = &s.kev[s.kev_next++];
        e->name = name;
        cudaEventRecord(e->a, st);
    }
    ~KScope() { if (e) cudaEventRecord(e->b, st); }
};

struct StageTimer {
    agpu_handle* h;
    Slot& s;
    size_t next = 0;
    StageTimer(agpu_handle* hh, Slot& ss) : h(hh), s(ss) {}
    void mark(cudaStream_t st = nullptr) {
        if (!h->profiling) return;
        if (next >= s.events.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            s.events.push_back(e);
        }
        cudaEventRecord(s.events[next++], st ? st : s.stream);
    }
};

// ------------------------------------------------------------------------------------------
// image + CC stages on `n` frames already on the device (shared by the pipeline and the stage hooks)
// ------------------------------------------------------------------------------------------
int run_image_stage(agpu_handle* h, Slot& sl, const uint8_t* d_src, int channels, int W, int H, size_t stride,
                    size_t frame_stride, int n, const Geom& g, const uint8_t** quad_im_out, size_t* q_pitch,
                    size_t* q_frame, const uint8_t** gray_full, size_t* gray_pitch, size_t* gray_frame,
                    bool* masks_written = nullptr) {
    // masks_written != null: the caller's next step is the CC pass and it can take the tile-major bit masks instead
    // of the threshold bytes (pipeline, decimate 1, no stage dumps)
    if (masks_written) *masks_written = false;
    const int f = h->prm.decim;
    const float sigma = h->cfg.quad_sigma;
    const uint8_t* src = d_src;
    size_t s_stride = stride, s_frame = frame_stride;
    int srcW = W, srcH = H;
    int F = f;
    // full-resolution gray image for refine_edges / decode.  BGR frames (the reference's input, tag_detector.py:25) at
    // decimate 1 / 2 / 4 without blur: the strip kernel converts, decimates and thresholds in one pass (`fused_bgr`);
    // other settings convert first (k_pack) and continue on the gray plane
    const bool fused_bgr = channels == 3 && (f == 1 || f == 2 || f == 4) && sigma == 0.0f && (g.wd >> 2) > 0 && (g.hd >> 2) > 0;
    const Geom gfull = make_geom(W, H, 1);
    if (fused_bgr) {
        CK(sl.d_gray.ensure(gfull.plane * n));
        *gray_full = sl.d_gray.as<uint8_t>();
        *gray_pitch = gfull.wp;
        *gray_frame = gfull.plane;
    } else if (channels == 3) {
        Geom gf = make_geom(W, H, 1);
        CK(sl.d_gray.ensure(gf.plane * n));
        size_t total = (size_t)n * gf.hd * (gf.wp >> 2);
        {
            KScope ks(h, sl, "k_pack(bgr)", sl.stream);
            k_pack<<<ceil_div(total, 256), 256, 0, sl.stream>>>(d_src, W, H, stride, frame_stride, 3, 1,
                                                                sl.d_gray.as<uint8_t>(), gf, n);
        }
        LAUNCH_CHECK("k_pack(bgr)");
        *gray_full = sl.d_gray.as<uint8_t>();
        *gray_pitch = gf.wp;
        *gray_frame = gf.plane;
        src = sl.d_gray.as<uint8_t>();
        s_stride = gf.wp;
        s_frame = gf.plane;
        channels = 1;
    } else {
        *gray_full = d_src;
        *gray_pitch = stride;
        *gray_frame = frame_stride;
    }
    const bool fast = (f == 1 || f == 2 || f == 4) && sigma == 0.0f;
    uint8_t* quad_im = nullptr;
    if (!fast) {
        // generic: decimate (+ blur) into quad_im, then threshold the quad image with F = 1
        CK(sl.d_quad_im.ensure(g.plane * n));
        quad_im = sl.d_quad_im.as<uint8_t>();
        BlurKernel bk;
        memset(&bk, 0, sizeof(bk));
        if (sigma != 0.0f) bk.ksz = gaussian_kernel_host(std::fabs(sigma), bk.k);
        if (bk.ksz > 1) {
            // U1 + U2 in one kernel: a shared-memory tile with its halo, 128-bit loads, row and column pass on chip
            const int vec_in = (s_stride % 16 == 0) && (s_frame % 16 == 0) && (((uintptr_t)src) % 16 == 0);
            dim3 grid(ceil_div(g.wp, BL_TW), ceil_div(g.hd, BL_TH), n);
            const size_t smem = bl_smem_bytes(bk.ksz);
            KScope ks(h, sl, "k_decimate_blur", sl.stream);
            // (7 taps at decimation 2: the ring of rows needs ~190 registers and the tile kernel is faster)
            const bool strip = vec_in && (f == 1 || f == 2) && (bk.ksz == 3 || bk.ksz == 5 || (bk.ksz == 7 && f == 1)) && h->tune.blur_strip;
            if (strip) {
                // the common case: register-resident strips, no shared memory
                const int nstrips = ceil_div(g.wp, 512), nsegs = ceil_div(g.hd, BLS_ROWS);
                const int blocks = (int)ceil_div((long long)nstrips * nsegs * n, 4);
#define LAUNCH_BLS(FF, KK, SS) k_decimate_blur_strip<FF, KK, SS><<<blocks, 128, 0, sl.stream>>>(src, s_stride, s_frame, quad_im, g, bk, nstrips, nsegs, n)
#define LAUNCH_BLS_K(FF, SS) do { if (bk.ksz == 3) LAUNCH_BLS(FF, 3, SS); else if (bk.ksz == 5) LAUNCH_BLS(FF, 5, SS); else LAUNCH_BLS(FF, 7, SS); } while (0)
                if (f == 1) { if (sigma < 0) LAUNCH_BLS_K(1, true); else LAUNCH_BLS_K(1, false); }
                else { if (sigma < 0) LAUNCH_BLS_K(2, true); else LAUNCH_BLS_K(2, false); }
#undef LAUNCH_BLS_K
#undef LAUNCH_BLS
            } else if (f == 1) k_decimate_blur<1><<<grid, 256, smem, sl.stream>>>(src, s_stride, s_frame, 1, vec_in, quad_im, g, bk, sigma < 0);
            else if (f == 2) k_decimate_blur<2><<<grid, 256, smem, sl.stream>>>(src, s_stride, s_frame, 2, vec_in, quad_im, g, bk, sigma < 0);
            else k_decimate_blur<0><<<grid, 256, smem, sl.stream>>>(src, s_stride, s_frame, f, 0, quad_im, g, bk, sigma < 0);
            LAUNCH_CHECK("k_decimate_blur");
        } else {
            size_t total = (size_t)n * g.hd * (g.wp >> 2);
            k_pack<<<ceil_div(total, 256), 256, 0, sl.stream>>>(src, srcW, srcH, s_stride, s_frame, 1, f, quad_im, g, n);
            LAUNCH_CHECK("k_pack");
        }
        src = quad_im;
        s_stride = g.wp;
        s_frame = g.plane;
        srcW = g.wd;
        srcH = g.hd;
        F = 1;
    }
    uint8_t* quad_out = nullptr;
    if (F > 1) {
        CK(sl.d_quad_im.ensure(g.plane * n));
        quad_out = sl.d_quad_im.as<uint8_t>();
    }
    if ((g.wd >> 2) == 0 || (g.hd >> 2) == 0) {
        CK(sl.d_thresh.ensure(g.plane * n));
        CK(cudaMemsetAsync(sl.d_thresh.p, 127, g.plane * n, sl.stream));
    } else {
        const int TPL = 4 / F;
        const int strip_px = 30 * TPL * 4;
        const int nstrips = ceil_div(g.wp, strip_px);
        const int th = g.hd >> 2;
        int seg_tiles = fused_bgr ? 16 : 8;   // (a segment re-converts its two halo tile rows: longer segments for BGR)
        if (h->tune.seg_tiles > 0) seg_tiles = h->tune.seg_tiles;
        const int nsegs = ceil_div(th, seg_tiles);
        const long long warps = (long long)n * nstrips * nsegs;
        const int blocks = ceil_div(warps * 32, 128);
        const int vec_ok = (s_stride % 16 == 0) && (s_frame % 16 == 0) && (((uintptr_t)src) % 16 == 0);
        const int md = h->prm.min_white_black_diff;
        const bool use_masks = masks_written && F == 1 && h->tune.masks && !fused_bgr;
        if (!use_masks) CK(sl.d_thresh.ensure(g.plane * n));
        if (use_masks) CK(sl.d_masks.ensure((size_t)cc_tiles_x(g) * cc_tiles_y(g) * n * 32 * sizeof(uint2)));
        {
        KScope ks(h, sl, "k_decimate_threshold", sl.stream);
        uint8_t* th_out = sl.d_thresh.as<uint8_t>();
        const int minb = h->tune.img_minb;
#define LAUNCH_DT(FF, MB) k_decimate_threshold<FF, MB><<<blocks, 128, 0, sl.stream>>>(src, srcW, srcH, s_stride, s_frame, quad_out, \
                                                                               th_out, g, nstrips, nsegs, seg_tiles, n, md, vec_ok)
        if (fused_bgr) {
#define LAUNCH_DT3(FF) k_decimate_threshold<FF, 3, false, 3><<<blocks, 128, 0, sl.stream>>>(src, srcW, srcH, s_stride, s_frame, quad_out, \
                                                              th_out, g, nstrips, nsegs, seg_tiles, n, md, vec_ok, nullptr,                \
                                                              sl.d_gray.as<uint8_t>(), (size_t)gfull.wp, gfull.plane)
            if (F == 1) LAUNCH_DT3(1); else if (F == 2) LAUNCH_DT3(2); else LAUNCH_DT3(4);
#undef LAUNCH_DT3
        } else if (use_masks) {
            k_decimate_threshold<1, 4, true><<<blocks, 128, 0, sl.stream>>>(src, srcW, srcH, s_stride, s_frame, nullptr, nullptr, g,
                                                                          nstrips, nsegs, seg_tiles, n, md, vec_ok,
                                                                          sl.d_masks.as<uint2>());
            *masks_written = true;
        } else if (F == 1) {
            if (minb == 4) LAUNCH_DT(1, 4); else if (minb == 5) LAUNCH_DT(1, 5); else if (minb == 6) LAUNCH_DT(1, 6); else LAUNCH_DT(1, 3);
        } else if (F == 2) {
            LAUNCH_DT(2, 4);
        } else {
            LAUNCH_DT(4, 4);
        }
#undef LAUNCH_DT
        LAUNCH_CHECK("k_decimate_threshold");
        }
    }
    if (F > 1) {
        *quad_im_out = quad_out; *q_pitch = g.wp; *q_frame = g.plane;
    } else if (fused_bgr) {
        *quad_im_out = sl.d_gray.as<uint8_t>(); *q_pitch = gfull.wp; *q_frame = gfull.plane;   // decimate 1: the gray plane
    } else {
        *quad_im_out = src; *q_pitch = s_stride; *q_frame = s_frame;
    }
    return AGPU_OK;
}

// root tables of `n` frames in one allocation (five u32 arrays of CC_SUBLISTS * sub_cap entries per frame)
int make_root_tables(agpu_handle* h, Slot& sl, int n, const Geom& g, int sub_cap, int* d_nroots, CcRoots& rt) {
    const int tx = cc_tiles_x(g), ty = cc_tiles_y(g);
    const size_t per = (size_t)CC_SUBLISTS * sub_cap * n;
    CK(sl.d_roottab.ensure(per * 5 * 4));
    CK(sl.d_tilebase.ensure((size_t)tx * ty * n * 4));
    uint32_t* p = sl.d_roottab.as<uint32_t>();
    rt.links = p; rt.sizes = p + per; rt.rootpix = p + 2 * per; rt.minpix = p + 3 * per; rt.dense = p + 4 * per;
    rt.tile_base = sl.d_tilebase.as<uint32_t>();
    rt.nroots = d_nroots;
    rt.sub_cap = sub_cap;
    rt.ntiles = tx * ty;
    return AGPU_OK;
}

int default_roots_cap(const Geom& g) {   // tile-local roots per sub-list: a sixteenth of the pixels per frame in all
    return std::max(1024, (int)(g.plane / 256));
}

int run_cc_stage(agpu_handle* h, Slot& sl, const uint8_t* d_thresh, int n, const Geom& g, int sub_cap, int* d_nroots,
                 int* d_ndense, bool canonical, CcRoots& rt, bool from_masks = false) {
    const int tx = cc_tiles_x(g), ty = cc_tiles_y(g);
    CK(sl.d_masks.ensure((size_t)tx * ty * n * 32 * sizeof(uint2)));
    CK(sl.d_l16.ensure((size_t)tx * ty * n * 1024 * sizeof(uint16_t)));
    int rc = make_root_tables(h, sl, n, g, sub_cap, d_nroots, rt);
    if (rc) return rc;
    dim3 grid(ceil_div(tx, CC_WARPS), ty, n);
    {
    KScope ks(h, sl, "k_cc_local", sl.stream);
    if (from_masks)
        k_cc_local<true><<<grid, CC_THREADS, 0, sl.stream>>>(nullptr, sl.d_masks.as<uint2>(), sl.d_l16.as<uint16_t>(), rt, g);
    else
        k_cc_local<false><<<grid, CC_THREADS, 0, sl.stream>>>(d_thresh, sl.d_masks.as<uint2>(), sl.d_l16.as<uint16_t>(), rt, g);
    }
    LAUNCH_CHECK("k_cc_local");
    {
        KScope ks(h, sl, "k_cc_boundary", sl.stream);
        const int bw = h->tune.boundary_warps;
        dim3 gridb(ceil_div(tx * ty, bw), 1, n);
#define LAUNCH_CCB(BW) k_cc_boundary<BW><<<gridb, BW * 32, 0, sl.stream>>>(sl.d_masks.as<uint2>(), sl.d_l16.as<uint16_t>(), rt, g)
        if (bw == 1) LAUNCH_CCB(1); else if (bw == 2) LAUNCH_CCB(2); else if (bw == 4) LAUNCH_CCB(4); else LAUNCH_CCB(8);
#undef LAUNCH_CCB
    }
    LAUNCH_CHECK("k_cc_boundary");
    // (grid-stride loops over a sub-list: any CTA count is correct; the count follows the longest sub-list the previous
    // chunk had, so that a clean frame's ~400 roots per sub-list do not launch eight CTAs of which six find nothing)
    const int roots_seen = h->roots_hint > 0 ? std::min(h->roots_hint, sub_cap) : sub_cap;
    dim3 grids(std::max(1, std::min(8, ceil_div(roots_seen, 256))), n * CC_SUBLISTS);
    {
        KScope ks(h, sl, "k_cc_sizes", sl.stream);
        k_cc_sizes<<<grids, 256, 0, sl.stream>>>(rt);
    }
    LAUNCH_CHECK("k_cc_sizes");
    if (d_ndense) {
        CK(sl.d_dense2rep.ensure((size_t)n * AGPU_MAX_DENSE * 4));
        {
            KScope ks(h, sl, "k_cc_dense", sl.stream);
            k_cc_dense<<<grids, 256, 0, sl.stream>>>(rt, sl.d_dense2rep.as<uint32_t>(), d_ndense);
        }
        LAUNCH_CHECK("k_cc_dense");
    }
    if (canonical) {   // stage dumps only
        CK(sl.d_canon.ensure(g.plane * n * 4));
        CK(sl.d_canon_sizes.ensure(g.plane * n * 4));
        dim3 gridf(ceil_div(g.wd, 32), ceil_div(g.hd, 8), n);
        k_cc_canonical<<<gridf, 256, 0, sl.stream>>>(sl.d_masks.as<uint2>(), sl.d_l16.as<uint16_t>(), rt, sl.d_canon.as<uint32_t>(),
                                                     sl.d_canon_sizes.as<uint32_t>(), g);
        LAUNCH_CHECK("k_cc_canonical");
    }
    return AGPU_OK;
}

struct PoseSpec {
    bool enabled = false;
    double K[9];
    double dist[8];
    int ndist = 0;
    double tag_size = 0;
};

void fill_pose_args(PoseArgs& pa, const PoseSpec& ps, int method) {
    pa.fx = ps.K[0]; pa.fy = ps.K[4]; pa.cx = ps.K[2]; pa.cy = ps.K[5];
    for (int i = 0; i < 8; i++) pa.dist[i] = ps.dist[i];
    pa.ndist = ps.ndist;
    pa.half = (double)(float)(ps.tag_size / 2);
    pa.method = method;
}

int parse_dist(agpu_handle* h, const double* dist, int ndist, PoseSpec& ps) {
    for (int i = 0; i < 8; i++) ps.dist[i] = 0;
    if (ndist < 0 || ndist > 8 || (ndist > 0 && !dist)) {
        h->set_err("dist: expected 0, 4, 5 or 8 coefficients");
        return AGPU_E_INVALID;
    }
    bool any = false;
    for (int i = 0; i < ndist; i++) {
        ps.dist[i] = dist[i];
        any |= dist[i] != 0.0;
    }
    ps.ndist = any ? ndist : 0;
    return AGPU_OK;
}

// ------------------------------------------------------------------------------------------
// the pipeline
// ------------------------------------------------------------------------------------------
struct CallCtx {   // constants of one agpu_detect* call
    const uint8_t* frames;
    int on_device, channels, B, W, H, stride;
    size_t frame_bytes;
    Geom g;
    int chunk, cap, maxcl, maxq, cap_out, key_bits;
    int cap_keys;     // cluster ids per frame the pair table can hand out
    int roots_cap;    // tile-local roots per sub-list of a frame (CC_SUBLISTS sub-lists)
    size_t ncnt;
    const PoseSpec* pose;
};

int alloc_slot(agpu_handle* h, Slot& s, const CallCtx& c) {
    int rc = init_slot(h, s);
    if (rc) return rc;
    const int chunk = c.chunk, cap = c.cap;
    if (!c.on_device) CK(s.d_in.ensure(c.frame_bytes * chunk));
    for (int i = 0; i < 2; i++) CK(s.d_recs[i].ensure((size_t)chunk * cap * 8));
    CK(s.d_dense2rep.ensure((size_t)chunk * AGPU_MAX_DENSE * 4));
    CK(s.d_pairkeys.ensure((size_t)chunk * c.cap_keys * 4));
    CK(s.d_pairslots.ensure((size_t)chunk * c.cap_keys * 4 * 8));   // open addressing at <= 25 % load
    CK(s.d_paircount.ensure((size_t)chunk * c.cap_keys * 4));
    CK(s.d_pairstart.ensure((size_t)chunk * c.cap_keys * 4));
    {   // quad-fit scratch: 56 bytes per point of the largest cluster of its tier for every persistent GROUP (re-used
        // cluster after cluster, so it lives in L2), not per edge point of the chunk
        const int max_cluster = 3 * (2 * c.g.wd + 2 * c.g.hd);
        size_t pts = 0;
        for (int t = 0; t < AGPU_NTIERS; t++) {
            const int cap_t = t == AGPU_NTIERS - 1 ? std::max(max_cluster, h->tune.tier_cap[t - 1] + 1) : h->tune.tier_cap[t];
            pts += (size_t)h->num_sms * h->tune.tier_ctas[t] * (t == 0 ? 8 : 1) * ((cap_t + 1) & ~1);   // (even: 16-byte aligned slices)
        }
        CK(s.d_qscratch.ensure(pts * 56));
    }
    {   // frames so large that a cluster of upstream's maximum size does not fit a CTA's shared memory (4K at decimate 1):
        // the last tier then sorts such clusters in a per-CTA global buffer
        const int max_cluster = 3 * (2 * c.g.wd + 2 * c.g.hd);
        if (max_cluster > QF_SMEM_CAP)
            CK(s.d_gsort.ensure((size_t)h->num_sms * h->tune.tier_ctas[AGPU_NTIERS - 1] * max_cluster * 8));
    }
    CK(s.d_counters.ensure(c.ncnt * 4));
    for (int t = 0; t < AGPU_NTIERS; t++) CK(s.d_clusters[t].ensure((size_t)chunk * c.maxcl * sizeof(ClusterRef)));
    if (h->cfg.debug) {
        CK(s.d_dbg_heads.ensure((size_t)chunk * cap / 4 * sizeof(ClusterRef)));
        CK(s.d_refined.ensure((size_t)chunk * c.maxq * 32));
    }
    CK(s.d_quads.ensure((size_t)chunk * c.maxq * sizeof(QuadRec)));
    CK(s.d_dets.ensure((size_t)chunk * REC_CAP * sizeof(DetRec)));
    CK(s.d_out.ensure((size_t)chunk * c.cap_out * sizeof(DetRec)));
    CK(s.h_out.ensure((size_t)chunk * c.cap_out * sizeof(DetRec)));
    CK(s.h_counts.ensure(c.ncnt * 4));
    if (c.pose->enabled) {
        CK(s.d_poses.ensure((size_t)chunk * c.cap_out * sizeof(PoseRec)));
        CK(s.h_poses.ensure((size_t)chunk * c.cap_out * sizeof(PoseRec)));
    }
    return AGPU_OK;
}

// enqueue the whole pipeline of frames [b0, b0+n) on the slot's stream (no host synchronisation)
int enqueue_chunk(agpu_handle* h, Slot& sl, const CallCtx& c, int b0, int n) {
    const Geom& g = c.g;
    const int chunk = c.chunk, cap = c.cap;
    int* d_cnt = sl.d_counters.as<int>();
    int* d_npts = d_cnt + CNT_FIXED;
    int* d_frame_quads = d_npts + chunk;
    int* d_ndets = d_frame_quads + chunk;
    int* d_out_counts = d_ndets + chunk;
    int* d_ndense = d_out_counts + chunk;
    int* d_nroots = d_ndense + chunk;          // CC_SUBLISTS counters per frame
    int* d_ndups = d_nroots + CC_SUBLISTS * chunk;   // merged duplicate points per frame (raw points = npts + ndups)
    int* d_ncl = d_ndups + chunk;                     // cluster ids handed out per frame
    StageTimer tm(h, sl);
    sl.kev_next = 0;
    tm.mark();  // 0
    const uint8_t* d_src;
    if (c.on_device) {
        d_src = c.frames + (size_t)b0 * c.frame_bytes;
    } else {
        CK(cudaMemcpyAsync(sl.d_in.p, c.frames + (size_t)b0 * c.frame_bytes, c.frame_bytes * n, cudaMemcpyHostToDevice,
                           sl.stream));
        d_src = sl.d_in.as<uint8_t>();
    }
    CK(cudaMemsetAsync(d_cnt, 0, c.ncnt * 4, sl.stream));
    CK(cudaMemsetAsync(sl.d_pairslots.p, 0xff, (size_t)n * c.cap_keys * 4 * 8, sl.stream));
    CK(cudaMemsetAsync(sl.d_paircount.p, 0, (size_t)n * c.cap_keys * 4, sl.stream));
    PairTable ptab;
    ptab.slots = sl.d_pairslots.as<unsigned long long>(); ptab.keys = sl.d_pairkeys.as<uint32_t>(); ptab.ncl = d_ncl;
    ptab.count = sl.d_paircount.as<uint32_t>(); ptab.start = sl.d_pairstart.as<uint32_t>();
    ptab.nslots = c.cap_keys * 4; ptab.cap_keys = c.cap_keys;
    tm.mark();  // 1: after H2D
    const uint8_t *quad_im, *gray_full;
    size_t q_pitch, q_frame, gray_pitch, gray_frame;
    bool from_masks = false;
    int rc = run_image_stage(h, sl, d_src, c.channels, c.W, c.H, c.stride, c.frame_bytes, n, g, &quad_im, &q_pitch,
                             &q_frame, &gray_full, &gray_pitch, &gray_frame, &from_masks);
    if (rc) return rc;
    tm.mark();  // 2: after image
    CcRoots rt;
    rc = run_cc_stage(h, sl, sl.d_thresh.as<uint8_t>(), n, g, c.roots_cap, d_nroots, d_ndense, h->cfg.debug != 0, rt, from_masks);
    if (rc) return rc;
    tm.mark();  // 3: after CC
    {
        const int ew = h->tune.edge_warps;
        dim3 grid(n, ceil_div(cc_tiles_x(g), ew), cc_tiles_y(g));
#define LAUNCH_EDGES(EW) k_edges<EW><<<grid, EW * 32, 0, sl.stream>>>(sl.d_masks.as<uint2>(), sl.d_l16.as<uint16_t>(), \
            rt, g, sl.d_recs[0].as<unsigned long long>(), d_npts, d_ndups, cap, ptab)
        {
            KScope ks(h, sl, "k_edges", sl.stream);
            if (ew == 1) LAUNCH_EDGES(1); else if (ew == 2) LAUNCH_EDGES(2); else if (ew == 4) LAUNCH_EDGES(4); else LAUNCH_EDGES(8);
        }
#undef LAUNCH_EDGES
        LAUNCH_CHECK("k_edges");
    }
    tm.mark();  // 4: after edges
    // ---- cluster starts + work lists from the per-id counts, then the one-sweep scatter by cluster id
    ClusterLists cl;
    const int max_cluster = 3 * (2 * g.wd + 2 * g.hd);   // upstream's cluster size limit
    int tier_cap[AGPU_NTIERS], tier_smem[AGPU_NTIERS];    // largest cluster of a tier / capacity of its shared-memory sort buffer
    for (int t = 0; t < AGPU_NTIERS; t++) {
        tier_cap[t] = t == AGPU_NTIERS - 1 ? std::max(max_cluster, h->tune.tier_cap[t - 1] + 1) : h->tune.tier_cap[t];
        tier_smem[t] = (std::min(tier_cap[t], QF_SMEM_CAP) + 1) & ~1;   // (even: the shared-memory areas behind the sort buffer are accessed 16 bytes at a time)
        cl.list[t] = sl.d_clusters[t].as<ClusterRef>();
        cl.cap[t] = tier_cap[t];
    }
    cl.counters = d_cnt;
    cl.cap_list = n * c.maxcl;
    cl.dbg_heads = h->cfg.debug ? sl.d_dbg_heads.as<ClusterRef>() : nullptr;
    cl.cap_dbg = (int)(sl.d_dbg_heads.bytes / sizeof(ClusterRef));
    {
        KScope ks(h, sl, "k_cluster_refs", sl.stream);
        k_cluster_refs<<<n, 256, 0, sl.stream>>>(ptab, g, std::max(h->prm.min_cluster_pixels, 24), cl, cap);
    }
    LAUNCH_CHECK("k_cluster_refs");
    {
        KScope ks(h, sl, "k_sort_scatter", sl.stream);
        dim3 grid(cap / RS_TILE, n);
        k_sort_scatter<<<grid, RS_THREADS, 0, sl.stream>>>(sl.d_recs[0].as<unsigned long long>(), sl.d_recs[1].as<unsigned long long>(),
                                                          d_npts, cap, ptab);
    }
    LAUNCH_CHECK("k_sort_scatter");
    const int cur = 1;
    tm.mark();  // 5: after sort
    const unsigned long long* srecs = sl.d_recs[cur].as<unsigned long long>();
    {
        QuadFitArgs qa;
        qa.recs = srecs; qa.dense2rep = sl.d_dense2rep.as<uint32_t>(); qa.pair_keys = sl.d_pairkeys.as<uint32_t>();
        qa.cap_keys = c.cap_keys; qa.cap = cap;
        qa.quad_im = quad_im; qa.q_pitch = q_pitch; qa.q_frame = q_frame;
        qa.g = g;
        qa.list_cap = n * c.maxcl;
        qa.quads = sl.d_quads.as<QuadRec>();
        qa.nquads = d_cnt + CNT_NQUADS;
        qa.cap_quads = n * c.maxq;
        qa.per_frame_quads = d_frame_quads;
        qa.oversize = d_cnt + CNT_OVERSIZE;
        qa.gsort = sl.d_gsort.as<unsigned long long>();
        qa.gsort_stride = (size_t)max_cluster;
        // the tiers are independent (own work list, atomic appends to the quad list): fork them over side
        // streams so that the latency-bound big-cluster warps overlap with the many small clusters
        CK(cudaEventRecord(sl.ev_mid, sl.stream));
        CK(cudaStreamWaitEvent(sl.tail, sl.ev_mid, 0));
        CK(cudaEventRecord(sl.ev_fork, sl.tail));
        // (profiling: the tiers run one after the other on the tail stream so that every kernel's event pair times
        // that kernel alone)
        const bool fork = !h->profiling;
        static const char* const tier_name[AGPU_NTIERS] = {"k_fit_quads<1>", "k_fit_quads<2>", "k_fit_quads<4>", "k_fit_quads<4>/6k",
                                                           "k_fit_quads<8>"};
        for (int t = AGPU_NTIERS - 1; t >= 0; t--) {
            cudaStream_t st = (t == 0 || !fork) ? sl.tail : sl.aux[t - 1];
            if (t > 0 && fork) CK(cudaStreamWaitEvent(st, sl.ev_fork, 0));
            qa.list = sl.d_clusters[t].as<ClusterRef>();
            qa.list_count = d_cnt + CNT_TIER0 + t;
            qa.cursor = d_cnt + CNT_CURSOR0 + t;
            // persistent groups: a full machine's worth for big chunks, no more than the chunk can keep busy for small ones
            // (a single 640x480 frame would otherwise launch 4000 CTAs whose only act is to find the work list empty)
            static const int per_frame[AGPU_NTIERS] = {32, 96, 48, 16, 12};   // per 320x240 working pixels
            const long long area = std::max<long long>(1, (long long)g.plane / 76800);
            const int nblk = std::max(1, (int)std::min<long long>((long long)h->num_sms * h->tune.tier_ctas[t], n * area * per_frame[t]));
            {   // this tier's slice of the per-group scratch (tiers are visited from the last to the first)
                size_t off = 0;
                for (int u = AGPU_NTIERS - 1; u > t; u--) off += (size_t)h->num_sms * h->tune.tier_ctas[u] * (u == 0 ? 8 : 1) * ((tier_cap[u] + 1) & ~1);
                qa.scratch = sl.d_qscratch.as<double>() + off * 7;
                qa.scratch_pts = (tier_cap[t] + 1) & ~1;   // (a group's slice starts 16-byte aligned: the moments are read 16 bytes at a time)
            }
            {
            KScope ks(h, sl, tier_name[t], st);
            if (t == 0) {
                const size_t smem = 8 * qf_smem_per_group(tier_smem[t], 1);
                k_fit_quads<1><<<nblk, 256, smem, st>>>(qa, h->prm, tier_smem[t]);
            } else if (t == 1) {
                k_fit_quads<2><<<nblk, 64, qf_smem_per_group(tier_smem[t], 2), st>>>(qa, h->prm, tier_smem[t]);
            } else if (t == 2 || t == 3) {
                k_fit_quads<4><<<nblk, 128, qf_smem_per_group(tier_smem[t], 4), st>>>(qa, h->prm, tier_smem[t]);
            } else {
                k_fit_quads<8><<<nblk, 256, qf_smem_per_group(tier_smem[t], 8), st>>>(qa, h->prm, tier_smem[t]);
            }
            }
            LAUNCH_CHECK("k_fit_quads");
            if (t > 0 && fork) {
                CK(cudaEventRecord(sl.ev_join[t - 1], st));
                CK(cudaStreamWaitEvent(sl.tail, sl.ev_join[t - 1], 0));
            }
        }
    }
    tm.mark(sl.tail);  // 6: after quads
    {
        DecodeArgs da;
        da.im = gray_full; da.pitch = gray_pitch; da.frame_stride = gray_frame;
        da.W = c.W; da.H = c.H;
        da.quads = sl.d_quads.as<QuadRec>();
        da.nquads = d_cnt + CNT_NQUADS;
        da.cap_quads = n * c.maxq;
        da.fams = h->d_fams.as<DevFamily>();
        da.codes = h->d_codes.as<unsigned long long>();
        da.dets = sl.d_dets.as<DetRec>();
        da.ndets = d_ndets;
        da.cap_dets = REC_CAP;
        da.dbg_refined = h->cfg.debug ? sl.d_refined.as<float>() : nullptr;
        {
            KScope ks(h, sl, "k_decode_quads", sl.tail);
            const int dec_blocks = (int)std::min<long long>((long long)h->num_sms * h->tune.decode_ctas * (128 / h->tune.tail_threads),
                                                            std::max(1LL, (long long)n * std::max<long long>(1, (long long)g.plane / 76800) * 256 * 32 / h->tune.tail_threads));
            k_decode_quads<<<dec_blocks, h->tune.tail_threads, 0, sl.tail>>>(da, h->prm);
        }
        LAUNCH_CHECK("k_decode_quads");
    }
    tm.mark(sl.tail);  // 7: after decode
    {
        KScope ks(h, sl, "k_reconcile", sl.tail);
        k_reconcile<<<n, REC_THREADS, 0, sl.tail>>>(sl.d_dets.as<DetRec>(), d_ndets, REC_CAP, n,
                                                          sl.d_out.as<DetRec>(), d_out_counts, c.cap_out);
    }
    LAUNCH_CHECK("k_reconcile");
    if (c.pose->enabled) {
        PoseArgs pa;
        fill_pose_args(pa, *c.pose, 0);
        pa.corners = reinterpret_cast<const double*>(sl.d_out.as<char>() + offsetof(DetRec, p));
        pa.corner_stride = (int)(sizeof(DetRec) / 8);
        pa.counts = d_out_counts;
        pa.per_frame = c.cap_out;
        pa.M = n * c.cap_out;
        pa.out = sl.d_poses.as<PoseRec>();
        {
            KScope ks(h, sl, "k_pose", sl.tail);
            k_pose<<<ceil_div((long long)pa.M * 4, h->tune.tail_threads), h->tune.tail_threads, 0, sl.tail>>>(pa);
        }
        LAUNCH_CHECK("k_pose");
    }
    tm.mark(sl.tail);  // 8: after reconcile/pose
    CK(cudaMemcpyAsync(sl.h_counts.p, d_cnt, c.ncnt * 4, cudaMemcpyDeviceToHost, sl.tail));
    CK(cudaMemcpyAsync(sl.h_out.p, sl.d_out.p, (size_t)n * c.cap_out * sizeof(DetRec), cudaMemcpyDeviceToHost, sl.tail));
    if (c.pose->enabled)
        CK(cudaMemcpyAsync(sl.h_poses.p, sl.d_poses.p, (size_t)n * c.cap_out * sizeof(PoseRec), cudaMemcpyDeviceToHost,
                           sl.tail));
    tm.mark(sl.tail);  // 9: after D2H
    sl.pending = true;
    sl.b0 = b0;
    sl.n = n;
    sl.sorted = cur;
    sl.from_masks = from_masks;
    return AGPU_OK;
}

unsigned long long slot_buffer_hash(const agpu_handle* h, const Slot& s) {
    const DevBuf* bufs[] = {&s.d_in, &s.d_gray, &s.d_quad_im, &s.d_thresh, &s.d_masks, &s.d_l16,
                            &s.d_canon, &s.d_canon_sizes, &s.d_roottab, &s.d_tilebase, &s.d_dense2rep, &s.d_recs[0], &s.d_recs[1],
                            &s.d_qscratch, &s.d_gsort, &s.d_pairslots, &s.d_pairkeys, &s.d_paircount, &s.d_pairstart, &s.d_counters, &s.d_clusters[0], &s.d_clusters[1], &s.d_clusters[2],
                            &s.d_clusters[3], &s.d_clusters[4], &s.d_quads, &s.d_dets, &s.d_out, &s.d_poses, &h->d_fams, &h->d_codes};
    unsigned long long x = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) { x = (x ^ v) * 1099511628211ull; };
    for (const DevBuf* b : bufs) mix((unsigned long long)(uintptr_t)b->p);
    mix((unsigned long long)(uintptr_t)s.h_out.p); mix((unsigned long long)(uintptr_t)s.h_counts.p); mix((unsigned long long)(uintptr_t)s.h_poses.p);
    return x;
}

// One chunk: either enqueue its ~30 kernels, or -- for the synchronous small-batch path -- replay them as one CUDA graph.
int launch_chunk(agpu_handle* h, Slot& sl, const CallCtx& c, int b0, int n) {
    sl.graph_launched = false;
    const bool graphable = h->tune.graph && !h->profiling && !h->cfg.debug && b0 == 0 && n == c.B &&
                           n <= h->tune.graph_max_frames && &sl == &h->slots[0];
    if (!graphable) return enqueue_chunk(h, sl, c, b0, n);
    // the frames are staged in the slot's own input buffer so that every address inside the graph is fixed
    CK(sl.d_in.ensure(c.frame_bytes * c.chunk));
    Slot::GraphKey key;
    memset(&key, 0, sizeof(key));
    key.W = c.W; key.H = c.H; key.stride = c.stride; key.channels = c.channels; key.n = n; key.cap = c.cap; key.maxcl = c.maxcl;
    key.maxq = c.maxq; key.cap_out = c.cap_out; key.cap_keys = c.cap_keys; key.roots_cap = c.roots_cap; key.pose = c.pose->enabled ? 1 : 0;
    if (c.pose->enabled) {
        memcpy(key.K, c.pose->K, sizeof(key.K)); memcpy(key.dist, c.pose->dist, sizeof(key.dist));
        key.tag_size = c.pose->tag_size; key.ndist = c.pose->ndist;
    }
    key.buffers = slot_buffer_hash(h, sl);
    const bool replay = sl.graph_exec && memcmp(&key, &sl.graph_key, sizeof(key)) == 0;
    const bool capture = !replay && sl.have_seen && memcmp(&key, &sl.seen_key, sizeof(key)) == 0;
    if (!replay && !capture) {   // first sighting of this key: a plain run (it also sizes every workspace)
        sl.seen_key = key;
        sl.have_seen = true;
        return enqueue_chunk(h, sl, c, b0, n);
    }
    CK(cudaMemcpyAsync(sl.d_in.p, c.frames, c.frame_bytes * n, c.on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                       sl.stream));
    if (capture) {
        if (sl.graph_exec) { cudaGraphExecDestroy(sl.graph_exec); sl.graph_exec = nullptr; }
        CallCtx cg = c;
        cg.frames = sl.d_in.as<uint8_t>();
        cg.on_device = 1;
        const long long before = h->launches;
        CK(cudaStreamBeginCapture(sl.stream, cudaStreamCaptureModeRelaxed));
        int rc = enqueue_chunk(h, sl, cg, 0, n);
        if (rc == AGPU_OK) {   // the back half ends on the tail stream: join it into the origin stream
            if (cudaEventRecord(sl.ev_end, sl.tail) != cudaSuccess || cudaStreamWaitEvent(sl.stream, sl.ev_end, 0) != cudaSuccess)
                rc = AGPU_E_CUDA;
        }
        cudaGraph_t graph = nullptr;
        const cudaError_t ee = cudaStreamEndCapture(sl.stream, &graph);
        sl.pending = false;
        if (rc != AGPU_OK || ee != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            if (rc == AGPU_OK) { h->set_err(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ee)); rc = AGPU_E_CUDA; }
            return rc;
        }
        const cudaError_t ie = cudaGraphInstantiate(&sl.graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) {
            sl.graph_exec = nullptr;
            h->set_err(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie));
            return AGPU_E_CUDA;
        }
        sl.graph_key = key;
        sl.graph_launches = h->launches - before;
        h->launches = before;
        sl.graph_sorted = sl.sorted;
        sl.graph_from_masks = sl.from_masks;
    }
    CK(cudaGraphLaunch(sl.graph_exec, sl.stream));
    h->launches += sl.graph_launches;
    sl.graph_launched = true;
    sl.pending = true;
    sl.b0 = b0;
    sl.n = n;
    sl.sorted = sl.graph_sorted;
    sl.from_masks = sl.graph_from_masks;
    return AGPU_OK;
}

struct Overflow {
    int max_pts = 0, max_cl_per_frame = 0, max_q_per_frame = 0, max_ncl = 0, max_roots = 0;
    bool any = false;
};

// wait for the slot's chunk and move its results into the caller's arrays; 1 = a work list overflowed (redo)
int finish_chunk(agpu_handle* h, Slot& sl, const CallCtx& c, agpu_detection* out, agpu_pose_t* poses, int* counts,
                 Overflow& ov, int& rc_final) {
    CK(cudaStreamSynchronize(sl.graph_launched ? sl.stream : sl.tail));   // (the tail is ordered after everything on sl.stream
                                                                           // through ev_mid; a graph is launched on sl.stream)
    sl.pending = false;
    const int n = sl.n, b0 = sl.b0, chunk = c.chunk;
    if (h->profiling) {
        for (int s = 0; s < AGPU_NUM_STAGES; s++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, sl.events[s], sl.events[s + 1]);
            h->stage_ms[s] += ms;
        }
        if (h->ev_t0) {
            h->timeline.push_back((float)sl.b0); h->timeline.push_back((float)sl.n);
            h->timeline.push_back((float)(&sl - h->slots.data()));
            for (int s = 0; s <= AGPU_NUM_STAGES; s++) {
                float ms = 0;
                cudaEventElapsedTime(&ms, h->ev_t0, sl.events[s]);
                h->timeline.push_back(ms);
            }
        }
        for (size_t i = 0; i < sl.kev_next; i++) {
            float kms = 0;
            if (cudaEventElapsedTime(&kms, sl.kev[i].a, sl.kev[i].b) != cudaSuccess) continue;
            auto& e = h->kernel_ms[sl.kev[i].name];
            e.first += kms;
            e.second += 1;
        }
    }
    const int* hc = sl.h_counts.as<int>();
    const int* h_npts = hc + CNT_FIXED;
    const int* h_nd = h_npts + 2 * chunk;
    const int* h_oc = h_nd + chunk;
    const int* h_ndense = h_oc + chunk;
    const int* h_ndups = h_ndense + chunk + CC_SUBLISTS * chunk;
    const int* h_ncl = h_ndups + chunk;
    int max_dense = 0;
    for (int i = 0; i < n; i++) max_dense = std::max(max_dense, h_ndense[i]);
    if (max_dense > AGPU_MAX_DENSE) {
        h->set_err("more than 65535 connected components of >= 25 pixels in one frame");
        return AGPU_E_WORKSPACE;
    }
    h->max_dense_seen = std::max(h->max_dense_seen, max_dense);
    int max_pts = 0, max_cl = 0;
    for (int i = 0; i < n; i++) max_pts = std::max(max_pts, h_npts[i]);
    for (int t = 0; t < AGPU_NTIERS; t++) max_cl = std::max(max_cl, hc[CNT_TIER0 + t]);
    int max_ncl = 0;
    for (int i = 0; i < n; i++) max_ncl = std::max(max_ncl, h_ncl[i]);
    h->max_clusters_seen = std::max(h->max_clusters_seen, max_ncl);
    bool redo = false;
    // the pair table ran out of cluster ids: run the chunk again with a larger one
    if (max_ncl > c.cap_keys) { ov.max_ncl = std::max(ov.max_ncl, max_ncl); redo = true; }
    // ... or a root sub-list was too short for the frame's tile-local roots
    int max_roots = 0;
    {
        const int* h_nroots = h_ndense + chunk;
        for (int i = 0; i < n * CC_SUBLISTS; i++) max_roots = std::max(max_roots, h_nroots[i]);
    }
    if (max_roots > c.roots_cap) { ov.max_roots = std::max(ov.max_roots, max_roots); redo = true; }
    h->roots_hint = std::max(1, max_roots);
    if (max_pts > c.cap) { ov.max_pts = std::max(ov.max_pts, max_pts); redo = true; }
    if (max_cl > n * c.maxcl) { ov.max_cl_per_frame = std::max(ov.max_cl_per_frame, (max_cl + n - 1) / n); redo = true; }
    if (hc[CNT_NQUADS] > n * c.maxq) { ov.max_q_per_frame = std::max(ov.max_q_per_frame, (hc[CNT_NQUADS] + n - 1) / n); redo = true; }
    if (redo) {
        ov.any = true;
        return 1;
    }
    const DetRec* ho = sl.h_out.as<DetRec>();
    for (int i = 0; i < n; i++) {
        const int cnt = h_oc[i];
        counts[b0 + i] = cnt;
        const int m = std::min(cnt, c.cap_out);
        memcpy(out + (size_t)(b0 + i) * c.cap_out, ho + (size_t)i * c.cap_out, (size_t)m * sizeof(DetRec));
        if (c.pose->enabled && poses)
            memcpy(poses + (size_t)(b0 + i) * c.cap_out, sl.h_poses.as<PoseRec>() + (size_t)i * c.cap_out,
                   (size_t)m * sizeof(PoseRec));
        if (cnt > c.cap_out && rc_final == AGPU_OK) {
            h->set_err("more detections than cap_per_frame; counts[] hold the true numbers");
            rc_final = AGPU_E_TRUNCATED;
        }
        h->counters[0] += h_npts[i] + h_ndups[i];   // raw edge points, as upstream counts them
        h->counters[3] += h_nd[i];
        if (h_nd[i] > REC_CAP && rc_final == AGPU_OK) {   // (upstream has no such limit; the frame keeps its first 1024 candidates)
            h->set_err("more than 1024 raw detections (before reconcile) in one frame: the surplus was dropped");
            rc_final = AGPU_E_TRUNCATED;
        }
    }
    for (int t = 0; t < AGPU_NTIERS; t++) h->counters[1] += hc[CNT_TIER0 + t];
    // public statistics: classes 0..3 = clusters of up to 256 / 1024 / 2048 / more records (the two large tiers together)
    for (int t = 0; t < AGPU_NTIERS; t++) {
        const int cls = std::min(t, 3);
        if (cls > 0) h->counters[4 + cls] += hc[CNT_TIER0 + t];
        h->tier_stats[cls] += hc[CNT_TIER0 + t];
        h->tier_stats[4 + cls] += hc[CNT_TIER_RECS0 + t];
    }
    h->counters[2] += hc[CNT_NQUADS];
    h->counters[4] += hc[CNT_OVERSIZE];
    h->last_slot = (int)(&sl - h->slots.data());
    h->last_chunk = n;
    h->last_cap = c.cap;
    h->last_cap_keys = c.cap_keys;
    h->last_from_masks = sl.from_masks;
    h->last_bgr = c.channels == 3;
    h->have_last = true;
    return 0;
}

// Wait for everything a slot still has in flight (kernels, the asynchronous copies from the caller's host frames and
// into the pinned result buffers) and forget its chunk.  Every error exit of a detect call goes through this: a chunk
// left pending would otherwise be "finished" by the NEXT call with that call's batch geometry.
void drain_slots(agpu_handle* h) {
    for (Slot& s : h->slots) {
        if (!s.pending) continue;
        if (s.stream) cudaStreamSynchronize(s.stream);
        for (int t = 0; t < AGPU_NTIERS - 1; t++)
            if (s.aux[t]) cudaStreamSynchronize(s.aux[t]);
        if (s.tail) cudaStreamSynchronize(s.tail);
        s.pending = false;
    }
}

int detect_run(agpu_handle* h, const uint8_t* frames, int on_device, int channels, int B, int W, int H, int stride,
               void* cuda_stream, const PoseSpec& pose, agpu_detection* out, agpu_pose_t* poses, int cap_out,
               int* counts) {
    if (!h) return AGPU_E_INVALID;
    if (!frames || !out || !counts || B <= 0 || W <= 0 || H <= 0 || cap_out <= 0 || (channels != 1 && channels != 3) ||
        stride < W * channels) {
        h->set_err("agpu_detect: invalid argument");
        return AGPU_E_INVALID;
    }
    if (2 * W + 1 > 16383 || 2 * H + 1 > 16383) {
        h->set_err("agpu_detect: frames larger than 8190 pixels per side are not supported");
        return AGPU_E_UNSUPPORTED;
    }
    CK(cudaSetDevice(h->device));
    h->launches = 0;
    for (int i = 0; i < AGPU_NUM_STAGES; i++) h->stage_ms[i] = 0;
    h->kernel_ms.clear();
    for (int i = 0; i < 8; i++) h->counters[i] = h->tier_stats[i] = 0;
    CallCtx c;
    c.frames = frames; c.on_device = on_device; c.channels = channels; c.B = B; c.W = W; c.H = H; c.stride = stride;
    c.frame_bytes = (size_t)H * stride;
    c.g = make_geom(W, H, h->prm.decim);
    c.cap_out = cap_out;
    c.pose = &pose;
    const Geom& g = c.g;
    h->geom = g;
    for (int b = 0; b < B; b++) counts[b] = 0;
    if (g.wd < 8 || g.hd < 8) return AGPU_OK;  // nothing detectable (and the tile grid would be empty)

    // chunking: ~256 Mpx of working image per pass (129 frames of 1080p), several chunks in flight; a batch that is
    // big enough is cut into at least three chunks so that the slots overlap
    int chunk = h->cfg.chunk_frames;
    if (chunk <= 0) {
        const size_t target_px = (size_t)256 << 20;
        chunk = (int)std::max<size_t>(1, target_px / g.plane);
        chunk = std::min(chunk, 256);
        if (B >= 48) chunk = std::min(chunk, (B + 2) / 3);
        // host frames: the call is bound by the host link, so what matters is that the link never idles and that
        // little work is left once the last copy has landed -- small chunks, one more slot
        if (!on_device) chunk = std::min(chunk, std::max(1, (int)(((size_t)64 << 20) / g.plane)));
    }
    chunk = std::min(chunk, B);
    c.chunk = chunk;
    int nslots = h->cfg.pipeline_slots > 0 ? h->cfg.pipeline_slots : (on_device ? 3 : 4);
    if (h->tune.slots > 0) nslots = h->tune.slots;
    nslots = std::min(nslots, 8);
    nslots = std::min(nslots, ceil_div(B, chunk));
    if ((int)h->slots.size() < nslots) h->slots.resize(nslots);   // (slots are only ever appended: buffers stay put)

    // per-frame list capacities: user limits are hard; automatic ones grow (the chunks are re-run) on overflow
    const bool auto_pts = h->cfg.max_points_per_frame <= 0, auto_cl = h->cfg.max_clusters_per_frame <= 0,
               auto_q = h->cfg.max_quads_per_frame <= 0;
    c.cap = auto_pts ? std::max(h->cap_points, (int)std::max<size_t>(262144, g.plane / 4)) : h->cfg.max_points_per_frame;
    c.cap = (c.cap + RS_TILE - 1) / RS_TILE * RS_TILE;
    c.maxcl = auto_cl ? std::max(h->cap_clusters, 8192) : h->cfg.max_clusters_per_frame;
    c.maxq = auto_q ? std::max(h->cap_quads, 1024) : h->cfg.max_quads_per_frame;
    c.ncnt = CNT_FIXED + (size_t)(7 + CC_SUBLISTS) * chunk;
    c.cap_keys = std::max(h->cap_keys, 4096);
    c.roots_cap = std::max(h->cap_roots, default_roots_cap(g));
    c.key_bits = [&] { int nb = 1; while (((size_t)1 << nb) < g.plane) nb++; return nb; }();

    if (on_device) {   // order every slot stream after the producer's stream
        if (!h->ev_user) CK(cudaEventCreateWithFlags(&h->ev_user, cudaEventDisableTiming));
        CK(cudaEventRecord(h->ev_user, (cudaStream_t)cuda_stream));
    }
    h->timeline.clear();
    if (h->profiling) {
        if (!h->ev_t0) CK(cudaEventCreate(&h->ev_t0));
        CK(cudaEventRecord(h->ev_t0, on_device ? (cudaStream_t)cuda_stream : h->slots[0].stream));
    }
    int rc_final = AGPU_OK;
    std::vector<std::pair<int, int>> todo, redo;
    for (int b0 = 0; b0 < B; b0 += chunk) todo.push_back({b0, std::min(chunk, B - b0)});
    while (!todo.empty()) {
        for (int s = 0; s < nslots; s++) {
            int rc = alloc_slot(h, h->slots[s], c);
            if (rc) return rc;
            if (on_device) CK(cudaStreamWaitEvent(h->slots[s].stream, h->ev_user, 0));
        }
        h->cap_points = c.cap; h->cap_clusters = c.maxcl; h->cap_quads = c.maxq; h->cap_keys = c.cap_keys; h->cap_roots = c.roots_cap;
        Overflow ov;
        redo.clear();
        size_t next = 0;
        int si = 0;
        bool any_pending = true;
        while (next < todo.size() || any_pending) {
            Slot& sl = h->slots[si];
            if (sl.pending) {
                const std::pair<int, int> ch = {sl.b0, sl.n};
                int rc = finish_chunk(h, sl, c, out, poses, counts, ov, rc_final);
                if (rc < 0) return rc;
                if (rc == 1) redo.push_back(ch);
            }
            if (next < todo.size()) {
                int rc = launch_chunk(h, sl, c, todo[next].first, todo[next].second);
                if (rc) return rc;
                next++;
            }
            si = (si + 1) % nslots;
            any_pending = false;
            for (int s = 0; s < nslots; s++) any_pending |= h->slots[s].pending;
        }
        if (redo.empty()) break;
        // a work list overflowed somewhere: grow it (automatic limits) and run those chunks again
        if (ov.max_pts > c.cap) {
            if (!auto_pts) { h->set_err("edge-point list overflow: raise agpu_config.max_points_per_frame"); return AGPU_E_WORKSPACE; }
            c.cap = (ov.max_pts + ov.max_pts / 4 + RS_TILE - 1) / RS_TILE * RS_TILE;
        }
        if (ov.max_cl_per_frame > c.maxcl) {
            if (!auto_cl) { h->set_err("cluster list overflow: raise agpu_config.max_clusters_per_frame"); return AGPU_E_WORKSPACE; }
            c.maxcl = ov.max_cl_per_frame * 2;
        }
        if (ov.max_roots > c.roots_cap) c.roots_cap = ov.max_roots + ov.max_roots / 4;
        if (ov.max_ncl > c.cap_keys) {
            int ck = c.cap_keys;
            while (ck < ov.max_ncl + ov.max_ncl / 4) ck *= 2;
            if (ck > (1 << 22)) { h->set_err("more than 4M edge clusters in one frame"); return AGPU_E_WORKSPACE; }
            c.cap_keys = ck;
        }
        if (ov.max_q_per_frame > c.maxq) {
            if (!auto_q) { h->set_err("quad list overflow: raise agpu_config.max_quads_per_frame"); return AGPU_E_WORKSPACE; }
            c.maxq = ov.max_q_per_frame * 2;
        }
        todo = redo;
    }
    return rc_final;
}

int detect_impl(agpu_handle* h, const uint8_t* frames, int on_device, int channels, int B, int W, int H, int stride,
                void* cuda_stream, const PoseSpec& pose, agpu_detection* out, agpu_pose_t* poses, int cap_out,
                int* counts) {
    if (h) drain_slots(h);   // (safety net: nothing may be pending when a call starts)
    const int rc = detect_run(h, frames, on_device, channels, B, W, H, stride, cuda_stream, pose, out, poses, cap_out, counts);
    if (h) drain_slots(h);   // error exits leave chunks in flight; after a successful call this is a no-op
    return rc;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int agpu_version(void) { return AGPU_VERSION; }

void agpu_default_config(agpu_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->families = "tag36h11";
    cfg->threads = 1;
    cfg->maxhamming = 1;
    cfg->quad_decimate = 2.0f;
    cfg->quad_sigma = 0.0f;
    cfg->refine_edges = 1;
    cfg->decode_sharpening = 0.25;
    cfg->debug = 0;
    cfg->device = 0;
}

const char* agpu_last_error(const agpu_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int agpu_create(const agpu_config* cfg, agpu_handle** out) {
    if (!cfg || !out || !cfg->families) {
        g_create_error = "agpu_create: NULL argument";
        return AGPU_E_INVALID;
    }
    *out = nullptr;
    agpu_handle* h = new agpu_handle();
    h->cfg = *cfg;
    h->families_str = cfg->families;
    h->cfg.families = h->families_str.c_str();
    if (const char* e = getenv("AGPU_PRIO")) h->tune.prio = atoi(e) != 0;
    if (const char* e = getenv("AGPU_MASKS")) h->tune.masks = atoi(e) != 0;
    if (const char* e = getenv("AGPU_GRAPH")) h->tune.graph = atoi(e) != 0;
    // (experiment knobs: parsed once, clamped to what the kernels are instantiated for -- nothing here can divide by zero)
    auto pow2_knob = [](const char* name, int dflt) {
        const char* e = getenv(name);
        if (!e) return dflt;
        const int v = atoi(e);
        return (v == 1 || v == 2 || v == 4 || v == 8) ? v : dflt;
    };
    h->tune.edge_warps = pow2_knob("AGPU_EDGE_WARPS", h->tune.edge_warps);
    h->tune.boundary_warps = pow2_knob("AGPU_BOUNDARY_WARPS", h->tune.boundary_warps);
    if (const char* e = getenv("AGPU_SEG_TILES")) h->tune.seg_tiles = std::max(0, std::min(256, atoi(e)));
    if (const char* e = getenv("AGPU_IMG_MINB")) { const int v = atoi(e); h->tune.img_minb = (v >= 3 && v <= 6) ? v : 4; }
    if (const char* e = getenv("AGPU_SLOTS")) h->tune.slots = std::max(0, std::min(8, atoi(e)));
    if (const char* e = getenv("AGPU_BLUR_STRIP")) h->tune.blur_strip = atoi(e) != 0;
    if (const char* e = getenv("AGPU_TAIL_THREADS")) h->tune.tail_threads = atoi(e) >= 128 ? 128 : (atoi(e) >= 64 ? 64 : 32);
    if (const char* e = getenv("AGPU_DECODE_CTAS")) h->tune.decode_ctas = std::max(1, std::min(16, atoi(e)));
    if (const char* e = getenv("AGPU_TIER_CAP")) {
        int v[3];
        if (sscanf(e, "%d,%d,%d", &v[0], &v[1], &v[2]) == 3 && v[0] >= 64 && v[0] < v[1] && v[1] < v[2] && v[2] <= 4096)
            for (int t = 0; t < 3; t++) h->tune.tier_cap[t] = v[t];
    }
    if (const char* e = getenv("AGPU_TIER_CTAS")) {
        int v[AGPU_NTIERS];
        if (sscanf(e, "%d,%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3], &v[4]) == 5)
            for (int t = 0; t < AGPU_NTIERS; t++) h->tune.tier_ctas[t] = std::max(1, std::min(32, v[t]));
    }
    auto fail = [&](int rc, const std::string& msg) {
        g_create_error = msg;
        delete h;
        return rc;
    };
    // families
    {
        std::string s = h->families_str, tok;
        size_t pos = 0;
        while (pos <= s.size()) {
            size_t e = s.find_first_of(" ,", pos);
            if (e == std::string::npos) e = s.size();
            tok = s.substr(pos, e - pos);
            pos = e + 1;
            if (tok.empty()) continue;
            const FamilyDef* f = find_family(tok);
            if (!f) return fail(AGPU_E_INVALID, "Unrecognized tag family name: " + tok);
            h->fams.push_back(f);
        }
        if (h->fams.empty()) return fail(AGPU_E_INVALID, "no tag family given");
        if ((int)h->fams.size() > AGPU_MAX_FAMILIES) return fail(AGPU_E_INVALID, "too many families");
    }
    if (!(cfg->quad_decimate >= 1.0f) || cfg->quad_decimate != std::floor(cfg->quad_decimate) || cfg->quad_decimate > 16)
        return fail(AGPU_E_UNSUPPORTED, "quad_decimate must be an integer factor in [1, 16]");
    if (cfg->maxhamming < 0 || cfg->maxhamming > 2) return fail(AGPU_E_INVALID, "maxhamming must be 0, 1 or 2");
    for (const FamilyDef* f : h->fams)
        if (2 * cfg->maxhamming >= f->h)
            return fail(AGPU_E_INVALID, std::string("maxhamming too large for family ") + f->name);

    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(AGPU_E_CUDA, std::string("no CUDA device: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "count = 0"));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(AGPU_E_INVALID, "device ordinal out of range");
    h->device = cfg->device;
    if ((ce = cudaSetDevice(h->device)) != cudaSuccess) return fail(AGPU_E_CUDA, cudaGetErrorString(ce));
    cudaDeviceProp prop;
    if ((ce = cudaGetDeviceProperties(&prop, h->device)) != cudaSuccess) return fail(AGPU_E_CUDA, cudaGetErrorString(ce));
    if (prop.major != 10)
        return fail(AGPU_E_CUDA, "libaprilgpu is built for sm_100a (B200) only; device is sm_" +
                                     std::to_string(prop.major) + std::to_string(prop.minor));
    h->num_sms = prop.multiProcessorCount;
    h->slots.reserve(8);
    h->slots.resize(1);
    if (init_slot(h, h->slots[0]) != AGPU_OK) return fail(AGPU_E_CUDA, h->err);

    // parameters
    DevParams& P = h->prm;
    memset(&P, 0, sizeof(P));
    P.quad_decimate = cfg->quad_decimate;
    P.decim = (int)cfg->quad_decimate;
    P.refine_edges = cfg->refine_edges ? 1 : 0;
    P.decode_sharpening = cfg->decode_sharpening;
    P.maxhamming = cfg->maxhamming;
    P.min_cluster_pixels = 5;
    P.max_nmaxima = 10;
    P.cos_critical_rad = (float)std::cos(10.0 * M_PI / 180.0);
    P.max_line_fit_mse = 10.0f;
    P.min_white_black_diff = 5;
    int min_w = 1000000;
    for (const FamilyDef* f : h->fams) {
        min_w = std::min(min_w, f->width_at_border);
        if (f->reversed_border) P.reversed_border = 1; else P.normal_border = 1;
    }
    min_w = (int)(min_w / cfg->quad_decimate);
    P.min_tag_width = std::max(3, min_w);
    P.nfamilies = (int)h->fams.size();
    for (int i = 0; i < 7; i++) {
        int j = i - 3;
        P.smooth_f[i] = (float)std::exp(-j * j / (2 * 1.0 * 1.0));
    }
    for (int r = 0; r < 4; r++) {
        double theta = r * M_PI / 2.0;
        P.rot_c[r] = std::cos(theta);
        P.rot_s[r] = std::sin(theta);
    }
    // family tables
    {
        std::vector<DevFamily> df(h->fams.size());
        std::vector<unsigned long long> codes;
        for (size_t i = 0; i < h->fams.size(); i++) {
            const FamilyDef* f = h->fams[i];
            memset(&df[i], 0, sizeof(DevFamily));
            df[i].nbits = f->nbits; df[i].ncodes = f->ncodes; df[i].width_at_border = f->width_at_border;
            df[i].total_width = f->total_width; df[i].reversed_border = f->reversed_border;
            df[i].code_offset = (int)codes.size();
            for (int b = 0; b < f->nbits; b++) { df[i].bit_x[b] = f->bit_x[b]; df[i].bit_y[b] = f->bit_y[b]; }
            codes.insert(codes.end(), f->codes, f->codes + f->ncodes);
        }
        if (h->d_fams.ensure(df.size() * sizeof(DevFamily)) != cudaSuccess || h->d_codes.ensure(codes.size() * 8) != cudaSuccess)
            return fail(AGPU_E_CUDA, "cudaMalloc(family tables) failed");
        cudaMemcpy(h->d_fams.p, df.data(), df.size() * sizeof(DevFamily), cudaMemcpyHostToDevice);
        cudaMemcpy(h->d_codes.p, codes.data(), codes.size() * 8, cudaMemcpyHostToDevice);
    }
    // the large quad-fit tiers need more than 48 KB of dynamic shared memory
    ce = cudaFuncSetAttribute(k_fit_quads<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)qf_smem_per_group(QF_SMEM_CAP, 8));
    if (ce == cudaSuccess)
        ce = cudaFuncSetAttribute(k_fit_quads<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(8 * qf_smem_per_group(h->tune.tier_cap[0], 1)));
    if (ce == cudaSuccess)
        ce = cudaFuncSetAttribute(k_fit_quads<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)qf_smem_per_group(h->tune.tier_cap[3], 4));
    if (ce != cudaSuccess) return fail(AGPU_E_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(ce));
    *out = h;
    return AGPU_OK;
}

int agpu_destroy(agpu_handle* h) {
    if (!h) return AGPU_E_INVALID;
    cudaSetDevice(h->device);
    for (Slot& s : h->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        if (s.tail) cudaStreamSynchronize(s.tail);
        s.release();
    }
    h->d_fams.release(); h->d_codes.release(); h->d_pose_in.release(); h->d_pose_out.release();
    if (h->ev_user) cudaEventDestroy(h->ev_user);
    if (h->ev_t0) cudaEventDestroy(h->ev_t0);
    delete h;
    return AGPU_OK;
}

int agpu_detect(agpu_handle* h, const uint8_t* frames, int on_device, int B, int W, int H, int stride, void* cuda_stream,
                agpu_detection* out, int cap_per_frame, int* counts) {
    PoseSpec ps;
    return detect_impl(h, frames, on_device, 1, B, W, H, stride, cuda_stream, ps, out, nullptr, cap_per_frame, counts);
}

int agpu_detect_bgr(agpu_handle* h, const uint8_t* frames, int on_device, int B, int W, int H, int stride,
                    void* cuda_stream, agpu_detection* out, int cap_per_frame, int* counts) {
    PoseSpec ps;
    return detect_impl(h, frames, on_device, 3, B, W, H, stride, cuda_stream, ps, out, nullptr, cap_per_frame, counts);
}

int agpu_detect_pose(agpu_handle* h, const uint8_t* frames, int on_device, int channels, int B, int W, int H, int stride,
                     void* cuda_stream, const double K[9], const double* dist, int ndist, double tag_size,
                     agpu_detection* out, agpu_pose_t* poses, int cap_per_frame, int* counts) {
    if (!h) return AGPU_E_INVALID;
    if (!K || !poses) {
        h->set_err("agpu_detect_pose: NULL K or poses");
        return AGPU_E_INVALID;
    }
    PoseSpec ps;
    ps.enabled = true;
    for (int i = 0; i < 9; i++) ps.K[i] = K[i];
    ps.tag_size = tag_size;
    int rc = parse_dist(h, dist, ndist, ps);
    if (rc) return rc;
    return detect_impl(h, frames, on_device, channels, B, W, H, stride, cuda_stream, ps, out, poses, cap_per_frame, counts);
}

int agpu_pose(agpu_handle* h, const double* corners, int M, const double K[9], const double* dist, int ndist,
              double tag_size, int method, agpu_pose_t* poses) {
    if (!h) return AGPU_E_INVALID;
    if (!corners || !K || !poses || M < 0 || (method != 0 && method != 1)) {
        h->set_err("agpu_pose: invalid argument");
        return AGPU_E_INVALID;
    }
    if (M == 0) return AGPU_OK;
    CK(cudaSetDevice(h->device));
    h->launches = 0;
    PoseSpec ps;
    ps.enabled = true;
    for (int i = 0; i < 9; i++) ps.K[i] = K[i];
    ps.tag_size = tag_size;
    int rc = parse_dist(h, dist, ndist, ps);
    if (rc) return rc;
    cudaStream_t stream = h->slots[0].stream;
    CK(h->d_pose_in.ensure((size_t)M * 64));
    CK(h->d_pose_out.ensure((size_t)M * sizeof(PoseRec)));
    CK(cudaMemcpyAsync(h->d_pose_in.p, corners, (size_t)M * 64, cudaMemcpyHostToDevice, stream));
    PoseArgs pa;
    fill_pose_args(pa, ps, method);
    pa.corners = h->d_pose_in.as<double>();
    pa.corner_stride = 8;
    pa.counts = nullptr;
    pa.per_frame = 0;
    pa.M = M;
    pa.out = h->d_pose_out.as<PoseRec>();
    k_pose<<<ceil_div((long long)M * 4, 128), 128, 0, stream>>>(pa);
    LAUNCH_CHECK("k_pose");
    CK(cudaMemcpyAsync(poses, h->d_pose_out.p, (size_t)M * sizeof(PoseRec), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return AGPU_OK;
}

int agpu_set_profiling(agpu_handle* h, int on) {
    if (!h) return AGPU_E_INVALID;
    h->profiling = on != 0;
    return AGPU_OK;
}

int agpu_get_stage_ms(agpu_handle* h, float* ms) {
    if (!h || !ms) return AGPU_E_INVALID;
    for (int i = 0; i < AGPU_NUM_STAGES; i++) ms[i] = h->stage_ms[i];
    return AGPU_OK;
}

int agpu_get_timeline(agpu_handle* h, float* out, int cap_floats) {
    if (!h) return AGPU_E_INVALID;
    const int n = (int)h->timeline.size();
    if (out) memcpy(out, h->timeline.data(), sizeof(float) * std::min(n, std::max(cap_floats, 0)));
    return n;
}

int agpu_get_kernel_ms(agpu_handle* h, const char* kernel, float* ms) {
    if (!h || !kernel || !ms) return AGPU_E_INVALID;
    auto it = h->kernel_ms.find(kernel);
    if (it == h->kernel_ms.end()) {
        h->set_err(std::string("agpu_get_kernel_ms: no launch of '") + kernel + "' was timed in the last call (profiling on?)");
        return AGPU_E_INVALID;
    }
    *ms = it->second.first;
    return AGPU_OK;
}

int agpu_get_kernel_table(agpu_handle* h, char* buf, int cap) {
    if (!h) return AGPU_E_INVALID;
    std::string t;
    char line[160];
    for (const auto& kv : h->kernel_ms) {
        snprintf(line, sizeof(line), "%s\t%.6f\t%d\n", kv.first.c_str(), kv.second.first, kv.second.second);
        t += line;
    }
    if (buf && cap > 0) {
        const size_t ncopy = std::min(t.size(), (size_t)cap - 1);
        memcpy(buf, t.data(), ncopy);
        buf[ncopy] = 0;
    }
    return (int)t.size() + 1;
}

int agpu_get_launch_count(agpu_handle* h, long long* launches) {
    if (!h || !launches) return AGPU_E_INVALID;
    *launches = h->launches;
    return AGPU_OK;
}

int agpu_get_counters(agpu_handle* h, long long* counters) {
    if (!h || !counters) return AGPU_E_INVALID;
    for (int i = 0; i < 8; i++) counters[i] = h->counters[i];
    return AGPU_OK;
}

int agpu_get_tier_stats(agpu_handle* h, long long* stats) {
    if (!h || !stats) return AGPU_E_INVALID;
    for (int i = 0; i < 8; i++) stats[i] = h->tier_stats[i];
    return AGPU_OK;
}

int agpu_family_info(const char* family, int* ncodes, int* ncodes_upstream) {
    if (!family) return AGPU_E_INVALID;
    const FamilyDef* f = find_family(family);
    if (!f) return AGPU_E_INVALID;
    if (ncodes) *ncodes = f->ncodes;
    // tagStandard41h12: upstream's table has 2115 code words; only ids 0..4 (read off the reference's own tag images) are
    // available offline, so ids >= 5 can never be reported
    if (ncodes_upstream) *ncodes_upstream = std::string(family) == "tagStandard41h12" ? 2115 : f->ncodes;
    return AGPU_OK;
}

int agpu_debug_dims(agpu_handle* h, int* wd, int* hd) {
    if (!h || !h->have_last) return AGPU_E_INVALID;
    if (wd) *wd = h->geom.wd;
    if (hd) *hd = h->geom.hd;
    return AGPU_OK;
}

// threshold image of one frame of the last chunk, pitched (g.wp): straight from the byte image, or rebuilt from the
// tile-major bit masks when the threshold kernel wrote those instead (white -> 255, black -> 0, neither -> 127)
static bool fetch_thresh_frame(agpu_handle* h, Slot& sl, int frame, std::vector<uint8_t>& th) {
    const Geom& g = h->geom;
    th.assign(g.plane, 127);
    if (!h->last_from_masks)
        return cudaMemcpy(th.data(), sl.d_thresh.as<uint8_t>() + (size_t)frame * g.plane, g.plane, cudaMemcpyDeviceToHost) == cudaSuccess;
    const int tx = cc_tiles_x(g), ty = cc_tiles_y(g);
    std::vector<uint2> m((size_t)tx * ty * 32);
    if (cudaMemcpy(m.data(), sl.d_masks.as<uint2>() + (size_t)frame * m.size(), m.size() * sizeof(uint2),
                   cudaMemcpyDeviceToHost) != cudaSuccess)
        return false;
    for (int y = 0; y < g.hd; y++)
        for (int x = 0; x < g.wd; x++) {
            const uint2 r = m[((size_t)(y >> 5) * tx + (x >> 5)) * 32 + (y & 31)];
            if ((r.x >> (x & 31)) & 1u) th[(size_t)y * g.wp + x] = 255;
            else if ((r.y >> (x & 31)) & 1u) th[(size_t)y * g.wp + x] = 0;
        }
    return true;
}

long long agpu_debug_fetch(agpu_handle* h, const char* what, int frame, void* host_out, long long cap_bytes) {
    if (!h || !what || !host_out) return AGPU_E_INVALID;
    if (!h->have_last || !h->cfg.debug || frame < 0 || frame >= h->last_chunk) {
        h->set_err("agpu_debug_fetch: no debug state (cfg.debug = 1 and a previous agpu_detect are required)");
        return AGPU_E_INVALID;
    }
    cudaSetDevice(h->device);
    const Geom& g = h->geom;
    Slot& sl = h->slots[h->last_slot];
    const std::string w = what;
    const size_t npx = (size_t)g.wd * g.hd;
    auto unpitch = [&](uint32_t id) { return (uint32_t)((id / g.wp) * g.wd + (id % g.wp)); };
    auto unpitch_key = [&](unsigned long long k) {
        return ((unsigned long long)unpitch((uint32_t)(k >> 32)) << 32) | unpitch((uint32_t)k);
    };
    if (w == "thresh") {
        if ((size_t)cap_bytes < npx) return (long long)npx;
        std::vector<uint8_t> th;
        if (!fetch_thresh_frame(h, sl, frame, th)) return AGPU_E_CUDA;
        for (int y = 0; y < g.hd; y++) memcpy((uint8_t*)host_out + (size_t)y * g.wd, th.data() + (size_t)y * g.wp, g.wd);
        return (long long)npx;
    }
    if (w == "gray") {   // full-resolution gray plane converted from the BGR frames of the last call
        const Geom gf = make_geom(g.W, g.H, 1);
        if (!h->last_bgr || !sl.d_gray.p) { h->set_err("agpu_debug_fetch: the last call had no BGR input"); return AGPU_E_INVALID; }
        const size_t nfull = (size_t)g.W * g.H;
        if ((size_t)cap_bytes < nfull) return (long long)nfull;
        if (cudaMemcpy2D(host_out, g.W, sl.d_gray.as<uint8_t>() + (size_t)frame * gf.plane, gf.wp, g.W, g.H, cudaMemcpyDeviceToHost) != cudaSuccess)
            return AGPU_E_CUDA;
        return (long long)nfull;
    }
    if (w == "quad_im") {
        const uint8_t* src = sl.d_quad_im.as<uint8_t>();
        if (!src) { h->set_err("agpu_debug_fetch: buffer not materialised (decimate = 1 keeps no quad_im copy)"); return AGPU_E_INVALID; }
        if ((size_t)cap_bytes < npx) return (long long)npx;
        if (cudaMemcpy2D(host_out, g.wd, src + (size_t)frame * g.plane, g.wp, g.wd, g.hd, cudaMemcpyDeviceToHost) != cudaSuccess)
            return AGPU_E_CUDA;
        return (long long)npx;
    }
    if (w == "labels" || w == "sizes") {
        if ((size_t)cap_bytes < npx * 4) return (long long)npx;
        std::vector<uint32_t> tmp(g.plane), lab;
        const uint32_t* src = (w == "labels" ? sl.d_canon.as<uint32_t>() : sl.d_canon_sizes.as<uint32_t>()) + (size_t)frame * g.plane;
        if (cudaMemcpy(tmp.data(), src, g.plane * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return AGPU_E_CUDA;
        uint32_t* o = (uint32_t*)host_out;
        if (w == "labels") {
            for (int y = 0; y < g.hd; y++)
                for (int x = 0; x < g.wd; x++) o[(size_t)y * g.wd + x] = unpitch(tmp[(size_t)y * g.wp + x]);
        } else {
            lab.resize(g.plane);
            std::vector<uint8_t> th;
            cudaMemcpy(lab.data(), sl.d_canon.as<uint32_t>() + (size_t)frame * g.plane, g.plane * 4, cudaMemcpyDeviceToHost);
            if (!fetch_thresh_frame(h, sl, frame, th)) return AGPU_E_CUDA;
            for (int y = 0; y < g.hd; y++)
                for (int x = 0; x < g.wd; x++) {
                    size_t id = (size_t)y * g.wp + x;
                    // sizes are defined at representatives; 127-pixels are singletons that the device never counts
                    uint32_t v = 0;
                    if (lab[id] == id) v = (th[id] == 127) ? 1u : tmp[id];
                    o[(size_t)y * g.wd + x] = v;
                }
        }
        return (long long)npx;
    }
    if (w == "cluster_keys" || w == "cluster_sizes") {
        std::vector<int> cnt(CNT_FIXED);
        if (cudaMemcpy(cnt.data(), sl.d_counters.p, CNT_FIXED * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return AGPU_E_CUDA;
        int nh = std::min<long long>(cnt[CNT_HEADS], (long long)(sl.d_dbg_heads.bytes / sizeof(ClusterRef)));
        std::vector<ClusterRef> heads(nh);
        if (nh) cudaMemcpy(heads.data(), sl.d_dbg_heads.p, (size_t)nh * sizeof(ClusterRef), cudaMemcpyDeviceToHost);
        std::vector<unsigned long long> recs((size_t)h->last_cap);
        cudaMemcpy(recs.data(), sl.d_recs[sl.sorted].as<unsigned long long>() + (size_t)frame * h->last_cap,
                   (size_t)h->last_cap * 8, cudaMemcpyDeviceToHost);
        std::vector<uint32_t> d2r(AGPU_MAX_DENSE);
        cudaMemcpy(d2r.data(), sl.d_dense2rep.as<uint32_t>() + (size_t)frame * AGPU_MAX_DENSE, AGPU_MAX_DENSE * 4,
                   cudaMemcpyDeviceToHost);
        std::vector<uint32_t> pk((size_t)h->last_cap_keys);
        cudaMemcpy(pk.data(), sl.d_pairkeys.as<uint32_t>() + (size_t)frame * h->last_cap_keys, (size_t)h->last_cap_keys * 4,
                   cudaMemcpyDeviceToHost);
        std::vector<std::pair<unsigned long long, int>> v;
        for (const ClusterRef& r : heads)
            if (r.frame == frame) {
                const uint32_t ck = pk[(uint32_t)(recs[r.start] >> 32)];   // cluster id -> pair of dense component ids
                const uint32_t ra = d2r[ck >> 16], rb = d2r[ck & 0xffffu];
                int raw = r.size;   // upstream's cluster size counts the duplicate points that k_edges merged
                for (int i = 0; i < r.size; i++) raw += ((uint32_t)recs[r.start + i] >> 28) >= 8u;
                v.push_back({unpitch_key(((unsigned long long)std::max(ra, rb) << 32) | std::min(ra, rb)), raw});
            }
        std::sort(v.begin(), v.end());
        long long n = (long long)v.size();
        if (w == "cluster_keys") {
            if (cap_bytes >= n * 8) for (long long i = 0; i < n; i++) ((unsigned long long*)host_out)[i] = v[i].first;
        } else {
            if (cap_bytes >= n * 4) for (long long i = 0; i < n; i++) ((int*)host_out)[i] = v[i].second;
        }
        return n;
    }
    if (w == "quads" || w == "quad_keys" || w == "quads_refined") {
        std::vector<int> cnt(CNT_FIXED);
        if (cudaMemcpy(cnt.data(), sl.d_counters.p, CNT_FIXED * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return AGPU_E_CUDA;
        int nq = std::min<long long>(cnt[CNT_NQUADS], (long long)(sl.d_quads.bytes / sizeof(QuadRec)));
        std::vector<QuadRec> q(nq);
        std::vector<float> ref((size_t)nq * 8);
        if (nq) {
            cudaMemcpy(q.data(), sl.d_quads.p, (size_t)nq * sizeof(QuadRec), cudaMemcpyDeviceToHost);
            cudaMemcpy(ref.data(), sl.d_refined.p, (size_t)nq * 32, cudaMemcpyDeviceToHost);
        }
        std::vector<int> idx;
        for (int i = 0; i < nq; i++)
            if (q[i].frame == frame) idx.push_back(i);
        std::sort(idx.begin(), idx.end(), [&](int a, int b) { return unpitch_key(q[a].key) < unpitch_key(q[b].key); });
        long long n = (long long)idx.size();
        if (w == "quads") {
            if (cap_bytes >= n * 36)
                for (long long i = 0; i < n; i++) {
                    float* o = (float*)host_out + i * 9;
                    memcpy(o, q[idx[i]].p, 32);
                    o[8] = (float)q[idx[i]].reversed_border;
                }
        } else if (w == "quads_refined") {
            if (cap_bytes >= n * 32)
                for (long long i = 0; i < n; i++) memcpy((float*)host_out + i * 8, &ref[(size_t)idx[i] * 8], 32);
        } else {
            if (cap_bytes >= n * 8)
                for (long long i = 0; i < n; i++) ((unsigned long long*)host_out)[i] = unpitch_key(q[idx[i]].key);
        }
        return n;
    }
    h->set_err("agpu_debug_fetch: unknown buffer name");
    return AGPU_E_INVALID;
}

int agpu_render(agpu_handle* h, const void* tags_host, const int* tag_offsets_host, const uint8_t* backgrounds_host, int B,
                int W, int H, uint8_t* frames_dev, void* cuda_stream) {
    if (!h) return AGPU_E_INVALID;
    if (!tags_host || !tag_offsets_host || !backgrounds_host || !frames_dev || B <= 0 || W <= 0 || H <= 0 || B > 65535) {
        h->set_err("agpu_render: invalid argument");
        return AGPU_E_INVALID;
    }
    CK(cudaSetDevice(h->device));
    const int ntags = tag_offsets_host[B];
    DevBuf d_tags, d_off, d_bg;
    CK(d_tags.ensure(std::max<size_t>(1, (size_t)ntags) * sizeof(RenderTag)));
    CK(d_off.ensure((size_t)(B + 1) * 4));
    CK(d_bg.ensure((size_t)B));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CK(cudaMemcpyAsync(d_tags.p, tags_host, (size_t)ntags * sizeof(RenderTag), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_off.p, tag_offsets_host, (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_bg.p, backgrounds_host, (size_t)B, cudaMemcpyHostToDevice, st));
    dim3 grid(ceil_div(W, 32), ceil_div(H, 8), B);
    k_render<<<grid, 256, 0, st>>>(d_tags.as<RenderTag>(), d_off.as<int>(), d_bg.as<uint8_t>(), frames_dev, W, H, B);
    h->launches = 1;
    cudaError_t le = cudaGetLastError();
    cudaError_t se = cudaStreamSynchronize(st);
    d_tags.release(); d_off.release(); d_bg.release();
    if (le != cudaSuccess || se != cudaSuccess) {
        h->set_err(std::string("agpu_render: ") + cudaGetErrorString(le != cudaSuccess ? le : se));
        return AGPU_E_CUDA;
    }
    return AGPU_OK;
}

int agpu_stage_threshold(agpu_handle* h, const uint8_t* im, int W, int H, uint8_t* quad_im_out, uint8_t* thresh_out) {
    if (!h || !im || !thresh_out || W <= 0 || H <= 0) return AGPU_E_INVALID;
    CK(cudaSetDevice(h->device));
    const Geom g = make_geom(W, H, h->prm.decim);
    Slot& sl = h->slots[0];
    CK(sl.d_in.ensure((size_t)W * H));
    CK(cudaMemcpyAsync(sl.d_in.p, im, (size_t)W * H, cudaMemcpyHostToDevice, sl.stream));
    const uint8_t *quad_im, *gray_full;
    size_t q_pitch, q_frame, gp, gf;
    int rc = run_image_stage(h, sl, sl.d_in.as<uint8_t>(), 1, W, H, W, (size_t)W * H, 1, g, &quad_im, &q_pitch, &q_frame,
                             &gray_full, &gp, &gf);
    if (rc) return rc;
    CK(cudaMemcpy2DAsync(thresh_out, g.wd, sl.d_thresh.p, g.wp, g.wd, g.hd, cudaMemcpyDeviceToHost, sl.stream));
    if (quad_im_out)
        CK(cudaMemcpy2DAsync(quad_im_out, g.wd, quad_im, q_pitch, g.wd, g.hd, cudaMemcpyDeviceToHost, sl.stream));
    CK(cudaStreamSynchronize(sl.stream));
    return AGPU_OK;
}

int agpu_stage_labels(agpu_handle* h, const uint8_t* thresh, int W, int H, uint32_t* labels_out, uint32_t* sizes_out) {
    if (!h || !thresh || !labels_out || W <= 0 || H <= 0) return AGPU_E_INVALID;
    CK(cudaSetDevice(h->device));
    const Geom g = make_geom(W, H, 1);
    Slot& sl = h->slots[0];
    CK(sl.d_thresh.ensure(g.plane));
    CK(cudaMemsetAsync(sl.d_thresh.p, 127, g.plane, sl.stream));
    CK(cudaMemcpy2DAsync(sl.d_thresh.p, g.wp, thresh, W, W, H, cudaMemcpyHostToDevice, sl.stream));
    CK(sl.d_counters.ensure(256));
    CK(cudaMemsetAsync(sl.d_counters.p, 0, 256, sl.stream));
    CcRoots rt;
    // (worst case: every other pixel of a tile row starts a run that is a root -- 512 per tile; a stage hook has no re-run)
    const int sub_cap = cc_tiles_x(g) * ceil_div(cc_tiles_y(g), CC_SUBLISTS) * 512;
    int rc = run_cc_stage(h, sl, sl.d_thresh.as<uint8_t>(), 1, g, sub_cap, sl.d_counters.as<int>(), nullptr, true, rt);
    if (rc) return rc;
    std::vector<uint32_t> lab(g.plane), sz(g.plane);
    CK(cudaMemcpyAsync(lab.data(), sl.d_canon.p, g.plane * 4, cudaMemcpyDeviceToHost, sl.stream));
    CK(cudaMemcpyAsync(sz.data(), sl.d_canon_sizes.p, g.plane * 4, cudaMemcpyDeviceToHost, sl.stream));
    CK(cudaStreamSynchronize(sl.stream));
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            size_t id = (size_t)y * g.wp + x;
            uint32_t l = lab[id];
            labels_out[(size_t)y * W + x] = (l / g.wp) * W + (l % g.wp);
            if (sizes_out) {
                uint32_t v = 0;
                if (l == id) v = (thresh[(size_t)y * W + x] == 127) ? 1u : sz[id];
                sizes_out[(size_t)y * W + x] = v;
            }
        }
    return AGPU_OK;
}


// ---- tag graph (the step after the path) ------------------------------------------------------
int agpu_graph_create(agpu_handle* h, int nstreams, int max_tag_id, agpu_graph** out) {
    if (!h || !out) return AGPU_E_INVALID;
    *out = nullptr;
    if (nstreams <= 0 || max_tag_id < 0 || max_tag_id > (1 << 20)) {
        h->set_err("agpu_graph_create: nstreams > 0 and 0 <= max_tag_id <= 2^20 expected");
        return AGPU_E_INVALID;
    }
    CK(cudaSetDevice(h->device));
    agpu_graph* g = new agpu_graph();
    g->h = h; g->S = nstreams; g->nid = max_tag_id + 1;
    const size_t S = nstreams, n = (size_t)nstreams * g->nid;
    cudaError_t e = cudaSuccess;
    auto need = [&](DevBuf& b, size_t bytes) { if (e == cudaSuccess) e = b.ensure(bytes); };
    need(g->d_coord, S * 4); need(g->d_est, S * 16 * 8); need(g->d_skipped, S * 4);
    need(g->d_present, n); need(g->d_updated, n); need(g->d_visible, n);
    need(g->d_reference, n * 4); need(g->d_weight, n * 4); need(g->d_local, n * 16 * 8); need(g->d_world, n * 16 * 8);
    if (e != cudaSuccess) {
        h->set_err(std::string("agpu_graph_create: ") + cudaGetErrorString(e));
        agpu_graph_destroy(g);
        return AGPU_E_CUDA;
    }
    *out = g;
    return agpu_graph_reset(g);
}

int agpu_graph_reset(agpu_graph* g) {
    if (!g) return AGPU_E_INVALID;
    agpu_handle* h = g->h;
    CK(cudaSetDevice(h->device));
    const size_t S = g->S, n = (size_t)g->S * g->nid;
    CK(cudaMemset(g->d_coord.p, 0xff, S * 4));   // coordinate_id = -1
    CK(cudaMemset(g->d_est.p, 0, S * 16 * 8));
    CK(cudaMemset(g->d_skipped.p, 0, S * 4));
    CK(cudaMemset(g->d_present.p, 0, n)); CK(cudaMemset(g->d_updated.p, 0, n)); CK(cudaMemset(g->d_visible.p, 0, n));
    CK(cudaMemset(g->d_reference.p, 0xff, n * 4)); CK(cudaMemset(g->d_weight.p, 0, n * 4));
    CK(cudaMemset(g->d_local.p, 0, n * 16 * 8)); CK(cudaMemset(g->d_world.p, 0, n * 16 * 8));
    return AGPU_OK;
}

int agpu_graph_destroy(agpu_graph* g) {
    if (!g) return AGPU_E_INVALID;
    cudaSetDevice(g->h->device);
    DevBuf* bufs[] = {&g->d_coord, &g->d_est, &g->d_present, &g->d_updated, &g->d_visible, &g->d_reference, &g->d_weight,
                      &g->d_local, &g->d_world, &g->d_skipped, &g->d_dets, &g->d_poses, &g->d_counts, &g->d_my_pose, &g->d_valid};
    for (DevBuf* b : bufs) b->release();
    delete g;
    return AGPU_OK;
}

int agpu_graph_update(agpu_graph* g, int F, const agpu_detection* dets, const agpu_pose_t* poses, const int* counts,
                      int cap_per_frame, double* my_pose, uint8_t* valid) {
    if (!g) return AGPU_E_INVALID;
    agpu_handle* h = g->h;
    if (F <= 0 || cap_per_frame <= 0 || !dets || !poses || !counts || !my_pose || !valid) {
        h->set_err("agpu_graph_update: invalid argument");
        return AGPU_E_INVALID;
    }
    CK(cudaSetDevice(h->device));
    const size_t SF = (size_t)g->S * F, recs = SF * cap_per_frame;
    CK(g->d_dets.ensure(recs * sizeof(DetRec))); CK(g->d_poses.ensure(recs * sizeof(PoseRec)));
    CK(g->d_counts.ensure(SF * 4)); CK(g->d_my_pose.ensure(SF * 16 * 8)); CK(g->d_valid.ensure(SF));
    cudaStream_t st = h->slots[0].stream;
    CK(cudaMemcpyAsync(g->d_dets.p, dets, recs * sizeof(DetRec), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(g->d_poses.p, poses, recs * sizeof(PoseRec), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(g->d_counts.p, counts, SF * 4, cudaMemcpyHostToDevice, st));
    GraphArgs a;
    a.S = g->S; a.F = F; a.cap = cap_per_frame; a.nid = g->nid;
    a.dets = g->d_dets.as<DetRec>(); a.poses = g->d_poses.as<PoseRec>(); a.counts = g->d_counts.as<int>();
    a.coordinate_id = g->d_coord.as<int>(); a.estimated_pose = g->d_est.as<double>();
    a.present = g->d_present.as<unsigned char>(); a.updated = g->d_updated.as<unsigned char>();
    a.visible = g->d_visible.as<unsigned char>(); a.reference = g->d_reference.as<int>(); a.weight = g->d_weight.as<int>();
    a.local = g->d_local.as<double>(); a.world = g->d_world.as<double>();
    a.my_pose = g->d_my_pose.as<double>(); a.valid = g->d_valid.as<unsigned char>(); a.skipped = g->d_skipped.as<int>();
    h->launches = 0;
    k_graph_update<<<ceil_div(g->S, 32), 32, 0, st>>>(a);
    LAUNCH_CHECK("k_graph_update");
    CK(cudaMemcpyAsync(my_pose, g->d_my_pose.p, SF * 16 * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(valid, g->d_valid.p, SF, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return AGPU_OK;
}

int agpu_graph_get(agpu_graph* g, int stream, int* coordinate_id, double* estimated_pose, uint8_t* present, int* reference,
                   int* weight, uint8_t* updated, uint8_t* visible, double* local, double* world, int* skipped) {
    if (!g) return AGPU_E_INVALID;
    agpu_handle* h = g->h;
    if (stream < 0 || stream >= g->S) {
        h->set_err("agpu_graph_get: stream out of range");
        return AGPU_E_INVALID;
    }
    CK(cudaSetDevice(h->device));
    const size_t n = g->nid, o = (size_t)stream * n;
    if (coordinate_id) CK(cudaMemcpy(coordinate_id, g->d_coord.as<int>() + stream, 4, cudaMemcpyDeviceToHost));
    if (skipped) CK(cudaMemcpy(skipped, g->d_skipped.as<int>() + stream, 4, cudaMemcpyDeviceToHost));
    if (estimated_pose) CK(cudaMemcpy(estimated_pose, g->d_est.as<double>() + (size_t)stream * 16, 128, cudaMemcpyDeviceToHost));
    if (present) CK(cudaMemcpy(present, g->d_present.as<unsigned char>() + o, n, cudaMemcpyDeviceToHost));
    if (updated) CK(cudaMemcpy(updated, g->d_updated.as<unsigned char>() + o, n, cudaMemcpyDeviceToHost));
    if (visible) CK(cudaMemcpy(visible, g->d_visible.as<unsigned char>() + o, n, cudaMemcpyDeviceToHost));
    if (reference) CK(cudaMemcpy(reference, g->d_reference.as<int>() + o, n * 4, cudaMemcpyDeviceToHost));
    if (weight) CK(cudaMemcpy(weight, g->d_weight.as<int>() + o, n * 4, cudaMemcpyDeviceToHost));
    if (local) CK(cudaMemcpy(local, g->d_local.as<double>() + o * 16, n * 128, cudaMemcpyDeviceToHost));
    if (world) CK(cudaMemcpy(world, g->d_world.as<double>() + o * 16, n * 128, cudaMemcpyDeviceToHost));
    return AGPU_OK;
}

}  // extern "C"
