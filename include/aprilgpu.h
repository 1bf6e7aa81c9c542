/* aprilgpu.h -- C ABI of libaprilgpu.so, the B200 (sm_100a) AprilTag detect + per-tag pose library.
 *
 * This is the drop-in boundary for the one hot path of mikostrzewa/AprilSLAM.  Each entry point
 * names the reference interface it replaces (paths relative to /root/reference):
 *
 *   agpu_create        <-  apriltag(tag_type)                     src/detection/tag_detector.py:18
 *                          (upstream apriltag_pywrap.c ctor kwargs: family, threads, maxhamming,
 *                           decimate, blur, refine_edges, debug)
 *   agpu_detect        <-  self.detector.detect(gray)             src/detection/tag_detector.py:26
 *   agpu_detect_bgr    <-  cv2.cvtColor(image, BGR2GRAY) + detect src/detection/tag_detector.py:25-26
 *   agpu_pose          <-  cv2.solvePnP(obj, corners, K, dist)    src/detection/tag_detector.py:41
 *                          + cv2.Rodrigues(rvec) -> 4x4 T         src/detection/tag_detector.py:45-52
 *   agpu_detect_pose   <-  the caller loop  detect(); for d in detections: get_pose(d)
 *                          src/core/slam.py:21-32, src/simulation/simulation_engine.py:219-223
 *   agpu_graph_update  <-  SLAMGraph.add_or_update_node + SLAM.my_pose (the consumer of the path)
 *                          src/core/slam_graph.py:29-70, src/core/slam.py:36-63
 *
 * Plain C: pointers and sizes only.  Nothing crosses this boundary as a C++ or torch type, and
 * no exception leaves the library.  Every function returns AGPU_OK (0) or a negative agpu_status;
 * agpu_last_error() gives the message.  There is no CPU fallback: if no CUDA device is usable,
 * agpu_create fails with AGPU_E_CUDA.
 *
 * Threading: a handle is bound to one CUDA device and one internal stream set; it is not
 * thread-safe.  Distinct handles are independent (one per GPU / per host thread).
 */
#ifndef APRILGPU_H
#define APRILGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGPU_VERSION 100

typedef enum agpu_status {
    AGPU_OK = 0,
    AGPU_E_INVALID = -1,    /* bad argument (NULL, non-positive size, unknown family, ...) */
    AGPU_E_CUDA = -2,       /* CUDA runtime error or no device */
    AGPU_E_TRUNCATED = -3,  /* more detections than cap_per_frame in at least one frame (counts[] hold the true numbers), or more
                               than 1024 decoded candidates in one frame before reconcile (the surplus was dropped) */
    AGPU_E_WORKSPACE = -4,  /* an internal per-frame work list overflowed (edge points / clusters / quads); raise the agpu_config limits */
    AGPU_E_UNSUPPORTED = -5
} agpu_status;

typedef struct agpu_handle agpu_handle;

/* Constructor arguments.  The first block mirrors upstream's Python ctor exactly (defaults in
 * parentheses are what AprilSLAM gets, because it passes only the family name). */
typedef struct agpu_config {
    const char* families;     /* space separated: tag36h11 tag25h9 tag16h5 tagStandard41h12 (ids 0..4 only, see agpu_family_info) */
    int threads;              /* (1)  accepted for API compatibility, ignored */
    int maxhamming;           /* (1)  0..2 */
    float quad_decimate;      /* (2.0) integer factors >= 1 */
    float quad_sigma;         /* (0.0) "blur": >0 Gaussian blur, <0 unsharp; fused with the decimation in one kernel */
    int refine_edges;         /* (1) */
    double decode_sharpening; /* (0.25) */
    int debug;                /* (0)  1: keep stage buffers of the last chunk for agpu_debug_fetch */
    /* B200 side */
    int device;               /* CUDA device ordinal */
    int chunk_frames;         /* frames processed per pipeline pass (0 = auto: ~256 Mpx of working image for device frames, ~64 Mpx for host frames) */
    int pipeline_slots;       /* chunks in flight on independent streams (0 = auto: 3 for device frames, 4 for host frames) */
    int max_points_per_frame; /* edge-point list capacity per frame (0 = auto: decimated pixels / 4, >= 65536) */
    int max_clusters_per_frame; /* (0 = auto) */
    int max_quads_per_frame;  /* (0 = auto: 1024) */
} agpu_config;

/* One detection.  Field meaning follows upstream's apriltag_detection_t; the Python layer turns it
 * into the dict the reference indexes ('id', 'lb-rb-rt-lt', ...; tag_detector.py:27,32). */
typedef struct agpu_detection {
    int32_t family;    /* index into the configured family list */
    int32_t id;
    int32_t hamming;
    float margin;      /* decision_margin */
    double c[2];       /* center */
    double p[4][2];    /* corners: lb, rb, rt, lt (image y down) */
    double H[9];       /* homography, row major: tag (+-1) -> pixels */
} agpu_detection;

/* One pose, as returned by TagDetector.get_pose (tag_detector.py:43): retval, rvec, tvec (+R of T). */
typedef struct agpu_pose_t {
    double rvec[3];
    double tvec[3];
    double R[9];       /* Rodrigues(rvec), row major: T = [R t; 0 1] (tag_detector.py:45-52) */
    double err;        /* final RMS reprojection error in pixels */
    int32_t ok;        /* solvePnP's retval */
    int32_t iters;
} agpu_pose_t;

int agpu_version(void);
/* Code words a built-in family ships with, and how many upstream's table has.  They differ for tagStandard41h12 (the
 * reference's default tag_type, tag_detector.py:17): upstream has 2115 code words, only ids 0..4 -- decoded from the
 * reference's own assets/tags/tag0..4.png -- are available offline, so a detector of that family reports ids 0..4 only
 * (everything the reference's simulator scene contains; NOT a drop-in for tags with id >= 5).  tag36h11 (587), tag25h9
 * (35) and tag16h5 (30) are complete. */
int agpu_family_info(const char* family, int* ncodes, int* ncodes_upstream);
void agpu_default_config(agpu_config* cfg);
int agpu_create(const agpu_config* cfg, agpu_handle** out);
int agpu_destroy(agpu_handle* h);
const char* agpu_last_error(const agpu_handle* h);  /* h may be NULL: error of the last failed agpu_create */

/* Detect on B gray frames of W x H, `stride` bytes between rows, frames contiguous
 * (frame b starts at frames + b*H*stride).  on_device: 0 = host memory (copied H2D inside the
 * call), 1 = device memory on cfg.device.  cuda_stream: a cudaStream_t the input is ready on
 * (NULL = default stream); the call returns after results are in host memory.
 * out: host array [B][cap_per_frame]; counts: host int[B] (true number found per frame). */
int agpu_detect(agpu_handle* h, const uint8_t* frames, int on_device, int B, int W, int H, int stride,
                void* cuda_stream, agpu_detection* out, int cap_per_frame, int* counts);

/* Same, on interleaved BGR frames [B][H][W][3] (stride = bytes per row, >= 3*W); gray conversion
 * is cv2.cvtColor(BGR2GRAY)'s fixed-point formula.  At quad_decimate 1, 2 and 4 without blur it is fused into the strip
 * kernel: gray conversion, decimation and threshold are ONE pass over the BGR bytes (k_decimate_threshold<F,.,.,3>);
 * other settings convert first (k_pack) and continue on the gray plane. */
int agpu_detect_bgr(agpu_handle* h, const uint8_t* frames, int on_device, int B, int W, int H, int stride,
                    void* cuda_stream, agpu_detection* out, int cap_per_frame, int* counts);

/* Detect + per-tag pose in one pass (poses[b][i] belongs to out[b][i]).
 * K: 3x3 row major camera matrix; dist: ndist (0,4,5,8) OpenCV distortion coefficients or NULL. */
int agpu_detect_pose(agpu_handle* h, const uint8_t* frames, int on_device, int channels, int B, int W, int H,
                     int stride, void* cuda_stream, const double K[9], const double* dist, int ndist,
                     double tag_size, agpu_detection* out, agpu_pose_t* poses, int cap_per_frame, int* counts);

/* Pose only: M tags, corners [M][4][2] (lb, rb, rt, lt) in host memory.
 * method 0: homography decomposition + Levenberg-Marquardt on the pixel reprojection error
 *           (lands on cv2.solvePnP(SOLVEPNP_ITERATIVE)'s minimiser -- the reference's path);
 * method 1: homography decomposition + orthogonal iteration (object-space error). */
int agpu_pose(agpu_handle* h, const double* corners, int M, const double K[9], const double* dist, int ndist,
              double tag_size, int method, agpu_pose_t* poses);

/* ---- instrumentation -------------------------------------------------------------------- */

enum { AGPU_STAGE_H2D = 0, AGPU_STAGE_IMAGE, AGPU_STAGE_CC, AGPU_STAGE_EDGES, AGPU_STAGE_SORT,
       AGPU_STAGE_QUADS, AGPU_STAGE_DECODE, AGPU_STAGE_RECONCILE_POSE, AGPU_STAGE_D2H, AGPU_NUM_STAGES };

/* CUDA-event timing of the stages of the last agpu_detect* call (milliseconds summed over its
 * chunks, measured on the library's own stream).  on != 0 enables it (adds event records only). */
int agpu_set_profiling(agpu_handle* h, int on);
int agpu_get_stage_ms(agpu_handle* h, float* ms /* [AGPU_NUM_STAGES] */);
/* Profiling on: EVERY kernel launch of a chunk is bracketed by its own pair of CUDA events on the stream it is launched
 * on (and the quad-fit size tiers, normally concurrent on side streams, run one after the other), so with
 * pipeline_slots = 1 each interval is that kernel alone.  agpu_get_kernel_ms: milliseconds of one kernel summed over
 * the chunks of the last call; names: k_decimate_threshold, k_decimate_blur, k_pack(bgr), k_cc_local, k_cc_boundary, k_cc_sizes,
 * k_cc_dense, k_edges, k_cluster_refs, k_sort_scatter, k_fit_quads<1>, <2>, <4>, <4>/6k, <8>,
 * k_decode_quads, k_reconcile, k_pose.  agpu_get_kernel_table: all of them as text lines "name\tms\tlaunches\n";
 * returns the bytes needed (terminator included) and copies at most cap. */
int agpu_get_kernel_ms(agpu_handle* h, const char* kernel, float* ms);
int agpu_get_kernel_table(agpu_handle* h, char* buf, int cap);
/* Timeline of the last call (profiling on): per finished chunk AGPU_NUM_STAGES + 4 floats {first frame, frames, slot,
 * stage boundary marks in ms since the start of the call}.  Returns the number of floats available; copies at most
 * cap_floats of them.  Shows how the chunks in flight overlap. */
int agpu_get_timeline(agpu_handle* h, float* out, int cap_floats);
/* Kernel launches issued by the last agpu_detect* / agpu_pose call. */
int agpu_get_launch_count(agpu_handle* h, long long* launches);
/* Work counters of the last call, summed over frames: [0] edge points, [1] clusters fitted,
 * [2] quads, [3] detections before reconcile, [4] clusters over upstream's size limit of 3(2w+2h) raw points
 * (dropped before fitting, exactly as upstream drops them), [5] / [6] / [7] clusters of up to 1024 / up to 2048 / more
 * records handed to the multi-warp tiers of the quad-fit kernel ([1] counts all tiers). */
int agpu_get_counters(agpu_handle* h, long long* counters /* [8] */);
/* Quad-fit size classes in the last call (clusters of up to 256 / 1024 / 2048 / more records: k_fit_quads<1>, <2>, <4>, and
 * the two large tiers together): [0..3] clusters, [4..7] edge-point records handed to each class (8 bytes each: the
 * algorithmic input of those kernels). */
int agpu_get_tier_stats(agpu_handle* h, long long* stats /* [8] */);

/* Stage dumps for parity tests (cfg.debug = 1): buffers of frame `frame` of the LAST chunk.
 * what: "gray" u8[H*W] (BGR input only: the converted full-resolution plane), "quad_im" u8[hd*wd], "thresh" u8[hd*wd], "labels" u32[hd*wd] (min-index representative),
 * "sizes" u32[hd*wd] (valid at representatives), "cluster_keys" u64[n], "cluster_sizes" i32[n],
 * "quads" f32[n*9] (8 corner coords + reversed flag), "quad_keys" u64[n].
 * Returns the number of ELEMENTS available (copying at most cap_bytes), or a negative status. */
long long agpu_debug_fetch(agpu_handle* h, const char* what, int frame, void* host_out, long long cap_bytes);
int agpu_debug_dims(agpu_handle* h, int* wd, int* hd);

/* Stand-alone image stages on host buffers (stage-level parity tests; each runs the same kernels
 * the pipeline uses). */
int agpu_stage_threshold(agpu_handle* h, const uint8_t* im, int W, int H, uint8_t* quad_im_out, uint8_t* thresh_out);
int agpu_stage_labels(agpu_handle* h, const uint8_t* thresh, int W, int H, uint32_t* labels_out, uint32_t* sizes_out);

/* ---- synthetic frame source (the step before the path) -------------------------- */

/* Renders B gray frames of W x H into DEVICE memory frames_dev[B][H][W] with the geometry of the reference's
 * OpenGL renderer (src/simulation/renderer.py:91-96,188-251,253-274), bit-identical to aprilslam_b200/synth.py.
 * tags_host: array of 112-byte records {double Gi[9]; uint64 cells[2]; int32 total_width, ppc, x0, x1, y0, y1}
 * (aprilslam_b200.render.TAG_DTYPE); frame b owns tags [tag_offsets[b], tag_offsets[b+1]); backgrounds: one gray
 * level per frame.  The call returns after the frames are complete. */
int agpu_render(agpu_handle* h, const void* tags_host, const int* tag_offsets_host, const uint8_t* backgrounds_host,
                int B, int W, int H, uint8_t* frames_dev, void* cuda_stream);

/* ---- tag graph + camera pose estimate (the step after the path) ------------------- */

/* The graph update of SLAMGraph.add_or_update_node / find_world / get_world (src/core/slam_graph.py:29-70) and
 * the weighted pose average of SLAM.my_pose (src/core/slam.py:36-63), batched over S independent camera streams
 * (multi-camera rigs, replayed sequences).  The update is sequential inside a stream, so the unit of parallelism is
 * the stream; every stream's graph stays in device memory between calls.  Tag ids 0..max_tag_id. */
typedef struct agpu_graph agpu_graph;
int agpu_graph_create(agpu_handle* h, int nstreams, int max_tag_id, agpu_graph** out);
int agpu_graph_reset(agpu_graph* g);
int agpu_graph_destroy(agpu_graph* g);

/* F consecutive frames for each of the S streams, as agpu_detect_pose returned them (records in ascending id order,
 * TagDetector.detect's order): dets / poses host arrays [S][F][cap_per_frame], counts [S][F].  Per frame the call
 * replays the reference's caller loop (simulation_engine.py:219-232): visible_tags = all ids of the frame; for every
 * detection whose pose succeeded, add_or_update_node(id, T, visible_tags) with T = [R t; 0 1]; then my_pose().
 * Outputs (host): my_pose [S][F][16] row-major 4x4, valid [S][F] (0 where my_pose() returns None). */
int agpu_graph_update(agpu_graph* g, int F, const agpu_detection* dets, const agpu_pose_t* poses, const int* counts,
                      int cap_per_frame, double* my_pose, uint8_t* valid);

/* Node table of one stream (any output pointer may be NULL): arrays of max_tag_id+1 entries
 * (present = the id is a key of SLAMGraph.graph; local / world 4x4 row major; reference, weight, updated, visible =
 * the Node fields, slam_graph.py:5-12), coordinate_id, estimated_pose[16], and the number of detections the graph
 * could not place so far (the reference prints "Cannot find world reference"). */
int agpu_graph_get(agpu_graph* g, int stream, int* coordinate_id, double* estimated_pose, uint8_t* present, int* reference,
                   int* weight, uint8_t* updated, uint8_t* visible, double* local, double* world, int* skipped);

#ifdef __cplusplus
}
#endif
#endif /* APRILGPU_H */
