// common.cuh -- shared device/host definitions of libaprilgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define AGPU_WARP 32
#define FULL_MASK 0xffffffffu

struct FamilyDef {  // layout expected by families_data.inc (tools/gen_codebooks.py)
    const char* name;
    int nbits, h, ncodes, width_at_border, total_width, reversed_border;
    const unsigned long long* codes;
    const int* bit_x;
    const int* bit_y;
};

// Family table as the decode kernel sees it (device memory).
struct DevFamily {
    int nbits, ncodes, width_at_border, total_width, reversed_border;
    int code_offset;      // into the concatenated code array
    int bit_x[64], bit_y[64];
};

#define AGPU_MAX_FAMILIES 4

// Quad-threshold / decode parameters (upstream defaults: SURVEY.md appendix A.1)
struct DevParams {
    float quad_decimate;
    int decim;                 // (int)quad_decimate
    int refine_edges;
    double decode_sharpening;
    int maxhamming;
    int min_cluster_pixels;    // 5
    int max_nmaxima;           // 10
    float cos_critical_rad;    // cos(10 deg)
    float max_line_fit_mse;    // 10
    int min_white_black_diff;  // 5
    int min_tag_width;         // min family width_at_border / decimate, >= 3
    int normal_border, reversed_border;
    int nfamilies;
    float smooth_f[7];         // (float)exp(-j*j/2), j=-3..3 -- computed on the host like upstream does
    double rot_c[4], rot_s[4]; // cos/sin(rotation*pi/2) from the host libm
};

// Per-frame geometry of the working (decimated) image.  u8 planes and the u32 label plane share
// the pitch `wp` (wd rounded up to 16) so that every row start is 16-byte aligned; a pixel id is
// y*wp + x everywhere on the device.
struct Geom {
    int W, H;        // source frame
    int wd, hd;      // decimated image
    int wp;          // pitch of quad_im / thresh / labels planes (elements)
    size_t plane;    // wp*hd
};

// packed edge point: px (14b) | py (14b) << 14 | kind (4b) << 28
//   kind 0..7 : one point, kind = dir | positive << 2   (dir 0..3 = offsets (1,0) (0,1) (-1,1) (1,1))
//   kind 8..11: the point of direction 2 at (x, y) MERGED with the identical point that direction 3 emits at
//               (x-1, y) -- always the same pair of components, see k_edges -- kind = 8 | positive(dir 2) |
//               positive(dir 3) << 1.  Upstream emits both and drops the duplicate after the slope sort; merging
//               them at emission keeps a third of all points out of both sorts.
__host__ __device__ inline uint32_t pack_point(int px, int py, int kind) {
    return (uint32_t)px | ((uint32_t)py << 14) | ((uint32_t)kind << 28);
}

struct ClusterRef {
    int frame;   // frame index within the chunk
    int start;   // offset inside the frame's point segment
    int size;
    int pad;
};

struct QuadRec {
    float p[4][2];   // full-resolution corners (decimation undone)
    int frame;
    int reversed_border;
    unsigned long long key;  // (rep_hi, rep_lo) of the cluster, pitched ids
};

struct DetRec {  // == agpu_detection
    int32_t family, id, hamming;
    float margin;
    double c[2];
    double p[4][2];
    double H[9];
};

struct PoseRec {  // == agpu_pose_t
    double rvec[3], tvec[3], R[9], err;
    int32_t ok, iters;
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL_MASK, v, src); }

__device__ __forceinline__ uint32_t float_orderable(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ unsigned long long double_orderable(double d) {
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
