// k_render.cuh -- synthetic frame source on the GPU: the per-pixel restatement of the
// reference's OpenGL renderer (/root/reference/src/simulation/renderer.py:91-96 projection, :188-195 view,
// :222-251 one textured quad per tag, :253-274 read-back) that aprilslam_b200/synth.py evaluates with numpy.
// Same arithmetic, same order of operations (double precision, no FMA contraction), so frames are
// bit-identical to synth.render(); it exists so that benchmark batches (1024 distinct 1080p frames) are
// generated in HBM instead of being rendered on the host and copied.  Not part of the detector.
#pragma once
#include "common.cuh"

struct RenderTag {            // one textured quad; == aprilslam_b200.render.TAG_DTYPE
    double Gi[9];             // pixel (x, y, 1) -> (s*q, t*q, q), q = 1/depth (synth.tag_records)
    unsigned long long cells[2];   // total_width^2 cell bits, row major, bit 0 = top-left cell, 1 = white
    int total_width, ppc;     // cells per side, texels per cell
    int x0, x1, y0, y1;       // pixel bounding box [x0, x1) x [y0, y1)
};

__global__ void __launch_bounds__(256)
k_render(const RenderTag* __restrict__ tags, const int* __restrict__ tag_offsets, const uint8_t* __restrict__ background,
         uint8_t* __restrict__ frames, int W, int H, int nframes) {
    const int frame = blockIdx.z;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const int t0 = tag_offsets[frame], t1 = tag_offsets[frame + 1];
    const double X = x + 0.5, Y = y + 0.5;
    double zbest = 0.0;
    int out = background[frame];
    for (int ti = t0; ti < t1; ti++) {
        const RenderTag& T = tags[ti];
        if (x < T.x0 || x >= T.x1 || y < T.y0 || y >= T.y1) continue;
        const double a = T.Gi[0] * X + T.Gi[1] * Y + T.Gi[2];
        const double b = T.Gi[3] * X + T.Gi[4] * Y + T.Gi[5];
        const double q = T.Gi[6] * X + T.Gi[7] * Y + T.Gi[8];
        if (!(q > 0)) continue;
        const double s = a / q, t = b / q;
        if (!(s >= 0 && s <= 1 && t >= 0 && t <= 1)) continue;
        if (!(q > zbest)) continue;
        // GL_LINEAR + GL_REPEAT in the (total_width*ppc)^2 virtual texture, t = 0 at the bottom
        const int tw = T.total_width, ppc = T.ppc, n = tw * ppc;
        const double xt = s * n - 0.5, yt = (1.0 - t) * n - 0.5;
        const double fx0 = floor(xt), fy0 = floor(yt);
        const double ax = xt - fx0, ay = yt - fy0;
        const long long ix0 = (long long)fx0, iy0 = (long long)fy0;
        auto wrap = [n](long long v) { long long m = v % n; return (int)(m < 0 ? m + n : m); };
        const int cx0 = wrap(ix0) / ppc, cx1 = wrap(ix0 + 1) / ppc, cy0 = wrap(iy0) / ppc, cy1 = wrap(iy0 + 1) / ppc;
        auto cell = [&](int cy, int cx) {
            const int bit = cy * tw + cx;
            return ((T.cells[bit >> 6] >> (bit & 63)) & 1ull) ? 255.0 : 0.0;
        };
        const double v = (1 - ax) * (1 - ay) * cell(cy0, cx0) + ax * (1 - ay) * cell(cy0, cx1) +
                         (1 - ax) * ay * cell(cy1, cx0) + ax * ay * cell(cy1, cx1);
        out = (int)floor(v + 0.5);
        zbest = q;
    }
    frames[((size_t)frame * H + y) * W + x] = (uint8_t)out;
}
