// k_decode.cuh -- refine_edges, homography, bit sampling + Hamming decode (one warp per quad), and
// per-frame reconcile + sort.  Upstream stages U8-U10, the last part of the
// native call at /root/reference/src/detection/tag_detector.py:26; the result order (ascending id)
// is what tag_detector.py:27 re-establishes with sorted(key=id).
//
// Lanes share the data-parallel parts (edge samples, border strips, bit cells, code-book search);
// every reduction that feeds a floating-point decision is accumulated in the same order as the
// sequential algorithm (values are exchanged with shuffles and summed identically on all lanes).
#pragma once
#include "common.cuh"

struct DecodeArgs {
    const uint8_t* im;          // full-resolution gray frames
    size_t pitch, frame_stride; // bytes
    int W, H;
    const QuadRec* quads;
    const int* nquads;
    int cap_quads;
    const DevFamily* fams;
    const unsigned long long* codes;
    DetRec* dets;               // [nframes][cap_dets]
    int* ndets;                 // [nframes]
    int cap_dets;
    float* dbg_refined;         // optional [cap_quads][8]
};

__device__ __forceinline__ double bilinear_at(const uint8_t* im, size_t pitch, int xi, int yi, double a, double b) {
    return (1 - a) * (1 - b) * im[(size_t)yi * pitch + xi] + a * (1 - b) * im[(size_t)yi * pitch + xi + 1] +
           (1 - a) * b * im[(size_t)(yi + 1) * pitch + xi] + a * b * im[(size_t)(yi + 1) * pitch + xi + 1];
}

// refine_edges: the (edge, sample) pairs of all four edges are dealt to the lanes as ONE flat list (a typical tag has
// 4 x 16 samples: two full rounds instead of four half-empty ones); every sample runs the sequential search along the
// edge normal on its own lane, and the per-edge moments are then accumulated in sample order (identical on every lane),
// exactly the order of the sequential algorithm.
// position of sample k's term in row m of the staging buffer: skewed by the row, so the nine lanes that walk the nine
// rows side by side -- each at the same k -- hit nine different banks
#define DEC_STAGE_AT(m, k) ((m) * 32 + (((k) + (m)) & 31))
__device__ void refine_edges_warp(const DevParams& P, const uint8_t* im, size_t pitch, int width, int height,
                                  float (&p)[4][2], int reversed, double* stage) {
    const int lane = threadIdx.x & 31;
    double enx[4], eny[4];
    int ens[4], eoff[5];
    eoff[0] = 0;
#pragma unroll
    for (int edge = 0; edge < 4; edge++) {
        const int a = edge, b = (edge + 1) & 3;
        double nx = p[b][1] - p[a][1];
        double ny = -p[b][0] + p[a][0];
        double mag = sqrt(nx * nx + ny * ny);
        nx /= mag;
        ny /= mag;
        if (reversed) { nx = -nx; ny = -ny; }
        enx[edge] = nx;
        eny[edge] = ny;
        ens[edge] = max(16, (int)(mag / 8));
        eoff[edge + 1] = eoff[edge] + ens[edge];
    }
    const int total = eoff[4];
    const double range = P.quad_decimate + 1;
    const int nsteps = (int)(2 * range * 4) + 1;
    double lines[4][4];
    double Mx = 0, My = 0, Mxx = 0, Mxy = 0, Myy = 0, N = 0;
    double macc = 0;      // lane m < 6: running sum of moment m of the current edge
    int cur = 0;          // edge whose moments are being accumulated (uniform over the warp)
    int cur_end = eoff[1];
    auto finish_edge = [&]() {
        double Ex = Mx / N, Ey = My / N;
        double Cxx = Mxx / N - Ex * Ex, Cxy = Mxy / N - Ex * Ey, Cyy = Myy / N - Ey * Ey;
        double normal_theta = .5 * atan2f((float)(-2 * Cxy), (float)(Cyy - Cxx));
        const double c = cosf((float)normal_theta), s = sinf((float)normal_theta);
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (e == cur) { lines[e][0] = Ex; lines[e][1] = Ey; lines[e][2] = c; lines[e][3] = s; }
        cur++;
        cur_end = cur == 1 ? eoff[2] : (cur == 2 ? eoff[3] : eoff[4]);
        Mx = My = Mxx = Mxy = Myy = N = 0;
    };
    for (int base = 0; base < total; base += 32) {
        const int item = base + lane;
        double bestx = 0, besty = 0;
        int valid = 0;
        if (item < total) {
            const int e = (item >= eoff[1]) + (item >= eoff[2]) + (item >= eoff[3]);
            double nx = enx[0], ny = eny[0];
            float pax = p[0][0], pay = p[0][1], pbx = p[1][0], pby = p[1][1];
            int nsamples = ens[0], s = item;
#pragma unroll
            for (int k = 1; k < 4; k++)
                if (e == k) {
                    nx = enx[k]; ny = eny[k]; nsamples = ens[k]; s = item - eoff[k];
                    pax = p[k][0]; pay = p[k][1]; pbx = p[(k + 1) & 3][0]; pby = p[(k + 1) & 3][1];
                }
            double alpha = (1.0 + s) / (nsamples + 1);
            double x0 = alpha * pax + (1 - alpha) * pbx;
            double y0 = alpha * pay + (1 - alpha) * pby;
            double Mn = 0, Mcount = 0;
            for (int k = 0; k < nsteps; k++) {
                double n = -range + 0.25 * k;
                double grange = 1;
                double x1 = x0 + (n + grange) * nx - 0.5;
                double y1 = y0 + (n + grange) * ny - 0.5;
                int x1i = (int)floor(x1), y1i = (int)floor(y1);
                double a1 = x1 - x1i, b1 = y1 - y1i;
                if (x1i < 0 || x1i + 1 >= width || y1i < 0 || y1i + 1 >= height) continue;
                double x2 = x0 + (n - grange) * nx - 0.5;
                double y2 = y0 + (n - grange) * ny - 0.5;
                int x2i = (int)floor(x2), y2i = (int)floor(y2);
                double a2 = x2 - x2i, b2 = y2 - y2i;
                if (x2i < 0 || x2i + 1 >= width || y2i < 0 || y2i + 1 >= height) continue;
                double g1 = bilinear_at(im, pitch, x1i, y1i, a1, b1);
                double g2 = bilinear_at(im, pitch, x2i, y2i, a2, b2);
                if (g1 < g2) continue;
                double weight = (g2 - g1) * (g2 - g1);
                Mn += weight * n;
                Mcount += weight;
            }
            if (Mcount != 0) {
                double n0 = Mn / Mcount;
                bestx = x0 + n0 * nx;
                besty = y0 + n0 * ny;
                valid = 1;
            }
        }
        // Sequential-order accumulation of the six moments (the order of the sequential algorithm -- the sums feed
        // decisions).  Every lane parks the six terms of ITS sample in the staging buffer; lane m < 6 then runs the
        // sequential sum of moment m over the samples (one LDS + one DADD per sample), and the six sums are handed to
        // all lanes only where an edge ends -- instead of every lane replaying every sample through five shuffles.
        const int cnt = min(32, total - base);
        {
            const double v1 = valid ? 1.0 : 0.0, bx = valid ? bestx : 0.0, by = valid ? besty : 0.0;
            stage[DEC_STAGE_AT(0, lane)] = bx; stage[DEC_STAGE_AT(1, lane)] = by; stage[DEC_STAGE_AT(2, lane)] = bx * bx;
            stage[DEC_STAGE_AT(3, lane)] = bx * by; stage[DEC_STAGE_AT(4, lane)] = by * by; stage[DEC_STAGE_AT(5, lane)] = v1;
        }
        __syncwarp();
        const int mrow = lane < 6 ? lane : 0;
        int k = 0;
        while (k < cnt) {
            if (base + k >= cur_end) {   // (uniform) the edge ended with the previous sample: its sums to all lanes
                Mx = __shfl_sync(FULL_MASK, macc, 0); My = __shfl_sync(FULL_MASK, macc, 1); Mxx = __shfl_sync(FULL_MASK, macc, 2);
                Mxy = __shfl_sync(FULL_MASK, macc, 3); Myy = __shfl_sync(FULL_MASK, macc, 4); N = __shfl_sync(FULL_MASK, macc, 5);
                finish_edge();
                macc = 0;
            }
            const int kend = min(cnt, cur_end - base);
#pragma unroll 4
            for (; k < kend; k++) macc += stage[DEC_STAGE_AT(mrow, k)];
        }
        __syncwarp();
    }
    Mx = __shfl_sync(FULL_MASK, macc, 0); My = __shfl_sync(FULL_MASK, macc, 1); Mxx = __shfl_sync(FULL_MASK, macc, 2);
    Mxy = __shfl_sync(FULL_MASK, macc, 3); Myy = __shfl_sync(FULL_MASK, macc, 4); N = __shfl_sync(FULL_MASK, macc, 5);
    finish_edge();   // the fourth edge (every edge has >= 16 samples, so edges 0..2 were closed inside the loop)
    float np[4][2];
#pragma unroll
    for (int i = 0; i < 4; i++) { np[i][0] = p[i][0]; np[i][1] = p[i][1]; }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double A00 = lines[i][3], A01 = -lines[(i + 1) & 3][3];
        double A10 = -lines[i][2], A11 = lines[(i + 1) & 3][2];
        double B0 = -lines[i][0] + lines[(i + 1) & 3][0];
        double B1 = -lines[i][1] + lines[(i + 1) & 3][1];
        double det = A00 * A11 - A10 * A01;
        if (fabs(det) > 0.001) {
            double W00 = A11 / det, W01 = -A01 / det;
            double L0 = W00 * B0 + W01 * B1;
            np[(i + 1) & 3][0] = (float)(lines[i][0] + L0 * A00);
            np[(i + 1) & 3][1] = (float)(lines[i][1] + L0 * A10);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) { p[i][0] = np[i][0]; p[i][1] = np[i][1]; }
}

// 8x9 Gaussian elimination with partial pivoting (upstream homography_compute2), spread over the warp: lane r (mod 8)
// holds ROW r in nine registers, pivot search / row swap / pivot-row broadcast are shuffles.  Same operations in the
// same order as the sequential algorithm (first largest pivot wins), so the result is bit-identical to it.
__device__ bool homography_dev(const float (&p)[4][2], double (&H)[9]) {
    const int row = threadIdx.x & 7;
    double A[9];
    {
        const int i = row >> 1;
        float c2f = p[0][0], c3f = p[0][1];
#pragma unroll
        for (int k = 1; k < 4; k++)
            if (i == k) { c2f = p[k][0]; c3f = p[k][1]; }
        const double c0 = (i == 0 || i == 3) ? -1 : 1, c1 = (i == 0 || i == 1) ? -1 : 1;
        const double c2 = c2f, c3 = c3f;
        if ((row & 1) == 0) {
            A[0] = c0; A[1] = c1; A[2] = 1; A[3] = 0; A[4] = 0; A[5] = 0;
            A[6] = -c0 * c2; A[7] = -c1 * c2; A[8] = c2;
        } else {
            A[0] = 0; A[1] = 0; A[2] = 0; A[3] = c0; A[4] = c1; A[5] = 1;
            A[6] = -c0 * c3; A[7] = -c1 * c3; A[8] = c3;
        }
    }
    const double epsilon = 1e-10;
#pragma unroll
    for (int col = 0; col < 8; col++) {
        // pivot: largest |A[r][col]| over rows r >= col, the first one on ties
        const double val = row >= col ? fabs(A[col]) : -1.0;
        double mx = val;
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) {
            const double o = __shfl_xor_sync(FULL_MASK, mx, off);
            mx = o > mx ? o : mx;
        }
        if (!(mx >= epsilon)) return false;
        const uint32_t cand = __ballot_sync(FULL_MASK, val == mx) & 0xffu;
        const int max_idx = __ffs(cand) - 1;
        if (max_idx != col) {   // (uniform) swap rows col and max_idx
            const int partner = row == col ? max_idx : (row == max_idx ? col : row);
#pragma unroll
            for (int i = col; i < 9; i++) A[i] = __shfl_sync(FULL_MASK, A[i], (threadIdx.x & 24) | partner);
        }
        double pv[9];
#pragma unroll
        for (int j = col; j < 9; j++) pv[j] = __shfl_sync(FULL_MASK, A[j], (threadIdx.x & 24) | col);
        if (row > col) {
            const double f = A[col] / pv[col];
            A[col] = 0;
#pragma unroll
            for (int j = col + 1; j < 9; j++) A[j] -= f * pv[j];
        }
    }
    double x[8];
#pragma unroll
    for (int col = 7; col >= 0; col--) {
        double sum = 0;
#pragma unroll
        for (int i = col + 1; i < 8; i++) sum += A[i] * x[i];
        const double v = (A[8] - sum) / A[col];   // meaningful on the lanes that hold row `col`
        x[col] = __shfl_sync(FULL_MASK, v, (threadIdx.x & 24) | col);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) H[i] = x[i];
    H[8] = 1;
    return true;
}

__device__ __forceinline__ void hproject(const double (&H)[9], double x, double y, double& ox, double& oy) {
    double xx = H[0] * x + H[1] * y + H[2];
    double yy = H[3] * x + H[4] * y + H[5];
    double zz = H[6] * x + H[7] * y + H[8];
    ox = xx / zz;
    oy = yy / zz;
}

struct GrayModelDev {
    double A00, A01, A02, A11, A12, A22, B0, B1, B2, C0, C1, C2;
    __device__ void init() { A00 = A01 = A02 = A11 = A12 = A22 = B0 = B1 = B2 = C0 = C1 = C2 = 0; }
    __device__ void add(double x, double y, double g) {
        A00 += x * x; A01 += x * y; A02 += x; A11 += y * y; A12 += y; A22 += 1;
        B0 += x * g; B1 += y * g; B2 += g;
    }
    __device__ void solve() {
        double L0 = sqrt(A00), L3 = A01 / L0, L6 = A02 / L0;
        double L4 = sqrt(A11 - L3 * L3), L7 = (A12 - L3 * L6) / L4;
        double L8 = sqrt(A22 - L6 * L6 - L7 * L7);
        double M0 = 1 / L0, M3 = -L3 * M0 / L4, M4 = 1 / L4;
        double M6 = (-L6 * M0 - L7 * M3) / L8, M7 = -L7 * M4 / L8, M8 = 1 / L8;
        double t0 = M0 * B0, t1 = M3 * B0 + M4 * B1, t2 = M6 * B0 + M7 * B1 + M8 * B2;
        C0 = M0 * t0 + M3 * t1 + M6 * t2;
        C1 = M4 * t1 + M7 * t2;
        C2 = M8 * t2;
    }
    __device__ double interp(double x, double y) const { return C0 * x + C1 * y + C2; }
};

__device__ __forceinline__ unsigned long long rotate90_dev(unsigned long long w, int numBits) {
    int p = numBits;
    unsigned long long l = 0;
    if (numBits % 4 == 1) { p = numBits - 1; l = 1; }
    w = ((w >> l) << (p / 4 + l)) | (w >> (3 * p / 4 + l) << l) | (w & l);
    w &= ((1ull << numBits) - 1);
    return w;
}

#define DEC_GRID_MAX 144  // total_width^2 <= 12*12

// returns decision margin (< 0: rejected); fills id / hamming / rotation
__device__ float quad_decode_warp(const DevParams& P, const DevFamily& fam, const unsigned long long* codes,
                                  const uint8_t* im, size_t pitch, int width, int height, const double (&H)[9],
                                  double* values, double* sharp, int& out_id, int& out_hamming, int& out_rot) {
    const int lane = threadIdx.x & 31;
    const int wb = fam.width_at_border;
    GrayModelDev white, black;
    white.init();
    black.init();
    double gacc = 0;      // lanes 0..8 / 9..17: running sum of one moment of the white / black model
    const int nsamp = 8 * wb;
    for (int base = 0; base < nsamp; base += 32) {
        const int t = base + lane;
        double tagx = 0, tagy = 0;
        int v = 0, valid = 0, is_white = 0;
        if (t < nsamp) {
            const int pi = t / wb, i = t - pi * wb;
            float p0, p1, p2, p3;
            switch (pi) {
                case 0: p0 = -0.5f; p1 = 0.5f; p2 = 0; p3 = 1; is_white = 1; break;
                case 1: p0 = 0.5f; p1 = 0.5f; p2 = 0; p3 = 1; is_white = 0; break;
                case 2: p0 = wb + 0.5f; p1 = .5f; p2 = 0; p3 = 1; is_white = 1; break;
                case 3: p0 = wb - 0.5f; p1 = .5f; p2 = 0; p3 = 1; is_white = 0; break;
                case 4: p0 = 0.5f; p1 = -0.5f; p2 = 1; p3 = 0; is_white = 1; break;
                case 5: p0 = 0.5f; p1 = 0.5f; p2 = 1; p3 = 0; is_white = 0; break;
                case 6: p0 = 0.5f; p1 = wb + 0.5f; p2 = 1; p3 = 0; is_white = 1; break;
                default: p0 = 0.5f; p1 = wb - 0.5f; p2 = 1; p3 = 0; is_white = 0; break;
            }
            double tagx01 = (p0 + i * p2) / wb;
            double tagy01 = (p1 + i * p3) / wb;
            tagx = 2 * (tagx01 - 0.5);
            tagy = 2 * (tagy01 - 0.5);
            double px, py;
            hproject(H, tagx, tagy, px, py);
            int ix = (int)px, iy = (int)py;
            if (!(ix < 0 || iy < 0 || ix >= width || iy >= height)) {
                v = im[(size_t)iy * pitch + ix];
                valid = 1;
            }
        }
        // sequential-order sums of the two models' nine moments: lane k parks the nine terms of its sample, lanes 0..8 run
        // the white model's sums and lanes 9..17 the black model's over the samples in order (a sample the model does not
        // own adds 0.0, which leaves a sum as it is)
        const int cnt = min(32, nsamp - base);
        const uint32_t wmask = __ballot_sync(FULL_MASK, valid && is_white), bmask = __ballot_sync(FULL_MASK, valid && !is_white);
        {
            const double g = (double)v;
            values[DEC_STAGE_AT(0, lane)] = tagx * tagx; values[DEC_STAGE_AT(1, lane)] = tagx * tagy; values[DEC_STAGE_AT(2, lane)] = tagx;
            values[DEC_STAGE_AT(3, lane)] = tagy * tagy; values[DEC_STAGE_AT(4, lane)] = tagy; values[DEC_STAGE_AT(5, lane)] = 1.0;
            values[DEC_STAGE_AT(6, lane)] = tagx * g; values[DEC_STAGE_AT(7, lane)] = tagy * g; values[DEC_STAGE_AT(8, lane)] = g;
        }
        __syncwarp();
        {
            const int mrow = lane < 9 ? lane : (lane < 18 ? lane - 9 : 0);
            const uint32_t mine = lane < 9 ? wmask : (lane < 18 ? bmask : 0u);
#pragma unroll 4
            for (int k = 0; k < cnt; k++) {
                const double t = values[DEC_STAGE_AT(mrow, k)];
                gacc += ((mine >> k) & 1u) ? t : 0.0;
            }
        }
        __syncwarp();
    }
    white.A00 = __shfl_sync(FULL_MASK, gacc, 0); white.A01 = __shfl_sync(FULL_MASK, gacc, 1); white.A02 = __shfl_sync(FULL_MASK, gacc, 2);
    white.A11 = __shfl_sync(FULL_MASK, gacc, 3); white.A12 = __shfl_sync(FULL_MASK, gacc, 4); white.A22 = __shfl_sync(FULL_MASK, gacc, 5);
    white.B0 = __shfl_sync(FULL_MASK, gacc, 6); white.B1 = __shfl_sync(FULL_MASK, gacc, 7); white.B2 = __shfl_sync(FULL_MASK, gacc, 8);
    black.A00 = __shfl_sync(FULL_MASK, gacc, 9); black.A01 = __shfl_sync(FULL_MASK, gacc, 10); black.A02 = __shfl_sync(FULL_MASK, gacc, 11);
    black.A11 = __shfl_sync(FULL_MASK, gacc, 12); black.A12 = __shfl_sync(FULL_MASK, gacc, 13); black.A22 = __shfl_sync(FULL_MASK, gacc, 14);
    black.B0 = __shfl_sync(FULL_MASK, gacc, 15); black.B1 = __shfl_sync(FULL_MASK, gacc, 16); black.B2 = __shfl_sync(FULL_MASK, gacc, 17);
    white.solve();
    black.solve();
    if ((white.interp(0, 0) - black.interp(0, 0) < 0) != (fam.reversed_border != 0)) return -1.0f;

    const int tw = fam.total_width;
    const int min_coord = (wb - tw) / 2;
    for (int i = lane; i < tw * tw; i += 32) values[i] = 0;
    __syncwarp();
    for (int i = lane; i < fam.nbits; i += 32) {
        const int bity = fam.bit_y[i], bitx = fam.bit_x[i];
        double tagx = 2 * ((bitx + 0.5) / wb - 0.5), tagy = 2 * ((bity + 0.5) / wb - 0.5);
        double px, py;
        hproject(H, tagx, tagy, px, py);
        int x1 = (int)floor(px - 0.5), x2 = (int)ceil(px - 0.5);
        double x = px - 0.5 - x1;
        int y1 = (int)floor(py - 0.5), y2 = (int)ceil(py - 0.5);
        double y = py - 0.5 - y1;
        if (x1 < 0 || x2 >= width || y1 < 0 || y2 >= height) continue;
        double v = im[(size_t)y1 * pitch + x1] * (1 - x) * (1 - y) + im[(size_t)y1 * pitch + x2] * x * (1 - y) +
                   im[(size_t)y2 * pitch + x1] * (1 - x) * y + im[(size_t)y2 * pitch + x2] * x * y;
        double thresh = (black.interp(tagx, tagy) + white.interp(tagx, tagy)) / 2.0;
        values[tw * (bity - min_coord) + bitx - min_coord] = v - thresh;
    }
    __syncwarp();
    for (int c = lane; c < tw * tw; c += 32) {
        const int y = c / tw, x = c - y * tw;
        double acc = 0;  // kernel order: (-1,0) up, (0,-1) left, centre*4, right, down -- row major like upstream
        if (y - 1 >= 0) acc += values[(y - 1) * tw + x] * -1.0;
        if (x - 1 >= 0) acc += values[y * tw + x - 1] * -1.0;
        acc += values[y * tw + x] * 4.0;
        if (x + 1 <= tw - 1) acc += values[y * tw + x + 1] * -1.0;
        if (y + 1 <= tw - 1) acc += values[(y + 1) * tw + x] * -1.0;
        sharp[c] = acc;
    }
    __syncwarp();
    for (int c = lane; c < tw * tw; c += 32) values[c] = values[c] + P.decode_sharpening * sharp[c];
    __syncwarp();

    float black_score = 0, white_score = 0, black_count = 1, white_count = 1;
    unsigned long long rcode = 0;
    for (int i = 0; i < fam.nbits; i++) {
        const int bity = fam.bit_y[i], bitx = fam.bit_x[i];
        rcode <<= 1;
        double v = values[(bity - min_coord) * tw + bitx - min_coord];
        if (v > 0) {
            white_score = (float)((double)white_score + v);
            white_count += 1;
            rcode |= 1;
        } else {
            black_score = (float)((double)black_score - v);
            black_count += 1;
        }
    }
    // quick decode: four rotations, nearest code word (lowest id on ties), accept <= maxhamming
    out_id = 65535; out_hamming = 255; out_rot = 0;
    const unsigned long long* fc = codes + fam.code_offset;
    for (int r = 0; r < 4; r++) {
        uint32_t best = 0xffffffffu;
        for (int i = lane; i < fam.ncodes; i += 32) {
            uint32_t hd = __popcll(fc[i] ^ rcode);
            best = min(best, (hd << 16) | (uint32_t)i);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) best = min(best, __shfl_xor_sync(FULL_MASK, best, off));
        if ((int)(best >> 16) <= P.maxhamming) {
            out_id = best & 0xffff; out_hamming = best >> 16; out_rot = r;
            break;
        }
        rcode = rotate90_dev(rcode, fam.nbits);
    }
    return fminf(white_score / white_count, black_score / black_count);
}

#ifndef DEC_MINB
#define DEC_MINB 4   // 128 registers: all 16 persistent one-warp CTAs per SM (decode_ctas = 4) are resident; at the 155
                     // the compiler takes unbounded only 12 were and a quarter of the grid ran as a second wave
                     // (0.42 -> 0.32 ms per 128 frames; 96 registers spill too much)
#endif
__global__ void __launch_bounds__(128, DEC_MINB)
k_decode_quads(DecodeArgs a, DevParams P) {
    __shared__ double s_buf[4][2 * DEC_GRID_MAX];   // per warp: bit grid + sharpened grid; before that the 9 x 32 staging buffer
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nq = min(*a.nquads, a.cap_quads);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    for (int qi = blockIdx.x * (blockDim.x >> 5) + w; qi < nq; qi += nwarps) {
        const QuadRec q = a.quads[qi];
        const uint8_t* im = a.im + (size_t)q.frame * a.frame_stride;
        float p[4][2];
#pragma unroll
        for (int i = 0; i < 4; i++) { p[i][0] = q.p[i][0]; p[i][1] = q.p[i][1]; }
        if (P.refine_edges) refine_edges_warp(P, im, a.pitch, a.W, a.H, p, q.reversed_border, s_buf[w]);
        if (a.dbg_refined && lane == 0) {
#pragma unroll
            for (int i = 0; i < 4; i++) { a.dbg_refined[qi * 8 + 2 * i] = p[i][0]; a.dbg_refined[qi * 8 + 2 * i + 1] = p[i][1]; }
        }
        double H[9];
        if (!homography_dev(p, H)) continue;
        for (int fi = 0; fi < P.nfamilies; fi++) {
            const DevFamily& fam = a.fams[fi];
            if ((fam.reversed_border != 0) != (q.reversed_border != 0)) continue;
            int id, hamming, rot;
            float margin = quad_decode_warp(P, fam, a.codes, im, a.pitch, a.W, a.H, H, s_buf[w], s_buf[w] + DEC_GRID_MAX, id,
                                            hamming, rot);
            if (margin >= 0 && hamming < 255 && lane == 0) {
                DetRec d;
                d.family = fi; d.id = id; d.hamming = hamming; d.margin = margin;
                const double c = P.rot_c[rot], s = P.rot_s[rot];
                const double R[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
                double Hd[9];
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int cc = 0; cc < 3; cc++) {
                        double acc = 0;
#pragma unroll
                        for (int k = 0; k < 3; k++) acc += H[r * 3 + k] * R[k * 3 + cc];
                        Hd[r * 3 + cc] = acc;
                    }
#pragma unroll
                for (int k = 0; k < 9; k++) d.H[k] = Hd[k];
                hproject(Hd, 0, 0, d.c[0], d.c[1]);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    int tcx = (i == 1 || i == 2) ? 1 : -1;
                    int tcy = (i < 2) ? 1 : -1;
                    hproject(Hd, tcx, tcy, d.p[i][0], d.p[i][1]);
                }
                int slot = atomicAdd(&a.ndets[q.frame], 1);
                if (slot < a.cap_dets) a.dets[(size_t)q.frame * a.cap_dets + slot] = d;
            }
            __syncwarp();
        }
    }
}

// ---- reconcile + sort: one warp per frame ---------------------------------------------------
__device__ __forceinline__ bool seg_intersect_dev(const double* a, const double* b, const double* c, const double* d) {
    double o1 = (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]);
    double o2 = (b[0] - a[0]) * (d[1] - a[1]) - (b[1] - a[1]) * (d[0] - a[0]);
    double o3 = (d[0] - c[0]) * (a[1] - c[1]) - (d[1] - c[1]) * (a[0] - c[0]);
    double o4 = (d[0] - c[0]) * (b[1] - c[1]) - (d[1] - c[1]) * (b[0] - c[0]);
    return ((o1 > 0) != (o2 > 0)) && ((o3 > 0) != (o4 > 0));
}
__device__ __forceinline__ bool poly_contains_dev(const double (*p)[2], const double* q) {
    bool in = false;
    for (int i = 0, j = 3; i < 4; j = i++) {
        if (((p[i][1] > q[1]) != (p[j][1] > q[1])) &&
            (q[0] < (p[j][0] - p[i][0]) * (q[1] - p[i][1]) / (p[j][1] - p[i][1]) + p[i][0]))
            in = !in;
    }
    return in;
}
__device__ bool polys_overlap_dev(const double (*a)[2], const double (*b)[2]) {
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            if (seg_intersect_dev(a[i], a[(i + 1) & 3], b[j], b[(j + 1) & 3])) return true;
    return poly_contains_dev(a, b[0]) || poly_contains_dev(b, a[0]);
}
__device__ int prefer_dev(const DetRec& a, const DetRec& b) {
    if (a.hamming != b.hamming) return a.hamming < b.hamming ? -1 : 1;
    if (a.margin != b.margin) return a.margin > b.margin ? -1 : 1;
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 2; k++)
            if (a.p[i][k] != b.p[i][k]) return a.p[i][k] < b.p[i][k] ? -1 : 1;
    return -1;
}

#define REC_CAP 1024  // internal per-frame detection capacity before reconcile (29 KB of shared-memory sort keys)

// One CTA of REC_THREADS threads per frame (a 4K frame carries a few hundred detections).  The sort keys (id, family, centre) are staged in shared memory once, so the
// two rank sorts (n^2 comparisons, dealt to all threads) run out of shared memory instead of re-reading the 168-byte
// records n times; the pairwise rule in between is sequential inside a run of equal ids (one thread per run).
#define REC_THREADS 256
__global__ void __launch_bounds__(REC_THREADS)
k_reconcile(const DetRec* __restrict__ dets, const int* __restrict__ ndets, int cap_dets, int nframes,
            DetRec* __restrict__ out, int* __restrict__ out_counts, int cap_out) {
    __shared__ int s_perm[REC_CAP];
    __shared__ unsigned char s_dead[REC_CAP];
    __shared__ int s_id[REC_CAP], s_fam[REC_CAP];
    __shared__ double s_cx[REC_CAP], s_cy[REC_CAP];
    __shared__ int s_alive;
    const int tid = threadIdx.x;
    const int frame = blockIdx.x;
    if (frame >= nframes) return;
    const int n = min(min(ndets[frame], cap_dets), REC_CAP);
    const DetRec* fd = dets + (size_t)frame * cap_dets;
    int* perm = s_perm;
    unsigned char* dead = s_dead;
    if (tid == 0) s_alive = 0;
    for (int i = tid; i < n; i += REC_THREADS) {
        s_id[i] = fd[i].id; s_fam[i] = fd[i].family; s_cx[i] = fd[i].c[0]; s_cy[i] = fd[i].c[1];
    }
    __syncthreads();
    // order 0: (id, family, cx, cy)   order 1: (id, cx, cy)
    auto less = [&](int a, int b, int order) {
        if (s_id[a] != s_id[b]) return s_id[a] < s_id[b];
        if (order == 0 && s_fam[a] != s_fam[b]) return s_fam[a] < s_fam[b];
        if (s_cx[a] != s_cx[b]) return s_cx[a] < s_cx[b];
        return s_cy[a] < s_cy[b];
    };
    for (int i = tid; i < n; i += REC_THREADS) {
        int rank = 0;
        for (int j = 0; j < n; j++)
            if (less(j, i, 0) || (!less(i, j, 0) && j < i)) rank++;
        perm[rank] = i;
        dead[i] = 0;
    }
    __syncthreads();
    // the pairwise rule only ever compares detections of one id: the runs of equal id (consecutive in the sorted order) are
    // independent, so every run is walked -- sequentially, as the rule is defined -- by its own thread
    for (int a0 = tid; a0 < n; a0 += REC_THREADS) {
        if (a0 > 0 && s_id[perm[a0 - 1]] == s_id[perm[a0]]) continue;   // not the first of its run
        for (int a = a0; a < n && s_id[perm[a]] == s_id[perm[a0]]; a++) {
            const int i = perm[a];
            if (dead[i]) continue;
            for (int b = a + 1; b < n; b++) {
                const int j = perm[b];
                if (s_id[j] != s_id[i]) break;
                if (dead[j] || s_fam[j] != s_fam[i]) continue;
                if (!polys_overlap_dev(fd[i].p, fd[j].p)) continue;
                if (prefer_dev(fd[i], fd[j]) < 0) dead[j] = 1;
                else { dead[i] = 1; break; }
            }
        }
    }
    __syncthreads();
    int alive = 0;
    for (int i = tid; i < n; i += REC_THREADS) {
        if (dead[i]) continue;
        int rank = 0;
        for (int j = 0; j < n; j++) {
            if (dead[j]) continue;
            if (less(j, i, 1) || (!less(i, j, 1) && j < i)) rank++;
        }
        if (rank < cap_out) out[(size_t)frame * cap_out + rank] = fd[i];
        alive++;
    }
    if (alive) atomicAdd(&s_alive, alive);
    __syncthreads();
    if (tid == 0) out_counts[frame] = s_alive;
}
