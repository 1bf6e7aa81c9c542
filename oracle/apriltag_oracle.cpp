// apriltag_oracle.cpp -- CPU restatement of the AprilTag-3 detector that AprilSLAM calls.
//
// *** TEST INFRASTRUCTURE, NOT PRODUCT CODE. ***  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load this library.  The product
// (aprilslam_b200/, libaprilgpu.so) never links, imports or calls it.
//
// What it restates: the native call behind
//     /root/reference/src/detection/tag_detector.py:18   apriltag(tag_type)
//     /root/reference/src/detection/tag_detector.py:26   self.detector.detect(gray)
// i.e. AprilRobotics/apriltag (apriltag_detector_detect + apriltag_pywrap.c).  That library is
// an UN-VENDORED, UNPINNED dependency of the reference (README.md:42-57 "git clone ...apriltag.git",
// requirements.txt:2, .gitignore:5) and its source is not present in /root/reference nor in this
// container, so this file restates its published algorithm from the stage specification in
// SURVEY.md section 8(a) rows U1-U10 and Appendix A.  PARITY UNPINNED at the upstream boundary:
// the reference has no tests or golden vectors for this path.  What pins this oracle instead
// (tests/test_oracle_*.py): code-book known-answer words, analytic ground-truth corners and poses
// on frames rendered with the reference's renderer geometry, cv2.aruco id cross-checks, and the
// reference's committed run log / CSV (soft vectors).
//
// Deterministic choices where upstream is order-dependent (documented in DESIGN.md):
//   * cluster points are ordered by (slope, y, x) -- upstream: stable sort on slope of an
//     insertion order that depends on its hash map / thread chunking;
//   * the border polarity sum (dot) is evaluated from exact integer sums;
//   * clusters are processed in ascending (rep_hi, rep_lo) key order, detections are reconciled
//     in (id, family, cx, cy) order and reported sorted by (id, cx, cy).
//
// Build: see oracle/Makefile  (g++ -O3 -ffp-contract=off: no FMA contraction, so float
// decisions are reproducible).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Family {
    std::string name;
    int nbits, h, ncodes, width_at_border, total_width, reversed_border;
    std::vector<uint64_t> codes;
    std::vector<int> bit_x, bit_y;
};

struct Params {
    float quad_decimate = 2.0f;
    float quad_sigma = 0.0f;
    int refine_edges = 1;
    double decode_sharpening = 0.25;
    int maxhamming = 1;
    // qtp defaults
    int min_cluster_pixels = 5;
    int max_nmaxima = 10;
    float cos_critical_rad = (float)std::cos(10.0 * M_PI / 180.0);
    float max_line_fit_mse = 10.0f;
    int min_white_black_diff = 5;
};

struct Quad {
    float p[4][2];
    int reversed_border;
    double H[9];
};

struct Pt {
    uint16_t x, y;
    int16_t gx, gy;
    float slope;
};

}  // namespace

extern "C" {
// POD shared with tests (same field order as agpu_detection in include/aprilgpu.h)
struct ao_detection {
    int32_t family, id, hamming;
    float margin;
    double c[2];
    double p[4][2];  // lb, rb, rt, lt
    double H[9];
};

struct ao_debug {  // every pointer optional (NULL = not wanted); sizes are caller-guaranteed
    uint8_t* quad_im;      // [hd*wd] decimated (+blurred) image
    uint8_t* thresh;       // [hd*wd]
    uint32_t* labels;      // [hd*wd] min-index representative of each pixel's component
    uint32_t* sizes;       // [hd*wd] component size at the representative, 0 elsewhere
    int wd, hd;            // out
    int npoints;           // out: edge points emitted (before per-cluster de-dup)
    int nclusters;         // out: number of distinct (rep_hi,rep_lo) keys
    uint64_t* cluster_keys;  // [cap_clusters]
    int32_t* cluster_sizes;  // [cap_clusters]
    int cap_clusters;
    int nquads;            // out
    float* quads;          // [cap_quads*9]: 8 corner floats (decimation undone) + reversed flag
    float* quads_refined;  // [cap_quads*8]
    uint64_t* quad_keys;   // [cap_quads]
    int cap_quads;
    int noversize;         // out: clusters dropped by upstream's size limit (more than 3(2w+2h) points)
};
}

namespace {

struct Detector {
    Params prm;
    std::vector<Family> fams;
    std::string err;
};

// ------------------------------------------------------------------------------------------
// U1 decimate: point sampling
// ------------------------------------------------------------------------------------------
void decimate(const uint8_t* im, int w, int h, int stride, int f, std::vector<uint8_t>& out, int& wd, int& hd) {
    wd = 1 + (w - 1) / f;
    hd = 1 + (h - 1) / f;
    out.resize((size_t)wd * hd);
    for (int y = 0; y < hd; y++)
        for (int x = 0; x < wd; x++) out[(size_t)y * wd + x] = im[(size_t)(y * f) * stride + x * f];
}

// ------------------------------------------------------------------------------------------
// U2 blur / sharpen
// ------------------------------------------------------------------------------------------
void convolve1d(const uint8_t* x, uint8_t* y, int sz, const uint8_t* k, int ksz) {
    for (int i = 0; i < ksz / 2 && i < sz; i++) y[i] = x[i];
    for (int i = 0; i < sz - ksz; i++) {
        uint32_t acc = 0;
        for (int j = 0; j < ksz; j++) acc += k[j] * x[i + j];
        y[ksz / 2 + i] = acc >> 8;
    }
    for (int i = sz - ksz + ksz / 2; i < sz; i++)
        if (i >= 0) y[i] = x[i];
}

int gaussian_kernel(float sigma, uint8_t* k /*>=64*/) {
    int ksz = (int)(4 * sigma);
    if ((ksz & 1) == 0) ksz++;
    if (ksz <= 1) return 0;
    double dk[64];
    double acc = 0;
    for (int i = 0; i < ksz; i++) {
        int x = -ksz / 2 + i;
        dk[i] = std::exp(-.5 * (x / (double)sigma) * (x / (double)sigma));
        acc += dk[i];
    }
    for (int i = 0; i < ksz; i++) k[i] = (uint8_t)(dk[i] / acc * 255.0);
    return ksz;
}

void gaussian_blur(std::vector<uint8_t>& im, int w, int h, const uint8_t* k, int ksz) {
    std::vector<uint8_t> a(std::max(w, h)), b(std::max(w, h));
    for (int y = 0; y < h; y++) {
        memcpy(a.data(), &im[(size_t)y * w], w);
        convolve1d(a.data(), b.data(), w, k, ksz);
        memcpy(&im[(size_t)y * w], b.data(), w);
    }
    for (int x = 0; x < w; x++) {
        for (int y = 0; y < h; y++) a[y] = im[(size_t)y * w + x];
        convolve1d(a.data(), b.data(), h, k, ksz);
        for (int y = 0; y < h; y++) im[(size_t)y * w + x] = b[y];
    }
}

void blur_stage(std::vector<uint8_t>& im, int w, int h, float quad_sigma) {
    if (quad_sigma == 0) return;
    uint8_t k[64];
    int ksz = gaussian_kernel(std::fabs(quad_sigma), k);
    if (ksz <= 1) return;
    if (quad_sigma > 0) {
        gaussian_blur(im, w, h, k, ksz);
    } else {
        std::vector<uint8_t> orig = im;
        gaussian_blur(im, w, h, k, ksz);
        for (size_t i = 0; i < im.size(); i++) {
            int v = 2 * orig[i] - im[i];
            im[i] = (uint8_t)std::min(255, std::max(0, v));
        }
    }
}

// ------------------------------------------------------------------------------------------
// U3 threshold
// ------------------------------------------------------------------------------------------
void threshold(const std::vector<uint8_t>& im, int w, int h, int min_wb_diff, std::vector<uint8_t>& out) {
    const int ts = 4;
    int tw = w / ts, th = h / ts;
    out.assign((size_t)w * h, 127);
    if (tw == 0 || th == 0) return;
    std::vector<uint8_t> mx((size_t)tw * th), mn((size_t)tw * th), mx2((size_t)tw * th), mn2((size_t)tw * th);
    for (int ty = 0; ty < th; ty++)
        for (int tx = 0; tx < tw; tx++) {
            uint8_t a = 255, b = 0;
            for (int dy = 0; dy < ts; dy++)
                for (int dx = 0; dx < ts; dx++) {
                    uint8_t v = im[(size_t)(ty * ts + dy) * w + tx * ts + dx];
                    a = std::min(a, v);
                    b = std::max(b, v);
                }
            mn[ty * tw + tx] = a;
            mx[ty * tw + tx] = b;
        }
    for (int ty = 0; ty < th; ty++)
        for (int tx = 0; tx < tw; tx++) {
            uint8_t a = 255, b = 0;
            for (int dy = -1; dy <= 1; dy++) {
                if (ty + dy < 0 || ty + dy >= th) continue;
                for (int dx = -1; dx <= 1; dx++) {
                    if (tx + dx < 0 || tx + dx >= tw) continue;
                    a = std::min(a, mn[(ty + dy) * tw + tx + dx]);
                    b = std::max(b, mx[(ty + dy) * tw + tx + dx]);
                }
            }
            mn2[ty * tw + tx] = a;
            mx2[ty * tw + tx] = b;
        }
    // full tiles: low-contrast tiles (max - min < min_white_black_diff) become 127 = "unknown"
    for (int ty = 0; ty < th; ty++)
        for (int tx = 0; tx < tw; tx++) {
            int a = mn2[ty * tw + tx], b = mx2[ty * tw + tx];
            if (b - a < min_wb_diff) continue;   // stays 127
            uint8_t thr = (uint8_t)(a + (b - a) / 2);
            for (int dy = 0; dy < ts; dy++)
                for (int dx = 0; dx < ts; dx++) {
                    size_t i = (size_t)(ty * ts + dy) * w + tx * ts + dx;
                    out[i] = im[i] > thr ? 255 : 0;
                }
        }
    // leftover right / bottom pixels (w or h not a multiple of 4): upstream's fix-up loop thresholds them against
    // the last full tile's (dilated) extrema and NEVER writes 127 there -- no low-contrast test on this path
    for (int y = 0; y < h; y++) {
        int x0 = y >= th * ts ? 0 : tw * ts;
        int ty = std::min(y / ts, th - 1);
        for (int x = x0; x < w; x++) {
            int tx = std::min(x / ts, tw - 1);
            int a = mn2[ty * tw + tx], b = mx2[ty * tw + tx];
            int thr = a + (b - a) / 2;
            out[(size_t)y * w + x] = im[(size_t)y * w + x] > thr ? 255 : 0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// U4 connected components
// ------------------------------------------------------------------------------------------
struct UF {
    std::vector<uint32_t> parent, size;
    explicit UF(size_t n) : parent(n), size(n, 1) {
        for (size_t i = 0; i < n; i++) parent[i] = (uint32_t)i;
    }
    uint32_t find(uint32_t a) {
        while (parent[a] != a) {
            parent[a] = parent[parent[a]];
            a = parent[a];
        }
        return a;
    }
    void unite(uint32_t a, uint32_t b) {
        a = find(a);
        b = find(b);
        if (a == b) return;
        if (size[a] < size[b]) std::swap(a, b);
        parent[b] = a;
        size[a] += size[b];
    }
};

// labels[i] = smallest pixel index of i's component; sizes[rep] = component size
void connected_components(const std::vector<uint8_t>& t, int w, int h, std::vector<uint32_t>& labels,
                          std::vector<uint32_t>& sizes) {
    size_t n = (size_t)w * h;
    UF uf(n);
    // upstream's do_unionfind_first_line / do_unionfind_line2, guards included.  The guards skip unions that
    // are implied by others -- EXCEPT the up-right one: it is skipped whenever up == up-right, and at
    // x = w-2 the implied link (w-1, y-1) ~ (w-2, y-1) does not exist (column w-1 never initiates a union),
    // so a white pixel of the last column joins a component only through a diagonal from (w-2, y+1) whose
    // upper neighbour is not white.  Restated literally so that the partition is upstream's.
    for (int y = 0; y < h; y++)
        for (int x = 1; x < w - 1; x++) {
            uint8_t v = t[(size_t)y * w + x];
            if (v == 127) continue;
            uint32_t id = (uint32_t)(y * w + x);
            if (t[id - 1] == v) uf.unite(id, id - 1);
            if (y == 0) continue;
            const uint8_t v_m1_0 = t[id - 1], v_m1_m1 = t[id - w - 1], v_0_m1 = t[id - w], v_1_m1 = t[id - w + 1];
            if (x == 1 || !(v_m1_0 == v_m1_m1 && v_m1_m1 == v_0_m1))
                if (v_0_m1 == v) uf.unite(id, id - w);
            if (v == 255) {
                if (x == 1 || !(v_m1_0 == v_m1_m1 || v_0_m1 == v_m1_m1))
                    if (v_m1_m1 == v) uf.unite(id, id - w - 1);
                if (!(v_0_m1 == v_1_m1))
                    if (v_1_m1 == v) uf.unite(id, id - w + 1);
            }
        }
    labels.assign(n, 0xffffffffu);
    sizes.assign(n, 0);
    std::vector<uint32_t> minidx(n, 0xffffffffu);
    for (size_t i = 0; i < n; i++) {
        uint32_t r = uf.find((uint32_t)i);
        if (minidx[r] == 0xffffffffu) minidx[r] = (uint32_t)i;  // first visit in raster order = min index
        labels[i] = minidx[r];
    }
    for (size_t i = 0; i < n; i++) sizes[labels[i]]++;
}

// ------------------------------------------------------------------------------------------
// U5 gradient clusters
// ------------------------------------------------------------------------------------------
struct KeyedPt {
    uint64_t key;
    Pt p;
};

void gradient_clusters(const std::vector<uint8_t>& t, int w, int h, const std::vector<uint32_t>& labels,
                       const std::vector<uint32_t>& sizes, std::vector<KeyedPt>& pts) {
    static const int off[4][2] = {{1, 0}, {0, 1}, {-1, 1}, {1, 1}};
    pts.clear();
    for (int y = 0; y < h - 1; y++)
        for (int x = 1; x < w - 1; x++) {
            uint8_t v0 = t[(size_t)y * w + x];
            if (v0 == 127) continue;
            uint32_t rep0 = labels[(size_t)y * w + x];
            if (sizes[rep0] < 25) continue;
            for (int k = 0; k < 4; k++) {
                int dx = off[k][0], dy = off[k][1];
                uint8_t v1 = t[(size_t)(y + dy) * w + x + dx];
                if (v0 + v1 != 255) continue;
                uint32_t rep1 = labels[(size_t)(y + dy) * w + x + dx];
                if (sizes[rep1] < 25) continue;
                KeyedPt kp;
                uint32_t hi = std::max(rep0, rep1), lo = std::min(rep0, rep1);
                kp.key = ((uint64_t)hi << 32) | lo;
                kp.p.x = (uint16_t)(2 * x + dx);
                kp.p.y = (uint16_t)(2 * y + dy);
                kp.p.gx = (int16_t)(dx * ((int)v1 - (int)v0));
                kp.p.gy = (int16_t)(dy * ((int)v1 - (int)v0));
                kp.p.slope = 0;
                pts.push_back(kp);
            }
        }
    std::stable_sort(pts.begin(), pts.end(), [](const KeyedPt& a, const KeyedPt& b) { return a.key < b.key; });
}

// ------------------------------------------------------------------------------------------
// U6 quad fit
// ------------------------------------------------------------------------------------------
struct LineFitPt {
    double Mx, My, Mxx, Mxy, Myy, W;
};

void fit_line(const LineFitPt* lfps, int sz, int i0, int i1, double* lineparm, double* err, double* mse) {
    double Mx, My, Mxx, Mxy, Myy, W;
    int N;
    if (i0 < i1) {
        N = i1 - i0 + 1;
        Mx = lfps[i1].Mx; My = lfps[i1].My; Mxx = lfps[i1].Mxx; Mxy = lfps[i1].Mxy; Myy = lfps[i1].Myy; W = lfps[i1].W;
        if (i0 > 0) {
            Mx -= lfps[i0 - 1].Mx; My -= lfps[i0 - 1].My; Mxx -= lfps[i0 - 1].Mxx;
            Mxy -= lfps[i0 - 1].Mxy; Myy -= lfps[i0 - 1].Myy; W -= lfps[i0 - 1].W;
        }
    } else {
        // wrap-around span [i0 .. sz-1] + [0 .. i1]; i0 > 0 here
        Mx = lfps[sz - 1].Mx - lfps[i0 - 1].Mx; My = lfps[sz - 1].My - lfps[i0 - 1].My;
        Mxx = lfps[sz - 1].Mxx - lfps[i0 - 1].Mxx; Mxy = lfps[sz - 1].Mxy - lfps[i0 - 1].Mxy;
        Myy = lfps[sz - 1].Myy - lfps[i0 - 1].Myy; W = lfps[sz - 1].W - lfps[i0 - 1].W;
        Mx += lfps[i1].Mx; My += lfps[i1].My; Mxx += lfps[i1].Mxx; Mxy += lfps[i1].Mxy; Myy += lfps[i1].Myy; W += lfps[i1].W;
        N = sz - i0 + i1 + 1;
    }
    double Ex = Mx / W, Ey = My / W;
    double Cxx = Mxx / W - Ex * Ex;
    double Cxy = Mxy / W - Ex * Ey;
    double Cyy = Myy / W - Ey * Ey;
    float disc = sqrtf((float)((Cxx - Cyy) * (Cxx - Cyy) + 4 * Cxy * Cxy));
    double eig_small = 0.5 * (Cxx + Cyy - disc);
    if (lineparm) {
        lineparm[0] = Ex;
        lineparm[1] = Ey;
        double eig = 0.5 * (Cxx + Cyy + disc);
        double nx1 = Cxx - eig, ny1 = Cxy, M1 = nx1 * nx1 + ny1 * ny1;
        double nx2 = Cxy, ny2 = Cyy - eig, M2 = nx2 * nx2 + ny2 * ny2;
        double nx, ny, M;
        if (M1 > M2) { nx = nx1; ny = ny1; M = M1; } else { nx = nx2; ny = ny2; M = M2; }
        double length = sqrtf((float)M);
        if (std::fabs(length) < 1e-12) {
            lineparm[2] = lineparm[3] = 0;
        } else {
            lineparm[2] = nx / length;
            lineparm[3] = ny / length;
        }
    }
    if (err) *err = N * eig_small;
    if (mse) *mse = eig_small;
}

bool quad_segment_maxima(const Params& prm, int sz, const LineFitPt* lfps, int indices[4]) {
    int ksz = std::min(20, sz / 12);
    if (ksz < 2) return false;
    std::vector<double> errs(sz), y(sz);
    for (int i = 0; i < sz; i++) fit_line(lfps, sz, (i + sz - ksz) % sz, (i + ksz) % sz, nullptr, &errs[i], nullptr);
    {
        const double sigma = 1, cutoff = 0.05;
        int fsz = (int)(std::sqrt(-std::log(cutoff) * 2 * sigma * sigma) + 1);
        fsz = 2 * fsz + 1;
        float f[16];
        for (int i = 0; i < fsz; i++) {
            int j = i - fsz / 2;
            f[i] = (float)std::exp(-j * j / (2 * sigma * sigma));
        }
        for (int iy = 0; iy < sz; iy++) {
            double acc = 0;
            for (int i = 0; i < fsz; i++) acc += errs[(iy + i - fsz / 2 + sz) % sz] * f[i];
            y[iy] = acc;
        }
        errs = y;
    }
    std::vector<int> maxima;
    std::vector<double> maxima_errs;
    for (int i = 0; i < sz; i++)
        if (errs[i] > errs[(i + 1) % sz] && errs[i] > errs[(i + sz - 1) % sz]) {
            maxima.push_back(i);
            maxima_errs.push_back(errs[i]);
        }
    int nmaxima = (int)maxima.size();
    if (nmaxima < 4) return false;
    if (nmaxima > prm.max_nmaxima) {
        std::vector<double> c = maxima_errs;
        std::sort(c.begin(), c.end(), [](double a, double b) { return a > b; });
        double thresh = c[prm.max_nmaxima];
        int out = 0;
        for (int in = 0; in < nmaxima; in++) {
            if (maxima_errs[in] <= thresh) continue;
            maxima[out++] = maxima[in];
        }
        nmaxima = out;
    }
    int best[4] = {0, 0, 0, 0};
    double best_error = HUGE_VALF;
    double err01, err12, err23, err30, mse01, mse12, mse23, mse30;
    double p01[4], p12[4], p23[4], p30[4];
    double max_dot = prm.cos_critical_rad;
    for (int m0 = 0; m0 < nmaxima - 3; m0++) {
        int i0 = maxima[m0];
        for (int m1 = m0 + 1; m1 < nmaxima - 2; m1++) {
            int i1 = maxima[m1];
            fit_line(lfps, sz, i0, i1, p01, &err01, &mse01);
            if (mse01 > prm.max_line_fit_mse) continue;
            for (int m2 = m1 + 1; m2 < nmaxima - 1; m2++) {
                int i2 = maxima[m2];
                fit_line(lfps, sz, i1, i2, p12, &err12, &mse12);
                if (mse12 > prm.max_line_fit_mse) continue;
                double dot = p01[2] * p12[2] + p01[3] * p12[3];
                if (std::fabs(dot) > max_dot) continue;
                for (int m3 = m2 + 1; m3 < nmaxima; m3++) {
                    int i3 = maxima[m3];
                    fit_line(lfps, sz, i2, i3, p23, &err23, &mse23);
                    if (mse23 > prm.max_line_fit_mse) continue;
                    fit_line(lfps, sz, i3, i0, p30, &err30, &mse30);
                    if (mse30 > prm.max_line_fit_mse) continue;
                    double err = err01 + err12 + err23 + err30;
                    if (err < best_error) {
                        best_error = err;
                        best[0] = i0; best[1] = i1; best[2] = i2; best[3] = i3;
                    }
                }
            }
        }
    }
    if (best_error == HUGE_VALF) return false;
    for (int i = 0; i < 4; i++) indices[i] = best[i];
    return best_error / sz < prm.max_line_fit_mse;
}

inline double sq(double v) { return v * v; }

// pts: the cluster (mutable: sorted / de-duplicated in place).  quad_im: decimated gray image.
bool fit_quad(const Params& prm, const uint8_t* quad_im, int w, int h, std::vector<Pt>& pts, int tag_width,
              bool normal_border, bool reversed_border, Quad& quad) {
    int sz = (int)pts.size();
    if (sz < 24) return false;
    int xmax = pts[0].x, xmin = xmax, ymax = pts[0].y, ymin = ymax;
    for (int i = 1; i < sz; i++) {
        xmax = std::max<int>(xmax, pts[i].x); xmin = std::min<int>(xmin, pts[i].x);
        ymax = std::max<int>(ymax, pts[i].y); ymin = std::min<int>(ymin, pts[i].y);
    }
    if ((xmax - xmin) * (ymax - ymin) < tag_width) return false;
    float cx = (xmin + xmax) * 0.5f + 0.05118f;
    float cy = (ymin + ymax) * 0.5f - 0.028581f;
    // polarity: sum over points of (p - c) . g, from exact integer sums (order independent)
    int64_t Sxgx = 0, Sgx = 0, Sygy = 0, Sgy = 0;
    static const float quadrants[2][2] = {{-1 * (2 << 15), 0}, {2 * (2 << 15), 2 << 15}};
    for (int i = 0; i < sz; i++) {
        Pt& p = pts[i];
        Sxgx += (int64_t)p.x * p.gx; Sgx += p.gx;
        Sygy += (int64_t)p.y * p.gy; Sgy += p.gy;
        float dx = p.x - cx, dy = p.y - cy;
        float q = quadrants[dy > 0][dx > 0];
        if (dy < 0) { dy = -dy; dx = -dx; }
        if (dx < 0) { float tmp = dx; dx = dy; dy = -tmp; }
        p.slope = q + dy / dx;
    }
    double dot = ((double)Sxgx - (double)cx * (double)Sgx) + ((double)Sygy - (double)cy * (double)Sgy);
    quad.reversed_border = dot < 0;
    if (!reversed_border && quad.reversed_border) return false;
    if (!normal_border && !quad.reversed_border) return false;

    std::sort(pts.begin(), pts.end(), [](const Pt& a, const Pt& b) {
        if (a.slope != b.slope) return a.slope < b.slope;
        if (a.y != b.y) return a.y < b.y;
        return a.x < b.x;
    });
    {
        int out = 1;
        for (int i = 1; i < sz; i++)
            if (pts[i].x != pts[out - 1].x || pts[i].y != pts[out - 1].y) pts[out++] = pts[i];
        pts.resize(out);
        sz = out;
    }
    if (sz < 24) return false;

    std::vector<LineFitPt> lfps(sz);
    for (int i = 0; i < sz; i++) {
        const Pt& p = pts[i];
        if (i > 0) lfps[i] = lfps[i - 1]; else lfps[i] = LineFitPt{0, 0, 0, 0, 0, 0};
        double delta = 0.5;
        double x = p.x * .5 + delta, y = p.y * .5 + delta;
        int ix = (int)x, iy = (int)y;
        double W = 1;
        if (ix > 0 && ix + 1 < w && iy > 0 && iy + 1 < h) {
            int grad_x = quad_im[iy * w + ix + 1] - quad_im[iy * w + ix - 1];
            int grad_y = quad_im[(iy + 1) * w + ix] - quad_im[(iy - 1) * w + ix];
            W = std::sqrt((double)(grad_x * grad_x + grad_y * grad_y)) + 1;
        }
        double fx = x, fy = y;
        lfps[i].Mx += W * fx;
        lfps[i].My += W * fy;
        lfps[i].Mxx += W * fx * fx;
        lfps[i].Mxy += W * fx * fy;
        lfps[i].Myy += W * fy * fy;
        lfps[i].W += W;
    }

    int indices[4];
    if (!quad_segment_maxima(prm, sz, lfps.data(), indices)) return false;

    double lines[4][4];
    for (int i = 0; i < 4; i++) {
        double mse;
        fit_line(lfps.data(), sz, indices[i], indices[(i + 1) & 3], lines[i], nullptr, &mse);
        if (mse > prm.max_line_fit_mse) return false;
    }
    for (int i = 0; i < 4; i++) {
        double A00 = lines[i][3], A01 = -lines[(i + 1) & 3][3];
        double A10 = -lines[i][2], A11 = lines[(i + 1) & 3][2];
        double B0 = -lines[i][0] + lines[(i + 1) & 3][0];
        double B1 = -lines[i][1] + lines[(i + 1) & 3][1];
        double det = A00 * A11 - A10 * A01;
        if (std::fabs(det) < 0.001) return false;
        double W00 = A11 / det, W01 = -A01 / det;
        double L0 = W00 * B0 + W01 * B1;
        quad.p[i][0] = (float)(lines[i][0] + L0 * A00);
        quad.p[i][1] = (float)(lines[i][1] + L0 * A10);
    }
    {
        double area = 0, length[3], p;
        for (int i = 0; i < 3; i++) {
            int a = i, b = (i + 1) % 3;
            length[i] = std::sqrt(sq(quad.p[b][0] - quad.p[a][0]) + sq(quad.p[b][1] - quad.p[a][1]));
        }
        p = (length[0] + length[1] + length[2]) / 2;
        area += std::sqrt(p * (p - length[0]) * (p - length[1]) * (p - length[2]));
        static const int idxs[4] = {2, 3, 0, 2};
        for (int i = 0; i < 3; i++) {
            int a = idxs[i], b = idxs[i + 1];
            length[i] = std::sqrt(sq(quad.p[b][0] - quad.p[a][0]) + sq(quad.p[b][1] - quad.p[a][1]));
        }
        p = (length[0] + length[1] + length[2]) / 2;
        area += std::sqrt(p * (p - length[0]) * (p - length[1]) * (p - length[2]));
        if (area < 0.95 * tag_width * tag_width) return false;
    }
    for (int i = 0; i < 4; i++) {
        int i0 = i, i1 = (i + 1) & 3, i2 = (i + 2) & 3;
        double dx1 = quad.p[i1][0] - quad.p[i0][0], dy1 = quad.p[i1][1] - quad.p[i0][1];
        double dx2 = quad.p[i2][0] - quad.p[i1][0], dy2 = quad.p[i2][1] - quad.p[i1][1];
        double cos_dtheta = (dx1 * dx2 + dy1 * dy2) / std::sqrt((dx1 * dx1 + dy1 * dy1) * (dx2 * dx2 + dy2 * dy2));
        if ((cos_dtheta > prm.cos_critical_rad || cos_dtheta < -prm.cos_critical_rad) || dx1 * dy2 < dy1 * dx2) return false;
    }
    return true;
}

// ------------------------------------------------------------------------------------------
// U8 refine_edges (bilinear variant)
// ------------------------------------------------------------------------------------------
void refine_edges(const Params& prm, const uint8_t* im, int width, int height, int stride, Quad& quad) {
    double lines[4][4];
    for (int edge = 0; edge < 4; edge++) {
        int a = edge, b = (edge + 1) & 3;
        double nx = quad.p[b][1] - quad.p[a][1];
        double ny = -quad.p[b][0] + quad.p[a][0];
        double mag = std::sqrt(nx * nx + ny * ny);
        nx /= mag;
        ny /= mag;
        if (quad.reversed_border) { nx = -nx; ny = -ny; }
        int nsamples = std::max(16, (int)(mag / 8));
        double Mx = 0, My = 0, Mxx = 0, Mxy = 0, Myy = 0, N = 0;
        for (int s = 0; s < nsamples; s++) {
            double alpha = (1.0 + s) / (nsamples + 1);
            double x0 = alpha * quad.p[a][0] + (1 - alpha) * quad.p[b][0];
            double y0 = alpha * quad.p[a][1] + (1 - alpha) * quad.p[b][1];
            double Mn = 0, Mcount = 0;
            double range = prm.quad_decimate + 1;
            int nsteps = (int)(2 * range * 4) + 1;  // n = -range + 0.25*k, k = 0..nsteps-1 (exact in binary)
            for (int k = 0; k < nsteps; k++) {
                double n = -range + 0.25 * k;
                double grange = 1;
                double x1 = x0 + (n + grange) * nx - 0.5;
                double y1 = y0 + (n + grange) * ny - 0.5;
                int x1i = (int)std::floor(x1), y1i = (int)std::floor(y1);
                double a1 = x1 - x1i, b1 = y1 - y1i;
                if (x1i < 0 || x1i + 1 >= width || y1i < 0 || y1i + 1 >= height) continue;
                double x2 = x0 + (n - grange) * nx - 0.5;
                double y2 = y0 + (n - grange) * ny - 0.5;
                int x2i = (int)std::floor(x2), y2i = (int)std::floor(y2);
                double a2 = x2 - x2i, b2 = y2 - y2i;
                if (x2i < 0 || x2i + 1 >= width || y2i < 0 || y2i + 1 >= height) continue;
                double g1 = (1 - a1) * (1 - b1) * im[y1i * stride + x1i] + a1 * (1 - b1) * im[y1i * stride + x1i + 1] +
                            (1 - a1) * b1 * im[(y1i + 1) * stride + x1i] + a1 * b1 * im[(y1i + 1) * stride + x1i + 1];
                double g2 = (1 - a2) * (1 - b2) * im[y2i * stride + x2i] + a2 * (1 - b2) * im[y2i * stride + x2i + 1] +
                            (1 - a2) * b2 * im[(y2i + 1) * stride + x2i] + a2 * b2 * im[(y2i + 1) * stride + x2i + 1];
                if (g1 < g2) continue;
                double weight = (g2 - g1) * (g2 - g1);
                Mn += weight * n;
                Mcount += weight;
            }
            if (Mcount == 0) continue;
            double n0 = Mn / Mcount;
            double bestx = x0 + n0 * nx, besty = y0 + n0 * ny;
            Mx += bestx; My += besty; Mxx += bestx * bestx; Mxy += bestx * besty; Myy += besty * besty; N++;
        }
        double Ex = Mx / N, Ey = My / N;
        double Cxx = Mxx / N - Ex * Ex, Cxy = Mxy / N - Ex * Ey, Cyy = Myy / N - Ey * Ey;
        double normal_theta = .5 * atan2f((float)(-2 * Cxy), (float)(Cyy - Cxx));
        nx = cosf((float)normal_theta);
        ny = sinf((float)normal_theta);
        lines[edge][0] = Ex; lines[edge][1] = Ey; lines[edge][2] = nx; lines[edge][3] = ny;
    }
    for (int i = 0; i < 4; i++) {
        double A00 = lines[i][3], A01 = -lines[(i + 1) & 3][3];
        double A10 = -lines[i][2], A11 = lines[(i + 1) & 3][2];
        double B0 = -lines[i][0] + lines[(i + 1) & 3][0];
        double B1 = -lines[i][1] + lines[(i + 1) & 3][1];
        double det = A00 * A11 - A10 * A01;
        if (std::fabs(det) > 0.001) {  // NaN (an edge with no samples) fails this test: corner kept
            double W00 = A11 / det, W01 = -A01 / det;
            double L0 = W00 * B0 + W01 * B1;
            quad.p[(i + 1) & 3][0] = (float)(lines[i][0] + L0 * A00);
            quad.p[(i + 1) & 3][1] = (float)(lines[i][1] + L0 * A10);
        }
    }
}

// ------------------------------------------------------------------------------------------
// U9 homography + decode
// ------------------------------------------------------------------------------------------
bool homography_compute2(const double c[4][4], double H[9]) {
    double A[72];
    for (int i = 0; i < 4; i++) {
        double* r0 = &A[(2 * i) * 9];
        double* r1 = &A[(2 * i + 1) * 9];
        r0[0] = c[i][0]; r0[1] = c[i][1]; r0[2] = 1; r0[3] = 0; r0[4] = 0; r0[5] = 0;
        r0[6] = -c[i][0] * c[i][2]; r0[7] = -c[i][1] * c[i][2]; r0[8] = c[i][2];
        r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = c[i][0]; r1[4] = c[i][1]; r1[5] = 1;
        r1[6] = -c[i][0] * c[i][3]; r1[7] = -c[i][1] * c[i][3]; r1[8] = c[i][3];
    }
    const double epsilon = 1e-10;
    for (int col = 0; col < 8; col++) {
        double max_val = 0;
        int max_idx = -1;
        for (int row = col; row < 8; row++) {
            double val = std::fabs(A[row * 9 + col]);
            if (val > max_val) { max_val = val; max_idx = row; }
        }
        if (max_idx < 0 || max_val < epsilon) return false;
        if (max_idx != col)
            for (int i = col; i < 9; i++) std::swap(A[col * 9 + i], A[max_idx * 9 + i]);
        for (int i = col + 1; i < 8; i++) {
            double f = A[i * 9 + col] / A[col * 9 + col];
            A[i * 9 + col] = 0;
            for (int j = col + 1; j < 9; j++) A[i * 9 + j] -= f * A[col * 9 + j];
        }
    }
    for (int col = 7; col >= 0; col--) {
        double sum = 0;
        for (int i = col + 1; i < 8; i++) sum += A[col * 9 + i] * A[i * 9 + 8];
        A[col * 9 + 8] = (A[col * 9 + 8] - sum) / A[col * 9 + col];
    }
    for (int i = 0; i < 8; i++) H[i] = A[i * 9 + 8];
    H[8] = 1;
    return true;
}

bool quad_update_homographies(Quad& q) {
    double corr[4][4];
    for (int i = 0; i < 4; i++) {
        corr[i][0] = (i == 0 || i == 3) ? -1 : 1;
        corr[i][1] = (i == 0 || i == 1) ? -1 : 1;
        corr[i][2] = q.p[i][0];
        corr[i][3] = q.p[i][1];
    }
    return homography_compute2(corr, q.H);
}

inline void homography_project(const double* H, double x, double y, double* ox, double* oy) {
    double xx = H[0] * x + H[1] * y + H[2];
    double yy = H[3] * x + H[4] * y + H[5];
    double zz = H[6] * x + H[7] * y + H[8];
    *ox = xx / zz;
    *oy = yy / zz;
}

struct GrayModel {
    double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    double B[3] = {0, 0, 0};
    double C[3] = {0, 0, 0};
    void add(double x, double y, double g) {
        A[0][0] += x * x; A[0][1] += x * y; A[0][2] += x; A[1][1] += y * y; A[1][2] += y; A[2][2] += 1;
        B[0] += x * g; B[1] += y * g; B[2] += g;
    }
    void solve() {  // symmetric 3x3 via Cholesky (upper entries of A only)
        const double a0 = A[0][0], a1 = A[0][1], a2 = A[0][2], a4 = A[1][1], a5 = A[1][2], a8 = A[2][2];
        double L0 = std::sqrt(a0), L3 = a1 / L0, L6 = a2 / L0;
        double L4 = std::sqrt(a4 - L3 * L3), L7 = (a5 - L3 * L6) / L4;
        double L8 = std::sqrt(a8 - L6 * L6 - L7 * L7);
        double M0 = 1 / L0, M3 = -L3 * M0 / L4, M4 = 1 / L4;
        double M6 = (-L6 * M0 - L7 * M3) / L8, M7 = -L7 * M4 / L8, M8 = 1 / L8;
        double t0 = M0 * B[0], t1 = M3 * B[0] + M4 * B[1], t2 = M6 * B[0] + M7 * B[1] + M8 * B[2];
        C[0] = M0 * t0 + M3 * t1 + M6 * t2;
        C[1] = M4 * t1 + M7 * t2;
        C[2] = M8 * t2;
    }
    double interp(double x, double y) const { return C[0] * x + C[1] * y + C[2]; }
};

double value_for_pixel(const uint8_t* im, int width, int height, int stride, double px, double py) {
    int x1 = (int)std::floor(px - 0.5), x2 = (int)std::ceil(px - 0.5);
    double x = px - 0.5 - x1;
    int y1 = (int)std::floor(py - 0.5), y2 = (int)std::ceil(py - 0.5);
    double y = py - 0.5 - y1;
    if (x1 < 0 || x2 >= width || y1 < 0 || y2 >= height) return -1;
    return im[y1 * stride + x1] * (1 - x) * (1 - y) + im[y1 * stride + x2] * x * (1 - y) +
           im[y2 * stride + x1] * (1 - x) * y + im[y2 * stride + x2] * x * y;
}

uint64_t rotate90(uint64_t w, int numBits) {
    int p = numBits;
    uint64_t l = 0;
    if (numBits % 4 == 1) { p = numBits - 1; l = 1; }
    w = ((w >> l) << (p / 4 + l)) | (w >> (3 * p / 4 + l) << l) | (w & l);
    w &= ((UINT64_C(1) << numBits) - 1);
    return w;
}

struct DecodeEntry {
    int id = 65535, hamming = 255, rotation = 0;
};

void quick_decode(const Family& f, int maxhamming, uint64_t rcode, DecodeEntry& e) {
    for (int r = 0; r < 4; r++) {
        int best = 256, bid = -1;
        for (int i = 0; i < f.ncodes; i++) {
            int hd = __builtin_popcountll(f.codes[i] ^ rcode);
            if (hd < best) { best = hd; bid = i; }
        }
        if (best <= maxhamming) {
            e.id = bid; e.hamming = best; e.rotation = r;
            return;
        }
        rcode = rotate90(rcode, f.nbits);
    }
    e = DecodeEntry();
}

float quad_decode(const Params& prm, const Family& fam, const uint8_t* im, int width, int height, int stride,
                  const Quad& quad, DecodeEntry& entry) {
    const int wb = fam.width_at_border;
    const float patterns[] = {
        -0.5f, 0.5f, 0, 1, 1,  0.5f, 0.5f, 0, 1, 0,  wb + 0.5f, .5f, 0, 1, 1,  wb - 0.5f, .5f, 0, 1, 0,
        0.5f, -0.5f, 1, 0, 1,  0.5f, 0.5f, 1, 0, 0,  0.5f, wb + 0.5f, 1, 0, 1,  0.5f, wb - 0.5f, 1, 0, 0};
    GrayModel white, black;
    for (int pi = 0; pi < 8; pi++) {
        const float* pat = &patterns[pi * 5];
        int is_white = (int)pat[4];
        for (int i = 0; i < wb; i++) {
            double tagx01 = (pat[0] + i * pat[2]) / wb;
            double tagy01 = (pat[1] + i * pat[3]) / wb;
            double tagx = 2 * (tagx01 - 0.5), tagy = 2 * (tagy01 - 0.5);
            double px, py;
            homography_project(quad.H, tagx, tagy, &px, &py);
            int ix = (int)px, iy = (int)py;
            if (ix < 0 || iy < 0 || ix >= width || iy >= height) continue;
            int v = im[iy * stride + ix];
            if (is_white) white.add(tagx, tagy, v); else black.add(tagx, tagy, v);
        }
    }
    white.solve();
    black.solve();
    if ((white.interp(0, 0) - black.interp(0, 0) < 0) != (fam.reversed_border != 0)) return -1;

    float black_score = 0, white_score = 0, black_count = 1, white_count = 1;
    const int tw = fam.total_width;
    std::vector<double> values((size_t)tw * tw, 0.0), sharp((size_t)tw * tw, 0.0);
    int min_coord = (wb - tw) / 2;
    for (int i = 0; i < fam.nbits; i++) {
        int bity = fam.bit_y[i], bitx = fam.bit_x[i];
        double tagx = 2 * ((bitx + 0.5) / wb - 0.5), tagy = 2 * ((bity + 0.5) / wb - 0.5);
        double px, py;
        homography_project(quad.H, tagx, tagy, &px, &py);
        double v = value_for_pixel(im, width, height, stride, px, py);
        if (v == -1) continue;
        double thresh = (black.interp(tagx, tagy) + white.interp(tagx, tagy)) / 2.0;
        values[tw * (bity - min_coord) + bitx - min_coord] = v - thresh;
    }
    static const double kernel[9] = {0, -1, 0, -1, 4, -1, 0, -1, 0};
    for (int y = 0; y < tw; y++)
        for (int x = 0; x < tw; x++) {
            double acc = 0;
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) {
                    if (y + i - 1 < 0 || y + i - 1 > tw - 1 || x + j - 1 < 0 || x + j - 1 > tw - 1) continue;
                    acc += values[(y + i - 1) * tw + (x + j - 1)] * kernel[i * 3 + j];
                }
            sharp[y * tw + x] = acc;
        }
    for (int i = 0; i < tw * tw; i++) values[i] = values[i] + prm.decode_sharpening * sharp[i];

    uint64_t rcode = 0;
    for (int i = 0; i < fam.nbits; i++) {
        int bity = fam.bit_y[i], bitx = fam.bit_x[i];
        rcode <<= 1;
        double v = values[(bity - min_coord) * tw + bitx - min_coord];
        if (v > 0) {
            white_score += v;
            white_count++;
            rcode |= 1;
        } else {
            black_score -= v;
            black_count++;
        }
    }
    quick_decode(fam, prm.maxhamming, rcode, entry);
    return std::fmin(white_score / white_count, black_score / black_count);
}

// ------------------------------------------------------------------------------------------
// U10 reconcile
// ------------------------------------------------------------------------------------------
bool seg_intersect(const double* a, const double* b, const double* c, const double* d) {
    auto orient = [](const double* p, const double* q, const double* r) {
        return (q[0] - p[0]) * (r[1] - p[1]) - (q[1] - p[1]) * (r[0] - p[0]);
    };
    double o1 = orient(a, b, c), o2 = orient(a, b, d), o3 = orient(c, d, a), o4 = orient(c, d, b);
    return ((o1 > 0) != (o2 > 0)) && ((o3 > 0) != (o4 > 0));
}

bool poly_contains(const double p[4][2], const double* q) {
    bool in = false;
    for (int i = 0, j = 3; i < 4; j = i++) {
        if (((p[i][1] > q[1]) != (p[j][1] > q[1])) &&
            (q[0] < (p[j][0] - p[i][0]) * (q[1] - p[i][1]) / (p[j][1] - p[i][1]) + p[i][0]))
            in = !in;
    }
    return in;
}

bool polys_overlap(const double a[4][2], const double b[4][2]) {
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            if (seg_intersect(a[i], a[(i + 1) & 3], b[j], b[(j + 1) & 3])) return true;
    return poly_contains(a, b[0]) || poly_contains(b, a[0]);
}

// <0: keep a, >0: keep b
int prefer(const ao_detection& a, const ao_detection& b) {
    if (a.hamming != b.hamming) return a.hamming < b.hamming ? -1 : 1;
    if (a.margin != b.margin) return a.margin > b.margin ? -1 : 1;
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 2; k++)
            if (a.p[i][k] != b.p[i][k]) return a.p[i][k] < b.p[i][k] ? -1 : 1;
    return -1;
}

void reconcile(std::vector<ao_detection>& dets) {
    std::sort(dets.begin(), dets.end(), [](const ao_detection& a, const ao_detection& b) {
        if (a.id != b.id) return a.id < b.id;
        if (a.family != b.family) return a.family < b.family;
        if (a.c[0] != b.c[0]) return a.c[0] < b.c[0];
        return a.c[1] < b.c[1];
    });
    std::vector<char> dead(dets.size(), 0);
    for (size_t i = 0; i < dets.size(); i++) {
        if (dead[i]) continue;
        for (size_t j = i + 1; j < dets.size() && dets[j].id == dets[i].id; j++) {
            if (dead[j] || dets[j].family != dets[i].family) continue;
            if (!polys_overlap(dets[i].p, dets[j].p)) continue;
            if (prefer(dets[i], dets[j]) < 0) {
                dead[j] = 1;
            } else {
                dead[i] = 1;
                break;
            }
        }
    }
    std::vector<ao_detection> out;
    for (size_t i = 0; i < dets.size(); i++)
        if (!dead[i]) out.push_back(dets[i]);
    std::sort(out.begin(), out.end(), [](const ao_detection& a, const ao_detection& b) {
        if (a.id != b.id) return a.id < b.id;
        if (a.c[0] != b.c[0]) return a.c[0] < b.c[0];
        return a.c[1] < b.c[1];
    });
    dets.swap(out);
}

// ------------------------------------------------------------------------------------------
// top level
// ------------------------------------------------------------------------------------------
int detect_one(const Detector& D, const uint8_t* im, int w, int h, int stride, std::vector<ao_detection>& dets,
               ao_debug* dbg) {
    const Params& prm = D.prm;
    dets.clear();
    int f = (int)prm.quad_decimate;
    if (f < 1) f = 1;
    std::vector<uint8_t> qim;
    int wd, hd;
    decimate(im, w, h, stride, f, qim, wd, hd);
    blur_stage(qim, wd, hd, prm.quad_sigma);
    std::vector<uint8_t> thr;
    threshold(qim, wd, hd, prm.min_white_black_diff, thr);
    std::vector<uint32_t> labels, sizes;
    connected_components(thr, wd, hd, labels, sizes);
    std::vector<KeyedPt> kpts;
    gradient_clusters(thr, wd, hd, labels, sizes, kpts);

    bool normal_border = false, reversed_border = false;
    int min_tag_width = 1000000;
    for (const Family& fm : D.fams) {
        min_tag_width = std::min(min_tag_width, fm.width_at_border);
        normal_border |= !fm.reversed_border;
        reversed_border |= fm.reversed_border != 0;
    }
    min_tag_width = (int)(min_tag_width / prm.quad_decimate);
    if (min_tag_width < 3) min_tag_width = 3;

    if (dbg) {
        dbg->wd = wd; dbg->hd = hd;
        if (dbg->quad_im) memcpy(dbg->quad_im, qim.data(), qim.size());
        if (dbg->thresh) memcpy(dbg->thresh, thr.data(), thr.size());
        if (dbg->labels) memcpy(dbg->labels, labels.data(), labels.size() * 4);
        if (dbg->sizes) memcpy(dbg->sizes, sizes.data(), sizes.size() * 4);
        dbg->npoints = (int)kpts.size();
        dbg->nclusters = 0;
        dbg->nquads = 0;
        dbg->noversize = 0;
    }

    std::vector<Quad> quads;
    std::vector<uint64_t> quad_keys;
    std::vector<Pt> cl;
    const int max_cluster = 3 * (2 * wd + 2 * hd);
    for (size_t s = 0; s < kpts.size();) {
        size_t e = s;
        while (e < kpts.size() && kpts[e].key == kpts[s].key) e++;
        int csz = (int)(e - s);
        if (dbg) {
            if (dbg->cluster_keys && dbg->nclusters < dbg->cap_clusters) {
                dbg->cluster_keys[dbg->nclusters] = kpts[s].key;
                dbg->cluster_sizes[dbg->nclusters] = csz;
            }
            dbg->nclusters++;
        }
        if (dbg && csz > max_cluster) dbg->noversize++;
        if (csz >= prm.min_cluster_pixels && csz <= max_cluster) {
            cl.clear();
            for (size_t i = s; i < e; i++) cl.push_back(kpts[i].p);
            Quad q;
            if (fit_quad(prm, qim.data(), wd, hd, cl, min_tag_width, normal_border, reversed_border, q)) {
                quads.push_back(q);
                quad_keys.push_back(kpts[s].key);
            }
        }
        s = e;
    }
    if (prm.quad_decimate > 1)
        for (Quad& q : quads)
            for (int j = 0; j < 4; j++) {
                q.p[j][0] = (q.p[j][0] - 0.5f) * prm.quad_decimate + 0.5f;
                q.p[j][1] = (q.p[j][1] - 0.5f) * prm.quad_decimate + 0.5f;
            }
    if (dbg) {
        dbg->nquads = (int)quads.size();
        for (int i = 0; i < (int)quads.size() && i < dbg->cap_quads; i++) {
            if (dbg->quads) {
                memcpy(&dbg->quads[i * 9], quads[i].p, 32);
                dbg->quads[i * 9 + 8] = (float)quads[i].reversed_border;
            }
            if (dbg->quad_keys) dbg->quad_keys[i] = quad_keys[i];
        }
    }
    for (size_t qi = 0; qi < quads.size(); qi++) {
        Quad& q = quads[qi];
        if (prm.refine_edges) refine_edges(prm, im, w, h, stride, q);
        if (dbg && dbg->quads_refined && (int)qi < dbg->cap_quads) memcpy(&dbg->quads_refined[qi * 8], q.p, 32);
        if (!quad_update_homographies(q)) continue;
        for (size_t fi = 0; fi < D.fams.size(); fi++) {
            const Family& fam = D.fams[fi];
            if ((fam.reversed_border != 0) != (q.reversed_border != 0)) continue;
            DecodeEntry entry;
            float margin = quad_decode(prm, fam, im, w, h, stride, q, entry);
            if (margin >= 0 && entry.hamming < 255) {
                ao_detection d;
                memset(&d, 0, sizeof(d));
                d.family = (int)fi; d.id = entry.id; d.hamming = entry.hamming; d.margin = margin;
                double theta = entry.rotation * M_PI / 2.0;
                double c = std::cos(theta), s = std::sin(theta);
                const double R[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
                for (int r = 0; r < 3; r++)
                    for (int cc = 0; cc < 3; cc++) {
                        double acc = 0;
                        for (int k = 0; k < 3; k++) acc += q.H[r * 3 + k] * R[k * 3 + cc];
                        d.H[r * 3 + cc] = acc;
                    }
                homography_project(d.H, 0, 0, &d.c[0], &d.c[1]);
                for (int i = 0; i < 4; i++) {
                    int tcx = (i == 1 || i == 2) ? 1 : -1;
                    int tcy = (i < 2) ? 1 : -1;
                    homography_project(d.H, tcx, tcy, &d.p[i][0], &d.p[i][1]);
                }
                dets.push_back(d);
            }
        }
    }
    reconcile(dets);
    return (int)dets.size();
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C interface (ctypes)
// ------------------------------------------------------------------------------------------
extern "C" {

void* ao_create(float decimate, float sigma, int refine_edges, double sharpening, int maxhamming) {
    Detector* D = new Detector();
    D->prm.quad_decimate = decimate;
    D->prm.quad_sigma = sigma;
    D->prm.refine_edges = refine_edges;
    D->prm.decode_sharpening = sharpening;
    D->prm.maxhamming = maxhamming;
    return D;
}

void ao_destroy(void* h) { delete (Detector*)h; }

int ao_add_family(void* h, const char* name, int nbits, int hmin, int ncodes, int wb, int tw, int reversed,
                  const uint64_t* codes, const int* bit_x, const int* bit_y) {
    Detector* D = (Detector*)h;
    Family f;
    f.name = name; f.nbits = nbits; f.h = hmin; f.ncodes = ncodes; f.width_at_border = wb; f.total_width = tw;
    f.reversed_border = reversed;
    f.codes.assign(codes, codes + ncodes);
    f.bit_x.assign(bit_x, bit_x + nbits);
    f.bit_y.assign(bit_y, bit_y + nbits);
    D->fams.push_back(f);
    return (int)D->fams.size() - 1;
}

int ao_detect(void* h, const uint8_t* im, int w, int hgt, int stride, ao_detection* out, int cap, ao_debug* dbg) {
    Detector* D = (Detector*)h;
    std::vector<ao_detection> dets;
    int n = detect_one(*D, im, w, hgt, stride, dets, dbg);
    for (int i = 0; i < n && i < cap; i++) out[i] = dets[i];
    return n;
}

// frames: [B][h][stride]; frames are independent units and are spread over nthreads host threads
int ao_detect_batch(void* h, const uint8_t* frames, int B, int w, int hgt, int stride, int nthreads,
                    ao_detection* out, int cap_per_frame, int* counts) {
    Detector* D = (Detector*)h;
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next(0);
    auto worker = [&]() {
        std::vector<ao_detection> dets;
        for (;;) {
            int b = next.fetch_add(1);
            if (b >= B) break;
            int n = detect_one(*D, frames + (size_t)b * hgt * stride, w, hgt, stride, dets, nullptr);
            counts[b] = n;
            for (int i = 0; i < n && i < cap_per_frame; i++) out[(size_t)b * cap_per_frame + i] = dets[i];
        }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; i++) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    return 0;
}

// single stages, for stage-level parity tests
void ao_stage_threshold(const uint8_t* im, int w, int h, int min_diff, uint8_t* out) {
    std::vector<uint8_t> in(im, im + (size_t)w * h), o;
    threshold(in, w, h, min_diff, o);
    memcpy(out, o.data(), o.size());
}

void ao_stage_labels(const uint8_t* thr, int w, int h, uint32_t* labels, uint32_t* sizes) {
    std::vector<uint8_t> in(thr, thr + (size_t)w * h);
    std::vector<uint32_t> l, s;
    connected_components(in, w, h, l, s);
    memcpy(labels, l.data(), l.size() * 4);
    memcpy(sizes, s.data(), s.size() * 4);
}

void ao_stage_blur(uint8_t* im, int w, int h, float sigma) {
    std::vector<uint8_t> v(im, im + (size_t)w * h);
    blur_stage(v, w, h, sigma);
    memcpy(im, v.data(), v.size());
}

uint64_t ao_rotate90(uint64_t w, int nbits) { return rotate90(w, nbits); }

int ao_version(void) { return 1; }
}
