// k_pose.cuh -- per-tag pose, one 4-lane thread group per tag (one lane per corner).
//
// Replaces cv2.solvePnP(obj_points, corners, K, dist) + cv2.Rodrigues of
// /root/reference/src/detection/tag_detector.py:30-52 for the tag's four coplanar corners
// (object points (+-s/2, +-s/2, 0) in lb, rb, rt, lt order, tag_detector.py:35-38; corners are
// cast to float32 first, tag_detector.py:32, and so are they here).
//
// method 0 (reference behaviour): the planar initialisation OpenCV's SOLVEPNP_ITERATIVE uses
//   (undistort -> 4-point homography -> [h1 h2 h1xh2] orthogonalised) followed by
//   Levenberg-Marquardt on the PIXEL reprojection error with the full distortion model, run to
//   convergence -- it lands on the same minimiser cv2.solvePnP reports.
// method 1: same initialisation + orthogonal iteration (Lu/Hager/Mjolsness, object-space error),
//   the scheme upstream apriltag's estimate_tag_pose uses.
#pragma once
#include "common.cuh"

struct PoseArgs {
    const double* corners;   // [M][4][2] (stride in doubles between tags given by corner_stride)
    int corner_stride;       // doubles between consecutive tags
    const int* counts;       // optional per-frame valid counts (detect_pose path), else null
    int per_frame;           // slots per frame when counts != null
    int M;
    double fx, fy, cx, cy;
    double dist[8];
    int ndist;
    double half;             // tag_size / 2 (after the float32 cast the reference applies)
    int method;
    PoseRec* out;
};

// sum over the 4 lanes of a tag group; `m` is the group's own lane mask, so groups of one warp may
// diverge from each other (different iteration counts) without dead-locking
__device__ __forceinline__ double grp_sum_m(unsigned m, double v) {
    v += __shfl_xor_sync(m, v, 1);
    v += __shfl_xor_sync(m, v, 2);
    return v;
}
#define grp_sum(v) grp_sum_m(gmask, (v))

__device__ __forceinline__ void distort(const PoseArgs& a, double x, double y, double& xd, double& yd,
                                        double (&J)[4]) {
    // OpenCV model: k1 k2 p1 p2 k3 k4 k5 k6
    const double k1 = a.dist[0], k2 = a.dist[1], p1 = a.dist[2], p2 = a.dist[3], k3 = a.dist[4];
    const double k4 = a.dist[5], k5 = a.dist[6], k6 = a.dist[7];
    double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    double num = 1 + k1 * r2 + k2 * r4 + k3 * r6, den = 1 + k4 * r2 + k5 * r4 + k6 * r6;
    double icd = 1.0 / den, cdist = num * icd;
    double dnum = k1 + 2 * k2 * r2 + 3 * k3 * r4, dden = k4 + 2 * k5 * r2 + 3 * k6 * r4;  // d/d(r2)
    double dc = (dnum * den - num * dden) * icd * icd;                                    // d cdist / d r2
    double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
    xd = x * cdist + p1 * a1 + p2 * a2;
    yd = y * cdist + p1 * a3 + p2 * a1;
    // d(xd,yd)/d(x,y)
    J[0] = cdist + x * dc * 2 * x + p1 * 2 * y + p2 * (2 * x + 4 * x);
    J[1] = x * dc * 2 * y + p1 * 2 * x + p2 * 2 * y;
    J[2] = y * dc * 2 * x + p1 * 2 * x + p2 * 2 * y;
    J[3] = cdist + y * dc * 2 * y + p1 * (2 * y + 4 * y) + p2 * 2 * x;
}

__device__ void rodrigues_to_vec(const double (&R)[9], double (&r)[3]) {  // cv::Rodrigues, matrix -> vector
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) { r[0] = r[1] = r[2] = 0; return; }
        double t = (R[0] + 1) * 0.5;
        rx = sqrt(fmax(t, 0.));
        t = (R[4] + 1) * 0.5;
        ry = sqrt(fmax(t, 0.)) * (R[1] < 0 ? -1. : 1.);
        t = (R[8] + 1) * 0.5;
        rz = sqrt(fmax(t, 0.)) * (R[2] < 0 ? -1. : 1.);
        if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
        theta /= sqrt(rx * rx + ry * ry + rz * rz);
        r[0] = rx * theta; r[1] = ry * theta; r[2] = rz * theta;
    } else {
        double vth = 1 / (2 * s);
        vth *= theta;
        r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
    }
}

__device__ void exp_so3(const double (&w)[3], double (&E)[9]) {
    double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
    double A, B;
    if (th < 1e-8) { A = 1 - th2 / 6; B = 0.5 - th2 / 24; }
    else { A = sin(th) / th; B = (1 - cos(th)) / th2; }
    const double wx = w[0], wy = w[1], wz = w[2];
    E[0] = 1 - B * (wy * wy + wz * wz); E[1] = -A * wz + B * wx * wy;      E[2] = A * wy + B * wx * wz;
    E[3] = A * wz + B * wx * wy;        E[4] = 1 - B * (wx * wx + wz * wz); E[5] = -A * wx + B * wy * wz;
    E[6] = -A * wy + B * wx * wz;       E[7] = A * wx + B * wy * wz;        E[8] = 1 - B * (wx * wx + wy * wy);
}

__device__ __forceinline__ void mat3_mul(const double (&A)[9], const double (&B)[9], double (&C)[9]) {
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}

__device__ bool inv3(const double (&M)[9], double (&I)[9]) {
    double c0 = M[4] * M[8] - M[5] * M[7], c1 = M[5] * M[6] - M[3] * M[8], c2 = M[3] * M[7] - M[4] * M[6];
    double det = M[0] * c0 + M[1] * c1 + M[2] * c2;
    if (fabs(det) < 1e-300) return false;
    double id = 1.0 / det;
    I[0] = c0 * id; I[1] = (M[2] * M[7] - M[1] * M[8]) * id; I[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    I[3] = c1 * id; I[4] = (M[0] * M[8] - M[2] * M[6]) * id; I[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    I[6] = c2 * id; I[7] = (M[1] * M[6] - M[0] * M[7]) * id; I[8] = (M[0] * M[4] - M[1] * M[3]) * id;
    return true;
}

// nearest rotation (polar factor U*Vt of M, det > 0) by Newton iteration R <- (R + R^-T)/2
__device__ bool orthogonalize(double (&R)[9]) {
    for (int it = 0; it < 30; it++) {
        double I[9];
        if (!inv3(R, I)) return false;
        double diff = 0;
        double N[9];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                N[r * 3 + c] = 0.5 * (R[r * 3 + c] + I[c * 3 + r]);
                diff += fabs(N[r * 3 + c] - R[r * 3 + c]);
            }
#pragma unroll
        for (int k = 0; k < 9; k++) R[k] = N[k];
        if (diff < 1e-15) break;
    }
    return true;
}

// 6x6 SPD solve (Cholesky) on the packed lower triangle S[i*(i+1)/2 + j], j <= i, with `lambda * (S_ii + 1e-12)` added
// to the diagonal (Levenberg-Marquardt damping); every loop has constant bounds, so all of it lives in registers.
// Returns false when the damped matrix is not positive definite.
#define SYM(i, j) ((i) * ((i) + 1) / 2 + (j))
__device__ __forceinline__ bool solve6_sym(const double (&S)[21], double lambda, const double (&b)[6], double (&x)[6]) {
    // (one reciprocal square root per column instead of a square root and up to eleven divisions: the FP64 divide chain
    // was 60 % of the kernel's stall samples.  The step only has to be a descent step; the fixed point is unchanged.)
    double L[21], inv[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
#pragma unroll
        for (int j = 0; j <= i; j++) {
            double s = S[SYM(i, j)];
            if (i == j) s += lambda * (s + 1e-12);
#pragma unroll
            for (int k = 0; k < j; k++) s -= L[SYM(i, k)] * L[SYM(j, k)];
            if (i == j) {
                if (!(s > 0)) return false;
                inv[i] = rsqrt(s);
                L[SYM(i, i)] = s * inv[i];
            } else {
                L[SYM(i, j)] = s * inv[j];
            }
        }
    }
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double s = b[i];
#pragma unroll
        for (int k = 0; k < i; k++) s -= L[SYM(i, k)] * y[k];
        y[i] = s * inv[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        double s = y[i];
#pragma unroll
        for (int k = i + 1; k < 6; k++) s -= L[SYM(k, i)] * x[k];
        x[i] = s * inv[i];
    }
    return true;
}

__global__ void __launch_bounds__(128)
k_pose(PoseArgs a) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int tag = gid >> 2, c = gid & 3;
    bool active = tag < a.M;
    if (active && a.counts) {
        int frame = tag / a.per_frame, slot = tag - frame * a.per_frame;
        active = slot < min(a.counts[frame], a.per_frame);
    }
    if (!active) return;  // whole 4-lane groups leave together; all shuffles below are group-masked
    const unsigned gmask = 0xFu << ((threadIdx.x & 31) & ~3);
    const double* cp = a.corners + (size_t)tag * a.corner_stride;
    // the reference casts corners and object points to float32 (tag_detector.py:32,35-38)
    const double u = active ? (double)(float)cp[2 * c] : 0.0, v = active ? (double)(float)cp[2 * c + 1] : 0.0;
    const double X = ((c == 1 || c == 2) ? 1.0 : -1.0) * a.half;
    const double Y = ((c >= 2) ? 1.0 : -1.0) * a.half;

    // ---- normalised, undistorted image point (fixed-point iteration like cv::undistortPoints)
    double xn = (u - a.cx) / a.fx, yn = (v - a.cy) / a.fy;
    if (a.ndist > 0) {
        const double x0 = xn, y0 = yn;
        for (int it = 0; it < 20; it++) {
            double r2 = xn * xn + yn * yn;
            double icdist = (1 + ((a.dist[7] * r2 + a.dist[6]) * r2 + a.dist[5]) * r2) /
                            (1 + ((a.dist[4] * r2 + a.dist[1]) * r2 + a.dist[0]) * r2);
            double dX = 2 * a.dist[2] * xn * yn + a.dist[3] * (r2 + 2 * xn * xn);
            double dY = a.dist[2] * (r2 + 2 * yn * yn) + 2 * a.dist[3] * xn * yn;
            xn = (x0 - dX) * icdist;
            yn = (y0 - dY) * icdist;
        }
    }
    // ---- homography (X, Y) -> (xn, yn).  The object points are the corners of a square, so the 4-point homography has
    //      a closed form (the projective square-to-quadrilateral map): ~40 flops in registers on every lane instead of
    //      an 8x9 elimination in local memory.  (The pose oracle is cv2.solvePnP, matched through the minimiser the
    //      iteration below converges to; the start value only has to be in its basin, as OpenCV's own is.)
    double h[9];
    bool ok;
    {
        const int gbase = (threadIdx.x & 31) & ~3;
        const double x0 = __shfl_sync(gmask, xn, gbase), y0 = __shfl_sync(gmask, yn, gbase);
        const double x1 = __shfl_sync(gmask, xn, gbase + 1), y1 = __shfl_sync(gmask, yn, gbase + 1);
        const double x2 = __shfl_sync(gmask, xn, gbase + 2), y2 = __shfl_sync(gmask, yn, gbase + 2);
        const double x3 = __shfl_sync(gmask, xn, gbase + 3), y3 = __shfl_sync(gmask, yn, gbase + 3);
        const double sx = x0 - x1 + x2 - x3, sy = y0 - y1 + y2 - y3;
        const double dx1 = x1 - x2, dx2 = x3 - x2, dy1 = y1 - y2, dy2 = y3 - y2;
        const double den = dx1 * dy2 - dy1 * dx2;
        ok = fabs(den) > 1e-300;
        const double id = ok ? 1.0 / den : 0.0;
        const double g = (sx * dy2 - sy * dx2) * id, hh = (dx1 * sy - dy1 * sx) * id;
        // unit square (s, t) -> image: [a b c; d e f; g hh 1], with s = X / (2 half) + 1/2, t = Y / (2 half) + 1/2
        const double ua = x1 - x0 + g * x1, ub = x3 - x0 + hh * x3, uc = x0;
        const double ud = y1 - y0 + g * y1, ue = y3 - y0 + hh * y3, uf = y0;
        const double i2 = 0.5 / a.half;
        const double w = 1.0 + 0.5 * (g + hh);
        const double iw = (ok && fabs(w) > 1e-300) ? 1.0 / w : 1.0;
        h[0] = ua * i2 * iw; h[1] = ub * i2 * iw; h[2] = (uc + 0.5 * (ua + ub)) * iw;
        h[3] = ud * i2 * iw; h[4] = ue * i2 * iw; h[5] = (uf + 0.5 * (ud + ue)) * iw;
        h[6] = g * i2 * iw;  h[7] = hh * i2 * iw; h[8] = 1.0;
        if (!ok) { h[0] = 1; h[1] = 0; h[2] = 0; h[3] = 0; h[4] = 1; h[5] = 0; h[6] = 0; h[7] = 0; }
    }
    // ---- decomposition as in OpenCV's planar branch: normalise h1, h2; t = h3 * 2/(|h1|+|h2|)
    double R[9], t[3];
    {
        double n1 = sqrt(h[0] * h[0] + h[3] * h[3] + h[6] * h[6]);
        double n2 = sqrt(h[1] * h[1] + h[4] * h[4] + h[7] * h[7]);
        double i1 = 1. / fmax(n1, 2.220446049250313e-16), i2 = 1. / fmax(n2, 2.220446049250313e-16);
        double a0 = h[0] * i1, a1 = h[3] * i1, a2 = h[6] * i1;
        double b0 = h[1] * i2, b1 = h[4] * i2, b2 = h[7] * i2;
        double sc = 2. / fmax(n1 + n2, 2.220446049250313e-16);
        t[0] = h[2] * sc; t[1] = h[5] * sc; t[2] = h[8] * sc;
        R[0] = a0; R[3] = a1; R[6] = a2;
        R[1] = b0; R[4] = b1; R[7] = b2;
        R[2] = a1 * b2 - a2 * b1; R[5] = a2 * b0 - a0 * b2; R[8] = a0 * b1 - a1 * b0;
        if (!ok || !orthogonalize(R)) {
            R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
            t[0] = 0; t[1] = 0; t[2] = 1;
        }
    }
    int iters = 0;
    double cost = 0;
    auto residual = [&](const double (&Rm)[9], const double (&tv)[3], double& ru, double& rv, double (&Jr)[12],
                        bool want_jac) {
        // camera-frame point (Z = 0 on the tag)
        double RX0 = Rm[0] * X + Rm[1] * Y, RX1 = Rm[3] * X + Rm[4] * Y, RX2 = Rm[6] * X + Rm[7] * Y;
        double px = RX0 + tv[0], py = RX1 + tv[1], pz = RX2 + tv[2];
        double iz = 1.0 / pz;
        double x = px * iz, y = py * iz;
        double xd = x, yd = y, Jd[4] = {1, 0, 0, 1};
        if (a.ndist > 0) distort(a, x, y, xd, yd, Jd);
        ru = a.fx * xd + a.cx - u;
        rv = a.fy * yd + a.cy - v;
        if (!want_jac) return;
        // d(x,y)/d(p)
        double dxp[3] = {iz, 0, -x * iz}, dyp[3] = {0, iz, -y * iz};
        double du[3], dv[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            du[k] = a.fx * (Jd[0] * dxp[k] + Jd[1] * dyp[k]);
            dv[k] = a.fy * (Jd[2] * dxp[k] + Jd[3] * dyp[k]);
        }
        // dp/d(omega) = -[RX]x  (left perturbation R <- exp(omega) R), dp/dt = I
        // du . (w x RX) = w . (RX x du)
        Jr[0] = RX1 * du[2] - RX2 * du[1];
        Jr[1] = RX2 * du[0] - RX0 * du[2];
        Jr[2] = RX0 * du[1] - RX1 * du[0];
        Jr[3] = du[0]; Jr[4] = du[1]; Jr[5] = du[2];
        Jr[6] = RX1 * dv[2] - RX2 * dv[1];
        Jr[7] = RX2 * dv[0] - RX0 * dv[2];
        Jr[8] = RX0 * dv[1] - RX1 * dv[0];
        Jr[9] = dv[0]; Jr[10] = dv[1]; Jr[11] = dv[2];
    };

    if (a.method == 0) {
        double lambda = 1e-3;
        double ru, rv, Jr[12];
        residual(R, t, ru, rv, Jr, true);
        cost = grp_sum(ru * ru + rv * rv);
        for (iters = 0; iters < 100; iters++) {
            double S[21], b[6];   // packed lower triangle of J^T J, and -J^T r
#pragma unroll
            for (int i = 0; i < 6; i++) {
                b[i] = -grp_sum(Jr[i] * ru + Jr[6 + i] * rv);
#pragma unroll
                for (int j = 0; j <= i; j++) S[SYM(i, j)] = grp_sum(Jr[i] * Jr[j] + Jr[6 + i] * Jr[6 + j]);
            }
            bool improved = false;
            double step_norm = 0;
            const double cost_before = cost;
            bool at_minimum = false;
            for (int tries = 0; tries < 12 && !improved; tries++) {
                double d[6];
                const bool solved = solve6_sym(S, lambda, b, d);
                double Rn[9], tn[3], E[9];
                double nru = 0, nrv = 0, ncost = 1e300;
                if (solved) {
                    double w3[3] = {d[0], d[1], d[2]};
                    exp_so3(w3, E);
                    mat3_mul(E, R, Rn);
                    tn[0] = t[0] + d[3]; tn[1] = t[1] + d[4]; tn[2] = t[2] + d[5];
                    double Jdummy[12];
                    residual(Rn, tn, nru, nrv, Jdummy, false);
                }
                double nc = grp_sum(solved ? nru * nru + nrv * nrv : 0.0);
                if (solved) ncost = nc;
                if (solved && ncost <= cost) {
                    step_norm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);
#pragma unroll
                    for (int k = 0; k < 9; k++) R[k] = Rn[k];
                    t[0] = tn[0]; t[1] = tn[1]; t[2] = tn[2];
                    cost = ncost;
                    lambda = fmax(lambda * 0.1, 1e-12);
                    improved = true;
                } else {
                    // a proposed step that is already below the resolution of the estimate and still does not lower the
                    // cost: what is left is rounding noise -- stop instead of raising the damping eleven more times
                    if (solved && sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3] + d[4] * d[4] + d[5] * d[5]) <
                                      1e-9 * (1 + sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]))) {
                        at_minimum = true;
                        break;
                    }
                    lambda *= 10;
                }
            }
            if (!improved || at_minimum) break;
            residual(R, t, ru, rv, Jr, true);
            double tn2 = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
            if (step_norm < 1e-11 * (1 + tn2)) { iters++; break; }   // quadratic convergence: the next step would be ~1e-20
            if (cost_before - cost <= 1e-14 * cost_before) { iters++; break; }   // at the minimum: only rounding noise is left
        }
        orthogonalize(R);
    } else {
        // orthogonal iteration: V_i = v v^T / (v^T v), t(R) = (I - mean V)^-1 mean((V_i - I) R p_i), R from the
        // polar factor of sum (q_i - qbar)(p_i - pbar)^T
        const double vx = xn, vy = yn, vz = 1.0;
        const double vv = vx * vx + vy * vy + vz * vz;
        double V[9] = {vx * vx / vv, vx * vy / vv, vx * vz / vv, vy * vx / vv, vy * vy / vv, vy * vz / vv,
                       vz * vx / vv, vz * vy / vv, vz * vz / vv};
        double Vm[9], G[9], Gi[9];
#pragma unroll
        for (int k = 0; k < 9; k++) Vm[k] = grp_sum(V[k]) * 0.25;
#pragma unroll
        for (int k = 0; k < 9; k++) G[k] = ((k == 0 || k == 4 || k == 8) ? 1.0 : 0.0) - Vm[k];
        bool gi_ok = inv3(G, Gi);
        double prev_err = 1e300;
        for (iters = 0; iters < 200 && gi_ok; iters++) {
            // optimal translation for the current R
            double Rp[3] = {R[0] * X + R[1] * Y, R[3] * X + R[4] * Y, R[6] * X + R[7] * Y};
            double w3[3];
#pragma unroll
            for (int r = 0; r < 3; r++)
                w3[r] = (V[r * 3] * Rp[0] + V[r * 3 + 1] * Rp[1] + V[r * 3 + 2] * Rp[2]) - Rp[r];
            double m3[3] = {grp_sum(w3[0]) * 0.25, grp_sum(w3[1]) * 0.25, grp_sum(w3[2]) * 0.25};
#pragma unroll
            for (int r = 0; r < 3; r++) t[r] = Gi[r * 3] * m3[0] + Gi[r * 3 + 1] * m3[1] + Gi[r * 3 + 2] * m3[2];
            // projected points q_i = V_i (R p_i + t)
            double P3[3] = {Rp[0] + t[0], Rp[1] + t[1], Rp[2] + t[2]};
            double q[3];
#pragma unroll
            for (int r = 0; r < 3; r++) q[r] = V[r * 3] * P3[0] + V[r * 3 + 1] * P3[1] + V[r * 3 + 2] * P3[2];
            double e0 = P3[0] - q[0], e1 = P3[1] - q[1], e2 = P3[2] - q[2];
            double err = grp_sum(e0 * e0 + e1 * e1 + e2 * e2);
            double qm[3] = {grp_sum(q[0]) * 0.25, grp_sum(q[1]) * 0.25, grp_sum(q[2]) * 0.25};
            // M = sum (q_i - qm) p_i^T   (object points are centred: pbar = 0, Z = 0)
            double M3[9];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                M3[r * 3] = grp_sum((q[r] - qm[r]) * X);
                M3[r * 3 + 1] = grp_sum((q[r] - qm[r]) * Y);
                M3[r * 3 + 2] = 0;
            }
            // rank-2 M: complete the third column with the cross product so that the polar factor is a rotation
            double c0[3] = {M3[0], M3[3], M3[6]}, c1[3] = {M3[1], M3[4], M3[7]};
            double n0 = sqrt(c0[0] * c0[0] + c0[1] * c0[1] + c0[2] * c0[2]);
            double n1 = sqrt(c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2]);
            double sc = sqrt(n0 * n1);
            double cr[3] = {c0[1] * c1[2] - c0[2] * c1[1], c0[2] * c1[0] - c0[0] * c1[2], c0[0] * c1[1] - c0[1] * c1[0]};
            double ncr = sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
            if (ncr < 1e-300) break;
            M3[2] = cr[0] / ncr * sc; M3[5] = cr[1] / ncr * sc; M3[8] = cr[2] / ncr * sc;
            double Rn[9];
#pragma unroll
            for (int k = 0; k < 9; k++) Rn[k] = M3[k];
            if (!orthogonalize(Rn)) break;
#pragma unroll
            for (int k = 0; k < 9; k++) R[k] = Rn[k];
            if (fabs(prev_err - err) < 1e-16 * (1 + err)) { iters++; break; }
            prev_err = err;
        }
        // final translation for the final R
        if (gi_ok) {
            double Rp[3] = {R[0] * X + R[1] * Y, R[3] * X + R[4] * Y, R[6] * X + R[7] * Y};
            double w3[3];
#pragma unroll
            for (int r = 0; r < 3; r++)
                w3[r] = (V[r * 3] * Rp[0] + V[r * 3 + 1] * Rp[1] + V[r * 3 + 2] * Rp[2]) - Rp[r];
            double m3[3] = {grp_sum(w3[0]) * 0.25, grp_sum(w3[1]) * 0.25, grp_sum(w3[2]) * 0.25};
#pragma unroll
            for (int r = 0; r < 3; r++) t[r] = Gi[r * 3] * m3[0] + Gi[r * 3 + 1] * m3[1] + Gi[r * 3 + 2] * m3[2];
        }
        double ru, rv, Jd[12];
        residual(R, t, ru, rv, Jd, false);
        cost = grp_sum(ru * ru + rv * rv);
    }
    if (active && c == 0) {
        PoseRec o;
        rodrigues_to_vec(R, o.rvec);
#pragma unroll
        for (int k = 0; k < 3; k++) o.tvec[k] = t[k];
#pragma unroll
        for (int k = 0; k < 9; k++) o.R[k] = R[k];
        o.err = sqrt(cost / 4.0);
        bool finite = isfinite(cost) && isfinite(t[0]) && isfinite(t[1]) && isfinite(t[2]);
        o.ok = (ok && finite) ? 1 : 0;
        o.iters = iters;
        a.out[tag] = o;
    }
}
