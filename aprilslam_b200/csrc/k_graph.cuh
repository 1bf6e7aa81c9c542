// k_graph.cuh -- the step AFTER the hot path: the tag graph update and the camera pose estimate
// of /root/reference/src/core/slam_graph.py:29-70 (SLAMGraph.add_or_update_node / find_world / get_world) and
// /root/reference/src/core/slam.py:36-63 (SLAM.my_pose), batched over independent camera streams.
//
// The update is sequential by construction (every detection reads the graph the previous one wrote), so the unit
// of parallelism is the STREAM: one thread walks the frames of one camera in order, over the detection + pose
// records exactly as agpu_detect_pose returns them; the graph of every stream stays resident in HBM between calls.
#pragma once
#include "common.cuh"

struct GraphArgs {
    int S, F, cap, nid;
    const DetRec* dets;      // [S][F][cap]
    const PoseRec* poses;    // [S][F][cap]
    const int* counts;       // [S][F]
    // graph state, per stream
    int* coordinate_id;      // [S]  (-1: none yet)
    double* estimated_pose;  // [S][16]
    unsigned char* present;  // [S][nid]
    unsigned char* updated;  // [S][nid]
    unsigned char* visible;  // [S][nid]
    int* reference;          // [S][nid]
    int* weight;             // [S][nid]
    double* local;           // [S][nid][16]
    double* world;           // [S][nid][16]
    // per frame results
    double* my_pose;         // [S][F][16]
    unsigned char* valid;    // [S][F]   0: my_pose() returned None
    int* skipped;            // [S]      detections the graph could not place ("Cannot find world reference") or ids out of range
};

__device__ inline void g_matmul(const double* A, const double* B, double* C) {
    double t[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double acc = 0;
            for (int k = 0; k < 4; k++) acc += A[i * 4 + k] * B[k * 4 + j];
            t[i * 4 + j] = acc;
        }
    for (int i = 0; i < 16; i++) C[i] = t[i];
}

// np.linalg.inv of a 4x4: LU with partial pivoting (what LAPACK getrf/getri do), here as Gauss-Jordan on [A | I]
__device__ inline void g_inv4(const double* A, double* inv) {
    double M[4][8];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            M[i][j] = A[i * 4 + j];
            M[i][4 + j] = i == j ? 1.0 : 0.0;
        }
    for (int c = 0; c < 4; c++) {
        int piv = c;
        double best = fabs(M[c][c]);
        for (int r = c + 1; r < 4; r++)
            if (fabs(M[r][c]) > best) { best = fabs(M[r][c]); piv = r; }
        if (piv != c)
            for (int j = 0; j < 8; j++) { double t = M[c][j]; M[c][j] = M[piv][j]; M[piv][j] = t; }
        const double d = 1.0 / M[c][c];
        for (int j = 0; j < 8; j++) M[c][j] *= d;
        for (int r = 0; r < 4; r++) {
            if (r == c) continue;
            const double f = M[r][c];
            for (int j = 0; j < 8; j++) M[r][j] -= f * M[c][j];
        }
    }
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) inv[i * 4 + j] = M[i][4 + j];
}

__global__ void __launch_bounds__(32)
k_graph_update(GraphArgs a) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.S) return;
    unsigned char* present = a.present + (size_t)s * a.nid;
    unsigned char* updated = a.updated + (size_t)s * a.nid;
    unsigned char* visible = a.visible + (size_t)s * a.nid;
    int* reference = a.reference + (size_t)s * a.nid;
    int* weight = a.weight + (size_t)s * a.nid;
    double* local = a.local + (size_t)s * a.nid * 16;
    double* world = a.world + (size_t)s * a.nid * 16;
    int coord = a.coordinate_id[s];
    int skipped = a.skipped[s];
    const double I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    for (int f = 0; f < a.F; f++) {
        const size_t base = ((size_t)s * a.F + f) * a.cap;
        const int n = min(max(a.counts[(size_t)s * a.F + f], 0), a.cap);
        // SLAM.detect (slam.py:21-25): visible_tags = the ids of ALL detections of the frame
        int ref_min = 0x7fffffff;
        for (int i = 0; i < n; i++) ref_min = min(ref_min, a.dets[base + i].id);
        // SLAM.get_pose per detection (slam.py:27-32): the graph is only touched when solvePnP succeeded
        for (int i = 0; i < n; i++) {
            const PoseRec& P = a.poses[base + i];
            if (!P.ok) continue;
            const int id = a.dets[base + i].id;
            if (id < 0 || id >= a.nid) { skipped++; continue; }
            double T[16] = {P.R[0], P.R[1], P.R[2], P.tvec[0], P.R[3], P.R[4], P.R[5], P.tvec[1],
                            P.R[6], P.R[7], P.R[8], P.tvec[2], 0, 0, 0, 1};
            double Ti[16];
            g_inv4(T, Ti);
            double* L = local + (size_t)id * 16;
            double* Wd = world + (size_t)id * 16;
            // SLAMGraph.add_or_update_node (slam_graph.py:29-60)
            if (coord == -1 || coord == id || id < coord) {
                // (id < coord also calls update_world(), which the reference leaves unimplemented: slam_graph.py:72-76)
                coord = id;
                for (int k = 0; k < 16; k++) { L[k] = Ti[k]; Wd[k] = I4[k]; }
                present[id] = 1; reference[id] = coord; weight[id] = 1; updated[id] = 1; visible[id] = 0;
            } else {
                const int ref = ref_min;
                if (ref == coord) {   // the world tag is in view: world = local(world tag) @ T
                    double Wn[16];
                    g_matmul(local + (size_t)ref * 16, T, Wn);
                    for (int k = 0; k < 16; k++) { L[k] = Ti[k]; Wd[k] = Wn[k]; }
                    present[id] = 1; reference[id] = coord; weight[id] = 1; updated[id] = 1; visible[id] = 0;
                } else if (present[id] && reference[id] == coord) {   // keep the world transform, refresh local
                    for (int k = 0; k < 16; k++) L[k] = Ti[k];
                    reference[id] = coord; updated[id] = 0; visible[id] = 0;
                } else if (ref != id && ref >= 0 && ref < a.nid && present[ref]) {   // chain through the lowest visible tag
                    double G[16], Wn[16];
                    g_matmul(local + (size_t)ref * 16, T, G);
                    g_matmul(world + (size_t)ref * 16, G, Wn);
                    const int w = weight[ref] + 1, nr = reference[ref];
                    const unsigned char up = updated[ref];
                    for (int k = 0; k < 16; k++) { L[k] = Ti[k]; Wd[k] = Wn[k]; }
                    present[id] = 1; reference[id] = nr; weight[id] = w; updated[id] = up; visible[id] = 0;
                } else {
                    skipped++;   // "Cannot find world reference"
                }
            }
        }
        // SLAM.my_pose (slam.py:36-63)
        unsigned char ok = 0;
        double out[16];
        for (int k = 0; k < 16; k++) out[k] = 0;
        if (n > 0) {
            for (int t = 0; t < a.nid; t++) visible[t] = 0;
            double count = 0;
            for (int i = 0; i < n; i++) {
                const int id = a.dets[base + i].id;
                if (id < 0 || id >= a.nid || !present[id]) continue;
                visible[id] = 1;
                double Tm[16];
                g_matmul(world + (size_t)id * 16, local + (size_t)id * 16, Tm);
                const double w = (double)weight[id];
                for (int k = 0; k < 16; k++) out[k] += Tm[k] / w;
                count += 1 / w;
            }
            if (count != 0) {
                for (int k = 0; k < 16; k++) out[k] = out[k] / count;
                ok = 1;
                for (int k = 0; k < 16; k++) a.estimated_pose[(size_t)s * 16 + k] = out[k];
            }
        }
        a.valid[(size_t)s * a.F + f] = ok;
        for (int k = 0; k < 16; k++) a.my_pose[((size_t)s * a.F + f) * 16 + k] = out[k];
    }
    a.coordinate_id[s] = coord;
    a.skipped[s] = skipped;
}
