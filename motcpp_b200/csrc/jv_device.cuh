// jv_device.cuh - the reference's dense LAPJV, step for step, run by ONE warp on the extended (n+m) x (n+m) matrix
//     [ C      L/2 ]
//     [ L/2     0  ]          L = thresh
// (reference include/motcpp/association/lap_solver.hpp:36-231 and :251-332; restated in oracle/lapjv.cpp, which is
// pinned to the real header).  It exists for ONE purpose: OC-SORT frames in which bit-identical "twin" tracks make the
// optimum non-unique.  There the reference's answer is whatever LAPJV's scan order yields over the whole dense matrix,
// non-candidate entries included, and only the same algorithm reproduces it.  The sparse solver (lap_device.cuh)
// handles every other frame.  Scans with first-index tie-breaking are spread over the 32 lanes (lexicographic
// (value, index) reductions give the sequential scan's result exactly); the order-dependent bookkeeping runs on
// lane 0.  fp64 throughout, like the reference.
#pragma once
#include "simt.cuh"

namespace mot {

constexpr double kJvBig = 1000000.0;          // lap_solver.hpp:24

struct JvCost {                               // the extended matrix, never materialised
    const float* dense;                       // (n x m) row-major fp32 costs, leading dimension ld
    int n, m, ld;
    double half;                              // thresh / 2 (lap_solver.hpp:299)
    __device__ __forceinline__ double at(int i, int j) const {
        if (i < n && j < m) return (double)dense[(size_t)i * ld + j];          // matching.cpp:31 cast
        return (i >= n && j >= m) ? 0.0 : half;
    }
};

struct JvWork {                               // N = n + m entries each, shared memory
    double* v;
    double* dist;
    int* x;
    int* y;
    int* fr;                                  // free rows
    int* order;
    int* pred;
    unsigned char* sole;
};

MOT_HD constexpr size_t jv_work_bytes(int n_max) {
    return ((size_t)n_max * (2 * sizeof(double) + 5 * sizeof(int) + 1) + 64 + 15) & ~(size_t)15;
}
__device__ __forceinline__ JvWork jv_carve(unsigned char* p, int n_max) {
    JvWork w;
    w.v = (double*)p;        p += sizeof(double) * (size_t)n_max;
    w.dist = (double*)p;     p += sizeof(double) * (size_t)n_max;
    w.x = (int*)p;           p += sizeof(int) * (size_t)n_max;
    w.y = (int*)p;           p += sizeof(int) * (size_t)n_max;
    w.fr = (int*)p;          p += sizeof(int) * (size_t)n_max;
    w.order = (int*)p;       p += sizeof(int) * (size_t)n_max;
    w.pred = (int*)p;        p += sizeof(int) * (size_t)n_max;
    w.sole = p;
    return w;
}

// lexicographic (value, index) minimum across the warp
__device__ __forceinline__ void warp_lexmin(double& val, int& idx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(kFullMask, val, o);
        const int oi = __shfl_xor_sync(kFullMask, idx, o);
        if (ov < val || (ov == val && oi < idx)) { val = ov; idx = oi; }
    }
}

// lap_solver.hpp:157-211 (find_path_dense + the augmentation), lane 0 only
static __device__ __noinline__ void jv_augment_from(const JvCost& c, int N, const JvWork& w, int start) {
    for (int j = 0; j < N; ++j) { w.order[j] = j; w.pred[j] = start; w.dist[j] = c.at(start, j) - w.v[j]; }
    int lo = 0, hi = 0, settled = 0, sink = -1;
    while (sink < 0) {
        if (lo == hi) {                                           // open the next distance level
            settled = lo;
            hi = lo + 1;
            double level = w.dist[w.order[lo]];
            for (int k = hi; k < N; ++k) {
                const int j = w.order[k];
                const double dj = w.dist[j];
                if (dj <= level) {
                    if (dj < level) { hi = lo; level = dj; }
                    w.order[k] = w.order[hi];
                    w.order[hi++] = j;
                }
            }
            for (int k = lo; k < hi; ++k)
                if (w.y[w.order[k]] < 0) sink = w.order[k];       // last free column of the level
        }
        if (sink < 0) {                                           // relax from the level's columns
            int slo = lo, shi = hi;
            bool hit = false;
            while (slo != shi && !hit) {
                const int jq = w.order[slo++];
                const int i = w.y[jq];
                const double level = w.dist[jq];
                const double base = c.at(i, jq) - w.v[jq] - level;
                for (int k = shi; k < N; ++k) {
                    const int j = w.order[k];
                    const double cand = c.at(i, j) - w.v[j] - base;
                    if (cand < w.dist[j]) {
                        w.dist[j] = cand;
                        w.pred[j] = i;
                        if (cand == level) {
                            if (w.y[j] < 0) { sink = j; hit = true; break; }
                            w.order[k] = w.order[shi];
                            w.order[shi++] = j;
                        }
                    }
                }
            }
            if (!hit) { lo = slo; hi = shi; }                     // on a hit the caller's lo/hi stay put (:152)
        }
    }
    const double level = w.dist[w.order[lo]];
    for (int k = 0; k < settled; ++k) {
        const int j = w.order[k];
        w.v[j] += w.dist[j] - level;
    }
    int j = sink, i;
    do {                                                          // flip the path (:203-208)
        i = w.pred[j];
        w.y[j] = i;
        const int prev = w.x[i];
        w.x[i] = j;
        j = prev;
    } while (i != start);
}

// All 32 lanes of one warp.  On return x[0..N) / y[0..N) hold the square assignment (visible to the warp).
static __device__ __noinline__ void warp_dense_lapjv(const JvCost c, int N, const JvWork w) {
    const int lane = lane_id();
    // ---- column reduction (lap_solver.hpp:36-52): per column the minimum over rows, lowest row on ties
    for (int j = lane; j < N; j += 32) {
        double vj = kJvBig;
        int yj = 0;
        for (int i = 0; i < N; ++i) {
            const double val = c.at(i, j);
            if (val < vj) { vj = val; yj = i; }
        }
        w.v[j] = vj; w.y[j] = yj; w.x[j] = -1; w.sole[j] = 1;
    }
    __syncwarp();
    int n_free = 0;
    if (lane == 0) {
        for (int j = N - 1; j >= 0; --j) {                        // right-to-left claim (:53-62)
            const int i = w.y[j];
            if (w.x[i] < 0) w.x[i] = j;
            else { w.sole[i] = 0; w.y[j] = -1; }
        }
    }
    __syncwarp();
    // ---- free rows and reduction transfer (:63-72): sequential over rows, the minimum over columns in parallel
    for (int i = 0; i < N; ++i) {
        const int xi = w.x[i];
        if (xi < 0) { if (lane == 0) w.fr[n_free] = i; ++n_free; continue; }
        if (!w.sole[i]) continue;
        double second = kJvBig;
        for (int k = lane; k < N; k += 32) {
            if (k == xi) continue;
            const double red = c.at(i, k) - w.v[k];
            if (red < second) second = red;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(kFullMask, second, o); if (t < second) second = t; }
        if (lane == 0) w.v[xi] -= second;
        __syncwarp();
    }
    __syncwarp();
    // ---- augmenting row reduction, at most two sweeps (:74-113, :221-224)
    for (int sweep = 0; sweep < 2 && n_free > 0; ++sweep) {
        const unsigned total = (unsigned)n_free;
        unsigned cur = 0, rounds = 0;
        int kept = 0;
        while (cur < total) {
            ++rounds;
            const int i = w.fr[cur++];
            // (u1, j1): minimum reduced cost, first index; (u2, j2): minimum over the other columns, first index
            double b1 = 1.0e300, b2 = 1.0e300;
            int i1 = 0x7fffffff, i2 = 0x7fffffff;
            for (int j = lane; j < N; j += 32) {
                const double red = c.at(i, j) - w.v[j];
                if (red < b1) { b2 = b1; i2 = i1; b1 = red; i1 = j; }
                else if (red < b2) { b2 = red; i2 = j; }
            }
            double u1 = b1; int j1 = i1;
            warp_lexmin(u1, j1);
            double u2 = (i1 == j1) ? b2 : b1;
            int j2 = (i1 == j1) ? i2 : i1;
            warp_lexmin(u2, j2);
            if (!(u2 < kJvBig)) { u2 = kJvBig; j2 = -1; }          // the scan only accepts a runner-up below LARGE
            __syncwarp();                                          // every lane has read fr[cur] before lane 0 rewrites it
            if (lane == 0) {
                int owner = w.y[j1];
                const double lowered = w.v[j1] - (u2 - u1);
                const bool strictly_lower = lowered < w.v[j1];
                if (rounds < cur * (unsigned)N) {                  // unsigned arithmetic as in :101
                    if (strictly_lower) w.v[j1] = lowered;
                    else if (owner >= 0 && j2 >= 0) { j1 = j2; owner = w.y[j2]; }
                    if (owner >= 0) {
                        if (strictly_lower) w.fr[--cur] = owner;   // re-process the displaced row now
                        else w.fr[kept++] = owner;
                    }
                } else if (owner >= 0) {
                    w.fr[kept++] = owner;
                }
                w.x[i] = j1;
                w.y[j1] = i;
            }
            cur = __shfl_sync(kFullMask, cur, 0);
            kept = __shfl_sync(kFullMask, kept, 0);
            __syncwarp();
        }
        n_free = kept;
    }
    // ---- shortest augmenting paths for what is still free (:195-211)
    if (lane == 0)
        for (int f = 0; f < n_free; ++f) jv_augment_from(c, N, w, w.fr[f]);
    __syncwarp();
}

}  // namespace mot
