// kf_xywh_device.cuh - BoT-SORT's KalmanFilterXYWH update, which inverts S with a general
// (LU, partial pivoting) inverse instead of a Cholesky solve
// (reference include/motcpp/motion/kalman_filters/xywh_kf.hpp:103-135, S.inverse() at :125).
// Operation order follows oracle/smallmat.hpp inverse_lu().
#pragma once
#include "kf_device.cuh"

namespace mot {

__device__ __forceinline__ void swap_rows4(float (&a)[4][4], int (&perm)[4], int r0, int r1, bool doit) {
    if (!doit) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float t = a[r0][j]; a[r0][j] = a[r1][j]; a[r1][j] = t; }
    const int t = perm[r0]; perm[r0] = perm[r1]; perm[r1] = t;
}

__device__ __forceinline__ void inverse4_lu(const float (&S)[4][4], float (&inv)[4][4]) {
    float lu[4][4];
    int perm[4] = {0, 1, 2, 3};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) lu[i][j] = S[i][j];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int piv = k;
        float best = fabsf(lu[k][k]);
#pragma unroll
        for (int i = k + 1; i < 4; ++i) {
            const float v = fabsf(lu[i][k]);
            if (v > best) { best = v; piv = i; }
        }
#pragma unroll
        for (int i = k + 1; i < 4; ++i) swap_rows4(lu, perm, k, i, piv == i);
#pragma unroll
        for (int i = k + 1; i < 4; ++i) {
            lu[i][k] = xdiv(lu[i][k], lu[k][k]);
#pragma unroll
            for (int j = k + 1; j < 4; ++j) lu[i][j] = xsub(lu[i][j], xmul(lu[i][k], lu[k][j]));
        }
    }
#pragma unroll
    for (int col = 0; col < 4; ++col) {
        float y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = (perm[i] == col) ? 1.0f : 0.0f;
#pragma unroll
            for (int p = 0; p < i; ++p) v = xsub(v, xmul(lu[i][p], y[p]));
            y[i] = v;
        }
#pragma unroll
        for (int i = 3; i >= 0; --i) {
            float v = y[i];
#pragma unroll
            for (int p = i + 1; p < 4; ++p) v = xsub(v, xmul(lu[i][p], y[p]));
            y[i] = xdiv(v, lu[i][i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) inv[i][col] = y[i];
    }
}

// KalmanFilterXYWH::update, row g of an 8-lane group
__device__ __forceinline__ bool kf_xywh_update(KfRow& s, int g, int base, const float (&z)[4]) {
    (void)g;
    float S[4][4];
    kf_gather_S(s, base, S);
    const float h = __shfl_sync(kFullMask, s.m, base + 3);
    const float sp = xmul(kf_wpos(), h);
    const float r = xmul(sp, sp);
#pragma unroll
    for (int a = 0; a < 4; ++a) S[a][a] = xadd(S[a][a], r);
    float innov[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) innov[a] = xsub(z[a], __shfl_sync(kFullMask, s.m, base + a));
    float inv[4][4];
    if (__all_sync(kFullMask, sym4_is_diagonal(S))) {
        // LU with partial pivoting of a diagonal matrix: no row swaps, zero multipliers, inverse = diag(1 / S_ii)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) inv[a][c] = (a == c) ? xdiv(1.0f, S[a][a]) : 0.0f;
    } else {
        inverse4_lu(S, inv);
    }
    float k[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        float acc = xmul(s.p[0], inv[0][a]);
        acc = xadd(acc, xmul(s.p[1], inv[1][a]));
        acc = xadd(acc, xmul(s.p[2], inv[2][a]));
        acc = xadd(acc, xmul(s.p[3], inv[3][a]));
        k[a] = acc;
    }
    kf_apply_gain8(s, base, k, S, innov);
    return true;
}

}  // namespace mot
