// gate_device.cuh - the cost terms of StrongSORT's association, exact fp32, reference op order.
//   Mahalanobis gate + motion blend   linear_assignment::gate_cost_matrix   reference src/trackers/strongsort.cpp:451-492
//                                      on KalmanFilterXYAH::gating_distance  reference src/motion/kalman_filter.cpp:148-176
//                                      ("maha" there is d^T S^-2 d: chol.solve(d) followed by a squared norm, :171-172)
//   tlwh IoU                           iou_matching::iou                     reference src/trackers/strongsort.cpp:502-536
// A track's projection S = H P H^T + R(h) and its Cholesky factor are computed ONCE per row (GateRow) and shared by
// every measurement of that row; the per-pair work is one 4x4 forward/backward substitution.
#pragma once
#include "kf_device.cuh"

namespace mot {

constexpr float kGatingThreshold = 9.4877f;        // chi2inv95[4] (strongsort.cpp:461)
constexpr float kInftyCost = 1e5f;                 // linear_assignment::INFTY_COST (strongsort.hpp)

struct GateRow {
    Chol4 L;                 // factor of the 4x4 projected covariance (conf = 0: no NSA scaling in the gate)
    float l00p, l10p, l11p;  // factor of its leading 2x2 block (only_position)
    float m[4];              // projected mean = mean[0:4]
    bool ok4, ok2;
};

// rec = XYAH record [mean 8 | cov 8x8 row-major]; BaseKalmanFilter::project with confidence 0 (kalman_filter.cpp:60-75)
__device__ __forceinline__ GateRow gate_prepare(const float* __restrict__ rec) {
    GateRow g;
    const float h = rec[3];
    float S[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) S[a][b] = rec[8 + 8 * a + b];
    const float sp = xmul(kf_wpos(), h);
    const float sa = xmul(1e-1f, 1.0f);
    S[0][0] = xadd(S[0][0], xmul(sp, sp));
    S[1][1] = xadd(S[1][1], xmul(sp, sp));
    S[2][2] = xadd(S[2][2], xmul(sa, sa));
    S[3][3] = xadd(S[3][3], xmul(sp, sp));
    g.ok4 = chol4(S, g.L);
    g.ok2 = false;
    g.l00p = 1.0f; g.l10p = 0.0f; g.l11p = 1.0f;
    if (S[0][0] > 0.0f) {
        g.l00p = xsqrt(S[0][0]);
        g.l10p = xdiv(S[1][0], g.l00p);
        const float x1 = xsub(S[1][1], xmul(g.l10p, g.l10p));
        if (x1 > 0.0f) { g.l11p = xsqrt(x1); g.ok2 = true; }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) g.m[a] = rec[a];
    return g;
}

// gating_distance(mean, cov, z, only_position, "maha") for one measurement z = (cx, cy, a, h)
__device__ __forceinline__ float gate_distance(const GateRow& g, float4 z, bool only_position) {
    float d[4] = {xsub(z.x, g.m[0]), xsub(z.y, g.m[1]), xsub(z.z, g.m[2]), xsub(z.w, g.m[3])};
    if (!only_position) {
        if (g.ok4) chol4_solve(g.L, d);
        float acc = xmul(d[0], d[0]);
        acc = xadd(acc, xmul(d[1], d[1]));
        acc = xadd(acc, xmul(d[2], d[2]));
        return xadd(acc, xmul(d[3], d[3]));
    }
    if (g.ok2) {
        d[0] = xdiv(d[0], g.l00p);
        d[1] = xdiv(xsub(d[1], xmul(g.l10p, d[0])), g.l11p);
        d[1] = xdiv(d[1], g.l11p);
        d[0] = xdiv(xsub(d[0], xmul(g.l10p, d[1])), g.l00p);
    }
    return xadd(xmul(d[0], d[0]), xmul(d[1], d[1]));
}

// one entry of gate_cost_matrix (strongsort.cpp:477-487)
__device__ __forceinline__ float gate_blend(float cost, float gd, float mc_lambda, float gated_cost) {
    if (gd > kGatingThreshold) cost = gated_cost;
    return xadd(xmul(mc_lambda, cost), xmul(xsub(1.0f, mc_lambda), gd));
}

// Detection::to_xyah (strongsort.cpp:33-40) of a tlwh box
__device__ __forceinline__ float4 tlwh2xyah_strong(float4 b) {
    return make_float4(xadd(b.x, xdiv(b.z, 2.0f)), xadd(b.y, xdiv(b.w, 2.0f)), xdiv(b.z, b.w), b.w);
}

// iou_matching::iou (strongsort.cpp:502-536) for one (track tlwh, candidate tlwh) pair
__device__ __forceinline__ float iou_tlwh_pair(float4 b, float4 c) {
    const float bx2 = xadd(b.x, b.z), by2 = xadd(b.y, b.w);
    const float cx2 = xadd(c.x, c.z), cy2 = xadd(c.y, c.w);
    const float w = fmaxf(0.0f, xsub(fminf(bx2, cx2), fmaxf(b.x, c.x)));
    const float h = fmaxf(0.0f, xsub(fminf(by2, cy2), fmaxf(b.y, c.y)));
    const float inter = xmul(w, h);
    const float uni = xsub(xadd(xmul(b.z, b.w), xmul(c.z, c.w)), inter);
    if (!(uni > 1e-6f)) return 0.0f;
    const bool z = (inter == 0.0f);                  // 0 / uni == +0: keep zero dividends off the divider's slow path
    const float q = xdiv(z ? 1.0f : inter, uni);
    return z ? 0.0f : q;
}

}  // namespace mot
