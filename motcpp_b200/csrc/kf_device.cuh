// kf_device.cuh - Kalman filters with one 8-lane group per track ("row-distributed"):
// lane g of the group (g = lane & 7) keeps row g of the covariance (8 registers) and element g of
// the mean; a warp advances four tracks at once and exchanges rows with shuffles.  State records
// are 72 contiguous floats [mean 8 | cov 8x8 row-major], so each lane's row is two float4 loads
// and a warp reads four whole 288-byte records - every 32-byte sector fully used.
//
// The arithmetic is the reference's, element by element, in the operation order fixed by the
// oracle (oracle/kalman.cpp + oracle/smallmat.hpp): ascending-k sums, one rounding per op, no
// FMA.  Results are bit-identical to the oracle; they agree with the stock Eigen build to fp32
// round-off (Eigen's internal summation order is unspecified).
//   XYAH  reference src/motion/kalman_filter.cpp:29-112, src/motion/kalman_filters/xyah_kf.cpp:14-62
//   XYSR  reference src/motion/kalman_filters/xysr_kf.cpp:10-112
//   XYWH  reference include/motcpp/motion/kalman_filters/xywh_kf.hpp:41-135
// All functions must be called by all 32 lanes of a warp (groups without work pass dummies).
#pragma once
#include "simt.cuh"

namespace mot {

constexpr int kRecFloats = 72;            // mean[8] + cov[64]

struct KfRow {
    float m;        // mean[g]
    float p[8];     // cov[g][0..7]
};

__device__ __forceinline__ float kf_wpos() { return 1.0f / 20.0f; }     // kalman_filter.cpp:13
__device__ __forceinline__ float kf_wvel() { return 1.0f / 160.0f; }    // kalman_filter.cpp:14

__device__ __forceinline__ void kf_load_row(const float* __restrict__ rec, int g, KfRow& s) {
    s.m = rec[g];
    const float4 a = *reinterpret_cast<const float4*>(rec + 8 + 8 * g);
    const float4 b = *reinterpret_cast<const float4*>(rec + 8 + 8 * g + 4);
    s.p[0] = a.x; s.p[1] = a.y; s.p[2] = a.z; s.p[3] = a.w;
    s.p[4] = b.x; s.p[5] = b.y; s.p[6] = b.z; s.p[7] = b.w;
}

__device__ __forceinline__ void kf_store_row(float* __restrict__ rec, int g, const KfRow& s) {
    rec[g] = s.m;
    *reinterpret_cast<float4*>(rec + 8 + 8 * g) = make_float4(s.p[0], s.p[1], s.p[2], s.p[3]);
    *reinterpret_cast<float4*>(rec + 8 + 8 * g + 4) = make_float4(s.p[4], s.p[5], s.p[6], s.p[7]);
}

// ---- constant-velocity motion step shared by XYAH / XYWH:  mean' = F mean, P' = F P F^T + diag(q2)
// (F P)(i,j) = P(i,j) + P(i+4,j) for i < 4; (.. F^T)(i,j) = T(i,j) + T(i,j+4) for j < 4.
__device__ __forceinline__ void kf_cv8_motion(KfRow& s, int g, int base, float q2) {
    const int partner = base + ((g + 4) & 7);
    const float mo = __shfl_sync(kFullMask, s.m, partner);
    if (g < 4) s.m = xadd(s.m, mo);
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float o = __shfl_sync(kFullMask, s.p[j], partner);
        t[j] = (g < 4) ? xadd(s.p[j], o) : s.p[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = xadd(t[j], t[j + 4]);
#pragma unroll
    for (int j = 0; j < 8; ++j) s.p[j] = (j == g) ? xadd(t[j], q2) : t[j];
}

// KalmanFilterXYAH::predict.  zero_vh: STrack::multi_predict zeroes mean[7] when the track is not
// in state Tracked (bytetrack.cpp:108-110).
__device__ __forceinline__ void kf_xyah_predict(KfRow& s, int g, int base, bool zero_vh) {
    if (zero_vh && g == 7) s.m = 0.0f;
    const float h = __shfl_sync(kFullMask, s.m, base + 3);     // mean(3) before the motion step
    const float sp = xmul(kf_wpos(), h), sv = xmul(kf_wvel(), h);
    float q = (g < 4) ? sp : sv;
    if (g == 2) q = 1e-2f;
    if (g == 6) q = 1e-5f;
    kf_cv8_motion(s, g, base, xmul(q, q));
}

__device__ __forceinline__ void kf_xywh_predict(KfRow& s, int g, int base) {
    const float h = __shfl_sync(kFullMask, s.m, base + 3);
    const float q = (g < 4) ? xmul(kf_wpos(), h) : xmul(kf_wvel(), h);
    kf_cv8_motion(s, g, base, xmul(q, q));
}

// 4x4 lower Cholesky factor, column by column (smallmat.hpp cholesky_lower).  S is read from its
// lower triangle.  Returns false when a pivot is not positive.
struct Chol4 {
    float l00, l10, l11, l20, l21, l22, l30, l31, l32, l33;
};

// a / l for l > 0.  The structure of the tracking filters makes most numerators here EXACT zeros
// (x, y, a, h evolve independently, so S and the gain are block-sparse), and a zero dividend sends
// the hardware's IEEE division down its slow path.  0 / l is the signed zero itself, so it is
// returned directly and the divider only ever sees a benign dividend: same bits, no slow path.
__device__ __forceinline__ float xdiv_pos(float a, float l) {
    const bool z = (a == 0.0f);
    const float q = xdiv(z ? 1.0f : a, l);
    return z ? a : q;
}

__device__ __forceinline__ bool chol4(const float (&S)[4][4], Chol4& L) {
    float x = S[0][0];
    bool ok = x > 0.0f;
    L.l00 = xsqrt(x);
    L.l10 = xdiv_pos(S[1][0], L.l00);
    L.l20 = xdiv_pos(S[2][0], L.l00);
    L.l30 = xdiv_pos(S[3][0], L.l00);
    x = xsub(S[1][1], xmul(L.l10, L.l10));
    ok = ok && (x > 0.0f);
    L.l11 = xsqrt(x);
    L.l21 = xdiv_pos(xsub(S[2][1], xmul(L.l20, L.l10)), L.l11);
    L.l31 = xdiv_pos(xsub(S[3][1], xmul(L.l30, L.l10)), L.l11);
    x = xsub(S[2][2], xadd(xmul(L.l20, L.l20), xmul(L.l21, L.l21)));
    ok = ok && (x > 0.0f);
    L.l22 = xsqrt(x);
    L.l32 = xdiv_pos(xsub(S[3][2], xadd(xmul(L.l30, L.l20), xmul(L.l31, L.l21))), L.l22);
    x = xsub(S[3][3], xadd(xadd(xmul(L.l30, L.l30), xmul(L.l31, L.l31)), xmul(L.l32, L.l32)));
    ok = ok && (x > 0.0f);
    L.l33 = xsqrt(x);
    return ok;
}

// The tracking filters keep x, y, a|s|w, h|r statistically independent (F and H never mix them), so the innovation
// covariance S is DIAGONAL for every state these trackers can reach: all off-diagonal entries are exact zeros.
// Then the factor is diag(sqrt(S_ii)) and each solve is two divisions - the very operations the general code
// performs on those entries (its other terms multiply or subtract exact zeros), so the values are identical.
// The test is made warp-uniform by the callers; a non-diagonal S takes the general path.
__device__ __forceinline__ bool sym4_is_diagonal(const float (&S)[4][4]) {
    return S[1][0] == 0.0f && S[2][0] == 0.0f && S[2][1] == 0.0f && S[3][0] == 0.0f && S[3][1] == 0.0f && S[3][2] == 0.0f &&
           S[0][1] == 0.0f && S[0][2] == 0.0f && S[1][2] == 0.0f && S[0][3] == 0.0f && S[1][3] == 0.0f && S[2][3] == 0.0f;
}
__device__ __forceinline__ bool chol4_diag(const float (&S)[4][4], Chol4& L) {
    const bool ok = S[0][0] > 0.0f && S[1][1] > 0.0f && S[2][2] > 0.0f && S[3][3] > 0.0f;
    L.l00 = xsqrt(S[0][0]); L.l11 = xsqrt(S[1][1]); L.l22 = xsqrt(S[2][2]); L.l33 = xsqrt(S[3][3]);
    L.l10 = 0.0f; L.l20 = 0.0f; L.l21 = 0.0f; L.l30 = 0.0f; L.l31 = 0.0f; L.l32 = 0.0f;
    return ok;
}
__device__ __forceinline__ void chol4_solve_diag(const Chol4& L, float (&b)[4]) {
    b[0] = xdiv_pos(xdiv_pos(b[0], L.l00), L.l00);
    b[1] = xdiv_pos(xdiv_pos(b[1], L.l11), L.l11);
    b[2] = xdiv_pos(xdiv_pos(b[2], L.l22), L.l22);
    b[3] = xdiv_pos(xdiv_pos(b[3], L.l33), L.l33);
}

// Solve (L L^T) x = b in place (smallmat.hpp cholesky_solve).
__device__ __forceinline__ void chol4_solve(const Chol4& L, float (&b)[4]) {
    b[0] = xdiv_pos(b[0], L.l00);
    b[1] = xdiv_pos(xsub(b[1], xmul(L.l10, b[0])), L.l11);
    b[2] = xdiv_pos(xsub(b[2], xadd(xmul(L.l20, b[0]), xmul(L.l21, b[1]))), L.l22);
    b[3] = xdiv_pos(xsub(b[3], xadd(xadd(xmul(L.l30, b[0]), xmul(L.l31, b[1])), xmul(L.l32, b[2]))), L.l33);
    b[3] = xdiv_pos(b[3], L.l33);
    b[2] = xdiv_pos(xsub(b[2], xmul(L.l32, b[3])), L.l22);
    b[1] = xdiv_pos(xsub(b[1], xadd(xmul(L.l21, b[2]), xmul(L.l31, b[3]))), L.l11);
    b[0] = xdiv_pos(xsub(b[0], xadd(xadd(xmul(L.l10, b[1]), xmul(L.l20, b[2])), xmul(L.l30, b[3]))), L.l00);
}

// Gather S = P[0:4,0:4] (all 16 entries) into every lane of the group.
__device__ __forceinline__ void kf_gather_S(const KfRow& s, int base, float (&S)[4][4]) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) S[a][b] = __shfl_sync(kFullMask, s.p[b], base + a);
}

// mean += K innov ; P -= (K S) K^T   with K row g in k[], for the 8-state filters.
__device__ __forceinline__ void kf_apply_gain8(KfRow& s, int base, const float (&k)[4], const float (&S)[4][4],
                                               const float (&innov)[4]) {
    float dm = xmul(k[0], innov[0]);
    dm = xadd(dm, xmul(k[1], innov[1]));
    dm = xadd(dm, xmul(k[2], innov[2]));
    dm = xadd(dm, xmul(k[3], innov[3]));
    s.m = xadd(s.m, dm);
    float ks[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        float acc = xmul(k[0], S[0][b]);
        acc = xadd(acc, xmul(k[1], S[1][b]));
        acc = xadd(acc, xmul(k[2], S[2][b]));
        acc = xadd(acc, xmul(k[3], S[3][b]));
        ks[b] = acc;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float kj0 = __shfl_sync(kFullMask, k[0], base + j);
        const float kj1 = __shfl_sync(kFullMask, k[1], base + j);
        const float kj2 = __shfl_sync(kFullMask, k[2], base + j);
        const float kj3 = __shfl_sync(kFullMask, k[3], base + j);
        float acc = xmul(ks[0], kj0);
        acc = xadd(acc, xmul(ks[1], kj1));
        acc = xadd(acc, xmul(ks[2], kj2));
        acc = xadd(acc, xmul(ks[3], kj3));
        s.p[j] = xsub(s.p[j], acc);
    }
}

// BaseKalmanFilter::update for XYAH (kalman_filter.cpp:77-112) with NSA confidence `conf`
// (ByteTrack passes 0).  Returns false (state untouched) where the reference would leave the
// Cholesky path for its pseudo-inverse fallback.
__device__ __forceinline__ bool kf_xyah_update(KfRow& s, int g, int base, const float (&z)[4], float conf) {
    (void)g;
    float S[4][4];
    kf_gather_S(s, base, S);
    const float h = __shfl_sync(kFullMask, s.m, base + 3);
    const float one_minus = xsub(1.0f, conf);
    const float sp = xmul(xmul(kf_wpos(), h), one_minus);
    const float sa = xmul(1e-1f, one_minus);
    S[0][0] = xadd(S[0][0], xmul(sp, sp));
    S[1][1] = xadd(S[1][1], xmul(sp, sp));
    S[2][2] = xadd(S[2][2], xmul(sa, sa));
    S[3][3] = xadd(S[3][3], xmul(sp, sp));
    float innov[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) innov[a] = xsub(z[a], __shfl_sync(kFullMask, s.m, base + a));
    Chol4 L;
    float k[4] = {s.p[0], s.p[1], s.p[2], s.p[3]};       // row g of P H^T
    bool ok;
    if (__all_sync(kFullMask, sym4_is_diagonal(S))) {
        ok = chol4_diag(S, L);
        chol4_solve_diag(L, k);
    } else {
        ok = chol4(S, L);
        chol4_solve(L, k);
    }
    if (!ok) return false;                               // uniform across the group
    kf_apply_gain8(s, base, k, S, innov);
    return true;
}

// ------------------------------------------------------------------ independent-coordinate form (fused trackers)
// F and H never mix x, y, a, h, and initiate() starts from a diagonal covariance, so every state the fused trackers can
// reach has P(i, j) = 0 unless i = j (mod 4): four independent (position, velocity) filters.  The dense expressions of
// kf_xyah_predict / kf_xyah_update applied to such a state only ever add, subtract or multiply EXACT zeros outside
// the four 2 x 2 blocks, so evaluating just the blocks - same operations, same order - gives the same values for the
// in-block entries and leaves the other 48 entries of the record at zero.  One thread per (track, coordinate), no
// shuffles in the arithmetic: the dependent chain is ~40 operations instead of ~60 shuffles + the dense update.
//   s.mc = mean[c], s.mv = mean[c+4]; pcc = P[c][c], pcv = P[c][c+4], pvc = P[c+4][c], pvv = P[c+4][c+4]
struct KfBlock {
    float mc, mv, pcc, pcv, pvc, pvv;
};
// Compact record of the independent-coordinate form: 24 floats (96 B, three 32-byte sectors) instead of 72,
//   [ mean 8 | P[c][c] x4 | P[c][c+4] x4 | P[c+4][c] x4 | P[c+4][c+4] x4 ]
// - the 48 entries it leaves out are structural zeros.  A quad's loads are four consecutive floats per array.
constexpr int kRecFloatsCompact = 24;
__device__ __forceinline__ void kfb_load(const float* __restrict__ rec, int c, KfBlock& s) {
    s.mc = rec[c]; s.mv = rec[c + 4];
    s.pcc = rec[8 + c]; s.pcv = rec[12 + c]; s.pvc = rec[16 + c]; s.pvv = rec[20 + c];
}
__device__ __forceinline__ void kfb_store(float* __restrict__ rec, int c, const KfBlock& s) {
    rec[c] = s.mc; rec[c + 4] = s.mv;
    rec[8 + c] = s.pcc; rec[12 + c] = s.pcv; rec[16 + c] = s.pvc; rec[20 + c] = s.pvv;
}
// KalmanFilterXYAH::initiate (kalman_filter.cpp:29-42, xyah_kf.cpp:14-29) for coordinate c: a diagonal covariance
__device__ __forceinline__ void kfb_xyah_initiate(KfBlock& s, int c, float zc, float h) {
    const float sp = (c == 2) ? 1e-2f : xmul(xmul(2.0f, kf_wpos()), h);
    const float sv = (c == 2) ? 1e-5f : xmul(xmul(10.0f, kf_wvel()), h);
    s.mc = zc; s.mv = 0.0f;
    s.pcc = xmul(sp, sp); s.pcv = 0.0f; s.pvc = 0.0f; s.pvv = xmul(sv, sv);
}
// dense 72-float record [mean 8 | cov 8x8 row-major] of a compact one (host side: state dumps)
MOT_HD inline void kfb_expand(const float* compact, float* dense72) {
    for (int k = 0; k < 72; ++k) dense72[k] = 0.0f;
    for (int k = 0; k < 8; ++k) dense72[k] = compact[k];
    for (int c = 0; c < 4; ++c) {
        dense72[8 + 9 * c] = compact[8 + c];
        dense72[8 + 9 * c + 4] = compact[12 + c];
        dense72[8 + 8 * (c + 4) + c] = compact[16 + c];
        dense72[8 + 9 * (c + 4)] = compact[20 + c];
    }
}
// kf_xyah_predict restricted to block c.  h = mean(3) BEFORE the motion step; zero_vh as in kf_xyah_predict.
__device__ __forceinline__ void kfb_xyah_predict(KfBlock& s, int c, float h, bool zero_vh) {
    if (zero_vh && c == 3) s.mv = 0.0f;
    const float sp = xmul(kf_wpos(), h), sv = xmul(kf_wvel(), h);
    const float qp = (c == 2) ? 1e-2f : sp, qv = (c == 2) ? 1e-5f : sv;
    s.mc = xadd(s.mc, s.mv);
    const float tcc = xadd(s.pcc, s.pvc), tcv = xadd(s.pcv, s.pvv);      // row c of F P
    const float ncc = xadd(xadd(tcc, tcv), xmul(qp, qp));
    const float nvc = xadd(s.pvc, s.pvv);
    const float nvv = xadd(s.pvv, xmul(qv, qv));
    s.pcc = ncc; s.pcv = tcv; s.pvc = nvc; s.pvv = nvv;
}
// kf_xyah_update restricted to block c.  h = mean(3) of the CURRENT (predicted) state, z = measurement of coordinate c.
// Returns false when this coordinate's pivot is not positive (the caller combines the four).
__device__ __forceinline__ bool kfb_xyah_update(KfBlock& s, int c, float h, float z, float conf) {
    const float one_minus = xsub(1.0f, conf);
    const float r = (c == 2) ? xmul(1e-1f, one_minus) : xmul(xmul(kf_wpos(), h), one_minus);
    const float scc = xadd(s.pcc, xmul(r, r));
    if (!(scc > 0.0f)) return false;
    const float innov = xsub(z, s.mc);
    const float l = xsqrt(scc);
    const float kc = xdiv_pos(xdiv_pos(s.pcc, l), l), kv = xdiv_pos(xdiv_pos(s.pvc, l), l);
    s.mc = xadd(s.mc, xmul(kc, innov));
    s.mv = xadd(s.mv, xmul(kv, innov));
    const float ksc = xmul(kc, scc), ksv = xmul(kv, scc);
    s.pcc = xsub(s.pcc, xmul(ksc, kc));
    s.pcv = xsub(s.pcv, xmul(ksc, kv));
    s.pvc = xsub(s.pvc, xmul(ksv, kc));
    s.pvv = xsub(s.pvv, xmul(ksv, kv));
    return true;
}

// KalmanFilterXYAH::initiate for row g (kalman_filter.cpp:29-42, xyah_kf.cpp:14-29)
__device__ __forceinline__ void kf_xyah_initiate(KfRow& s, int g, const float (&z)[4]) {
    const float h = z[3];
    float sd = (g < 4) ? xmul(xmul(2.0f, kf_wpos()), h) : xmul(xmul(10.0f, kf_wvel()), h);
    if (g == 2) sd = 1e-2f;
    if (g == 6) sd = 1e-5f;
    s.m = (g < 4) ? z[g & 3] : 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s.p[j] = (j == g) ? xmul(sd, sd) : 0.0f;
}

__device__ __forceinline__ void kf_xywh_initiate(KfRow& s, int g, const float (&z)[4]) {
    const float h = z[3];
    const float sd = (g < 4) ? xmul(xmul(2.0f, kf_wpos()), h) : xmul(xmul(10.0f, kf_wvel()), h);
    s.m = (g < 4) ? z[g & 3] : 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s.p[j] = (j == g) ? xmul(sd, sd) : 0.0f;
}

// ------------------------------------------------------------------ XYSR (7 states; lane 7 idles)
// records are 56 floats [x 7 | P 7x7 row-major]
constexpr int kRecFloatsXYSR = 56;

struct KfRow7 {
    float m;
    float p[7];
};

__device__ __forceinline__ void kf7_load_row(const float* __restrict__ rec, int g, KfRow7& s) {
    if (g < 7) {
        s.m = rec[g];
#pragma unroll
        for (int j = 0; j < 7; ++j) s.p[j] = rec[7 + 7 * g + j];
    } else {
        s.m = 0.0f;
#pragma unroll
        for (int j = 0; j < 7; ++j) s.p[j] = 0.0f;
    }
}

__device__ __forceinline__ void kf7_store_row(float* __restrict__ rec, int g, const KfRow7& s) {
    if (g < 7) {
        rec[g] = s.m;
#pragma unroll
        for (int j = 0; j < 7; ++j) rec[7 + 7 * g + j] = s.p[j];
    }
}

// KalmanFilterXYSR ctor state for a new track (xysr_kf.cpp:49-55, sort.cpp:21-41)
__device__ __forceinline__ void kf_xysr_init(KfRow7& s, int g, const float (&z)[4]) {
    s.m = (g < 4) ? z[g & 3] : 0.0f;
#pragma unroll
    for (int j = 0; j < 7; ++j) s.p[j] = (j == g) ? ((g < 4) ? 10.0f : 1000.0f) : 0.0f;
}

// predict (xysr_kf.cpp:71-77): q44 = q55 = fl(0.01f * Q_xy_scaling), q66 = fl(0.0001f * Q_s_scaling)
__device__ __forceinline__ void kf_xysr_predict(KfRow7& s, int g, int base, float q44, float q66) {
    const int partner = base + ((g + 4) & 7);
    const float mo = __shfl_sync(kFullMask, s.m, partner);
    if (g < 3) s.m = xadd(s.m, mo);
    float t[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const float o = __shfl_sync(kFullMask, s.p[j], partner);
        t[j] = (g < 3) ? xadd(s.p[j], o) : s.p[j];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) t[j] = xadd(t[j], t[j + 4]);
    const float q = (g < 4) ? 1.0f : ((g < 6) ? q44 : q66);
#pragma unroll
    for (int j = 0; j < 7; ++j) s.p[j] = (j == g) ? xadd(t[j], q) : t[j];
}

// update (xysr_kf.cpp:79-112), Joseph form.  Returns false where the reference would fall back to
// its pseudo-inverse (state untouched).
__device__ __forceinline__ bool kf_xysr_update(KfRow7& s, int g, int base, const float (&z)[4]) {
    float S[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) S[a][b] = __shfl_sync(kFullMask, s.p[b], base + a);
    S[0][0] = xadd(S[0][0], 1.0f);
    S[1][1] = xadd(S[1][1], 1.0f);
    S[2][2] = xadd(S[2][2], 10.0f);
    S[3][3] = xadd(S[3][3], 10.0f);
    float y[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) y[a] = xsub(z[a], __shfl_sync(kFullMask, s.m, base + a));
    Chol4 L;
    bool ok;
    float sinv[4][4];                                    // chol.solve(Identity), column by column
    if (__all_sync(kFullMask, sym4_is_diagonal(S))) {
        ok = chol4_diag(S, L);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) sinv[a][c] = 0.0f;
        sinv[0][0] = xdiv(xdiv(1.0f, L.l00), L.l00); sinv[1][1] = xdiv(xdiv(1.0f, L.l11), L.l11);
        sinv[2][2] = xdiv(xdiv(1.0f, L.l22), L.l22); sinv[3][3] = xdiv(xdiv(1.0f, L.l33), L.l33);
    } else {
        ok = chol4(S, L);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float b[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            b[c] = 1.0f;
            chol4_solve(L, b);
#pragma unroll
            for (int a = 0; a < 4; ++a) sinv[a][c] = b[a];
        }
    }
    if (!ok) return false;
    float k[4];                                          // K(g,:) = (P H^T)(g,:) * Sinv
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        float acc = xmul(s.p[0], sinv[0][a]);
        acc = xadd(acc, xmul(s.p[1], sinv[1][a]));
        acc = xadd(acc, xmul(s.p[2], sinv[2][a]));
        acc = xadd(acc, xmul(s.p[3], sinv[3][a]));
        k[a] = acc;
    }
    float dm = xmul(k[0], y[0]);
    dm = xadd(dm, xmul(k[1], y[1]));
    dm = xadd(dm, xmul(k[2], y[2]));
    dm = xadd(dm, xmul(k[3], y[3]));
    const float new_m = xadd(s.m, dm);
    // I - K H, row g: columns 0..3 are delta - K(g,c); columns 4..6 are delta
    float ikh[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) ikh[c] = xsub((c == g) ? 1.0f : 0.0f, k[c]);
    // A = (I - K H) P : A(g,j) = sum_{k<4} ikh[k] P(k,j)  (+ P(g,j) when g >= 4, added after them)
    float A[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        float acc = xmul(ikh[0], __shfl_sync(kFullMask, s.p[j], base + 0));
        acc = xadd(acc, xmul(ikh[1], __shfl_sync(kFullMask, s.p[j], base + 1)));
        acc = xadd(acc, xmul(ikh[2], __shfl_sync(kFullMask, s.p[j], base + 2)));
        acc = xadd(acc, xmul(ikh[3], __shfl_sync(kFullMask, s.p[j], base + 3)));
        if (g >= 4) acc = xadd(acc, s.p[j]);
        A[j] = acc;
    }
    // B = A (I - K H)^T ; C = (K R) K^T ; P' = B + C
    const float kr[4] = {xmul(k[0], 1.0f), xmul(k[1], 1.0f), xmul(k[2], 10.0f), xmul(k[3], 10.0f)};
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const float kj0 = __shfl_sync(kFullMask, k[0], base + j);
        const float kj1 = __shfl_sync(kFullMask, k[1], base + j);
        const float kj2 = __shfl_sync(kFullMask, k[2], base + j);
        const float kj3 = __shfl_sync(kFullMask, k[3], base + j);
        float acc = xmul(A[0], xsub((j == 0) ? 1.0f : 0.0f, kj0));
        acc = xadd(acc, xmul(A[1], xsub((j == 1) ? 1.0f : 0.0f, kj1)));
        acc = xadd(acc, xmul(A[2], xsub((j == 2) ? 1.0f : 0.0f, kj2)));
        acc = xadd(acc, xmul(A[3], xsub((j == 3) ? 1.0f : 0.0f, kj3)));
        if (j >= 4) acc = xadd(acc, A[j]);
        float cc = xmul(kr[0], kj0);
        cc = xadd(cc, xmul(kr[1], kj1));
        cc = xadd(cc, xmul(kr[2], kj2));
        cc = xadd(cc, xmul(kr[3], kj3));
        s.p[j] = xadd(acc, cc);
    }
    s.m = new_m;
    return true;
}

}  // namespace mot
