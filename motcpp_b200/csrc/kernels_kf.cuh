// kernels_kf.cuh - standalone batched Kalman kernels behind mot_kf_*().  One 8-lane group per
// track (a warp advances four tracks), records are contiguous [mean | cov] blocks so a warp's
// loads and stores cover whole 32-byte sectors.  HBM-bound: 288 B in + 288 B out per XYAH/XYWH
// track (224 + 224 for XYSR); see DESIGN.md for the roofline arithmetic.
// Reference: BaseKalmanFilter::{initiate,predict,update} src/motion/kalman_filter.cpp:29-112,
// KalmanFilterXYSR src/motion/kalman_filters/xysr_kf.cpp:71-112, KalmanFilterXYWH xywh_kf.hpp:41-135.
#pragma once
#include "kf_device.cuh"
#include "kf_xywh_device.cuh"

namespace mot {

enum : int { kKfXYAH = 0, kKfXYSR = 1, kKfXYWH = 2 };

// ---- predict.  flags (nullable): bit0 = zero the height velocity first (ByteTrack, non-Tracked).
template <int KIND>
__global__ void __launch_bounds__(256) kf_predict_kernel(float* __restrict__ recs, const unsigned char* __restrict__ flags,
                                                         long long n, float q44, float q66) {
    const int lane = lane_id(), g = lane & 7, base = lane & ~7;
    const long long groups = ((long long)gridDim.x * blockDim.x) >> 3;
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const long long rounds = (n + groups - 1) / groups;
    for (long long it = 0; it < rounds; ++it) {
        const long long k = it * groups + gid;
        const bool live = k < n;
        if (KIND == kKfXYSR) {
            float* rec = recs + (live ? k : 0) * kRecFloatsXYSR;
            KfRow7 s;
            kf7_load_row(rec, live ? g : 7, s);
            kf_xysr_predict(s, g, base, q44, q66);
            if (live) kf7_store_row(rec, g, s);
        } else {
            float* rec = recs + (live ? k : 0) * kRecFloats;
            KfRow s;
            if (live) kf_load_row(rec, g, s);
            else { s.m = 0.0f; for (int j = 0; j < 8; ++j) s.p[j] = 0.0f; }
            if (KIND == kKfXYAH) kf_xyah_predict(s, g, base, live && flags && (flags[k] & 1));
            else kf_xywh_predict(s, g, base);
            if (live) kf_store_row(rec, g, s);
        }
    }
}

// ---- update with measurement z[k][4] (and NSA confidence conf[k] for XYAH, nullable => 0).
// fail[k] (nullable) is set to 1 where the reference would have left the Cholesky path.
template <int KIND>
__global__ void __launch_bounds__(256) kf_update_kernel(float* __restrict__ recs, const float* __restrict__ z,
                                                        const float* __restrict__ conf, long long n,
                                                        unsigned char* __restrict__ fail) {
    const int lane = lane_id(), g = lane & 7, base = lane & ~7;
    const long long groups = ((long long)gridDim.x * blockDim.x) >> 3;
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const long long rounds = (n + groups - 1) / groups;
    for (long long it = 0; it < rounds; ++it) {
        const long long k = it * groups + gid;
        const bool live = k < n;
        float zz[4] = {0.0f, 0.0f, 0.0f, 1.0f};
        if (live) {
            const float4 q = *reinterpret_cast<const float4*>(z + 4 * k);
            zz[0] = q.x; zz[1] = q.y; zz[2] = q.z; zz[3] = q.w;
        }
        bool ok = true;
        if (KIND == kKfXYSR) {
            float* rec = recs + (live ? k : 0) * kRecFloatsXYSR;
            KfRow7 s;
            kf7_load_row(rec, live ? g : 7, s);
            if (!live) { s.m = 1.0f; for (int j = 0; j < 7; ++j) s.p[j] = (j == g) ? 1.0f : 0.0f; }
            ok = kf_xysr_update(s, g, base, zz);
            if (live && ok) kf7_store_row(rec, g, s);
        } else {
            float* rec = recs + (live ? k : 0) * kRecFloats;
            KfRow s;
            if (live) kf_load_row(rec, g, s);
            else { s.m = 1.0f; for (int j = 0; j < 8; ++j) s.p[j] = (j == g) ? 1.0f : 0.0f; }
            if (KIND == kKfXYAH) ok = kf_xyah_update(s, g, base, zz, (live && conf) ? conf[k] : 0.0f);
            else ok = kf_xywh_update(s, g, base, zz);
            if (live && ok) kf_store_row(rec, g, s);
        }
        if (live && fail && g == 0) fail[k] = ok ? 0 : 1;
    }
}

// ---- initiate from measurement z[k][4]
template <int KIND>
__global__ void __launch_bounds__(256) kf_initiate_kernel(float* __restrict__ recs, const float* __restrict__ z, long long n) {
    const int g = lane_id() & 7;
    const long long groups = ((long long)gridDim.x * blockDim.x) >> 3;
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    for (long long k = gid; k < n; k += groups) {
        const float4 q = *reinterpret_cast<const float4*>(z + 4 * k);
        const float zz[4] = {q.x, q.y, q.z, q.w};
        if (KIND == kKfXYSR) {
            KfRow7 s;
            kf_xysr_init(s, g, zz);
            kf7_store_row(recs + k * kRecFloatsXYSR, g, s);
        } else {
            KfRow s;
            if (KIND == kKfXYAH) kf_xyah_initiate(s, g, zz);
            else kf_xywh_initiate(s, g, zz);
            kf_store_row(recs + k * kRecFloats, g, s);
        }
    }
}

// ---- gating distances: one thread per (track, measurement) pair, 4x4 solve in registers.
// XYAH "maha" reproduces the reference's d^T S^-2 d (kalman_filter.cpp:166-172); XYWH is the true
// Mahalanobis form with the full 4x4 inverse, whose top-left 2x2 serves only_position (xywh_kf.hpp:160-171).
// out is (n_tracks x n_meas) row-major.
template <int KIND>
__global__ void __launch_bounds__(256) kf_gating_kernel(const float* __restrict__ recs, int n_tracks,
                                                        const float* __restrict__ meas, int n_meas, int only_position,
                                                        int metric, float* __restrict__ out) {
    const long long total = (long long)n_tracks * n_meas;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(idx / n_meas), j = (int)(idx % n_meas);
        const float* rec = recs + (size_t)t * kRecFloats;
        const int dim = only_position ? 2 : 4;
        const float h = rec[3];
        float S[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) S[a][b] = rec[8 + 8 * a + b];
        const float sp = xmul(kf_wpos(), h);
        S[0][0] = xadd(S[0][0], xmul(sp, sp));
        S[1][1] = xadd(S[1][1], xmul(sp, sp));
        if (KIND == kKfXYAH) { const float sa = xmul(1e-1f, 1.0f); S[2][2] = xadd(S[2][2], xmul(sa, sa)); }
        else S[2][2] = xadd(S[2][2], xmul(sp, sp));
        S[3][3] = xadd(S[3][3], xmul(sp, sp));
        float d[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) d[a] = xsub(meas[(size_t)j * 4 + a], rec[a]);
        float res;
        if (KIND == kKfXYAH) {
            bool solved = false;
            if (metric == 0) {
                if (dim == 4) {
                    Chol4 L;
                    if (chol4(S, L)) { chol4_solve(L, d); solved = true; }
                } else {
                    // 2x2 Cholesky of the position block
                    const float x0 = S[0][0];
                    if (x0 > 0.0f) {
                        const float l00 = xsqrt(x0), l10 = xdiv(S[1][0], l00);
                        const float x1 = xsub(S[1][1], xmul(l10, l10));
                        if (x1 > 0.0f) {
                            const float l11 = xsqrt(x1);
                            d[0] = xdiv(d[0], l00);
                            d[1] = xdiv(xsub(d[1], xmul(l10, d[0])), l11);
                            d[1] = xdiv(d[1], l11);
                            d[0] = xdiv(xsub(d[0], xmul(l10, d[1])), l00);
                            solved = true;
                        }
                    }
                }
            }
            (void)solved;     // on a failed factorisation the reference falls back to |d|^2 as well
            float acc = xmul(d[0], d[0]);
            acc = xadd(acc, xmul(d[1], d[1]));
            if (dim == 4) { acc = xadd(acc, xmul(d[2], d[2])); acc = xadd(acc, xmul(d[3], d[3])); }
            res = acc;
        } else {
            float inv[4][4];
            inverse4_lu(S, inv);
            float tt[4];
            for (int c = 0; c < dim; ++c) {
                float acc = xmul(d[0], inv[0][c]);
                for (int a = 1; a < dim; ++a) acc = xadd(acc, xmul(d[a], inv[a][c]));
                tt[c] = acc;
            }
            float acc = xmul(tt[0], d[0]);
            for (int c = 1; c < dim; ++c) acc = xadd(acc, xmul(tt[c], d[c]));
            res = acc;
        }
        out[idx] = res;
    }
}

// KalmanFilterXYSR::apply_affine_correction (reference src/motion/kalman_filters/xysr_kf.cpp:114-141): the camera-motion
// warp x[0:2] = m x[0:2] + t, x[4:6] = m x[4:6], P blocks (0,0), (4,4), (0,4) -> m B m^T, (4,0) = (0,4)^T, for n XYSR
// records [x 7 | P 7x7] in place.  One thread per track (20 of the record's 56 floats change).  aff6 = [m00 m01 m10 m11 t0 t1].
struct Affine6 { float m00, m01, m10, m11, t0, t1; };
__global__ void __launch_bounds__(256) kf_xysr_affine_kernel(float* __restrict__ recs, long long n, Affine6 a) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        float* x = recs + k * kRecFloatsXYSR;
        float* P = x + 7;
        const float c0 = xadd(xmul(a.m00, x[0]), xmul(a.m01, x[1])), c1 = xadd(xmul(a.m10, x[0]), xmul(a.m11, x[1]));
        const float v0 = xadd(xmul(a.m00, x[4]), xmul(a.m01, x[5])), v1 = xadd(xmul(a.m10, x[4]), xmul(a.m11, x[5]));
        x[0] = xadd(c0, a.t0); x[1] = xadd(c1, a.t1);
        x[4] = v0; x[5] = v1;
        auto block = [&](int r0, int q0, float (&o)[4]) {
            const float b00 = P[r0 * 7 + q0], b01 = P[r0 * 7 + q0 + 1], b10 = P[(r0 + 1) * 7 + q0], b11 = P[(r0 + 1) * 7 + q0 + 1];
            const float a00 = xadd(xmul(a.m00, b00), xmul(a.m01, b10)), a01 = xadd(xmul(a.m00, b01), xmul(a.m01, b11));
            const float a10 = xadd(xmul(a.m10, b00), xmul(a.m11, b10)), a11 = xadd(xmul(a.m10, b01), xmul(a.m11, b11));
            o[0] = xadd(xmul(a00, a.m00), xmul(a01, a.m01)); o[1] = xadd(xmul(a00, a.m10), xmul(a01, a.m11));
            o[2] = xadd(xmul(a10, a.m00), xmul(a11, a.m01)); o[3] = xadd(xmul(a10, a.m10), xmul(a11, a.m11));
        };
        float pp[4], vv[4], pv[4];
        block(0, 0, pp);
        block(4, 4, vv);
        block(0, 4, pv);
        P[0] = pp[0]; P[1] = pp[1]; P[7] = pp[2]; P[8] = pp[3];
        P[32] = vv[0]; P[33] = vv[1]; P[39] = vv[2]; P[40] = vv[3];
        P[4] = pv[0]; P[5] = pv[1]; P[11] = pv[2]; P[12] = pv[3];
        P[28] = pv[0]; P[35] = pv[1]; P[29] = pv[2]; P[36] = pv[3];
    }
}

}  // namespace mot
