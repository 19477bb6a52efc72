// TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's three Kalman filters.
//   XYAH: src/motion/kalman_filter.cpp:10-112,148-176 + src/motion/kalman_filters/xyah_kf.cpp:14-62
//   XYSR: src/motion/kalman_filters/xysr_kf.cpp:10-112
//   XYWH: include/motcpp/motion/kalman_filters/xywh_kf.hpp:19-177
// Written as the same dense matrix expressions the reference hands to Eigen, evaluated
// left to right with ascending-k accumulation (smallmat.hpp).
#include "oracle.h"
#include "smallmat.hpp"

using orc::Mat;

namespace {

const float kWPos = 1.0f / 20.0f;    // kalman_filter.cpp:13
const float kWVel = 1.0f / 160.0f;   // kalman_filter.cpp:14

Mat motion8() {                      // kalman_filter.cpp:17-20 (dt = 1)
    Mat f = Mat::identity(8);
    for (int i = 0; i < 4; ++i) f(i, 4 + i) = 1.0f;
    return f;
}
Mat observe8() {                     // kalman_filter.cpp:23-26
    Mat h(4, 8);
    for (int i = 0; i < 4; ++i) h(i, i) = 1.0f;
    return h;
}

Mat col_from(const float* v, int n) { return Mat::from(v, n, 1); }

// ---- XYAH noise models (xyah_kf.cpp:14-62) ----
void xyah_init_std(const float* z, float* s) {
    const float h = z[3];
    s[0] = 2.0f * kWPos * h; s[1] = 2.0f * kWPos * h; s[2] = 1e-2f; s[3] = 2.0f * kWPos * h;
    s[4] = 10.0f * kWVel * h; s[5] = 10.0f * kWVel * h; s[6] = 1e-5f; s[7] = 10.0f * kWVel * h;
}
void xyah_process_std(const float* mean, float* s) {
    const float h = mean[3];
    s[0] = kWPos * h; s[1] = kWPos * h; s[2] = 1e-2f; s[3] = kWPos * h;
    s[4] = kWVel * h; s[5] = kWVel * h; s[6] = 1e-5f; s[7] = kWVel * h;
}
void xyah_meas_std(const float* mean, float* s) {
    const float h = mean[3];
    s[0] = kWPos * h; s[1] = kWPos * h; s[2] = 1e-1f; s[3] = kWPos * h;
}

// kalman_filter.cpp:60-75
void xyah_project(const float* mean, const Mat& cov, float conf, Mat& pm, Mat& pc) {
    float s[4];
    xyah_meas_std(mean, s);
    for (int i = 0; i < 4; ++i) s[i] = s[i] * (1.0f - conf);   // NSA scaling, :67
    const Mat h = observe8();
    pm = orc::mul(h, col_from(mean, 8));
    pc = orc::add(orc::mul_bt(orc::mul(h, cov), h), orc::diag_sq(s, 4));
}

}  // namespace

extern "C" {

// kalman_filter.cpp:29-42
void orc_kf_xyah_initiate(const float* z4, float* mean8, float* cov64) {
    for (int i = 0; i < 4; ++i) { mean8[i] = z4[i]; mean8[4 + i] = 0.0f; }
    float s[8];
    xyah_init_std(z4, s);
    orc::diag_sq(s, 8).to(cov64);
}

// kalman_filter.cpp:44-58
void orc_kf_xyah_predict(float* mean8, float* cov64) {
    float s[8];
    xyah_process_std(mean8, s);                        // h = mean(3) BEFORE the motion step
    const Mat f = motion8();
    const Mat q = orc::diag_sq(s, 8);
    const Mat nm = orc::mul(f, col_from(mean8, 8));
    const Mat p = Mat::from(cov64, 8, 8);
    const Mat np = orc::add(orc::mul_bt(orc::mul(f, p), f), q);
    nm.to(mean8);
    np.to(cov64);
}

void orc_kf_xyah_project(const float* mean8, const float* cov64, float conf, float* pm4, float* pc16) {
    Mat pm, pc;
    xyah_project(mean8, Mat::from(cov64, 8, 8), conf, pm, pc);
    pm.to(pm4);
    pc.to(pc16);
}

// kalman_filter.cpp:77-112
int orc_kf_xyah_update(float* mean8, float* cov64, const float* z4, float conf) {
    const Mat p = Mat::from(cov64, 8, 8);
    Mat pm, s;
    xyah_project(mean8, p, conf, pm, s);
    Mat l;
    if (!orc::cholesky_lower(s, l)) return 1;          // reference: pseudo-inverse fallback (:86-94)
    const Mat h = observe8();
    const Mat pht = orc::mul_bt(p, h);                  // (8 x 4) = covariance * H^T
    Mat k(8, 4);
    for (int i = 0; i < 8; ++i) {                      // row-by-row solves, :103-105
        float b[4] = {pht(i, 0), pht(i, 1), pht(i, 2), pht(i, 3)};
        orc::cholesky_solve(l, b);
        for (int a = 0; a < 4; ++a) k(i, a) = b[a];
    }
    Mat innov(4, 1);
    for (int a = 0; a < 4; ++a) innov(a, 0) = z4[a] - pm(a, 0);
    const Mat nm = orc::add(col_from(mean8, 8), orc::mul(k, innov));
    const Mat np = orc::sub(p, orc::mul_bt(orc::mul(k, s), k));
    nm.to(mean8);
    np.to(cov64);
    return 0;
}

// kalman_filter.cpp:148-176.  NOTE the reference's "maha" branch solves S z = d and returns
// |z|^2 = d^T S^-2 d (not d^T S^-1 d); reproduced on purpose.
void orc_kf_xyah_gating(const float* mean8, const float* cov64, const float* meas, int m,
                        int only_position, int metric, float* out) {
    Mat pm, pc;
    xyah_project(mean8, Mat::from(cov64, 8, 8), 0.0f, pm, pc);
    const int dim = only_position ? 2 : 4;
    Mat sub(dim, dim);
    for (int i = 0; i < dim; ++i)
        for (int j = 0; j < dim; ++j) sub(i, j) = pc(i, j);
    Mat l;
    const bool ok = (metric == 0) && orc::cholesky_lower(sub, l);
    for (int r = 0; r < m; ++r) {
        float d[4];
        for (int a = 0; a < dim; ++a) d[a] = meas[r * 4 + a] - pm(a, 0);
        if (ok) orc::cholesky_solve(l, d);              // l is dim x dim
        float acc = d[0] * d[0];
        for (int a = 1; a < dim; ++a) acc = acc + d[a] * d[a];
        out[r] = acc;
    }
}

// ------------------------------------------------------------------ XYSR (xysr_kf.cpp)
void orc_kf_xysr_init(const float* z4, float* x7, float* P49) {      // sort.cpp:21-41 + xysr_kf.cpp:49-55
    for (int i = 0; i < 7; ++i) x7[i] = 0.0f;
    for (int i = 0; i < 4; ++i) x7[i] = z4[i];
    Mat p = Mat::identity(7);
    for (auto& v : p.d) v = v * 10.0f;
    for (int i = 4; i < 7; ++i)
        for (int j = 4; j < 7; ++j) p(i, j) = p(i, j) * 100.0f;
    p.to(P49);
}

namespace {
Mat xysr_F() {                                         // xysr_kf.cpp:33-36
    Mat f = Mat::identity(7);
    f(0, 4) = 1.0f; f(1, 5) = 1.0f; f(2, 6) = 1.0f;
    return f;
}
Mat xysr_H() {
    Mat h(4, 7);
    for (int i = 0; i < 4; ++i) h(i, i) = 1.0f;
    return h;
}
Mat xysr_R() {                                         // xysr_kf.cpp:64-65
    Mat r = Mat::identity(4);
    for (int i = 2; i < 4; ++i)
        for (int j = 2; j < 4; ++j) r(i, j) = r(i, j) * 10.0f;
    return r;
}
}  // namespace

void orc_kf_xysr_predict(float* x7, float* P49, float q_xy_scale, float q_s_scale) {   // :71-77
    Mat q = Mat::identity(7);                          // :58-61
    q(4, 4) = 0.01f; q(5, 5) = 0.01f; q(6, 6) = 0.0001f;
    // OC-SORT: Q(4,4)*=Q_xy_scaling, Q(5,5)*=Q_xy_scaling, Q(6,6)*=Q_s_scaling (ocsort.cpp:77-79)
    q(4, 4) = q(4, 4) * q_xy_scale; q(5, 5) = q(5, 5) * q_xy_scale; q(6, 6) = q(6, 6) * q_s_scale;
    const Mat f = xysr_F();
    const Mat nx = orc::mul(f, col_from(x7, 7));
    const Mat np = orc::add(orc::mul_bt(orc::mul(f, Mat::from(P49, 7, 7)), f), q);
    nx.to(x7);
    np.to(P49);
}

int orc_kf_xysr_update(float* x7, float* P49, const float* z4) {     // :79-112
    const Mat h = xysr_H();
    const Mat r = xysr_R();
    const Mat p = Mat::from(P49, 7, 7);
    const Mat hx = orc::mul(h, col_from(x7, 7));
    Mat y(4, 1);
    for (int a = 0; a < 4; ++a) y(a, 0) = z4[a] - hx(a, 0);
    const Mat s = orc::add(orc::mul_bt(orc::mul(h, p), h), r);
    Mat l;
    if (!orc::cholesky_lower(s, l)) return 1;          // reference: COD pseudo-inverse (:100-104)
    Mat sinv(4, 4);                                    // chol.solve(Identity), column by column
    for (int c = 0; c < 4; ++c) {
        float b[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        b[c] = 1.0f;
        orc::cholesky_solve(l, b);
        for (int a = 0; a < 4; ++a) sinv(a, c) = b[a];
    }
    const Mat k = orc::mul(orc::mul_bt(p, h), sinv);  // (7 x 4)
    const Mat nx = orc::add(col_from(x7, 7), orc::mul(k, y));
    const Mat ikh = orc::sub(Mat::identity(7), orc::mul(k, h));
    const Mat np = orc::add(orc::mul_bt(orc::mul(ikh, p), ikh), orc::mul_bt(orc::mul(k, r), k));   // Joseph form
    nx.to(x7);
    np.to(P49);
    return 0;
}

// ------------------------------------------------------------------ XYWH (xywh_kf.hpp)
void orc_kf_xywh_initiate(const float* z4, float* mean8, float* cov64) {   // :41-63
    for (int i = 0; i < 4; ++i) { mean8[i] = z4[i]; mean8[4 + i] = 0.0f; }
    const float h = z4[3];
    float s[8];
    for (int i = 0; i < 4; ++i) { s[i] = 2.0f * kWPos * h; s[4 + i] = 10.0f * kWVel * h; }
    orc::diag_sq(s, 8).to(cov64);
}

void orc_kf_xywh_predict(float* mean8, float* cov64) {                     // :70-94
    const float h = mean8[3];
    float s[8];
    for (int i = 0; i < 4; ++i) { s[i] = kWPos * h; s[4 + i] = kWVel * h; }
    const Mat f = motion8();
    const Mat nm = orc::mul(f, col_from(mean8, 8));
    const Mat np = orc::add(orc::mul_bt(orc::mul(f, Mat::from(cov64, 8, 8)), f), orc::diag_sq(s, 8));
    nm.to(mean8);
    np.to(cov64);
}

namespace {
void xywh_S(const float* mean8, const Mat& p, Mat& pm, Mat& s) {           // :109-124
    const float h = mean8[3];
    float sd[4];
    for (int i = 0; i < 4; ++i) sd[i] = kWPos * h;
    const Mat hm = observe8();
    pm = orc::mul(hm, col_from(mean8, 8));
    s = orc::add(orc::mul_bt(orc::mul(hm, p), hm), orc::diag_sq(sd, 4));
}
}  // namespace

void orc_kf_xywh_update(float* mean8, float* cov64, const float* z4) {     // :103-135
    const Mat p = Mat::from(cov64, 8, 8);
    Mat pm, s;
    xywh_S(mean8, p, pm, s);
    const Mat k = orc::mul(orc::mul_bt(p, observe8()), orc::inverse_lu(s));   // general inverse, :125
    Mat innov(4, 1);
    for (int a = 0; a < 4; ++a) innov(a, 0) = z4[a] - pm(a, 0);
    const Mat nm = orc::add(col_from(mean8, 8), orc::mul(k, innov));
    const Mat np = orc::sub(p, orc::mul_bt(orc::mul(k, s), k));
    nm.to(mean8);
    np.to(cov64);
}

// :140-177.  only_position uses the top-left 2x2 of the FULL 4x4 inverse (:168-171).
void orc_kf_xywh_gating(const float* mean8, const float* cov64, const float* meas, int m,
                        int only_position, float* out) {
    Mat pm, s;
    xywh_S(mean8, Mat::from(cov64, 8, 8), pm, s);
    const Mat sinv = orc::inverse_lu(s);
    const int dim = only_position ? 2 : 4;
    for (int r = 0; r < m; ++r) {
        float d[4];
        for (int a = 0; a < dim; ++a) d[a] = meas[r * 4 + a] - pm(a, 0);
        // (d^T * Sinv) * d
        float t[4];
        for (int j = 0; j < dim; ++j) {
            float acc = d[0] * sinv(0, j);
            for (int a = 1; a < dim; ++a) acc = acc + d[a] * sinv(a, j);
            t[j] = acc;
        }
        float acc = t[0] * d[0];
        for (int j = 1; j < dim; ++j) acc = acc + t[j] * d[j];
        out[r] = acc;
    }
}

}  // extern "C"
