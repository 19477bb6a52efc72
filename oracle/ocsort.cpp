// TEST INFRASTRUCTURE (oracle): restatement of OC-SORT's per-frame state machine and its
// observation-centric-momentum (OCM) association cost.
//   k_previous_obs            src/trackers/ocsort.cpp:24-51
//   KalmanBoxTracker          src/trackers/ocsort.cpp:53-156, include/motcpp/trackers/ocsort.hpp:31-83
//   speed_direction / convert_x_to_bbox (free functions)  src/trackers/ocsort.cpp:159-186
//   OCSort::update            src/trackers/ocsort.cpp:285-606
//   ocsort_assoc::associate   src/trackers/ocsort.cpp:610-737
// ID counter is per tracker instance (reference: process-global static, ocsort.hpp:33-36).
//
// acosf: the reference calls std::acos(float) (ocsort.cpp:657), i.e. whatever libm the build
// links.  The contract here is the CORRECTLY ROUNDED fp32 arc cosine, obtained by evaluating
// acos in fp64 with a fixed sequence of IEEE operations (+,-,*,/,sqrt; no FMA) and rounding once
// to fp32 - a sequence the CUDA kernels repeat operation for operation (csrc/ocm_device.cuh).
// tests/test_oracle_kats.py measures how often this differs from this box's libm acosf.
#include "oracle.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

namespace {

// fdlibm-style rational approximation R(z) ~ (asin(sqrt z)/sqrt z - 1)/z on [0, 0.25] (public algorithm,
// W. Kahan / Sun fdlibm e_acos.c); evaluated by Horner, ascending rounding per operation.
inline double acos_R(double z) {
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
                 pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
                 pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05;
    const double qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
                 qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    const double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    return p / q;
}

}  // namespace

extern "C" {

float orc_acosf(float xf) {
    const double pi = 3.14159265358979311600e+00, pio2 = 1.57079632679489655800e+00;
    const double x = (double)xf;
    if (!(x == x)) return xf;
    if (x >= 1.0) return 0.0f;
    if (x <= -1.0) return (float)pi;
    double r;
    if (x >= -0.5 && x <= 0.5) {
        r = pio2 - (x + x * acos_R(x * x));
    } else if (x > 0.5) {
        const double z = (1.0 - x) * 0.5;
        const double s = std::sqrt(z);
        r = 2.0 * (s + s * acos_R(z));
    } else {
        const double z = (1.0 + x) * 0.5;
        const double s = std::sqrt(z);
        r = pi - 2.0 * (s + s * acos_R(z));
    }
    return (float)r;
}

// ocsort.cpp:610-700.  dets5 (n_dets x 5) [xyxy, score], trks4 (n_trks x 4) predicted boxes, vel2 (n_trks x 2)
// (dy, dx), prev5 (n_trks x 5) k_previous_obs rows.  out_cost / out_iou are (n_dets x n_trks) row-major.
}  // extern "C"

namespace {
// the AssociationFunction of the tracker (iou.hpp:371-411) for one pair: 0 = iou_batch, 6 = centroid_batch (the variant
// whose expression is defined for every N x M; hmiou / giou / diou / ciou only line up when the second set has one row)
inline float asso_pair(int asso, float norm, const float* d, const float* t) {
    if (asso == 6) {                                                         // iou.hpp:298-330
        const float dx = (d[0] + d[2]) / 2.0f - (t[0] + t[2]) / 2.0f;
        const float dy = (d[1] + d[3]) / 2.0f - (t[1] + t[3]) / 2.0f;
        const float dist = std::sqrt(dx * dx + dy * dy);
        return 1.0f - dist / norm;
    }
    const float area_d = (d[2] - d[0]) * (d[3] - d[1]), area_t = (t[2] - t[0]) * (t[3] - t[1]);
    const float w = std::max(0.0f, std::min(d[2], t[2]) - std::max(d[0], t[0]));
    const float h = std::max(0.0f, std::min(d[3], t[3]) - std::max(d[1], t[1]));
    const float inter = w * h;
    const float uni = area_d + area_t - inter;
    return (uni > 0.0f) ? (inter / uni) : 0.0f;
}
inline float asso_norm(int w, int h) { return static_cast<float>(std::sqrt(w * w + h * h)); }   // iou.hpp:325
void asso_matrix(int asso, float norm, const float* a, int n, const float* b, int m, float* out) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) out[(size_t)i * m + j] = asso_pair(asso, norm, a + 4 * i, b + 4 * j);
}
void ocm_cost_impl(const float* dets5, int n_dets, const float* trks4, const float* vel2, const float* prev5,
                   int n_trks, float inertia, int asso, float asso_nrm, float* out_cost, float* out_iou);
}  // namespace

extern "C" {

void orc_ocm_cost(const float* dets5, int n_dets, const float* trks4, const float* vel2, const float* prev5,
                  int n_trks, float inertia, float* out_cost, float* out_iou) {
    ocm_cost_impl(dets5, n_dets, trks4, vel2, prev5, n_trks, inertia, 0, 1.0f, out_cost, out_iou);
}

}  // extern "C"

namespace {
void ocm_cost_impl(const float* dets5, int n_dets, const float* trks4, const float* vel2, const float* prev5,
                   int n_trks, float inertia, int asso, float asso_nrm, float* out_cost, float* out_iou) {
    const float PI = 3.14159265358979323846f;
    for (int j = 0; j < n_dets; ++j) {
        const float* d = dets5 + 5 * j;
        const float cx1 = (d[0] + d[2]) / 2.0f, cy1 = (d[1] + d[3]) / 2.0f;
        for (int i = 0; i < n_trks; ++i) {
            const float* p = prev5 + 5 * i;
            const float cx2 = (p[0] + p[2]) / 2.0f, cy2 = (p[1] + p[3]) / 2.0f;
            const float dx = cx1 - cx2, dy = cy1 - cy2;                      // :632-633
            const float norm = std::sqrt(dx * dx + dy * dy) + 1e-6f;
            const float Y = dy / norm, X = dx / norm;
            float c = vel2[2 * i + 1] * X + vel2[2 * i + 0] * Y;             // :644 inertia_X*X + inertia_Y*Y
            c = std::min(std::max(c, -1.0f), 1.0f);                          // :645
            const float ang = (PI / 2.0f - std::fabs(orc_acosf(c))) / PI;    // :647-650
            const float valid = (p[4] >= 0.0f) ? 1.0f : 0.0f;                // :653
            float ac = (valid * ang) * inertia;                              // :667
            ac = ac * d[4];                                                  // :669
            // asso_func(detections, trackers): iou_batch iou.hpp:63-100 by default
            const float iou = asso_pair(asso, asso_nrm, d, trks4 + 4 * i);
            if (out_iou) out_iou[(size_t)j * n_trks + i] = iou;
            out_cost[(size_t)j * n_trks + i] = -(iou + ac);                  // :700
        }
    }
}
}  // namespace

namespace {

const float kPlaceholder[5] = {-1.0f, -1.0f, -1.0f, -1.0f, -1.0f};

inline float sum4(const float* b) { return ((b[0] + b[1]) + b[2]) + b[3]; }

// free function speed_direction (ocsort.cpp:159-170): unit (dy, dx) from bbox1's centre to bbox2's
inline void speed_direction(const float* b1, const float* b2, float* out) {
    const float cx1 = (b1[0] + b1[2]) / 2.0f, cy1 = (b1[1] + b1[3]) / 2.0f;
    const float cx2 = (b2[0] + b2[2]) / 2.0f, cy2 = (b2[1] + b2[3]) / 2.0f;
    const float dy = cy2 - cy1, dx = cx2 - cx1;
    const float norm = std::sqrt(dy * dy + dx * dx) + 1e-6f;
    out[0] = dy / norm; out[1] = dx / norm;
}

struct KalmanBoxTracker {
    int id, age = 0, hits = 0, hit_streak = 0, time_since_update = 0;
    float conf;
    int cls, det_ind, delta_t;
    float q_xy, q_s;
    float x[7], P[49];
    float last_observation[5];
    std::map<int, std::array<float, 5>> observations;          // age -> bbox, never pruned (ocsort.hpp:79)
    float velocity[2] = {0.0f, 0.0f};

    // ocsort.cpp:53-87; det7 = [x1,y1,x2,y2,conf,cls,det_ind]
    KalmanBoxTracker(const float* det7, int new_id, int dt, float qxy, float qs)
        : id(new_id), conf(det7[4]), cls((int)det7[5]), det_ind((int)det7[6]), delta_t(dt), q_xy(qxy), q_s(qs) {
        std::memcpy(last_observation, kPlaceholder, sizeof(kPlaceholder));
        float z[4];
        orc_xyxy2xysr(det7, z);
        orc_kf_xysr_init(z, x, P);
    }

    // ocsort.cpp:24-51
    void k_previous_obs(int k, float* out) const {
        if (observations.empty()) { std::memcpy(out, kPlaceholder, sizeof(kPlaceholder)); return; }
        for (int i = 0; i < k; ++i) {
            const int dt = k - i;
            auto it = observations.find(age - dt);
            if (it != observations.end()) { std::memcpy(out, it->second.data(), 5 * sizeof(float)); return; }
        }
        std::memcpy(out, observations.rbegin()->second.data(), 5 * sizeof(float));   // max age key
    }

    // ocsort.cpp:89-132 with a real box
    void update(const float* det7) {
        det_ind = (int)det7[6];
        conf = det7[4];
        cls = (int)det7[5];
        if (sum4(last_observation) >= 0.0f) {
            float prev[5];
            k_previous_obs(delta_t, prev);
            if (sum4(prev) >= 0.0f) speed_direction(prev, det7, velocity);
            else speed_direction(last_observation, det7, velocity);
        }
        std::memcpy(last_observation, det7, 4 * sizeof(float));
        last_observation[4] = conf;
        std::array<float, 5> o;
        std::memcpy(o.data(), last_observation, sizeof(last_observation));
        observations[age] = o;
        time_since_update = 0;
        ++hits;
        ++hit_streak;
        float z[4];
        orc_xyxy2xysr(det7, z);
        orc_kf_xysr_update(x, P, z);
    }
    // update(None): ocsort.cpp:90 and :128-131 (kf.update with a wrong-size z returns at xysr_kf.cpp:80-82)
    void update_none() { det_ind = 0; }

    // ocsort.cpp:134-151
    void predict(float* box) {
        if ((x[6] + x[2]) <= 0.0f) x[6] = 0.0f;
        orc_kf_xysr_predict(x, P, q_xy, q_s);
        ++age;
        if (time_since_update > 0) hit_streak = 0;
        ++time_since_update;
        orc_xysr2xyxy(x, box);
    }
    void state(float* box) const { orc_xysr2xyxy(x, box); }
};

struct Assoc {
    std::vector<std::array<int, 2>> matches;    // (det, trk)
    std::vector<int> unmatched_dets, unmatched_trks;
};

// ocsort.cpp:610-737
Assoc associate(const std::vector<float>& dets5, int n_dets, const std::vector<float>& trks4, int n_trks,
                float iou_threshold, const std::vector<float>& vel2, const std::vector<float>& prev5, float vdc_weight,
                int* used_lap, std::vector<float>* keep_cost, int tie_mode, int asso, float norm) {
    Assoc r;
    *used_lap = 0;
    if (n_trks == 0) {
        for (int i = 0; i < n_dets; ++i) r.unmatched_dets.push_back(i);
        return r;
    }
    if (n_dets > 0) {
        std::vector<float> cost((size_t)n_dets * n_trks), iou((size_t)n_dets * n_trks);
        ocm_cost_impl(dets5.data(), n_dets, trks4.data(), vel2.data(), prev5.data(), n_trks, vdc_weight, asso, norm, cost.data(),
                      iou.data());
        if (keep_cost) *keep_cost = cost;
        int max_row = 0, max_col = 0;                                          // :676-678
        std::vector<int> col_sum(n_trks, 0);
        for (int i = 0; i < n_dets; ++i) {
            int rs = 0;
            for (int j = 0; j < n_trks; ++j)
                if (iou[(size_t)i * n_trks + j] > iou_threshold) { ++rs; ++col_sum[j]; }
            max_row = std::max(max_row, rs);
        }
        for (int j = 0; j < n_trks; ++j) max_col = std::max(max_col, col_sum[j]);
        if (max_row == 1 && max_col == 1) {                                    // :680-689
            for (int i = 0; i < n_dets; ++i)
                for (int j = 0; j < n_trks; ++j)
                    if (iou[(size_t)i * n_trks + j] > iou_threshold) r.matches.push_back({i, j});
        } else {                                                               // :690-712
            *used_lap = 1;
            std::vector<int> r2c(n_dets), c2r(n_trks);
            if (tie_mode == 0 || (tie_mode == 2 && n_dets + n_trks <= 384))
                orc_linear_assignment(cost.data(), n_dets, n_trks, n_trks, -iou_threshold, r2c.data(), c2r.data());
            else   // exact ties (bit-identical twin tracks) resolved towards the higher column, as the CUDA kernel does
                orc_linear_assignment_biased(cost.data(), n_dets, n_trks, n_trks, -iou_threshold, r2c.data(), c2r.data());
            for (int i = 0; i < n_dets; ++i) {
                const int j = r2c[i];
                if (j < 0) continue;
                if (iou[(size_t)i * n_trks + j] >= iou_threshold) r.matches.push_back({i, j});
                else { r.unmatched_dets.push_back(i); r.unmatched_trks.push_back(j); }
            }
        }
    }
    std::vector<char> md(n_dets, 0), mt(n_trks, 0);                            // :715-735
    for (const auto& m : r.matches) { md[m[0]] = 1; mt[m[1]] = 1; }
    for (int i = 0; i < n_dets; ++i) if (!md[i]) r.unmatched_dets.push_back(i);
    for (int i = 0; i < n_trks; ++i) if (!mt[i]) r.unmatched_trks.push_back(i);
    return r;
}

void remove_values(std::vector<int>& v, const std::vector<int>& gone) {
    v.erase(std::remove_if(v.begin(), v.end(), [&](int x) { return std::find(gone.begin(), gone.end(), x) != gone.end(); }),
            v.end());
}

}  // namespace

struct OrcOcSort {
    float det_thresh, iou_threshold, min_conf, inertia, q_xy, q_s;
    int max_age, min_hits, delta_t, use_byte;
    int frame_count = 0;
    int id_counter = 0;
    int last_sizes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int asso = 0;                               // AssociationFunction: 0 iou, 6 centroid (orc_ocsort_set_asso)
    float asso_norm = 1.0f;                     // sqrt(w^2 + h^2) of the frames (centroid)
    int tie_mode = 0;                           // 0 = the reference's LAPJV scan order, 1 = prefer the higher column,
                                                // 2 = the CUDA kernel's policy: 0 while rows + columns <= 384, else 1
    bool capture = false;                       // tests: keep the first-association cost matrix of the last update()
    std::vector<float> last_cost;
    std::vector<KalmanBoxTracker> tracks;
};

extern "C" {

OrcOcSort* orc_ocsort_create(float det_thresh, int max_age, int max_obs, int min_hits, float iou_threshold,
                             float min_conf, int delta_t, float inertia, int use_byte, float q_xy_scaling,
                             float q_s_scaling) {
    (void)max_obs;
    auto* s = new OrcOcSort();
    s->det_thresh = det_thresh; s->max_age = max_age; s->min_hits = min_hits; s->iou_threshold = iou_threshold;
    s->min_conf = min_conf; s->delta_t = delta_t; s->inertia = inertia; s->use_byte = use_byte;
    s->q_xy = q_xy_scaling; s->q_s = q_s_scaling;
    return s;
}
// asso_func constructor argument (ocsort.hpp:93) + the size of the frames update() is given (the reference builds its
// AssociationFunction from img.cols / img.rows in every call, ocsort.cpp:413, :438, :494): 0 "iou", 6 "centroid"
int orc_ocsort_set_asso(OrcOcSort* s, int asso, int frame_w, int frame_h) {
    if (asso != 0 && asso != 6) return -1;
    s->asso = asso;
    s->asso_norm = asso_norm(frame_w, frame_h);
    return 0;
}
void orc_ocsort_destroy(OrcOcSort* s) { delete s; }
void orc_ocsort_reset(OrcOcSort* s) { s->frame_count = 0; s->tracks.clear(); }      // ocsort.cpp:222-227

int orc_ocsort_update(OrcOcSort* s, const float* dets, int n, float* out, int out_cap) {
    ++s->frame_count;                                                              // :293
    std::vector<int> second, remain;                                               // :311-320
    for (int i = 0; i < n; ++i) {
        const float c = dets[6 * i + 4];
        if (c > s->min_conf && c < s->det_thresh) second.push_back(i);
        if (c > s->det_thresh) remain.push_back(i);
    }
    auto det7 = [&](int src, float* row) { std::memcpy(row, dets + 6 * src, 6 * sizeof(float)); row[6] = (float)src; };
    auto spawn = [&](int src) {
        float row[7];
        det7(src, row);
        s->tracks.emplace_back(row, ++s->id_counter, s->delta_t, s->q_xy, s->q_s);
    };
    const int n_high = (int)remain.size(), n_second = (int)second.size();

    // :337-365 predict every track, erase the ones whose box has a NaN
    std::vector<float> trks4;
    {
        std::vector<KalmanBoxTracker> alive;
        for (auto& t : s->tracks) {
            float b[4];
            t.predict(b);
            if (std::isnan(b[0]) || std::isnan(b[1]) || std::isnan(b[2]) || std::isnan(b[3])) continue;
            alive.push_back(t);
            trks4.insert(trks4.end(), b, b + 4);
        }
        s->tracks.swap(alive);
    }
    const int n_trk = (int)s->tracks.size();
    for (int k = 0; k < 8; ++k) s->last_sizes[k] = 0;
    s->last_cost.clear();
    s->last_sizes[0] = n_high; s->last_sizes[1] = n_trk;
    if (n_trk == 0) {                                                              // :367-384
        for (int j = 0; j < n_high; ++j) spawn(remain[j]);
        return 0;
    }
    std::vector<float> vel2((size_t)n_trk * 2), prev5((size_t)n_trk * 5);          // :394-410
    for (int t = 0; t < n_trk; ++t) {
        vel2[2 * t] = s->tracks[t].velocity[0]; vel2[2 * t + 1] = s->tracks[t].velocity[1];
        s->tracks[t].k_previous_obs(s->delta_t, &prev5[5 * t]);
    }
    std::vector<float> dets5((size_t)n_high * 5);
    for (int j = 0; j < n_high; ++j) std::memcpy(&dets5[5 * j], dets + 6 * remain[j], 5 * sizeof(float));
    int used_lap = 0;
    Assoc a = associate(dets5, n_high, trks4, n_trk, s->iou_threshold, vel2, prev5, s->inertia, &used_lap,
                        s->capture ? &s->last_cost : nullptr, s->tie_mode, s->asso, s->asso_norm);   // :413-420
    s->last_sizes[2] = used_lap;
    s->last_sizes[3] = (int)a.matches.size();
    for (const auto& m : a.matches) {                                              // :423-430
        float row[7];
        det7(remain[m[0]], row);
        s->tracks[m[1]].update(row);
    }

    // :433-479 BYTE-style second association on the low-confidence detections
    if (s->use_byte && n_second > 0 && !a.unmatched_trks.empty()) {
        const int nu = (int)a.unmatched_trks.size();
        std::vector<float> ub((size_t)nu * 4), sb((size_t)n_second * 4);
        for (int k = 0; k < nu; ++k) std::memcpy(&ub[4 * k], &trks4[4 * a.unmatched_trks[k]], 4 * sizeof(float));
        for (int j = 0; j < n_second; ++j) std::memcpy(&sb[4 * j], dets + 6 * second[j], 4 * sizeof(float));
        std::vector<float> iou((size_t)n_second * nu);
        asso_matrix(s->asso, s->asso_norm, sb.data(), n_second, ub.data(), nu, iou.data());            // :438-439
        float mx = iou[0];
        for (float v : iou) mx = std::max(mx, v);
        if (mx > s->iou_threshold) {
            std::vector<float> cost(iou.size());
            for (size_t k = 0; k < iou.size(); ++k) cost[k] = -iou[k];
            std::vector<int> r2c(n_second), c2r(nu);
            if (s->tie_mode == 0 || (s->tie_mode == 2 && n_second + nu <= 384))
                orc_linear_assignment(cost.data(), n_second, nu, nu, -s->iou_threshold, r2c.data(), c2r.data());
            else orc_linear_assignment_biased(cost.data(), n_second, nu, nu, -s->iou_threshold, r2c.data(), c2r.data());
            std::vector<int> gone;
            for (int j = 0; j < n_second; ++j) {
                const int u = r2c[j];
                if (u < 0) continue;
                if (iou[(size_t)j * nu + u] < s->iou_threshold) continue;
                const int trk = a.unmatched_trks[u];
                float row[7];
                det7(second[j], row);
                s->tracks[trk].update(row);
                gone.push_back(trk);
            }
            remove_values(a.unmatched_trks, gone);
        }
    }

    // :482-545 re-match leftover detections against the leftover tracks' LAST OBSERVATIONS
    if (!a.unmatched_dets.empty() && !a.unmatched_trks.empty()) {
        const int nd = (int)a.unmatched_dets.size(), nu = (int)a.unmatched_trks.size();
        s->last_sizes[4] = nd; s->last_sizes[5] = nu;
        std::vector<float> db((size_t)nd * 4), tb((size_t)nu * 4);
        for (int k = 0; k < nd; ++k) std::memcpy(&db[4 * k], dets + 6 * remain[a.unmatched_dets[k]], 4 * sizeof(float));
        for (int k = 0; k < nu; ++k)
            std::memcpy(&tb[4 * k], s->tracks[a.unmatched_trks[k]].last_observation, 4 * sizeof(float));
        std::vector<float> iou((size_t)nd * nu);
        asso_matrix(s->asso, s->asso_norm, db.data(), nd, tb.data(), nu, iou.data());                  // :494-495
        float mx = iou[0];
        for (float v : iou) mx = std::max(mx, v);
        if (mx > s->iou_threshold) {
            std::vector<float> cost(iou.size());
            for (size_t k = 0; k < iou.size(); ++k) cost[k] = -iou[k];
            std::vector<int> r2c(nd), c2r(nu);
            if (s->tie_mode == 0 || (s->tie_mode == 2 && nd + nu <= 384))
                orc_linear_assignment(cost.data(), nd, nu, nu, -s->iou_threshold, r2c.data(), c2r.data());
            else orc_linear_assignment_biased(cost.data(), nd, nu, nu, -s->iou_threshold, r2c.data(), c2r.data());
            std::vector<int> gone_d, gone_t;
            for (int k = 0; k < nd; ++k) {
                const int u = r2c[k];
                if (u < 0) continue;
                if (iou[(size_t)k * nu + u] < s->iou_threshold) continue;
                const int det = a.unmatched_dets[k], trk = a.unmatched_trks[u];
                float row[7];
                det7(remain[det], row);
                s->tracks[trk].update(row);
                gone_d.push_back(det);
                gone_t.push_back(trk);
                ++s->last_sizes[6];
            }
            remove_values(a.unmatched_dets, gone_d);
            remove_values(a.unmatched_trks, gone_t);
        }
    }
    for (int trk : a.unmatched_trks) s->tracks[trk].update_none();                 // :548-550
    for (int det : a.unmatched_dets) spawn(remain[det]);                           // :553-561 (duplicates included)
    s->last_sizes[7] = (int)a.unmatched_dets.size();

    // :564-592 output walks the tracks in REVERSE, erasing the dead ones on the way
    int rows = 0;
    std::vector<float> buf;
    for (int k = (int)s->tracks.size() - 1; k >= 0; --k) {
        const KalmanBoxTracker& t = s->tracks[k];
        if (t.time_since_update < 1 && (t.hit_streak >= s->min_hits || s->frame_count <= s->min_hits)) {
            float d[4];
            if (sum4(t.last_observation) < 0.0f) t.state(d);
            else std::memcpy(d, t.last_observation, sizeof(d));
            const float row[8] = {d[0], d[1], d[2], d[3], (float)(t.id + 1), t.conf, (float)t.cls, (float)t.det_ind};
            buf.insert(buf.end(), row, row + 8);
            ++rows;
        }
        if (t.time_since_update > s->max_age) s->tracks.erase(s->tracks.begin() + k);
    }
    if (rows > out_cap) return -rows;
    if (rows) std::memcpy(out, buf.data(), buf.size() * sizeof(float));
    return rows;
}

/* tests: first-association cost matrix (n_high x n_trk, sizes in last_sizes[0..1]) of the last update() */
void orc_ocsort_capture(OrcOcSort* s, int on) { s->capture = on != 0; }
void orc_ocsort_set_tie_mode(OrcOcSort* s, int mode) { s->tie_mode = mode; }
int orc_ocsort_last_cost(const OrcOcSort* s, float* out, int cap) {
    const int n = (int)s->last_cost.size();
    if (out && cap >= n && n) std::memcpy(out, s->last_cost.data(), n * sizeof(float));
    return n;
}
int orc_ocsort_count(const OrcOcSort* s) { return (int)s->tracks.size(); }
void orc_ocsort_last_sizes(const OrcOcSort* s, int* sizes8) { std::memcpy(sizes8, s->last_sizes, sizeof(s->last_sizes)); }

/* rows of [id, age, hits, hit_streak, time_since_update, conf, cls, det_ind, last_obs 5, velocity 2,
 * k_previous_obs(delta_t) 5, x 7, P 49] = 76 floats */
int orc_ocsort_dump(const OrcOcSort* s, float* out, int cap_rows) {
    int k = 0;
    for (const auto& t : s->tracks) {
        if (k >= cap_rows) break;
        float* o = out + (size_t)76 * k++;
        o[0] = (float)t.id; o[1] = (float)t.age; o[2] = (float)t.hits; o[3] = (float)t.hit_streak;
        o[4] = (float)t.time_since_update; o[5] = t.conf; o[6] = (float)t.cls; o[7] = (float)t.det_ind;
        std::memcpy(o + 8, t.last_observation, 5 * sizeof(float));
        o[13] = t.velocity[0]; o[14] = t.velocity[1];
        t.k_previous_obs(s->delta_t, o + 15);
        std::memcpy(o + 20, t.x, 7 * sizeof(float));
        std::memcpy(o + 27, t.P, 49 * sizeof(float));
    }
    return k;
}

}  // extern "C"
