// TEST INFRASTRUCTURE (oracle): restatement of BoostTrack (second half) with its default options.
//   BoostKalmanFilter                       src/trackers/boosttrack.cpp:22-125   (8-state [cx, cy, h, r | velocities])
//   BoostTrack (one track)                  src/trackers/boosttrack.cpp:136-211
//   get_iou_matrix / get_mh_dist_matrix     src/trackers/boosttrack.cpp:297-358
//   dlo_confidence_boost                    src/trackers/boosttrack.cpp:361-426  (basic rule and use_vt; use_sb needs powf and is not restated)
//   BoostTrackTracker::update               src/trackers/boosttrack.cpp:465-699
// Camera-motion compensation (use_ecc) and ReID (with_reid) are image processing outside the hot path: ECC is the
// identity warp, embeddings are off.  Pinned against the reference's OWN boosttrack.cpp compiled in place
// (oracle/_ref/libref_core*.so, tests/test_ref_pin.py).  Dense Eigen expressions are restated as ascending-index sums
// (oracle/smallmat.hpp); the 4 x 4 innovation covariance goes through the partial-pivot LU inverse Eigen applies to a
// dynamic matrix (:68) - it is diagonal here, so the inverse is the reciprocal of the diagonal.
#include "oracle.h"
#include "smallmat.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

using orc::Mat;

inline void bbox_to_z(const float* b, float* z) {                 // :127-134
    const float w = b[2] - b[0], h = b[3] - b[1];
    z[0] = b[0] + w / 2.0f;
    z[1] = b[1] + h / 2.0f;
    z[2] = h;
    z[3] = (h > 1e-6f) ? w / h : 0.0f;
}

struct Track {
    int id, cls, det_ind, tsu = 0, age = 0, streak = 0;
    float conf;
    float x[8];
    Mat P;
    void state(float* b) const {                                  // :107-115
        const float w = x[3] * x[2];
        b[0] = x[0] - w / 2; b[1] = x[1] - x[2] / 2; b[2] = x[0] + w / 2; b[3] = x[1] + x[2] / 2;
    }
};

const Mat& F() {
    static const Mat m = [] { Mat f = Mat::identity(8); for (int i = 0; i < 4; ++i) f(i, i + 4) = 1.0f; return f; }();
    return m;
}
const Mat& H() {
    static const Mat m = [] { Mat h(4, 8); for (int i = 0; i < 4; ++i) h(i, i) = 1.0f; return h; }();
    return m;
}
const Mat& Q() {
    static const Mat m = [] { Mat q = Mat::identity(8); for (int i = 0; i < 4; ++i) { q(i, i) = 1.0f * 10.0f; q(i + 4, i + 4) = 1.0f * 0.01f; } return q; }();
    return m;
}
const Mat& R() {
    static const Mat m = [] { Mat r = Mat::identity(4); r(2, 2) = 10.0f; r(3, 3) = 0.01f; return r; }();
    return m;
}

// inverse of a small square matrix the way Eigen's PartialPivLU does it: LU with row pivoting, then solve for the identity
Mat lu_inverse(const Mat& a) {
    const int n = a.r;
    Mat lu = a;
    std::vector<int> perm(n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int k = 0; k < n; ++k) {
        int piv = k;
        float best = std::fabs(lu(k, k));
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(lu(i, k)) > best) { best = std::fabs(lu(i, k)); piv = i; }
        if (piv != k) {
            for (int j = 0; j < n; ++j) std::swap(lu(k, j), lu(piv, j));
            std::swap(perm[k], perm[piv]);
        }
        for (int i = k + 1; i < n; ++i) {
            lu(i, k) = lu(i, k) / lu(k, k);
            for (int j = k + 1; j < n; ++j) lu(i, j) = lu(i, j) - lu(i, k) * lu(k, j);
        }
    }
    Mat inv(n, n);
    for (int c = 0; c < n; ++c) {
        std::vector<float> y(n);
        for (int i = 0; i < n; ++i) {                             // L y = P e_c
            float acc = (perm[i] == c) ? 1.0f : 0.0f;
            for (int j = 0; j < i; ++j) acc = acc - lu(i, j) * y[j];
            y[i] = acc;
        }
        for (int i = n - 1; i >= 0; --i) {                        // U x = y
            float acc = y[i];
            for (int j = i + 1; j < n; ++j) acc = acc - lu(i, j) * inv(j, c);
            inv(i, c) = acc / lu(i, i);
        }
    }
    return inv;
}

void kf_init(Track& t, const float* z) {                          // :22-54
    for (int i = 0; i < 8; ++i) t.x[i] = i < 4 ? z[i] : 0.0f;
    t.P = Mat::identity(8);
    for (int i = 0; i < 8; ++i) t.P(i, i) = 1.0f * 10.0f;
    for (int i = 4; i < 8; ++i) t.P(i, i) = t.P(i, i) * 1000.0f;
}
void kf_predict(Track& t) {                                       // :56-59
    Mat xv = Mat::from(t.x, 8, 1);
    orc::mul(F(), xv).to(t.x);
    t.P = orc::add(orc::mul_bt(orc::mul(F(), t.P), F()), Q());
}
void kf_update(Track& t, const float* z) {                        // :61-75
    const Mat xv = Mat::from(t.x, 8, 1);
    const Mat pm = orc::mul(H(), xv);
    const Mat S = orc::add(orc::mul_bt(orc::mul(H(), t.P), H()), R());
    const Mat K = orc::mul(orc::mul_bt(t.P, H()), lu_inverse(S));
    Mat innov(4, 1);
    for (int i = 0; i < 4; ++i) innov(i, 0) = z[i] - pm(i, 0);
    const Mat dx = orc::mul(K, innov);
    for (int i = 0; i < 8; ++i) t.x[i] = t.x[i] + dx(i, 0);
    t.P = orc::sub(t.P, orc::mul_bt(orc::mul(K, S), K));
}

}  // namespace

struct OrcBoostTrack {
    float det_thresh, iou_threshold, aspect_ratio_thresh, lambda_iou, lambda_mhd, lambda_shape, dlo_boost_coef;
    int max_age, min_hits, min_box_area, use_dlo_boost, use_vt;
    int frame_count = 0, next_id = 0;
    int last_sizes[4] = {0, 0, 0, 0};                              // filtered detections, tracks, matches, spawned
    std::vector<Track> tracks;
};

extern "C" {

// get_iou_matrix (:297-329): 1 - IoU with the union > 1e-6 guard; trk4 = get_state() boxes
void orc_boost_iou_dist(const float* dets4, int n, const float* trk4, int m, float* out) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) {
            const float* d = dets4 + 4 * i;
            const float* t = trk4 + 4 * j;
            const float x1 = std::max(d[0], t[0]), y1 = std::max(d[1], t[1]), x2 = std::min(d[2], t[2]), y2 = std::min(d[3], t[3]);
            const float inter = std::max(0.0f, x2 - x1) * std::max(0.0f, y2 - y1);
            const float da = (d[2] - d[0]) * (d[3] - d[1]), ta = (t[2] - t[0]) * (t[3] - t[1]);
            const float uni = da + ta - inter;
            const float iou = (uni > 1e-6f) ? inter / uni : 0.0f;
            out[(size_t)i * m + j] = 1.0f - iou;
        }
}

// get_mh_dist_matrix (:331-358): diagonal Mahalanobis distance between convert_bbox_to_z(det) and the first 4 state
// components; mean4 (m x 4) = x.head(4), var4 (m x 4) = covariance.diagonal().head(4)
void orc_boost_mh_dist(const float* dets4, int n, const float* mean4, const float* var4, int m, float* out) {
    for (int i = 0; i < n; ++i) {
        float z[4];
        bbox_to_z(dets4 + 4 * i, z);
        for (int j = 0; j < m; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < 4; ++k) {
                const float diff = z[k] - mean4[4 * j + k];
                const float inv = 1.0f / var4[4 * j + k];
                const float term = (diff * diff) * inv;
                acc = (k == 0) ? term : acc + term;
            }
            out[(size_t)i * m + j] = acc;
        }
    }
}

// the blend of update() (:571-626): cost = iou_dist - lambda_mhd * mh_sim [- lambda_emb * (emb + 1) / 2]
void orc_boost_cost(const float* iou_dist, const float* mh_dist, const float* emb /* nullable */, int n, int m, float lambda_iou,
                    float lambda_mhd, float lambda_shape, float* out) {
    const float limit = 13.2767f;
    const float lambda_emb = (1.0f + lambda_iou + lambda_shape + lambda_mhd) * 1.5f;
    for (size_t k = 0; k < (size_t)n * m; ++k) {
        float s = mh_dist[k];
        if (s > limit) s = limit;
        s = (limit - s) / limit;
        float c = iou_dist[k] - lambda_mhd * s;
        if (emb) c = c - lambda_emb * ((emb[k] + 1.0f) / 2.0f);
        out[k] = c;
    }
}

OrcBoostTrack* orc_boosttrack_create(float det_thresh, int max_age, int max_obs, int min_hits, float iou_threshold, int min_box_area,
                                     float aspect_ratio_thresh, float lambda_iou, float lambda_mhd, float lambda_shape,
                                     int use_dlo_boost, float dlo_boost_coef, int use_vt) {
    auto* s = new OrcBoostTrack();
    s->det_thresh = det_thresh; s->max_age = max_age; s->min_hits = min_hits; s->iou_threshold = iou_threshold;
    s->min_box_area = min_box_area; s->aspect_ratio_thresh = aspect_ratio_thresh; s->lambda_iou = lambda_iou;
    s->lambda_mhd = lambda_mhd; s->lambda_shape = lambda_shape; s->use_dlo_boost = use_dlo_boost;
    s->dlo_boost_coef = dlo_boost_coef; s->use_vt = use_vt;
    (void)max_obs;                                                 // observation history only feeds plotting
    return s;
}
void orc_boosttrack_destroy(OrcBoostTrack* s) { delete s; }
void orc_boosttrack_reset(OrcBoostTrack* s) { s->tracks.clear(); s->frame_count = 0; s->next_id = 0; }   // :272-277
void orc_boosttrack_last_sizes(const OrcBoostTrack* s, int* out4) { std::memcpy(out4, s->last_sizes, sizeof(s->last_sizes)); }
int orc_boosttrack_count(const OrcBoostTrack* s) { return (int)s->tracks.size(); }

// dets (n x 6); out rows [x1,y1,x2,y2,id,conf,cls,det_ind]
int orc_boosttrack_update(OrcBoostTrack* s, const float* dets, int n, float* out, int out_cap) {
    ++s->frame_count;
    std::memset(s->last_sizes, 0, sizeof(s->last_sizes));
    const int nt = (int)s->tracks.size();
    std::vector<float> trk4((size_t)nt * 4);
    for (int j = 0; j < nt; ++j) {                                 // predict (:497-513, :156-163)
        Track& t = s->tracks[j];
        kf_predict(t);
        ++t.age;
        if (t.tsu > 0) t.streak = 0;
        ++t.tsu;
        t.state(&trk4[4 * j]);
    }
    // confidence boost (:520-526, :361-426)
    std::vector<float> conf(n);
    for (int i = 0; i < n; ++i) conf[i] = dets[6 * i + 4];
    if (s->use_dlo_boost && n > 0 && nt > 0) {
        std::vector<float> d4((size_t)n * 4), S((size_t)n * nt);
        for (int i = 0; i < n; ++i) std::memcpy(&d4[4 * i], dets + 6 * i, 4 * sizeof(float));
        orc_iou_batch(d4.data(), n, trk4.data(), nt, S.data());
        if (!s->use_vt) {
            for (int i = 0; i < n; ++i) {
                float mx = S[(size_t)i * nt];
                for (int j = 1; j < nt; ++j) mx = std::max(mx, S[(size_t)i * nt + j]);
                conf[i] = std::max(conf[i], mx * s->dlo_boost_coef);
            }
        } else {
            for (int i = 0; i < n; ++i) {
                bool boost = false;
                for (int j = 0; j < nt && !boost; ++j) {
                    const float th = std::max(0.95f - static_cast<float>(s->tracks[j].tsu - 1), 0.8f);
                    if (S[(size_t)i * nt + j] > th) boost = true;
                }
                if (boost) conf[i] = std::max(conf[i], s->det_thresh + 1e-5f);
            }
        }
    }
    std::vector<int> keep;                                         // :532-538
    for (int i = 0; i < n; ++i)
        if (conf[i] >= s->det_thresh) keep.push_back(i);
    const int nd = (int)keep.size();
    s->last_sizes[0] = nd; s->last_sizes[1] = nt;
    std::vector<int> r2c(nd, -1), c2r(nt, -1);
    if (nd > 0 && nt > 0) {                                        // :568-633
        std::vector<float> d4((size_t)nd * 4), mean4((size_t)nt * 4), var4((size_t)nt * 4);
        for (int i = 0; i < nd; ++i) std::memcpy(&d4[4 * i], dets + 6 * keep[i], 4 * sizeof(float));
        for (int j = 0; j < nt; ++j)
            for (int k = 0; k < 4; ++k) { mean4[4 * j + k] = s->tracks[j].x[k]; var4[4 * j + k] = s->tracks[j].P(k, k); }
        std::vector<float> iou((size_t)nd * nt), mh((size_t)nd * nt), cost((size_t)nd * nt);
        orc_boost_iou_dist(d4.data(), nd, trk4.data(), nt, iou.data());
        orc_boost_mh_dist(d4.data(), nd, mean4.data(), var4.data(), nt, mh.data());
        orc_boost_cost(iou.data(), mh.data(), nullptr, nd, nt, s->lambda_iou, s->lambda_mhd, s->lambda_shape, cost.data());
        orc_linear_assignment(cost.data(), nd, nt, nt, s->iou_threshold, r2c.data(), c2r.data());
    }
    for (int i = 0; i < nd; ++i) {                                 // matched tracks in detection order (:646-654)
        const int j = r2c[i];
        if (j < 0) continue;
        Track& t = s->tracks[j];
        const float* d = dets + 6 * keep[i];
        t.tsu = 0;
        ++t.streak;
        float z[4];
        bbox_to_z(d, z);
        kf_update(t, z);
        t.conf = conf[keep[i]];
        t.cls = (int)d[5];
        t.det_ind = keep[i];
        ++s->last_sizes[2];
    }
    for (int i = 0; i < nd; ++i) {                                 // new tracks (:657-666)
        if (r2c[i] >= 0) continue;
        const float* d = dets + 6 * keep[i];
        Track t;
        float z[4];
        bbox_to_z(d, z);
        kf_init(t, z);
        t.id = ++s->next_id;
        t.conf = conf[keep[i]];
        t.cls = (int)d[5];
        t.det_ind = keep[i];
        s->tracks.push_back(std::move(t));
        ++s->last_sizes[3];
    }
    // outputs in track order, then filter_outputs (:669-698, :434-463)
    int rows = 0;
    for (const Track& t : s->tracks) {
        if (!(t.tsu < 1 && (t.streak >= s->min_hits || s->frame_count <= s->min_hits))) continue;
        float b[4];
        t.state(b);
        const float w = b[2] - b[0], h = b[3] - b[1];
        const float area = w * h, ar = w / (h + 1e-6f);
        if (!(ar <= s->aspect_ratio_thresh && area > (float)s->min_box_area)) continue;
        if (rows < out_cap) {
            float* o = out + 8 * (size_t)rows;
            o[0] = b[0]; o[1] = b[1]; o[2] = b[2]; o[3] = b[3];
            o[4] = (float)t.id; o[5] = t.conf; o[6] = (float)t.cls; o[7] = (float)t.det_ind;
        }
        ++rows;
    }
    s->tracks.erase(std::remove_if(s->tracks.begin(), s->tracks.end(), [&](const Track& t) { return t.tsu > s->max_age; }),
                    s->tracks.end());
    return rows <= out_cap ? rows : -rows;
}

// rows of [id, age, streak, tsu, conf, cls, det_ind, 0, x 8, P 64] = 80 floats
int orc_boosttrack_dump(const OrcBoostTrack* s, float* out, int cap_rows) {
    int k = 0;
    for (const Track& t : s->tracks) {
        if (k >= cap_rows) break;
        float* o = out + 80 * (size_t)k;
        o[0] = (float)t.id; o[1] = (float)t.age; o[2] = (float)t.streak; o[3] = (float)t.tsu; o[4] = t.conf; o[5] = (float)t.cls;
        o[6] = (float)t.det_ind; o[7] = 0.0f;
        std::memcpy(o + 8, t.x, 8 * sizeof(float));
        std::memcpy(o + 16, t.P.d.data(), 64 * sizeof(float));
        ++k;
    }
    return k;
}

}  // extern "C"
