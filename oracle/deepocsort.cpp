// TEST INFRASTRUCTURE (oracle): restatement of DeepOC-SORT's per-frame state machine.
//   DeepOCSortKalmanBoxTracker       src/trackers/deepocsort.cpp:50-236, include/motcpp/trackers/deepocsort.hpp:21-93
//   compute_aw_max_metric            src/trackers/deepocsort.cpp:294-345   (orc_aw_max_metric, strongsort_ops.cpp)
//   deepocsort_assoc::associate      src/trackers/deepocsort.cpp:348-504
//   DeepOCSort::update               src/trackers/deepocsort.cpp:589-944
// Camera-motion compensation is off (cmc_off = true: image processing is outside the hot path) and the ReID network
// is replaced by the `embs` argument, exactly as the reference does when embeddings are passed in (:629-633).
// Pinned against the reference's OWN deepocsort.cpp compiled in place (oracle/_ref/libref_core*.so,
// tests/test_ref_pin.py).  Reference behaviours kept on purpose:
//   * every detection the assignment leaves unmatched is listed TWICE (the assignment's unmatched_a at :476-478 and the
//     final sweep at :491-495), so it spawns two bit-identical tracks whenever the assignment branch runs;
//   * detection embeddings enter the GEMM un-normalised (:759), only the tracks' copies are normalised (:75-80, :154-158);
//   * the second-stage embedding cost is computed and never used (:835-846).
// Sums follow the "textbook" order of oracle/smallmat.hpp (ascending index, one rounding per operation).
#include "oracle.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

namespace {

const float kNone5[5] = {-1.0f, -1.0f, -1.0f, -1.0f, -1.0f};

inline float sum4(const float* b) { return ((b[0] + b[1]) + b[2]) + b[3]; }

// speed_direction_impl (deepocsort.cpp:241-252)
inline void speed_direction(const float* b1, const float* b2, float* out) {
    const float cx1 = (b1[0] + b1[2]) / 2.0f, cy1 = (b1[1] + b1[3]) / 2.0f;
    const float cx2 = (b2[0] + b2[2]) / 2.0f, cy2 = (b2[1] + b2[3]) / 2.0f;
    const float dy = cy2 - cy1, dx = cx2 - cx1;
    const float norm = std::sqrt(dy * dy + dx * dx) + 1e-6f;
    out[0] = dy / norm; out[1] = dx / norm;
}

inline float norm_seq(const std::vector<float>& v) {
    if (v.empty()) return 0.0f;
    float acc = v[0] * v[0];
    for (size_t k = 1; k < v.size(); ++k) acc = acc + v[k] * v[k];
    return std::sqrt(acc);
}

struct Track {
    int id, age = 0, hits = 0, hit_streak = 0, time_since_update = 0;
    float conf;
    int cls, det_ind, delta_t;
    float q_xy, q_s;
    float x[7], P[49];
    float last_observation[5];
    std::map<int, std::array<float, 5>> observations;
    float velocity[2] = {0.0f, 0.0f};
    std::vector<float> emb;

    // deepocsort.cpp:50-93; det7 = [x1,y1,x2,y2,conf,cls,det_ind]
    Track(const float* det7, int new_id, const float* e, int dim, int dt, float qxy, float qs)
        : id(new_id), conf(det7[4]), cls((int)det7[5]), det_ind((int)det7[6]), delta_t(dt), q_xy(qxy), q_s(qs) {
        std::memcpy(last_observation, kNone5, sizeof(kNone5));
        if (e && dim > 0) {
            emb.assign(e, e + dim);
            const float n = norm_seq(emb);
            if (n > 1e-6f)
                for (float& v : emb) v = v / n;
        }
        float z[4];
        orc_xyxy2xysr(det7, z);
        orc_kf_xysr_init(z, x, P);
    }
    // k_previous_obs_impl (:26-47)
    void k_previous_obs(int k, float* out) const {
        if (observations.empty()) { std::memcpy(out, kNone5, sizeof(kNone5)); return; }
        for (int i = 0; i < k; ++i) {
            auto it = observations.find(age - (k - i));
            if (it != observations.end()) { std::memcpy(out, it->second.data(), 5 * sizeof(float)); return; }
        }
        std::memcpy(out, observations.rbegin()->second.data(), 5 * sizeof(float));
    }
    // update with a box (:95-141)
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
    void update_none() { det_ind = 0; }          // :96, :136-140 (kf.update of an empty vector returns at once)
    // update_emb (:143-161)
    void update_emb(const float* e, int dim, float alpha) {
        if (dim == 0) return;
        if (emb.empty()) emb.assign(e, e + dim);
        else
            for (int k = 0; k < dim; ++k) emb[k] = alpha * emb[k] + (1.0f - alpha) * e[k];
        const float n = norm_seq(emb);
        if (n > 1e-6f)
            for (float& v : emb) v = v / n;
    }
    // predict (:163-179)
    void predict(float* box) {
        if ((x[6] + x[2]) <= 0.0f) x[6] = 0.0f;
        orc_kf_xysr_predict(x, P, q_xy, q_s);
        ++age;
        if (time_since_update > 0) hit_streak = 0;
        ++time_since_update;
        orc_xysr2xyxy(x, box);
    }
};

struct Assoc {
    std::vector<std::array<int, 2>> matches;
    std::vector<int> unmatched_dets, unmatched_trks;
};

}  // namespace

struct OrcDeepOcSort {
    float det_thresh, iou_threshold, inertia, w_assoc_emb, alpha_fixed_emb, aw_param, q_xy, q_s;
    int max_age, min_hits, delta_t, embedding_off, aw_off;
    int frame_count = 0, id_counter = 0;
    int last_sizes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<Track> tracks;
};

namespace {

// deepocsort_assoc::associate (:348-504); dets5 rows [xyxy, score], trks4 predicted boxes, emb (n_dets x n_trks) or empty
Assoc associate(const OrcDeepOcSort* s, const std::vector<float>& dets5, int n_dets, const std::vector<float>& trks4, int n_trks,
                const std::vector<float>& vel2, const std::vector<float>& prev5, const std::vector<float>& emb, int* used_lap) {
    Assoc r;
    *used_lap = 0;
    if (n_trks == 0) {
        for (int i = 0; i < n_dets; ++i) r.unmatched_dets.push_back(i);
        return r;
    }
    const float PI = 3.14159265358979323846f;
    std::vector<float> iou((size_t)n_dets * n_trks), angle((size_t)n_dets * n_trks), e((size_t)n_dets * n_trks, 0.0f);
    if (n_dets > 0) {
        std::vector<float> d4((size_t)n_dets * 4);
        for (int i = 0; i < n_dets; ++i) std::memcpy(&d4[4 * i], &dets5[5 * i], 4 * sizeof(float));
        orc_iou_batch(d4.data(), n_dets, trks4.data(), n_trks, iou.data());
    }
    for (int j = 0; j < n_dets; ++j) {
        const float* d = &dets5[5 * j];
        const float cx1 = (d[0] + d[2]) / 2.0f, cy1 = (d[1] + d[3]) / 2.0f;
        for (int i = 0; i < n_trks; ++i) {
            const float* p = &prev5[5 * i];
            const float cx2 = (p[0] + p[2]) / 2.0f, cy2 = (p[1] + p[3]) / 2.0f;
            const float dx = cx1 - cx2, dy = cy1 - cy2;
            const float norm = std::sqrt(dx * dx + dy * dy) + 1e-6f;
            const float Y = dy / norm, X = dx / norm;
            float c = vel2[2 * i + 1] * X + vel2[2 * i + 0] * Y;                     // :399
            c = std::min(std::max(c, -1.0f), 1.0f);
            const float ang = (PI / 2.0f - std::fabs(orc_acosf(c))) / PI;            // :402-405
            const float valid = (p[4] >= 0.0f) ? 1.0f : 0.0f;
            angle[(size_t)j * n_trks + i] = ((valid * ang) * s->inertia) * d[4];     // :415-417
        }
    }
    if (!emb.empty()) {                                                              // :420-440
        for (size_t k = 0; k < e.size(); ++k) e[k] = (iou[k] <= 0.0f) ? 0.0f : emb[k];
        if (!s->aw_off) {
            std::vector<float> w(e.size());
            orc_aw_max_metric(e.data(), n_dets, n_trks, n_trks, s->w_assoc_emb, s->aw_param, w.data(), n_trks);
            e.swap(w);
        } else {
            for (float& v : e) v = v * s->w_assoc_emb;
        }
    }
    if (n_dets > 0) {
        // trivial one-to-one case (:443-456)
        int max_row = 0, max_col = 0;
        std::vector<int> colsum(n_trks, 0);
        for (int i = 0; i < n_dets; ++i) {
            int rs = 0;
            for (int j = 0; j < n_trks; ++j)
                if (iou[(size_t)i * n_trks + j] > s->iou_threshold) { ++rs; ++colsum[j]; }
            max_row = std::max(max_row, rs);
        }
        for (int j = 0; j < n_trks; ++j) max_col = std::max(max_col, colsum[j]);
        if (max_row == 1 && max_col == 1) {
            for (int i = 0; i < n_dets; ++i)
                for (int j = 0; j < n_trks; ++j)
                    if (iou[(size_t)i * n_trks + j] > s->iou_threshold) r.matches.push_back({i, j});
        } else {
            *used_lap = 1;
            std::vector<float> cost(iou.size());
            for (size_t k = 0; k < cost.size(); ++k) cost[k] = -((iou[k] + angle[k]) + e[k]);   // :459
            std::vector<int> r2c(n_dets), c2r(n_trks);
            orc_linear_assignment(cost.data(), n_dets, n_trks, n_trks, -s->iou_threshold, r2c.data(), c2r.data());
            for (int i = 0; i < n_dets; ++i) {
                const int j = r2c[i];
                if (j < 0) continue;
                if (iou[(size_t)i * n_trks + j] >= s->iou_threshold) r.matches.push_back({i, j});
                else { r.unmatched_dets.push_back(i); r.unmatched_trks.push_back(j); }
            }
            for (int i = 0; i < n_dets; ++i) if (r2c[i] < 0) r.unmatched_dets.push_back(i);   // :476-478 (then AGAIN below)
            for (int j = 0; j < n_trks; ++j) if (c2r[j] < 0) r.unmatched_trks.push_back(j);
        }
    }
    std::vector<char> md(n_dets, 0), mt(n_trks, 0);                                  // :485-501
    for (const auto& m : r.matches) { md[m[0]] = 1; mt[m[1]] = 1; }
    for (int i = 0; i < n_dets; ++i) if (!md[i]) r.unmatched_dets.push_back(i);
    for (int j = 0; j < n_trks; ++j) if (!mt[j]) r.unmatched_trks.push_back(j);
    return r;
}

void remove_values(std::vector<int>& v, const std::vector<int>& gone) {
    v.erase(std::remove_if(v.begin(), v.end(), [&](int x) { return std::find(gone.begin(), gone.end(), x) != gone.end(); }), v.end());
}

}  // namespace

extern "C" {

OrcDeepOcSort* orc_deepocsort_create(float det_thresh, int max_age, int max_obs, int min_hits, float iou_threshold, int delta_t,
                                     float inertia, float w_association_emb, float alpha_fixed_emb, float aw_param, int embedding_off,
                                     int aw_off, float q_xy_scaling, float q_s_scaling) {
    auto* s = new OrcDeepOcSort();
    s->det_thresh = det_thresh; s->iou_threshold = iou_threshold; s->inertia = inertia; s->w_assoc_emb = w_association_emb;
    s->alpha_fixed_emb = alpha_fixed_emb; s->aw_param = aw_param; s->q_xy = q_xy_scaling; s->q_s = q_s_scaling;
    s->max_age = max_age; s->min_hits = min_hits; s->delta_t = delta_t; s->embedding_off = embedding_off; s->aw_off = aw_off;
    (void)max_obs;                               // BaseTracker's fix-up (src/tracker.cpp:37-39) only feeds plotting history
    return s;
}
void orc_deepocsort_destroy(OrcDeepOcSort* s) { delete s; }
void orc_deepocsort_reset(OrcDeepOcSort* s) { s->frame_count = 0; s->tracks.clear(); }   // ids keep counting (:575-579)
int orc_deepocsort_count(const OrcDeepOcSort* s) { return (int)s->tracks.size(); }
void orc_deepocsort_last_sizes(const OrcDeepOcSort* s, int* out8) { std::memcpy(out8, s->last_sizes, sizeof(s->last_sizes)); }

// dets (n x 6), embs (n x dim) or NULL (then embedding_off must be set); out rows [x1,y1,x2,y2,id,conf,cls,det_ind]
int orc_deepocsort_update(OrcDeepOcSort* s, const float* dets, int n, const float* embs, int dim, float* out, int out_cap) {
    ++s->frame_count;
    std::memset(s->last_sizes, 0, sizeof(s->last_sizes));
    const bool use_emb = !s->embedding_off && embs != nullptr && dim > 0;
    std::vector<int> remain;
    for (int i = 0; i < n; ++i)
        if (dets[6 * i + 4] > s->det_thresh) remain.push_back(i);                   // :611-615
    const int nd = (int)remain.size();
    s->last_sizes[0] = nd;
    std::vector<float> ones(1, 1.0f);
    const int edim = use_emb ? dim : 1;                                              // dets_embs = Ones(n, 1) when off (:624-626)
    auto det_emb = [&](int k) -> const float* { return use_emb ? embs + (size_t)remain[k] * dim : ones.data(); };
    auto det7 = [&](int k, float* row) {
        std::memcpy(row, dets + 6 * remain[k], 6 * sizeof(float));
        row[6] = (float)remain[k];
    };
    auto spawn = [&](int k) {
        float row[7];
        det7(k, row);
        s->tracks.emplace_back(row, ++s->id_counter, det_emb(k), edim, s->delta_t, s->q_xy, s->q_s);
    };
    // dets_alpha (:650-652)
    std::vector<float> alpha(nd);
    for (int k = 0; k < nd; ++k) {
        const float trust = (dets[6 * remain[k] + 4] - s->det_thresh) / (1.0f - s->det_thresh);
        alpha[k] = s->alpha_fixed_emb + (1.0f - s->alpha_fixed_emb) * (1.0f - trust);
    }
    if (s->tracks.empty()) {                                                         // :655-668
        for (int k = 0; k < nd; ++k) spawn(k);
        return 0;
    }
    // predict, drop NaN boxes (:671-695)
    std::vector<float> trks4;
    {
        std::vector<Track> keep;
        keep.reserve(s->tracks.size());
        for (auto& t : s->tracks) {
            float b[4];
            t.predict(b);
            if (std::isnan(b[0]) || std::isnan(b[1]) || std::isnan(b[2]) || std::isnan(b[3])) continue;
            trks4.insert(trks4.end(), b, b + 4);
            keep.push_back(std::move(t));
        }
        s->tracks.swap(keep);
    }
    const int nt = (int)s->tracks.size();
    if (nt == 0) {                                                                   // :697-710
        for (int k = 0; k < nd; ++k) spawn(k);
        return 0;
    }
    std::vector<float> vel2((size_t)nt * 2), prev5((size_t)nt * 5), dets5((size_t)nd * 5);
    for (int t = 0; t < nt; ++t) {
        vel2[2 * t] = s->tracks[t].velocity[0]; vel2[2 * t + 1] = s->tracks[t].velocity[1];
        s->tracks[t].k_previous_obs(s->delta_t, &prev5[5 * t]);
    }
    for (int k = 0; k < nd; ++k) std::memcpy(&dets5[5 * k], dets + 6 * remain[k], 5 * sizeof(float));
    // emb_cost = dets_embs * trk_embs^T (:756-766), ascending-k sums
    std::vector<float> emb;
    if (!s->embedding_off && nd > 0) {
        const bool dims_ok = !s->tracks[0].emb.empty() && (int)s->tracks[0].emb.size() == edim;
        emb.assign((size_t)nd * nt, 0.0f);
        if (dims_ok)
            for (int i = 0; i < nd; ++i) {
                const float* a = det_emb(i);
                for (int j = 0; j < nt; ++j) {
                    const std::vector<float>& b = s->tracks[j].emb;
                    float acc = a[0] * b[0];
                    for (int k = 1; k < edim; ++k) acc = acc + a[k] * b[k];
                    emb[(size_t)i * nt + j] = acc;
                }
            }
    }
    int used_lap = 0;
    Assoc a = associate(s, dets5, nd, trks4, nt, vel2, prev5, emb, &used_lap);
    s->last_sizes[0] = nd; s->last_sizes[1] = nt; s->last_sizes[2] = used_lap; s->last_sizes[3] = (int)a.matches.size();
    for (const auto& m : a.matches) {                                                // :789-799
        float row[7];
        det7(m[0], row);
        s->tracks[m[1]].update(row);
        s->tracks[m[1]].update_emb(det_emb(m[0]), edim, alpha[m[0]]);
    }
    // second stage on last observations (:802-878)
    if (!a.unmatched_dets.empty() && !a.unmatched_trks.empty()) {
        const int n_d = (int)a.unmatched_dets.size(), n_u = (int)a.unmatched_trks.size();
        s->last_sizes[4] = n_d; s->last_sizes[5] = n_u;
        std::vector<float> db((size_t)n_d * 4), tb((size_t)n_u * 4);
        for (int k = 0; k < n_d; ++k) std::memcpy(&db[4 * k], dets + 6 * remain[a.unmatched_dets[k]], 4 * sizeof(float));
        for (int k = 0; k < n_u; ++k) std::memcpy(&tb[4 * k], s->tracks[a.unmatched_trks[k]].last_observation, 4 * sizeof(float));
        std::vector<float> iou((size_t)n_d * n_u);
        orc_iou_batch(db.data(), n_d, tb.data(), n_u, iou.data());
        float mx = iou[0];
        for (float v : iou) mx = std::max(mx, v);
        if (mx > s->iou_threshold) {
            std::vector<float> cost(iou.size());
            for (size_t k = 0; k < iou.size(); ++k) cost[k] = -iou[k];
            std::vector<int> r2c(n_d), c2r(n_u);
            orc_linear_assignment(cost.data(), n_d, n_u, n_u, -s->iou_threshold, r2c.data(), c2r.data());
            std::vector<int> gone_d, gone_t;
            for (int k = 0; k < n_d; ++k) {
                const int u = r2c[k];
                if (u < 0 || iou[(size_t)k * n_u + u] < s->iou_threshold) continue;
                const int det = a.unmatched_dets[k], trk = a.unmatched_trks[u];
                float row[7];
                det7(det, row);
                s->tracks[trk].update(row);
                s->tracks[trk].update_emb(det_emb(det), edim, alpha[det]);
                gone_d.push_back(det);
                gone_t.push_back(trk);
                ++s->last_sizes[6];
            }
            remove_values(a.unmatched_dets, gone_d);
            remove_values(a.unmatched_trks, gone_t);
        }
    }
    for (int trk : a.unmatched_trks) s->tracks[trk].update_none();                   // :881-883
    for (int det : a.unmatched_dets) spawn(det);                                     // :886-896 (duplicates included)
    s->last_sizes[7] = (int)a.unmatched_dets.size();
    // output in reverse, erase the aged-out tracks on the way (:899-925)
    int rows = 0;
    for (int k = (int)s->tracks.size() - 1; k >= 0; --k) {
        const Track& t = s->tracks[k];
        float d[4];
        if (sum4(t.last_observation) < 0.0f) orc_xysr2xyxy(t.x, d);
        else std::memcpy(d, t.last_observation, sizeof(d));
        if (t.time_since_update < 1 && (t.hit_streak >= s->min_hits || s->frame_count <= s->min_hits)) {
            if (rows < out_cap) {
                float* o = out + 8 * (size_t)rows;
                o[0] = d[0]; o[1] = d[1]; o[2] = d[2]; o[3] = d[3];
                o[4] = (float)t.id; o[5] = t.conf; o[6] = (float)t.cls; o[7] = (float)t.det_ind;
            }
            ++rows;
        }
        if (t.time_since_update > s->max_age) s->tracks.erase(s->tracks.begin() + k);
    }
    return rows <= out_cap ? rows : -rows;
}

// rows of [id, age, hits, hit_streak, time_since_update, conf, cls, det_ind, last_obs 5, velocity 2, x 7, P 49] = 71 floats;
// embs (nullable): one row of dim floats per track
int orc_deepocsort_dump(const OrcDeepOcSort* s, float* out, float* embs, int dim, int cap_rows) {
    int k = 0;
    for (const Track& t : s->tracks) {
        if (k >= cap_rows) break;
        float* o = out + 71 * (size_t)k;
        o[0] = (float)t.id; o[1] = (float)t.age; o[2] = (float)t.hits; o[3] = (float)t.hit_streak;
        o[4] = (float)t.time_since_update; o[5] = t.conf; o[6] = (float)t.cls; o[7] = (float)t.det_ind;
        std::memcpy(o + 8, t.last_observation, 5 * sizeof(float));
        o[13] = t.velocity[0]; o[14] = t.velocity[1];
        std::memcpy(o + 15, t.x, 7 * sizeof(float));
        std::memcpy(o + 22, t.P, 49 * sizeof(float));
        if (embs && dim > 0) {
            for (int q = 0; q < dim; ++q) embs[(size_t)k * dim + q] = q < (int)t.emb.size() ? t.emb[q] : 0.0f;
        }
        ++k;
    }
    return k;
}

}  // extern "C"
