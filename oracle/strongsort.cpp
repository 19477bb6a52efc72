// TEST INFRASTRUCTURE (oracle): restatement of StrongSORT's per-frame state machine.
//   strongsort::Detection / Track              src/trackers/strongsort.cpp:23-198
//   NearestNeighborDistanceMetric              src/trackers/strongsort.cpp:200-334
//   min_cost_matching / matching_cascade / gate_cost_matrix / iou_cost   :344-585
//   strongsort::Tracker::{predict,update,match,initiate_track}           :591-774
//   StrongSORT::update (public)                                          :829-987
// Outside the hot path: the ReID network (embeddings are passed in) and ECC camera-motion estimation.  The warp is the
// IDENTITY - what motion::ECC::apply returns on its first call and whenever cv::findTransformECC throws, e.g. on a
// featureless image (src/motion/cmc/ecc.cpp:32-34,80) - but Track::camera_update (:111-132) is still applied, because
// its xyah -> tlbr -> xyah round trip re-rounds the mean every frame.
// Reference behaviours kept on purpose ("empty index list means ALL", :358-365, :433-440, :547-555):
//   q1  no confirmed track  => matching_cascade runs over ALL tracks (every cost 1e5: no match), all of them come back
//       unmatched, and the IoU stage receives every (tentative) track TWICE: unconfirmed ++ unmatched-with-tsu-1.
//       A track whose second copy stays unmatched is mark_missed() right after its update => a matched tentative
//       track is deleted unless BOTH copies found a detection; the second copy's detection is swallowed (no new track).
//   q2  no IoU candidate    => the IoU stage runs over ALL tracks (rows with time_since_update > 1 cost 1e5);
//   q3  every detection matched by appearance => the IoU stage runs over ALL detections, and its leftover list
//       (which then contains appearance-matched detections) spawns new tracks.
//   Tracks start Tentative (the GITHUB_ACTIONS test-mode branch of the ctor, :62-72, is not mirrored).
// Vector sums use the "lanes32" order of oracle/botsort.cpp (Eigen's is unspecified).  Exact cost ties only arise from
// q1's duplicate rows; tie_mode chooses who resolves them (see orc_strongsort_set_tie_mode).
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <set>
#include <unordered_map>
#include <vector>

extern "C" int orc_linear_assignment_rowdup(const float* cost, int n, int m, int ld, float thresh, int n_first,
                                            int* row2col, int* col2row);

namespace {

enum SState { Tentative = 1, Confirmed = 2, Deleted = 3 };
constexpr float kInfty = 1e5f;                    // linear_assignment::INFTY_COST
constexpr int kJvMax = 384;                       // csrc/strongsort_kernel.cuh kJvMax

float l32_dot(const float* x, const float* y, int n) {      // "lanes32" (see oracle/botsort.cpp)
    float part[32];
    for (int l = 0; l < 32; ++l) part[l] = 0.0f;
    for (int k = 0; k < n; ++k) {
        const int l = (k >> 2) & 31;
        part[l] = part[l] + x[k] * y[k];
    }
    for (int o = 16; o > 0; o >>= 1)
        for (int l = 0; l < o; ++l) part[l] = part[l] + part[l + o];
    return part[0];
}
float l32_norm(const float* x, int n) { return std::sqrt(l32_dot(x, x, n)); }
// v / |v| when |v| > 1e-10, else v (strongsort.cpp:318-329)
std::vector<float> unit_or_same(const float* v, int n) {
    std::vector<float> o(v, v + n);
    const float nrm = l32_norm(v, n);
    if (nrm > 1e-10f)
        for (auto& e : o) e = e / nrm;
    return o;
}

struct SDet {
    float tlwh[4];
    float conf;
    int cls, det_ind;
    const float* feat;      // raw, dim floats, or nullptr
    void to_xyah(float* o) const {                             // :33-40
        o[0] = tlwh[0] + tlwh[2] / 2.0f; o[1] = tlwh[1] + tlwh[3] / 2.0f; o[2] = tlwh[2] / tlwh[3]; o[3] = tlwh[3];
    }
};

struct STrack {
    int id;
    float mean[8], cov[64];
    std::vector<float> feat;            // features.back() (the one smoothed feature), empty = none
    float conf;
    int cls, det_ind, hits, age, tsu, state;

    STrack(const SDet& d, int id_, int dim) : id(id_), conf(d.conf), cls(d.cls), det_ind(d.det_ind), hits(1), age(1), tsu(0), state(Tentative) {   // :46-91
        float z[4];
        d.to_xyah(z);
        orc_kf_xyah_initiate(z, mean, cov);
        if (d.feat && dim > 0) {
            const float nrm = l32_norm(d.feat, dim);
            if (nrm > 1e-10f) {
                feat.assign(d.feat, d.feat + dim);
                for (auto& e : feat) e = e / nrm;
            }
        }
    }
    void to_tlwh(float* o) const {                             // :93-99
        o[2] = mean[2] * mean[3]; o[3] = mean[3];
        o[0] = mean[0] - o[2] / 2.0f; o[1] = mean[1] - o[3] / 2.0f;
    }
    void to_tlbr(float* o) const {                             // :101-109
        float t[4];
        to_tlwh(t);
        o[0] = t[0]; o[1] = t[1]; o[2] = t[0] + t[2]; o[3] = t[1] + t[3];
    }
    void camera_update_identity() {                            // :111-132 with warp = [I | 0]
        float b[4];
        to_tlbr(b);
        const float x1 = (1.0f * b[0] + 0.0f * b[1]) + 0.0f * 1.0f, y1 = (0.0f * b[0] + 1.0f * b[1]) + 0.0f * 1.0f;
        const float x2 = (1.0f * b[2] + 0.0f * b[3]) + 0.0f * 1.0f, y2 = (0.0f * b[2] + 1.0f * b[3]) + 0.0f * 1.0f;
        const float w = x2 - x1, h = y2 - y1;
        mean[0] = x1 + w / 2.0f; mean[1] = y1 + h / 2.0f; mean[2] = w / h; mean[3] = h;
    }
    void predict() { orc_kf_xyah_predict(mean, cov); ++age; ++tsu; }          // :139-145
    void update(const SDet& d, int dim, float alpha, int n_init) {                        // :147-187
        float z[4];
        d.to_xyah(z);
        conf = d.conf; cls = d.cls; det_ind = d.det_ind;
        orc_kf_xyah_update(mean, cov, z, conf);
        if (d.feat && dim > 0) {
            const float nrm = l32_norm(d.feat, dim);
            if (!(nrm < 1e-10f)) {
                std::vector<float> fn(d.feat, d.feat + dim);
                for (auto& e : fn) e = e / nrm;
                if (!feat.empty()) {
                    const float beta = 1.0f - alpha;
                    std::vector<float> sm(dim);
                    for (int k = 0; k < dim; ++k) sm[k] = alpha * feat[k] + beta * fn[k];
                    const float sn = l32_norm(sm.data(), dim);
                    if (sn > 1e-10f) {
                        for (auto& e : sm) e = e / sn;
                        feat = sm;
                    }
                } else {
                    feat = fn;
                }
            }
        }
        ++hits;
        tsu = 0;
        if (state == Tentative && hits >= n_init) state = Confirmed;
    }
    void mark_missed(int max_age) {                                           // :189-195
        if (state == Tentative) state = Deleted;
        else if (tsu > max_age) state = Deleted;
    }
};

}  // namespace

struct OrcStrongSort {
    float min_conf, max_cos_dist, max_iou_dist, mc_lambda, ema_alpha;
    int max_age, n_init, nn_budget;
    int tie_mode = 0;
    int next_id = 1;
    std::vector<STrack> tracks;
    std::unordered_map<int, std::deque<std::vector<float>>> samples;
    int last_sizes[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    // min_cost_matching (:344-416) over explicit index lists; cost(rows x cols) supplied by the caller.
    // Returns matches (track, det) in ascending row order, unmatched rows' tracks, unmatched cols' dets.
    void lap_stage(std::vector<float>& cost, const std::vector<int>& track_idx, const std::vector<int>& det_idx, float max_distance,
                   int dup_first, std::vector<std::pair<int, int>>& matches, std::vector<int>& un_t, std::vector<int>& un_d) {
        const int n = (int)track_idx.size(), m = (int)det_idx.size();
        orc_clamp_cost(cost.data(), n, m, m, max_distance);                   // :372-377
        std::vector<int> r2c(n), c2r(m);
        const bool dup = dup_first > 0 && dup_first < n;
        if (dup && (tie_mode == 1 || (tie_mode == 2 && n + m > kJvMax)))
            orc_linear_assignment_rowdup(cost.data(), n, m, m, max_distance, dup_first, r2c.data(), c2r.data());
        else
            orc_linear_assignment(cost.data(), n, m, m, max_distance, r2c.data(), c2r.data());
        std::vector<char> mr(n, 0), mc(m, 0);
        for (int r = 0; r < n; ++r) {
            const int c = r2c[r];
            if (c >= 0 && cost[(size_t)r * m + c] <= max_distance) {          // :389-399
                if (getenv("ORC_SS_DEBUG")) fprintf(stderr, "ORC match row=%d trk=%d det=%d cost=%g thr=%g\n", r, track_idx[r], det_idx[c], cost[(size_t)r * m + c], max_distance);
                matches.push_back({track_idx[r], det_idx[c]});
                mr[r] = 1; mc[c] = 1;
            }
        }
        for (int r = 0; r < n; ++r) if (!mr[r]) un_t.push_back(track_idx[r]);
        for (int c = 0; c < m; ++c) if (!mc[c]) un_d.push_back(det_idx[c]);
    }

    // Tracker::match (:664-766)
    void match(const std::vector<SDet>& dets, int dim, std::vector<std::pair<int, int>>& matches, std::vector<int>& unmatched_tracks,
               std::vector<int>& unmatched_dets) {
        const int nt = (int)tracks.size(), nd = (int)dets.size();
        std::vector<int> confirmed, unconfirmed;
        for (int i = 0; i < nt; ++i) (tracks[i].state == Confirmed ? confirmed : unconfirmed).push_back(i);
        auto all_tracks = [&] { std::vector<int> v(nt); for (int i = 0; i < nt; ++i) v[i] = i; return v; };
        auto all_dets = [&] { std::vector<int> v(nd); for (int i = 0; i < nd; ++i) v[i] = i; return v; };

        // ---- appearance stage: matching_cascade -> min_cost_matching with the gated metric (:419-449, :667-726)
        std::vector<int> ta = confirmed.empty() ? all_tracks() : confirmed;   // "empty means all" (:433-436)
        std::vector<int> da = all_dets();
        std::vector<std::pair<int, int>> matches_a;
        std::vector<int> un_ta, un_da;
        last_sizes[0] = (int)ta.size(); last_sizes[1] = (int)da.size();
        if (ta.empty() || da.empty()) {                                        // :367-369
            un_ta = ta; un_da = da;
        } else {
            const int n = (int)ta.size(), m = nd;
            std::vector<float> cost((size_t)n * m, 1e5f);
            bool any_feat = false;
            for (int j = 0; j < m; ++j) any_feat = any_feat || (dets[j].feat && dim > 0);
            if (any_feat) {                                                    // feat_dim > 0 (:683-697)
                std::vector<std::vector<float>> fb(m);
                for (int j = 0; j < m; ++j) fb[j] = unit_or_same(dets[j].feat, dim);
                std::vector<float> meas((size_t)m * 4), recs((size_t)n * 72);
                for (int j = 0; j < m; ++j) dets[j].to_xyah(&meas[(size_t)j * 4]);
                for (int r = 0; r < n; ++r) {
                    const STrack& t = tracks[ta[r]];
                    std::memcpy(&recs[(size_t)r * 72], t.mean, sizeof(t.mean));
                    std::memcpy(&recs[(size_t)r * 72 + 8], t.cov, sizeof(t.cov));
                    auto it = samples.find(t.id);
                    if (it == samples.end() || it->second.empty()) continue;  // row stays 1e5 (:271)
                    for (int j = 0; j < m; ++j) {
                        float best = 0.0f;
                        bool first = true;
                        for (const auto& s : it->second) {
                            const std::vector<float> sa = unit_or_same(s.data(), dim);
                            const float d = 1.0f - l32_dot(sa.data(), fb[j].data(), dim);   // :333
                            if (first || d < best) { best = d; first = false; }
                        }
                        cost[(size_t)r * m + j] = best;
                    }
                }
                orc_gate_cost_matrix(cost.data(), m, recs.data(), n, meas.data(), m, mc_lambda, kInfty, 0);   // :722-724
            }
            lap_stage(cost, ta, da, max_cos_dist, 0, matches_a, un_ta, un_da);
        }

        // ---- IoU stage on unconfirmed ++ just-missed tracks (:728-763)
        std::vector<int> cand = unconfirmed, un_ta_filtered;
        for (int k : un_ta) (tracks[k].tsu == 1 ? cand : un_ta_filtered).push_back(k);
        // duplicates can only come from q1: un_ta then repeats the unconfirmed tracks
        const int dup_first = (confirmed.empty() && !unconfirmed.empty() && cand.size() > unconfirmed.size()) ? (int)unconfirmed.size() : 0;
        std::vector<int> tb = cand.empty() ? all_tracks() : cand;             // :358-361
        std::vector<int> db = un_da.empty() ? all_dets() : un_da;             // :362-365
        std::vector<std::pair<int, int>> matches_b;
        std::vector<int> un_tb, un_db;
        last_sizes[2] = (int)tb.size(); last_sizes[3] = (int)db.size();
        if (tb.empty() || db.empty()) {
            un_tb = tb; un_db = db;
        } else {
            const int n = (int)tb.size(), m = (int)db.size();
            std::vector<float> trk((size_t)n * 4), det((size_t)m * 4);
            std::vector<int> tsu(n);
            for (int r = 0; r < n; ++r) { tracks[tb[r]].to_tlwh(&trk[(size_t)r * 4]); tsu[r] = tracks[tb[r]].tsu; }
            for (int c = 0; c < m; ++c) std::memcpy(&det[(size_t)c * 4], dets[db[c]].tlwh, 16);
            std::vector<float> cost((size_t)n * m);
            orc_iou_cost_tlwh(trk.data(), tsu.data(), n, det.data(), m, cost.data());
            lap_stage(cost, tb, db, max_iou_dist, cand.empty() ? 0 : dup_first, matches_b, un_tb, un_db);
        }
        unmatched_dets = un_db;                                                // :741

        matches = matches_a;                                                   // :743-759
        std::set<int> mt, md;
        for (auto& p : matches_a) { mt.insert(p.first); md.insert(p.second); }
        for (auto& p : matches_b)
            if (!mt.count(p.first) && !md.count(p.second)) { matches.push_back(p); mt.insert(p.first); md.insert(p.second); }
        std::set<int> us(un_ta_filtered.begin(), un_ta_filtered.end());        // :761-765
        us.insert(un_tb.begin(), un_tb.end());
        unmatched_tracks.assign(us.begin(), us.end());
        last_sizes[4] = (int)matches_a.size(); last_sizes[5] = (int)matches.size() - (int)matches_a.size();
        last_sizes[6] = dup_first;
    }

    // Tracker::update (:614-662)
    void tracker_update(const std::vector<SDet>& dets, int dim) {
        std::vector<std::pair<int, int>> matches;
        std::vector<int> un_t, un_d;
        match(dets, dim, matches, un_t, un_d);
        for (auto& p : matches) tracks[p.first].update(dets[p.second], dim, ema_alpha, n_init);
        for (int k : un_t) tracks[k].mark_missed(max_age);
        for (int d : un_d) tracks.emplace_back(dets[d], next_id++, dim);       // initiate_track (:768-770)
        last_sizes[7] = (int)un_d.size();
        tracks.erase(std::remove_if(tracks.begin(), tracks.end(), [](const STrack& t) { return t.state == Deleted; }), tracks.end());
        std::vector<int> active;
        bool any = false;
        for (const auto& t : tracks)
            if (t.state == Confirmed) { active.push_back(t.id); any = any || !t.feat.empty(); }
        if (any) {                                                             // partial_fit (:213-238)
            for (const auto& t : tracks)
                if (t.state == Confirmed && !t.feat.empty()) {
                    auto& q = samples[t.id];
                    q.push_back(t.feat);
                    if (nn_budget > 0 && (int)q.size() > nn_budget) q.pop_front();
                }
            std::unordered_map<int, std::deque<std::vector<float>>> kept;
            for (int id : active) {
                auto it = samples.find(id);
                if (it != samples.end()) kept[id] = std::move(it->second);
            }
            samples = std::move(kept);
        }
    }
};

extern "C" {

OrcStrongSort* orc_strongsort_create(int max_age, float min_conf, float max_cos_dist, float max_iou_dist, int n_init, int nn_budget,
                                     float mc_lambda, float ema_alpha) {
    auto* s = new OrcStrongSort();
    s->max_age = max_age; s->min_conf = min_conf; s->max_cos_dist = max_cos_dist; s->max_iou_dist = max_iou_dist;
    s->n_init = n_init; s->nn_budget = nn_budget; s->mc_lambda = mc_lambda; s->ema_alpha = ema_alpha;
    return s;
}
void orc_strongsort_destroy(OrcStrongSort* s) { delete s; }
void orc_strongsort_reset(OrcStrongSort* s) { s->tracks.clear(); s->next_id = 1; s->samples.clear(); }    // :772-778
void orc_strongsort_set_tie_mode(OrcStrongSort* s, int mode) { s->tie_mode = mode; }
void orc_strongsort_last_sizes(const OrcStrongSort* s, int* out8) { std::memcpy(out8, s->last_sizes, sizeof(s->last_sizes)); }
int orc_strongsort_count(const OrcStrongSort* s) { return (int)s->tracks.size(); }

// StrongSORT::update (:829-987).  dets (n x 6), embs (n x dim) or nullptr.
int orc_strongsort_update(OrcStrongSort* s, const float* dets, int n, const float* embs, int dim, float* out, int out_cap) {
    std::vector<SDet> ds;
    for (int i = 0; i < n; ++i) {
        const float* r = dets + (size_t)i * 6;
        if (!(r[4] >= s->min_conf)) continue;                                  // :849-854
        SDet d;
        d.tlwh[0] = r[0]; d.tlwh[1] = r[1]; d.tlwh[2] = r[2] - r[0]; d.tlwh[3] = r[3] - r[1];      // :923-932
        d.conf = r[4]; d.cls = (int)r[5]; d.det_ind = i;
        d.feat = (embs && dim > 0) ? embs + (size_t)i * dim : nullptr;
        ds.push_back(d);
    }
    std::memset(s->last_sizes, 0, sizeof(s->last_sizes));
    if (ds.empty()) {                                                          // :833-837, :856-860: no camera update
        for (auto& t : s->tracks) t.predict();
        s->tracker_update({}, dim);
        return 0;
    }
    for (auto& t : s->tracks) t.camera_update_identity();                      // :873-878
    for (auto& t : s->tracks) t.predict();                                     // :942
    s->tracker_update(ds, dim);
    int rows = 0;
    for (const auto& t : s->tracks) {                                          // :946-972
        if (t.state != Confirmed || t.tsu >= 1) continue;
        if (rows < out_cap) {
            float b[4];
            t.to_tlbr(b);
            float* o = out + (size_t)rows * 8;
            o[0] = b[0]; o[1] = b[1]; o[2] = b[2]; o[3] = b[3];
            o[4] = (float)t.id; o[5] = t.conf; o[6] = (float)t.cls; o[7] = (float)t.det_ind;
        }
        ++rows;
    }
    return rows <= out_cap ? rows : -rows;
}

// rows of [id, state, hits, age, tsu, conf, cls, det_ind, has_feat, n_samples, mean 8, cov 64] = 82 floats, track-list order;
// feats (nullable): dim floats per row (the smoothed feature, zeros when none)
int orc_strongsort_dump(const OrcStrongSort* s, float* rows82, float* feats, int dim, int cap_rows) {
    int k = 0;
    for (const auto& t : s->tracks) {
        if (k >= cap_rows) break;
        float* o = rows82 + (size_t)k * 82;
        auto it = s->samples.find(t.id);
        o[0] = (float)t.id; o[1] = (float)t.state; o[2] = (float)t.hits; o[3] = (float)t.age; o[4] = (float)t.tsu;
        o[5] = t.conf; o[6] = (float)t.cls; o[7] = (float)t.det_ind; o[8] = t.feat.empty() ? 0.0f : 1.0f;
        o[9] = (it == s->samples.end()) ? 0.0f : (float)it->second.size();
        std::memcpy(o + 10, t.mean, sizeof(t.mean));
        std::memcpy(o + 18, t.cov, sizeof(t.cov));
        if (feats) {
            if (t.feat.empty()) std::memset(feats + (size_t)k * dim, 0, sizeof(float) * dim);
            else std::memcpy(feats + (size_t)k * dim, t.feat.data(), sizeof(float) * dim);
        }
        ++k;
    }
    return k;
}

}  // extern "C"
