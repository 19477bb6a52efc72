// TEST INFRASTRUCTURE (oracle): restatement of BoT-SORT's per-frame state machine (CMC and the ReID
// network are outside the hot path: cmc_method = "none", embeddings are passed in).
//   BotSTrack                      src/trackers/botsort.cpp:15-193, include/motcpp/trackers/botsort.hpp:27-101
//   BotSort::update and its stages src/trackers/botsort.cpp:260-764
//   pointer-vector cost templates  include/motcpp/utils/matching.hpp:129-234
//   KalmanFilterXYWH               include/motcpp/motion/kalman_filters/xywh_kf.hpp:17-135
// Reference behaviours kept on purpose:
//   - update() returns at once on an empty detection matrix, without advancing frame_count_ (:267-269);
//   - unconfirmed tracks are never predicted (only strack_pool is, :311);
//   - a LOST track that is re-found is re-activated in place inside lost_stracks_, its id then filters it
//     out of the new lost list, and nothing ever copies it to active_tracks_: it vanishes (:712-744);
//   - unmatched tracked tracks become Lost only when the second association actually runs (:525-559);
//   - remove_duplicate_stracks exists (:810) but is never called.
// Vector sums (.dot(), .norm()) use the fixed "lanes32" order below; Eigen's order is unspecified, so feature
// values agree with the stock build to fp32 round-off, not bit for bit.
// ID counter is per tracker instance and restarts at 0 on reset (botsort.cpp:249,257).
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

enum State { New = 0, Tracked = 1, Lost = 2, Removed = 3 };

// Summation order of every feature dot product / norm ("lanes32", shared with the CUDA kernel so that a warp can
// evaluate it with coalesced loads): the products of elements 4q..4q+3 go, in that order, into partial sum q % 32
// (each partial starts at +0 and takes its quadruples in ascending q); the 32 partials are then combined by the
// butterfly p[l] += p[l ^ o], o = 16, 8, 4, 2, 1.  Eigen's own .dot()/.norm() order is vectorised and
// unspecified, so any fixed order agrees with the stock build to fp32 round-off only.
float seq_dot(const float* x, const float* y, int n) {
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
float seq_norm(const float* x, int n) { return std::sqrt(seq_dot(x, x, n)); }

struct BotTrack {
    int id = 0, frame_id = 0, start_frame = 0, tracklet_len = 0;
    float xywh[4];
    float conf = 0.0f;
    int cls = 0, det_ind = -1;
    bool has_mean = false;
    float mean[8], cov[64];
    int state = New;
    bool is_activated = false;
    std::vector<float> curr_feat, smooth_feat;

    // botsort.cpp:23-46; det7 = [x1,y1,x2,y2,conf,cls,det_ind]
    BotTrack(const float* det7, const float* feat, int dim) {
        const float w = det7[2] - det7[0], h = det7[3] - det7[1];
        xywh[0] = det7[0] + w / 2.0f; xywh[1] = det7[1] + h / 2.0f; xywh[2] = w; xywh[3] = h;
        conf = det7[4]; cls = (int)det7[5]; det_ind = (int)det7[6];
        if (feat && dim > 0) {
            curr_feat.assign(feat, feat + dim);
            smooth_feat = curr_feat;
            normalise();
        }
    }
    void normalise() {                                                   // :42-45, :165-168
        const int n = (int)smooth_feat.size();
        const float nrm = seq_norm(smooth_feat.data(), n);
        if (nrm > 0.0f)
            for (auto& v : smooth_feat) v = v / nrm;
    }
    void update_features(const std::vector<float>& feat) {               // :158-169
        curr_feat = feat;
        if (smooth_feat.empty()) smooth_feat = feat;
        else {
            const float alpha = 0.9f, beta = 1.0f - alpha;
            for (size_t k = 0; k < smooth_feat.size(); ++k) smooth_feat[k] = alpha * smooth_feat[k] + beta * feat[k];
        }
        normalise();
    }
    void xyxy(float* b) const {                                          // :171-181
        const float* s = has_mean ? mean : xywh;
        b[0] = s[0] - s[2] / 2; b[1] = s[1] - s[3] / 2; b[2] = s[0] + s[2] / 2; b[3] = s[1] + s[3] / 2;
    }
    void predict() { orc_kf_xywh_predict(mean, cov); }                  // :48-52
    void activate(int new_id, int frame) {                               // :94-110
        id = new_id;
        orc_kf_xywh_initiate(xywh, mean, cov);
        has_mean = true;
        tracklet_len = 0;
        state = Tracked;
        if (frame == 1) is_activated = true;
        frame_id = frame; start_frame = frame;
    }
    void absorb(const BotTrack& det) {
        orc_kf_xywh_update(mean, cov, det.xywh);
        if (!det.curr_feat.empty()) update_features(det.curr_feat);
        state = Tracked; is_activated = true;
        conf = det.conf; cls = det.cls; det_ind = det.det_ind;
    }
    void re_activate(const BotTrack& det, int frame) {                   // :112-133 (new_id = false)
        absorb(det);
        tracklet_len = 0;
        frame_id = frame;
    }
    void update(const BotTrack& det, int frame) {                        // :135-156
        frame_id = frame;
        ++tracklet_len;
        absorb(det);
    }
};

}  // namespace

struct OrcBotSort {
    float track_high_thresh, track_low_thresh, new_track_thresh, match_thresh, proximity_thresh, appearance_thresh;
    int max_time_lost, fuse_first, with_reid, dim;
    int frame_count = 0, id_counter = 0;
    int last_sizes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<BotTrack> active, lost;
};

namespace {

// cost matrix of first_association / handle_unconfirmed_tracks (:438-466, :598-620)
std::vector<float> fused_cost(const OrcBotSort* s, const std::vector<BotTrack*>& trks, const std::vector<BotTrack>& dets,
                              bool fuse) {
    const int n = (int)trks.size(), m = (int)dets.size();
    std::vector<float> a((size_t)n * 4), b((size_t)m * 4), conf(m);
    for (int i = 0; i < n; ++i) trks[i]->xyxy(&a[4 * i]);
    for (int j = 0; j < m; ++j) { dets[j].xyxy(&b[4 * j]); conf[j] = dets[j].conf; }
    std::vector<float> d((size_t)n * m);
    orc_iou_distance(a.data(), n, b.data(), m, d.data());                // matching.hpp:129-182 (both sides non-empty here)
    std::vector<char> mask((size_t)n * m);
    for (size_t k = 0; k < d.size(); ++k) mask[k] = d[k] > s->proximity_thresh;
    if (fuse) orc_fuse_score(d.data(), n, m, conf.data());
    if (s->with_reid) {
        std::vector<float> tn(n), dn(m);
        for (int i = 0; i < n; ++i) tn[i] = seq_norm(trks[i]->smooth_feat.data(), (int)trks[i]->smooth_feat.size());
        for (int j = 0; j < m; ++j) dn[j] = seq_norm(dets[j].smooth_feat.data(), (int)dets[j].smooth_feat.size());
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < m; ++j) {
                const size_t k = (size_t)i * m + j;
                float e = 1.0f;
                if (!mask[k]) {                      // masked pairs are overwritten with 1 anyway (:458-460)
                    const int dim = (int)std::min(trks[i]->smooth_feat.size(), dets[j].smooth_feat.size());
                    const float sim = seq_dot(trks[i]->smooth_feat.data(), dets[j].smooth_feat.data(), dim) /
                                      (tn[i] * dn[j] + 1e-10f);         // matching.cpp:89
                    e = std::max(0.0f, 1.0f - sim) / 2.0f;              // matching.cpp:90, botsort.cpp:450
                    if (e > s->appearance_thresh) e = 1.0f;             // :455-457
                }
                d[k] = std::min(d[k], e);                               // :465
            }
    }
    return d;
}

}  // namespace

extern "C" {

OrcBotSort* orc_botsort_create(float track_high_thresh, float track_low_thresh, float new_track_thresh, int track_buffer,
                               float match_thresh, float proximity_thresh, float appearance_thresh, int frame_rate,
                               int fuse_first_associate, int with_reid) {
    auto* s = new OrcBotSort();
    s->track_high_thresh = track_high_thresh; s->track_low_thresh = track_low_thresh;
    s->new_track_thresh = new_track_thresh; s->match_thresh = match_thresh;
    s->proximity_thresh = proximity_thresh; s->appearance_thresh = appearance_thresh;
    s->max_time_lost = (int)(frame_rate / 30.0f * track_buffer);          // botsort.cpp:235-236
    s->fuse_first = fuse_first_associate; s->with_reid = with_reid; s->dim = 0;
    return s;
}
void orc_botsort_destroy(OrcBotSort* s) { delete s; }
void orc_botsort_reset(OrcBotSort* s) { s->frame_count = 0; s->id_counter = 0; s->active.clear(); s->lost.clear(); }

/* dets (n,6); embs (n,dim) or NULL / dim 0; out rows [x1,y1,x2,y2,id,conf,cls,det_ind] */
int orc_botsort_update(OrcBotSort* s, const float* dets, int n, const float* embs, int dim, float* out, int out_cap) {
    if (n == 0) return 0;                                                  // :267-269
    ++s->frame_count;
    const int frame = s->frame_count;
    for (int k = 0; k < 8; ++k) s->last_sizes[k] = 0;

    // split_detections + create_detections (:336-400)
    std::vector<BotTrack> detections, detections_second;
    for (int i = 0; i < n; ++i) {
        float row[7];
        std::memcpy(row, dets + 6 * i, 6 * sizeof(float));
        row[6] = (float)i;
        const float c = row[4];
        if (c > s->track_high_thresh)
            detections.emplace_back(row, (s->with_reid && embs && dim > 0) ? embs + (size_t)i * dim : nullptr, dim);
        else if (c > s->track_low_thresh)
            detections_second.emplace_back(row, nullptr, 0);
    }
    s->active.reserve(s->active.size() + detections.size() + 10);
    s->lost.reserve(s->lost.size() + s->active.size() + 10);
    std::vector<BotTrack*> unconfirmed, tracked, pool;
    for (auto& t : s->active) (t.is_activated ? tracked : unconfirmed).push_back(&t);
    pool = tracked;
    for (auto& t : s->lost) pool.push_back(&t);                            // joint_stracks (:766-785): ids are distinct
    for (auto* t : pool) t->predict();                                     // :311

    // first association (:402-495)
    std::vector<int> u_track, u_det;
    const int n1 = (int)pool.size(), m1 = (int)detections.size();
    s->last_sizes[0] = n1; s->last_sizes[1] = m1;
    {
        std::vector<int> r2c(n1, -1), c2r(m1, -1);
        if (n1 > 0 && m1 > 0) {
            std::vector<float> d = fused_cost(s, pool, detections, s->fuse_first != 0);
            orc_linear_assignment(d.data(), n1, m1, m1, s->match_thresh, r2c.data(), c2r.data());
        }
        for (int i = 0; i < n1; ++i) {
            if (r2c[i] < 0) { u_track.push_back(i); continue; }
            if (pool[i]->state == Tracked) pool[i]->update(detections[r2c[i]], frame);
            else pool[i]->re_activate(detections[r2c[i]], frame);          // a re-found lost track (vanishes below)
        }
        for (int j = 0; j < m1; ++j) if (c2r[j] < 0) u_det.push_back(j);
    }

    // second association (:497-562)
    std::vector<BotTrack*> newly_lost;
    {
        std::vector<BotTrack*> r_tracked;
        for (int i : u_track) if (pool[i]->state == Tracked) r_tracked.push_back(pool[i]);
        const int n2 = (int)r_tracked.size(), m2 = (int)detections_second.size();
        if (n2 > 0 && m2 > 0) {
            s->last_sizes[2] = n2; s->last_sizes[3] = m2;
            std::vector<float> a((size_t)n2 * 4), b((size_t)m2 * 4), d((size_t)n2 * m2);
            for (int i = 0; i < n2; ++i) r_tracked[i]->xyxy(&a[4 * i]);
            for (int j = 0; j < m2; ++j) detections_second[j].xyxy(&b[4 * j]);
            orc_iou_distance(a.data(), n2, b.data(), m2, d.data());
            std::vector<int> r2c(n2), c2r(m2);
            orc_linear_assignment(d.data(), n2, m2, m2, 0.5f, r2c.data(), c2r.data());
            for (int i = 0; i < n2; ++i) {
                if (r2c[i] >= 0) r_tracked[i]->update(detections_second[r2c[i]], frame);
                else { r_tracked[i]->state = Lost; newly_lost.push_back(r_tracked[i]); }
            }
        }
    }

    // unconfirmed tracks (:564-647); u_det_final indexes `remaining`
    std::vector<BotTrack> remaining;
    for (int j : u_det) remaining.push_back(detections[j]);
    std::vector<int> u_det_final;
    if (unconfirmed.empty() || remaining.empty()) {
        for (size_t k = 0; k < remaining.size(); ++k) u_det_final.push_back((int)k);
    } else {
        const int n3 = (int)unconfirmed.size(), m3 = (int)remaining.size();
        s->last_sizes[4] = n3; s->last_sizes[5] = m3;
        std::vector<float> d = fused_cost(s, unconfirmed, remaining, true);
        std::vector<int> r2c(n3), c2r(m3);
        orc_linear_assignment(d.data(), n3, m3, m3, 0.7f, r2c.data(), c2r.data());
        for (int i = 0; i < n3; ++i) {
            if (r2c[i] >= 0) unconfirmed[i]->update(remaining[r2c[i]], frame);
            else unconfirmed[i]->state = Removed;
        }
        for (int j = 0; j < m3; ++j) if (c2r[j] < 0) u_det_final.push_back(j);
    }

    // new tracks (:649-667)
    for (int j : u_det_final) {
        if (remaining[j].conf < s->new_track_thresh) continue;
        s->active.push_back(remaining[j]);
        s->active.back().activate(++s->id_counter, frame);
        ++s->last_sizes[6];
    }
    // expire lost tracks (:669-676); end_frame_ == frame_id_ at all times
    for (auto& t : s->lost)
        if (frame - t.frame_id > s->max_time_lost) t.state = Removed;

    // prepare_output (:678-764): re-found lost tracks are Tracked now and drop out of BOTH lists
    std::vector<BotTrack> new_lost, new_active;
    for (auto& t : s->lost)
        if (t.state == Lost) new_lost.push_back(t);
    for (auto* t : newly_lost) new_lost.push_back(*t);
    for (auto& t : s->active)
        if (t.state == Tracked) new_active.push_back(t);
    s->active.swap(new_active);
    s->lost.swap(new_lost);
    s->last_sizes[7] = (int)s->lost.size();

    int rows = 0;
    for (const auto& t : s->active) if (t.is_activated) ++rows;
    if (rows > out_cap) return -rows;
    int k = 0;
    for (const auto& t : s->active) {
        if (!t.is_activated) continue;
        float* o = out + 8 * k++;
        t.xyxy(o);
        o[4] = (float)t.id; o[5] = t.conf; o[6] = (float)t.cls; o[7] = (float)t.det_ind;
    }
    return rows;
}

int orc_botsort_counts(const OrcBotSort* s, int* n_active, int* n_lost) {
    *n_active = (int)s->active.size(); *n_lost = (int)s->lost.size();
    return s->frame_count;
}
void orc_botsort_last_sizes(const OrcBotSort* s, int* sizes8) { std::memcpy(sizes8, s->last_sizes, sizeof(s->last_sizes)); }

/* list `which` (0 active, 1 lost): rows of [id, state, is_activated, frame_id, start_frame, tracklet_len, conf, cls,
 * det_ind, has_feat, mean 8, cov 64] = 82 floats; feats (nullable) receives smooth_feat rows of `dim` floats */
int orc_botsort_dump(const OrcBotSort* s, int which, float* out, float* feats, int dim, int cap_rows) {
    const auto& v = which == 0 ? s->active : s->lost;
    int k = 0;
    for (const auto& t : v) {
        if (k >= cap_rows) break;
        float* o = out + (size_t)82 * k;
        o[0] = (float)t.id; o[1] = (float)t.state; o[2] = t.is_activated ? 1.0f : 0.0f; o[3] = (float)t.frame_id;
        o[4] = (float)t.start_frame; o[5] = (float)t.tracklet_len; o[6] = t.conf; o[7] = (float)t.cls;
        o[8] = (float)t.det_ind; o[9] = t.smooth_feat.empty() ? 0.0f : 1.0f;
        std::memcpy(o + 10, t.mean, sizeof(t.mean));
        std::memcpy(o + 18, t.cov, sizeof(t.cov));
        if (feats && dim > 0) {
            float* f = feats + (size_t)dim * k;
            for (int e = 0; e < dim; ++e) f[e] = e < (int)t.smooth_feat.size() ? t.smooth_feat[e] : 0.0f;
        }
        ++k;
    }
    return k;
}

}  // extern "C"
