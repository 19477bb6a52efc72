// TEST INFRASTRUCTURE (oracle): restatement of ByteTrack's per-frame state machine.
//   STrack            include/motcpp/trackers/bytetrack.hpp:31-90, src/trackers/bytetrack.cpp:15-128
//   ByteTrack::update src/trackers/bytetrack.cpp:166-621
//   joint/sub/remove_duplicate_stracks  src/trackers/bytetrack.cpp:623-706
// All twelve "parity traps" of SURVEY.md section 8 that concern ByteTrack are reproduced:
// predictions live on copies and are written back only for matched tracks; the second
// association uses un-predicted boxes; unmatched tracked tracks are marked Lost only when
// low-confidence detections exist; etc.  One deliberate difference: the ID counter is
// per tracker instance (the reference's is a process-global static, bytetrack.hpp:33-36),
// which is what "one engine stream" means on the GPU side.
#include "oracle.h"

#include <algorithm>
#include <cstring>
#include <unordered_set>
#include <vector>

namespace {

enum State { kNew = 0, kTracked = 1, kLost = 2, kRemoved = 3 };

struct STrack {
    float xywh[4], tlwh[4], xyah[4];
    float conf = 0.0f;
    int cls = 0, det_ind = 0;
    int id = 0;
    int state = kNew;
    bool activated = false;
    int tracklet_len = 0, frame_id = 0, start_frame = 0;
    bool has_state = false;
    float mean[8];
    float cov[64];

    // bytetrack.cpp:15-33: det row = [x1,y1,x2,y2,conf,cls,det_ind]
    static STrack from_det(const float* det7) {
        STrack t;
        orc_xyxy2xywh(det7, t.xywh);
        orc_xywh2tlwh(t.xywh, t.tlwh);
        orc_tlwh2xyah(t.tlwh, t.xyah);
        t.conf = det7[4];
        t.cls = static_cast<int>(det7[5]);
        t.det_ind = static_cast<int>(det7[6]);
        return t;
    }

    // bytetrack.cpp:118-128
    void xyxy(float* out) const {
        if (!has_state) { orc_xywh2xyxy(xywh, out); return; }
        float wh[4];
        orc_xyah2xywh(mean, wh);
        orc_xywh2xyxy(wh, out);
    }

    // bytetrack.cpp:35-49
    void activate(int new_id, int frame) {
        id = new_id;
        orc_kf_xyah_initiate(xyah, mean, cov);
        has_state = true;
        tracklet_len = 0;
        state = kTracked;
        if (frame == 1) activated = true;
        frame_id = frame;
        start_frame = frame;
    }

    // bytetrack.cpp:51-66 (new_id is always false at the call sites)
    void re_activate(const STrack& det, int frame) {
        orc_kf_xyah_update(mean, cov, det.xyah, 0.0f);
        tracklet_len = 0;
        state = kTracked;
        activated = true;
        frame_id = frame;
        conf = det.conf; cls = det.cls; det_ind = det.det_ind;
    }

    // bytetrack.cpp:68-85
    void update(const STrack& det, int frame) {
        frame_id = frame;
        tracklet_len += 1;
        orc_kf_xyah_update(mean, cov, det.xyah, 0.0f);
        state = kTracked;
        activated = true;
        conf = det.conf; cls = det.cls; det_ind = det.det_ind;
    }
};

using List = std::vector<STrack>;

List joint(const List& a, const List& b) {                 // bytetrack.cpp:623-640
    std::unordered_set<int> seen;
    List res = a;
    for (const auto& t : a) seen.insert(t.id);
    for (const auto& t : b)
        if (seen.insert(t.id).second) res.push_back(t);
    return res;
}

List subtract(const List& a, const List& b) {              // bytetrack.cpp:642-657
    std::unordered_set<int> drop;
    for (const auto& t : b) drop.insert(t.id);
    List res;
    for (const auto& t : a)
        if (!drop.count(t.id)) res.push_back(t);
    return res;
}

std::vector<float> boxes_of(const std::vector<const STrack*>& ts) {
    std::vector<float> out(ts.size() * 4);
    for (size_t i = 0; i < ts.size(); ++i) ts[i]->xyxy(&out[4 * i]);
    return out;
}

struct Assignment {
    std::vector<int> row2col, col2row;
};

Assignment assign(const std::vector<float>& cost, int n, int m, float thresh) {
    Assignment a;
    a.row2col.assign(n, -1);
    a.col2row.assign(m, -1);
    if (n > 0 && m > 0) orc_linear_assignment(cost.data(), n, m, m, thresh, a.row2col.data(), a.col2row.data());
    return a;
}

}  // namespace

struct OrcByteTrack {
    float det_thresh, min_conf, track_thresh, match_thresh;
    int max_time_lost;
    int frame_count = 0, frame_id = 0;
    int id_counter = 0;
    List active, lost;
    int last_sizes[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    int update(const float* dets, int n, float* out, int out_cap);
    void dedupe();
};

// bytetrack.cpp:659-706
void OrcByteTrack::dedupe() {
    last_sizes[6] = static_cast<int>(active.size());
    last_sizes[7] = static_cast<int>(lost.size());
    if (active.empty() || lost.empty()) return;
    std::vector<const STrack*> pa, pb;
    for (const auto& t : active) pa.push_back(&t);
    for (const auto& t : lost) pb.push_back(&t);
    const std::vector<float> ba = boxes_of(pa), bb = boxes_of(pb);
    const int na = static_cast<int>(pa.size()), nb = static_cast<int>(pb.size());
    std::vector<float> dist((size_t)na * nb);
    orc_iou_distance(ba.data(), na, bb.data(), nb, dist.data());
    std::vector<char> dup_a(na, 0), dup_b(nb, 0);
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j)
            if (dist[(size_t)i * nb + j] < 0.15f) {
                const int age_a = active[i].frame_id - active[i].start_frame;
                const int age_b = lost[j].frame_id - lost[j].start_frame;
                if (age_a > age_b) dup_b[j] = 1; else dup_a[i] = 1;
            }
    List ra, rb;
    for (int i = 0; i < na; ++i) if (!dup_a[i]) ra.push_back(active[i]);
    for (int j = 0; j < nb; ++j) if (!dup_b[j]) rb.push_back(lost[j]);
    active.swap(ra);
    lost.swap(rb);
}

int OrcByteTrack::update(const float* dets, int n, float* out, int out_cap) {
    ++frame_count;                                                   // :181-182
    ++frame_id;

    // :185-219 split by confidence (strict comparisons on both sides), build detection STracks
    std::vector<STrack> det_high, det_second;
    for (int i = 0; i < n; ++i) {
        float row[7];
        std::memcpy(row, dets + 6 * i, 6 * sizeof(float));
        row[6] = static_cast<float>(i);
        const float cf = row[4];
        if (cf > track_thresh) det_high.push_back(STrack::from_det(row));
        if (cf > min_conf && cf < track_thresh) det_second.push_back(STrack::from_det(row));
    }

    // :228-247 unconfirmed / tracked partition of the active list (indices into `active`)
    std::vector<int> unconfirmed_idx, tracked_idx;
    for (int i = 0; i < (int)active.size(); ++i)
        (active[i].activated ? tracked_idx : unconfirmed_idx).push_back(i);

    // :251-265 pool = tracked ++ lost (deduplicated by id), predicted on COPIES
    struct PoolRef { int idx; bool from_active; };
    List pool;
    std::vector<PoolRef> origin;
    {
        std::unordered_set<int> seen;
        for (int idx : tracked_idx) { pool.push_back(active[idx]); origin.push_back({idx, true}); seen.insert(active[idx].id); }
        for (int i = 0; i < (int)lost.size(); ++i)
            if (seen.insert(lost[i].id).second) { pool.push_back(lost[i]); origin.push_back({i, false}); }
    }
    for (auto& t : pool) {                                            // multi_predict :97-116
        if (t.state != kTracked) t.mean[7] = 0.0f;
        orc_kf_xyah_predict(t.mean, t.cov);
    }

    // :267-309 first association: IoU distance fused with detection score
    const int n1 = (int)pool.size(), m1 = (int)det_high.size();
    last_sizes[0] = n1; last_sizes[1] = m1;
    std::vector<const STrack*> pp, pd;
    for (const auto& t : pool) pp.push_back(&t);
    for (const auto& t : det_high) pd.push_back(&t);
    const std::vector<float> pool_boxes = boxes_of(pp), high_boxes = boxes_of(pd);
    std::vector<float> cost1((size_t)n1 * m1);
    // utils::iou_distance on ptr vectors returns Ones for an empty side; nothing to fill then
    if (n1 > 0 && m1 > 0) {
        orc_iou_distance(pool_boxes.data(), n1, high_boxes.data(), m1, cost1.data());
        std::vector<float> confs(m1);
        for (int j = 0; j < m1; ++j) confs[j] = det_high[j].conf;
        orc_fuse_score(cost1.data(), n1, m1, confs.data());
    }
    const Assignment a1 = assign(cost1, n1, m1, match_thresh);

    List activated_new;       // only genuinely new entries matter for joint(); see below
    List refind;
    List lost_new;
    std::vector<int> removed_ids;

    for (int r = 0; r < n1; ++r) {                                    // :337-365, ascending row order
        const int dj = a1.row2col[r];
        if (dj < 0) continue;
        STrack& orig = origin[r].from_active ? active[origin[r].idx] : lost[origin[r].idx];
        std::memcpy(orig.mean, pool[r].mean, sizeof(orig.mean));      // write back the prediction
        std::memcpy(orig.cov, pool[r].cov, sizeof(orig.cov));
        if (orig.state == kTracked) {
            orig.update(det_high[dj], frame_id);
        } else {
            orig.re_activate(det_high[dj], frame_id);
            refind.push_back(orig);
        }
    }

    // :367-442 second association: still-Tracked leftovers (from the active list) vs low-score dets,
    // boxes taken from the ORIGINAL (un-predicted) tracks
    std::vector<int> r_tracked_pool_rows;
    for (int r = 0; r < n1; ++r)
        if (a1.row2col[r] < 0 && pool[r].state == kTracked && origin[r].from_active) r_tracked_pool_rows.push_back(r);
    const int n2 = (int)r_tracked_pool_rows.size(), m2 = (int)det_second.size();
    last_sizes[2] = n2; last_sizes[3] = m2;
    if (n2 > 0 && m2 > 0) {
        std::vector<const STrack*> rt, d2;
        for (int r : r_tracked_pool_rows) rt.push_back(&active[origin[r].idx]);
        for (const auto& t : det_second) d2.push_back(&t);
        const std::vector<float> rb = boxes_of(rt), db = boxes_of(d2);
        std::vector<float> cost2((size_t)n2 * m2);
        orc_iou_distance(rb.data(), n2, db.data(), m2, cost2.data());
        const Assignment a2 = assign(cost2, n2, m2, 0.5f);
        for (int i = 0; i < n2; ++i) {
            const int r = r_tracked_pool_rows[i];
            STrack& trk = active[origin[r].idx];
            const int dj = a2.row2col[i];
            if (dj >= 0) {
                std::memcpy(trk.mean, pool[r].mean, sizeof(trk.mean));
                std::memcpy(trk.cov, pool[r].cov, sizeof(trk.cov));
                trk.update(det_second[dj], frame_id);                 // state is Tracked here
            }
        }
        for (int i = 0; i < n2; ++i) {                                // :435-441
            if (a2.row2col[i] >= 0) continue;
            STrack& trk = active[origin[r_tracked_pool_rows[i]].idx];
            if (trk.state != kLost) { trk.state = kLost; lost_new.push_back(trk); }
        }
    }

    // :448-542 unconfirmed tracks vs detections left over from the first association
    std::vector<int> u_detection;
    for (int j = 0; j < m1; ++j) if (a1.col2row[j] < 0) u_detection.push_back(j);
    std::vector<int> u_detection_final;
    const int n3 = (int)unconfirmed_idx.size(), m3 = (int)u_detection.size();
    last_sizes[4] = n3; last_sizes[5] = m3;
    if (n3 > 0 && m3 > 0) {
        std::vector<const STrack*> ut, rd;
        for (int idx : unconfirmed_idx) ut.push_back(&active[idx]);
        for (int j : u_detection) rd.push_back(&det_high[j]);
        const std::vector<float> ub = boxes_of(ut), db = boxes_of(rd);
        std::vector<float> cost3((size_t)n3 * m3);
        orc_iou_distance(ub.data(), n3, db.data(), m3, cost3.data());
        std::vector<float> confs(m3);
        for (int j = 0; j < m3; ++j) confs[j] = det_high[u_detection[j]].conf;
        orc_fuse_score(cost3.data(), n3, m3, confs.data());
        const Assignment a3 = assign(cost3, n3, m3, 0.7f);
        for (int j = 0; j < m3; ++j) if (a3.col2row[j] < 0) u_detection_final.push_back(u_detection[j]);
        for (int i = 0; i < n3; ++i) {
            const int dj = a3.row2col[i];
            if (dj >= 0) active[unconfirmed_idx[i]].update(det_high[u_detection[dj]], frame_id);   // no predict
        }
        for (int i = 0; i < n3; ++i)
            if (a3.row2col[i] < 0) {
                active[unconfirmed_idx[i]].state = kRemoved;
                removed_ids.push_back(active[unconfirmed_idx[i]].id);
            }
    } else {
        u_detection_final = u_detection;                              // :539-542
    }

    // :546-554 new tracks, IDs handed out in u_detection_final order
    for (int j : u_detection_final) {
        STrack& d = det_high[j];
        if (d.conf >= det_thresh) {
            d.activate(++id_counter, frame_id);
            activated_new.push_back(d);
        }
    }

    // :557-562 expire lost tracks
    for (auto& t : lost)
        if (frame_count - t.frame_id > max_time_lost) { t.state = kRemoved; removed_ids.push_back(t.id); }

    // :565-578 list algebra.  Tracks updated in place already live in `active`, so joint() with the
    // reference's `activated_stracks` only ever appends the new ones, then the re-found ones.
    List keep;
    for (const auto& t : active) if (t.state == kTracked) keep.push_back(t);
    active = joint(joint(keep, activated_new), refind);
    lost = subtract(lost, active);
    lost.insert(lost.end(), lost_new.begin(), lost_new.end());
    {
        std::unordered_set<int> drop(removed_ids.begin(), removed_ids.end());
        List res;
        for (const auto& t : lost) if (!drop.count(t.id)) res.push_back(t);
        lost.swap(res);
    }

    dedupe();                                                         // :581-585

    // :589-620 output rows for activated tracks, in list order
    int rows = 0;
    for (const auto& t : active) if (t.activated) ++rows;
    if (rows > out_cap) return -rows;
    int k = 0;
    for (const auto& t : active) {
        if (!t.activated) continue;
        float* o = out + 8 * k++;
        t.xyxy(o);
        o[4] = static_cast<float>(t.id);
        o[5] = t.conf;
        o[6] = static_cast<float>(t.cls);
        o[7] = static_cast<float>(t.det_ind);
    }
    return rows;
}

extern "C" {

OrcByteTrack* orc_bytetrack_create(float det_thresh, int max_age, int max_obs, int min_hits,
                                   float iou_threshold, float min_conf, float track_thresh,
                                   float match_thresh, int track_buffer, int frame_rate) {
    (void)det_thresh; (void)max_age; (void)max_obs; (void)min_hits; (void)iou_threshold;
    auto* t = new OrcByteTrack();
    t->min_conf = min_conf;
    t->track_thresh = track_thresh;
    t->match_thresh = match_thresh;
    t->det_thresh = track_thresh;                                     // bytetrack.cpp:145
    t->max_time_lost = static_cast<int>(frame_rate / 30.0f * track_buffer);   // :141-142
    return t;
}
void orc_bytetrack_destroy(OrcByteTrack* t) { delete t; }
void orc_bytetrack_reset(OrcByteTrack* t) {                           // :157-164 (ID counter NOT reset)
    t->frame_count = 0; t->frame_id = 0; t->active.clear(); t->lost.clear();
}
int orc_bytetrack_update(OrcByteTrack* t, const float* dets, int n, float* out, int out_cap) {
    return t->update(dets, n, out, out_cap);
}
int orc_bytetrack_counts(const OrcByteTrack* t, int* n_active, int* n_lost) {
    *n_active = (int)t->active.size(); *n_lost = (int)t->lost.size();
    return t->frame_count;
}
int orc_bytetrack_dump(const OrcByteTrack* t, int which, float* out, int cap_rows) {
    const List& l = which == 0 ? t->active : t->lost;
    int k = 0;
    for (const auto& s : l) {
        if (k >= cap_rows) break;
        float* o = out + 78 * k++;
        o[0] = (float)s.id; o[1] = (float)s.state; o[2] = s.activated ? 1.0f : 0.0f;
        o[3] = (float)s.frame_id; o[4] = (float)s.start_frame; o[5] = (float)s.tracklet_len;
        std::memcpy(o + 6, s.mean, sizeof(s.mean));
        std::memcpy(o + 14, s.cov, sizeof(s.cov));
    }
    return k;
}
void orc_bytetrack_last_sizes(const OrcByteTrack* t, int* sizes8) { std::memcpy(sizes8, t->last_sizes, sizeof(t->last_sizes)); }

}  // extern "C"
