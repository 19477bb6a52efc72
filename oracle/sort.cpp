// TEST INFRASTRUCTURE (oracle): restatement of SORT's per-frame state machine.
//   SortTrack   src/trackers/sort.cpp:16-82
//   Sort::update src/trackers/sort.cpp:102-255
// ID counter is per tracker instance (reference: process-global static, sort.cpp:16-19).
#include "oracle.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace {

struct SortTrack {
    int id = 0;
    float conf = 0.0f;
    int cls = 0, det_ind = -1;
    int hits = 1, time_since_update = 0, age = 1;
    float x[7];
    float P[49];

    // sort.cpp:21-41: det row = [x1,y1,x2,y2,conf,cls,det_ind]
    SortTrack(const float* det7, int new_id) : id(new_id), conf(det7[4]), cls((int)det7[5]), det_ind((int)det7[6]) {
        float z[4];
        orc_xyxy2xysr(det7, z);
        orc_kf_xysr_init(z, x, P);
    }
    void predict() {                                   // sort.cpp:43-51
        orc_kf_xysr_predict(x, P, 1.0f, 1.0f);
        ++age;
        ++time_since_update;
    }
    void update(const float* det7) {                   // sort.cpp:53-70
        conf = det7[4]; cls = (int)det7[5]; det_ind = (int)det7[6];
        float z[4];
        orc_xyxy2xysr(det7, z);
        orc_kf_xysr_update(x, P, z);
        ++hits;
        time_since_update = 0;
    }
    void state(float* out) const { orc_xysr2xyxy(x, out); }   // sort.cpp:72-76
};

}  // namespace

struct OrcSort {
    float det_thresh, iou_threshold;
    int max_age, min_hits;
    int frame_count = 0;
    int id_counter = 0;
    std::vector<SortTrack> trackers;
};

extern "C" {

OrcSort* orc_sort_create(float det_thresh, int max_age, int max_obs, int min_hits, float iou_threshold) {
    (void)max_obs;
    auto* s = new OrcSort();
    s->det_thresh = det_thresh; s->iou_threshold = iou_threshold; s->max_age = max_age; s->min_hits = min_hits;
    return s;
}
void orc_sort_destroy(OrcSort* s) { delete s; }
void orc_sort_reset(OrcSort* s) { s->trackers.clear(); s->frame_count = 0; }   // sort.cpp:97-100

int orc_sort_update(OrcSort* s, const float* dets, int n, float* out, int out_cap) {
    ++s->frame_count;                                               // :108
    std::vector<int> valid;                                         // :111-121
    for (int i = 0; i < n; ++i)
        if (dets[6 * i + 4] >= s->det_thresh) valid.push_back(i);
    const int m = (int)valid.size();
    std::vector<float> det_boxes((size_t)m * 4);
    for (int j = 0; j < m; ++j) std::memcpy(&det_boxes[4 * j], dets + 6 * valid[j], 4 * sizeof(float));

    // :124-150 predict, drop tracks whose predicted box has a NaN
    std::vector<SortTrack> alive;
    std::vector<float> trk_boxes;
    for (auto& t : s->trackers) {
        t.predict();
        float b[4];
        t.state(b);
        if (std::isnan(b[0] + b[1] + b[2] + b[3])) continue;
        alive.push_back(t);
        trk_boxes.insert(trk_boxes.end(), b, b + 4);
    }
    s->trackers.swap(alive);
    const int nt = (int)s->trackers.size();

    std::vector<int> row2col(nt, -1), col2row(m, -1);               // :152-181
    if (nt > 0 && m > 0) {
        std::vector<float> cost((size_t)nt * m);
        orc_iou_distance(trk_boxes.data(), nt, det_boxes.data(), m, cost.data());
        orc_linear_assignment(cost.data(), nt, m, m, 1.0f - s->iou_threshold, row2col.data(), col2row.data());
    }
    for (int i = 0; i < nt; ++i) {                                  // :184-193
        const int j = row2col[i];
        if (j < 0) continue;
        float row[7];
        std::memcpy(row, dets + 6 * valid[j], 6 * sizeof(float));
        row[6] = (float)valid[j];
        s->trackers[i].update(row);
    }
    for (int j = 0; j < m; ++j) {                                   // :196-204 spawn in unmatched_dets order
        if (col2row[j] >= 0) continue;
        float row[7];
        std::memcpy(row, dets + 6 * valid[j], 6 * sizeof(float));
        row[6] = (float)valid[j];
        s->trackers.emplace_back(row, ++s->id_counter);
    }
    std::vector<SortTrack> keep;                                    // :207-216
    for (const auto& t : s->trackers)
        if (t.time_since_update <= s->max_age) keep.push_back(t);
    s->trackers.swap(keep);

    int rows = 0;                                                   // :219-243
    for (const auto& t : s->trackers)
        if (t.time_since_update == 0 && (t.hits >= s->min_hits || s->frame_count <= s->min_hits)) ++rows;
    if (rows > out_cap) return -rows;
    int k = 0;
    for (const auto& t : s->trackers) {
        if (!(t.time_since_update == 0 && (t.hits >= s->min_hits || s->frame_count <= s->min_hits))) continue;
        float* o = out + 8 * k++;
        t.state(o);
        o[4] = (float)t.id; o[5] = t.conf; o[6] = (float)t.cls; o[7] = (float)t.det_ind;
    }
    return rows;
}

}  // extern "C"
