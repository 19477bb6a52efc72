// trackers.hpp - C++17 binding of the C ABI (include/motb200.h) with motcpp's own class surface:
//   motcpp_b200::ByteTrack(det_thresh, max_age, ..., frame_rate).update(dets, img[, embs]) / reset()
// Constructor arguments, defaults, return layout and exceptions follow
// include/motcpp/trackers/bytetrack.hpp:97-110, include/motcpp/tracker.hpp:47-74 and
// src/tracker.cpp:108-125.  Header-only; link with -lmotb200.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../motb200.h"
#include "compat.hpp"

namespace motcpp_b200 {

inline void throw_on(int rc) {
    if (rc == MOT_OK) return;
    const std::string msg = mot_last_error();
    if (rc == MOT_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}

// Base: one engine stream behind BaseTracker's interface
class BaseTracker {
public:
    virtual ~BaseTracker() { mot_engine_destroy(engine_); }
    BaseTracker(const BaseTracker&) = delete;
    BaseTracker& operator=(const BaseTracker&) = delete;

    // dets (N,6) [x1,y1,x2,y2,conf,cls] -> (M,8) [x1,y1,x2,y2,id,conf,cls,det_ind]
    virtual Eigen::MatrixXf update(const Eigen::MatrixXf& dets, const cv::Mat& img,
                                   const Eigen::MatrixXf& embs = Eigen::MatrixXf()) {
        check_inputs(dets, img, embs);
        const int n = static_cast<int>(dets.rows());
        if (n > max_dets_) throw std::invalid_argument("more detections than max_dets");
        for (int i = 0; i < n; ++i)                       // Eigen is column-major, the ABI row-major
            for (int c = 0; c < 6; ++c) dets_rm_[static_cast<size_t>(i) * 6 + c] = dets(i, c);
        int n_out = 0;
        throw_on(mot_engine_update_host(engine_, 1, dets_rm_.data(), &n, max_dets_, out_rm_.data(), &n_out, cap_));
        throw_on(mot_engine_check(engine_, nullptr));
        Eigen::MatrixXf out(n_out, 8);
        for (int i = 0; i < n_out; ++i)
            for (int c = 0; c < 8; ++c) out(i, c) = out_rm_[static_cast<size_t>(i) * 8 + c];
        return out;
    }
    virtual void reset() { throw_on(mot_engine_reset(engine_)); }

    // src/tracker.cpp:108-125
    void check_inputs(const Eigen::MatrixXf& dets, const cv::Mat& img, const Eigen::MatrixXf& embs) const {
        if (dets.rows() > 0 && dets.cols() != 6 && dets.cols() != 7)
            throw std::invalid_argument("Detections must have 6 (AABB) or 7 (OBB) columns");
        if (img.empty()) throw std::invalid_argument("Image cannot be empty");
        if (embs.rows() > 0 && dets.rows() != embs.rows())
            throw std::invalid_argument("Detections and embeddings must have same number of rows");
        if (dets.rows() > 0 && dets.cols() == 7)
            throw std::invalid_argument("OBB detections are outside the accelerated hot path");
    }

protected:
    explicit BaseTracker(const mot_engine_config& cfg) {
        throw_on(mot_engine_create(&cfg, &engine_));
        int threads = 0, smem = 0, ctas = 0, bytes = 0;
        mot_engine_info(engine_, &threads, &smem, &ctas, &bytes);
        max_dets_ = cfg.max_dets > 0 ? cfg.max_dets : 512;
        cap_ = cfg.track_capacity > 0 ? cfg.track_capacity : 1536;
        dets_rm_.resize(static_cast<size_t>(max_dets_) * 6);
        out_rm_.resize(static_cast<size_t>(cap_) * 8);
    }
    mot_engine* engine_ = nullptr;
    int max_dets_ = 512, cap_ = 1536;
    std::vector<float> dets_rm_, out_rm_;
};

class ByteTrack : public BaseTracker {
public:
    ByteTrack(float det_thresh = 0.3f, int max_age = 30, int max_obs = 50, int min_hits = 3,
              float iou_threshold = 0.3f, bool per_class = false, int nr_classes = 80,
              const std::string& asso_func = "iou", bool is_obb = false, float min_conf = 0.1f,
              float track_thresh = 0.45f, float match_thresh = 0.8f, int track_buffer = 25, int frame_rate = 30,
              int track_capacity = 0, int max_dets = 0, int device = 0)
        : BaseTracker(make(det_thresh, max_age, max_obs, min_hits, iou_threshold, per_class, nr_classes, asso_func,
                           is_obb, min_conf, track_thresh, match_thresh, track_buffer, frame_rate, track_capacity,
                           max_dets, device)) {}

private:
    static mot_engine_config make(float det_thresh, int max_age, int max_obs, int min_hits, float iou_threshold,
                                  bool per_class, int /*nr_classes*/, const std::string& asso_func, bool is_obb,
                                  float min_conf, float track_thresh, float match_thresh, int track_buffer,
                                  int frame_rate, int track_capacity, int max_dets, int device) {
        if (asso_func != "iou") throw std::invalid_argument("Invalid association mode: " + asso_func);   // iou.hpp:407
        if (per_class || is_obb) throw std::invalid_argument("per_class / OBB are outside the accelerated hot path");
        mot_engine_config c;
        throw_on(mot_engine_default_config(MOT_TRACKER_BYTETRACK, &c));
        c.n_streams = 1; c.track_capacity = track_capacity; c.max_dets = max_dets; c.device = device;
        c.det_thresh = det_thresh; c.max_age = max_age; c.max_obs = max_obs; c.min_hits = min_hits;
        c.iou_threshold = iou_threshold; c.min_conf = min_conf; c.track_thresh = track_thresh;
        c.match_thresh = match_thresh; c.track_buffer = track_buffer; c.frame_rate = frame_rate;
        return c;
    }
};

}  // namespace motcpp_b200
