// trackers.hpp - C++17 binding of the C ABI (include/motb200.h) with motcpp's own class surface:
//   motcpp_b200::{Sort, ByteTrack, OCSort, BotSort, StrongSORT, DeepOCSort, BoostTrackTracker}(<the reference's positional
//   ctor arguments>)
//       .update(dets, img[, embs]) / .reset()
// Constructor arguments, defaults, return layout and exceptions follow include/motcpp/trackers/sort.hpp:69-77,
// bytetrack.hpp:97-110, ocsort.hpp:88-102, botsort.hpp:108-134, include/motcpp/tracker.hpp:47-74 and
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
        if (validate_) check_inputs(dets, img, embs);
        const int n = static_cast<int>(dets.rows());
        if (n > max_dets_) throw std::invalid_argument("more detections than max_dets");
        for (int i = 0; i < n; ++i)                       // Eigen is column-major, the ABI row-major
            for (int c = 0; c < 6; ++c) dets_rm_[static_cast<size_t>(i) * 6 + c] = dets(i, c);
        const float* embs_ptr = nullptr;
        // The reference uses whatever `embs` it is handed; here the feature dimension is fixed at construction, so a
        // mismatch is an error - never a silent fall-back to IoU-only association.
        if (embs.rows() > 0 && embs.cols() > 0 && n > 0 && (embs.rows() != dets.rows() || embs.cols() != emb_dim_))
            throw std::invalid_argument("embeddings are (" + std::to_string(embs.rows()) + ", " + std::to_string(embs.cols()) +
                                        ") but the tracker was constructed for emb_dim = " + std::to_string(emb_dim_) +
                                        " and this frame has " + std::to_string(n) + " detections");
        if (emb_dim_ > 0 && embs.rows() == dets.rows() && embs.cols() == emb_dim_ && n > 0) {
            for (int i = 0; i < n; ++i)
                for (int c = 0; c < emb_dim_; ++c) embs_rm_[static_cast<size_t>(i) * emb_dim_ + c] = embs(i, c);
            embs_ptr = embs_rm_.data();
        }
        if (need_embs_) {
            // DeepOC-SORT with embeddings on: the reference would run its ReID network on a frame without `embs`
            if (n > 0 && !embs_ptr)
                throw std::invalid_argument("ReID inference is outside the accelerated hot path: pass embeddings to update()");
            embs_ptr = embs_rm_.data();
        }
        int n_out = 0;
        throw_on(mot_engine_update_host_embs(engine_, 1, dets_rm_.data(), &n, max_dets_, embs_ptr, out_rm_.data(), &n_out,
                                             cap_));
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
        emb_dim_ = (cfg.kind == MOT_TRACKER_BOTSORT || cfg.kind == MOT_TRACKER_STRONGSORT) ? cfg.emb_dim : 0;
        if (cfg.kind == MOT_TRACKER_DEEPOCSORT && !cfg.embedding_off) { emb_dim_ = cfg.emb_dim; need_embs_ = true; }
        dets_rm_.resize(static_cast<size_t>(max_dets_) * 6);
        out_rm_.resize(static_cast<size_t>(cap_) * 8);
        embs_rm_.resize(static_cast<size_t>(max_dets_) * static_cast<size_t>(emb_dim_));
    }
    static void only_iou_aabb(const std::string& asso_func, bool per_class, bool is_obb) {
        if (asso_func != "iou") throw std::invalid_argument("Invalid association mode: " + asso_func);   // iou.hpp:407
        if (per_class || is_obb) throw std::invalid_argument("per_class / OBB are outside the accelerated hot path");
    }
    mot_engine* engine_ = nullptr;
    int max_dets_ = 512, cap_ = 1536, emb_dim_ = 0;
    bool validate_ = true;
    bool need_embs_ = false;
    std::vector<float> dets_rm_, out_rm_, embs_rm_;
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
        only_iou_aabb(asso_func, per_class, is_obb);
        mot_engine_config c;
        throw_on(mot_engine_default_config(MOT_TRACKER_BYTETRACK, &c));
        c.n_streams = 1; c.track_capacity = track_capacity; c.max_dets = max_dets; c.device = device;
        c.det_thresh = det_thresh; c.max_age = max_age; c.max_obs = max_obs; c.min_hits = min_hits;
        c.iou_threshold = iou_threshold; c.min_conf = min_conf; c.track_thresh = track_thresh;
        c.match_thresh = match_thresh; c.track_buffer = track_buffer; c.frame_rate = frame_rate;
        return c;
    }
};

// motcpp::trackers::Sort (include/motcpp/trackers/sort.hpp:69-77).  Like the reference it never validates its
// input (Sort::update does not call check_inputs, src/trackers/sort.cpp:102-108).
class Sort : public BaseTracker {
public:
    Sort(float det_thresh = 0.3f, int max_age = 1, int max_obs = 50, int min_hits = 3, float iou_threshold = 0.3f,
         bool per_class = false, int nr_classes = 80, const std::string& asso_func = "iou", bool is_obb = false,
         int track_capacity = 256, int max_dets = 64, int device = 0)
        : BaseTracker(make(det_thresh, max_age, max_obs, min_hits, iou_threshold, per_class, nr_classes, asso_func, is_obb,
                           track_capacity, max_dets, device)) { validate_ = false; }

private:
    static mot_engine_config make(float det_thresh, int max_age, int max_obs, int min_hits, float iou_threshold,
                                  bool per_class, int /*nr_classes*/, const std::string& asso_func, bool is_obb,
                                  int track_capacity, int max_dets, int device) {
        only_iou_aabb(asso_func, per_class, is_obb);
        mot_engine_config c;
        throw_on(mot_engine_default_config(MOT_TRACKER_SORT, &c));
        c.n_streams = 1; c.track_capacity = track_capacity; c.max_dets = max_dets; c.device = device;
        c.det_thresh = det_thresh; c.max_age = max_age; c.max_obs = max_obs; c.min_hits = min_hits;
        c.iou_threshold = iou_threshold;
        return c;
    }
};

// motcpp::trackers::OCSort (include/motcpp/trackers/ocsort.hpp:88-102)
// asso_func: "iou" or "centroid" (AssociationFunction, iou.hpp:371-411; the reference's other variants are only defined
// for one-row box sets).  "centroid" normalises by the frame diagonal, which the reference reads from every img
// (ocsort.cpp:413): pass the frame size as frame_width / frame_height; update() checks img against it.
class OCSort : public BaseTracker {
public:
    OCSort(float det_thresh = 0.2f, int max_age = 30, int max_obs = 50, int min_hits = 3, float iou_threshold = 0.3f,
           bool per_class = false, int nr_classes = 80, const std::string& asso_func = "iou", bool is_obb = false,
           float min_conf = 0.1f, int delta_t = 3, float inertia = 0.2f, bool use_byte = false,
           float Q_xy_scaling = 0.01f, float Q_s_scaling = 0.0001f, int track_capacity = 0, int max_dets = 0,
           int device = 0, int frame_width = 0, int frame_height = 0)
        : BaseTracker(make(det_thresh, max_age, max_obs, min_hits, iou_threshold, per_class, nr_classes, asso_func, is_obb,
                           min_conf, delta_t, inertia, use_byte, Q_xy_scaling, Q_s_scaling, track_capacity, max_dets,
                           device, frame_width, frame_height)),
          centroid_(asso_func == "centroid"), frame_w_(frame_width), frame_h_(frame_height) {}

    Eigen::MatrixXf update(const Eigen::MatrixXf& dets, const cv::Mat& img,
                           const Eigen::MatrixXf& embs = Eigen::MatrixXf()) override {
        if (centroid_ && !img.empty() && (img.cols != frame_w_ || img.rows != frame_h_))
            throw std::invalid_argument("asso_func \"centroid\": the frame size differs from the frame_width / frame_height the tracker was built for");
        return BaseTracker::update(dets, img, embs);
    }

private:
    bool centroid_;
    int frame_w_, frame_h_;
    static mot_engine_config make(float det_thresh, int max_age, int max_obs, int min_hits, float iou_threshold,
                                  bool per_class, int /*nr_classes*/, const std::string& asso_func, bool is_obb,
                                  float min_conf, int delta_t, float inertia, bool use_byte, float q_xy, float q_s,
                                  int track_capacity, int max_dets, int device, int frame_width, int frame_height) {
        if (asso_func != "iou" && asso_func != "centroid") throw std::invalid_argument("Invalid association mode: " + asso_func);   // iou.hpp:407
        if (per_class || is_obb) throw std::invalid_argument("per_class / OBB are outside the accelerated hot path");
        mot_engine_config c;
        throw_on(mot_engine_default_config(MOT_TRACKER_OCSORT, &c));
        c.n_streams = 1; c.track_capacity = track_capacity; c.max_dets = max_dets; c.device = device;
        c.det_thresh = det_thresh; c.max_age = max_age; c.max_obs = max_obs; c.min_hits = min_hits;
        c.iou_threshold = iou_threshold; c.min_conf = min_conf; c.delta_t = delta_t; c.inertia = inertia;
        c.use_byte = use_byte ? 1 : 0; c.q_xy_scaling = q_xy; c.q_s_scaling = q_s;
        if (asso_func == "centroid") { c.asso_func = 6; c.frame_width = frame_width; c.frame_height = frame_height; }
        return c;
    }
};

// motcpp::trackers::BotSort (include/motcpp/trackers/botsort.hpp:108-134).  Camera-motion compensation and ReID
// inference are image processing outside the association hot path: cmc_method must be "none", reid_weights empty,
// and the embeddings are passed to update() - the reference's own `embs` argument (botsort.cpp:276-283).
class BotSort : public BaseTracker {
public:
    BotSort(const std::string& reid_weights = "", bool use_half = false, bool use_gpu = false, float det_thresh = 0.3f,
            int max_age = 30, int max_obs = 50, int min_hits = 3, float iou_threshold = 0.3f, bool per_class = false,
            int nr_classes = 80, const std::string& asso_func = "iou", bool is_obb = false,
            float track_high_thresh = 0.5f, float track_low_thresh = 0.1f, float new_track_thresh = 0.6f,
            int track_buffer = 30, float match_thresh = 0.8f, float proximity_thresh = 0.5f,
            float appearance_thresh = 0.25f, const std::string& cmc_method = "none", int frame_rate = 30,
            bool fuse_first_associate = false, bool with_reid = true, int emb_dim = 0, int track_capacity = 0,
            int max_dets = 0, int device = 0)
        : BaseTracker(make(reid_weights, use_half, use_gpu, det_thresh, max_age, max_obs, min_hits, iou_threshold, per_class,
                           asso_func, is_obb, track_high_thresh, track_low_thresh, new_track_thresh, track_buffer,
                           match_thresh, proximity_thresh, appearance_thresh, cmc_method, frame_rate, fuse_first_associate,
                           with_reid, emb_dim, track_capacity, max_dets, device)) {}

    // BotSort::update validates (dets, img) only and returns at once on an empty frame (botsort.cpp:265-269)
    Eigen::MatrixXf update(const Eigen::MatrixXf& dets, const cv::Mat& img,
                           const Eigen::MatrixXf& embs = Eigen::MatrixXf()) override {
        check_inputs(dets, img, Eigen::MatrixXf());
        if (dets.rows() == 0) return Eigen::MatrixXf(0, 8);
        if (embs.rows() > 0 && (embs.rows() != dets.rows() || embs.cols() != emb_dim_))
            throw std::invalid_argument("Detections and embeddings must have same number of rows");
        validate_ = false;
        return BaseTracker::update(dets, img, embs);
    }

private:
    static mot_engine_config make(const std::string& reid_weights, bool, bool, float det_thresh, int max_age, int max_obs,
                                  int min_hits, float iou_threshold, bool per_class, const std::string& asso_func,
                                  bool is_obb, float high, float low, float new_thresh, int track_buffer,
                                  float match_thresh, float prox, float app, const std::string& cmc_method,
                                  int frame_rate, bool fuse_first, bool with_reid, int emb_dim, int track_capacity,
                                  int max_dets, int device) {
        only_iou_aabb(asso_func, per_class, is_obb);
        if (!(cmc_method.empty() || cmc_method == "none"))
            throw std::invalid_argument("camera-motion compensation is outside the accelerated hot path (cmc_method must be \"none\")");
        if (!reid_weights.empty())
            throw std::invalid_argument("ReID inference is outside the accelerated hot path: pass embeddings to update()");
        mot_engine_config c;
        throw_on(mot_engine_default_config(MOT_TRACKER_BOTSORT, &c));
        c.n_streams = 1; c.track_capacity = track_capacity; c.max_dets = max_dets; c.device = device;
        c.det_thresh = det_thresh; c.max_age = max_age; c.max_obs = max_obs; c.min_hits = min_hits;
        c.iou_threshold = iou_threshold; c.track_high_thresh = high; c.track_low_thresh = low;
        c.new_track_thresh = new_thresh; c.track_buffer = track_buffer; c.match_thresh = match_thresh;
        c.proximity_thresh = prox; c.appearance_thresh = app; c.frame_rate = frame_rate;
        c.fuse_first_associate = fuse_first ? 1 : 0; c.with_reid = with_reid ? 1 : 0; c.emb_dim = emb_dim;
        return c;
    }
};

// motcpp::trackers::StrongSORT (include/motcpp/trackers/strongsort.hpp:287-305).  ReID inference and ECC camera-motion
// estimation are image processing outside the association hot path: reid_weights must be empty, the embeddings are passed
// to update() (the reference's own `embs` argument, strongsort.cpp:880-906) and the camera warp is the identity (what
// motion::ECC::apply yields on a static / featureless image).  nn_budget is the per-track gallery ring size (>= 1).
class StrongSORT : public BaseTracker {
public:
    StrongSORT(const std::string& reid_weights = "", bool use_half = false, bool use_gpu = false, float det_thresh = 0.3f,
               int max_age = 30, int max_obs = 50, int min_hits = 3, float iou_threshold = 0.3f, bool per_class = false,
               int nr_classes = 80, const std::string& asso_func = "iou", bool is_obb = false, float min_conf = 0.1f,
               float max_cos_dist = 0.2f, float max_iou_dist = 0.7f, int n_init = 3, int nn_budget = 100,
               float mc_lambda = 0.98f, float ema_alpha = 0.9f, int emb_dim = 0, int track_capacity = 0, int max_dets = 0,
               int device = 0)
        : BaseTracker(make(reid_weights, use_half, use_gpu, det_thresh, max_age, max_obs, min_hits, iou_threshold, per_class,
                           nr_classes, asso_func, is_obb, min_conf, max_cos_dist, max_iou_dist, n_init, nn_budget, mc_lambda,
                           ema_alpha, emb_dim, track_capacity, max_dets, device)) {}

    // an empty frame still advances the tracker: predict + every track missed (strongsort.cpp:833-837)
    Eigen::MatrixXf update(const Eigen::MatrixXf& dets, const cv::Mat& img,
                           const Eigen::MatrixXf& embs = Eigen::MatrixXf()) override {
        check_inputs(dets, img, embs);
        validate_ = false;
        return BaseTracker::update(dets, img, embs);
    }

private:
    static mot_engine_config make(const std::string& reid_weights, bool, bool, float det_thresh, int max_age, int max_obs,
                                  int min_hits, float iou_threshold, bool per_class, int, const std::string&, bool is_obb,
                                  float min_conf, float max_cos_dist, float max_iou_dist, int n_init, int nn_budget,
                                  float mc_lambda, float ema_alpha, int emb_dim, int track_capacity, int max_dets, int device) {
        if (per_class || is_obb) throw std::invalid_argument("per_class / OBB are outside the accelerated hot path");
        if (!reid_weights.empty())
            throw std::invalid_argument("ReID inference is outside the accelerated hot path: pass embeddings to update()");
        mot_engine_config c;
        throw_on(mot_engine_default_config(MOT_TRACKER_STRONGSORT, &c));
        c.n_streams = 1; c.track_capacity = track_capacity; c.max_dets = max_dets; c.device = device;
        c.det_thresh = det_thresh; c.max_age = max_age; c.max_obs = max_obs; c.min_hits = min_hits;
        c.iou_threshold = iou_threshold; c.min_conf = min_conf; c.max_cos_dist = max_cos_dist; c.max_iou_dist = max_iou_dist;
        c.n_init = n_init; c.nn_budget = nn_budget; c.mc_lambda = mc_lambda; c.ema_alpha = ema_alpha; c.emb_dim = emb_dim;
        return c;
    }
};

// motcpp::trackers::DeepOCSort (include/motcpp/trackers/deepocsort.hpp:93-117).  ReID inference and camera-motion
// compensation are image processing outside the association hot path: reid_weights must be empty (unless embedding_off),
// cmc_off must be true, and the embeddings are passed to update() (the reference's own `embs` argument,
// deepocsort.cpp:629-633); emb_dim fixes their width at construction.  The reference lists unmatched detections and
// tracks twice (:476-481, :485-501): size track_capacity for 2 x the tracks that can be unmatched in one frame.
class DeepOCSort : public BaseTracker {
public:
    DeepOCSort(const std::string& reid_weights = "", bool use_half = false, bool use_gpu = false, float det_thresh = 0.3f,
               int max_age = 30, int max_obs = 50, int min_hits = 3, float iou_threshold = 0.3f, bool per_class = false,
               int nr_classes = 80, const std::string& asso_func = "iou", bool is_obb = false, int delta_t = 3,
               float inertia = 0.2f, float w_association_emb = 0.5f, float alpha_fixed_emb = 0.95f, float aw_param = 0.5f,
               bool embedding_off = false, bool cmc_off = true, bool aw_off = false, float Q_xy_scaling = 0.01f,
               float Q_s_scaling = 0.0001f, int emb_dim = 0, int track_capacity = 0, int max_dets = 0, int device = 0)
        : BaseTracker(make(reid_weights, use_half, use_gpu, det_thresh, max_age, max_obs, min_hits, iou_threshold, per_class,
                           nr_classes, asso_func, is_obb, delta_t, inertia, w_association_emb, alpha_fixed_emb, aw_param,
                           embedding_off, cmc_off, aw_off, Q_xy_scaling, Q_s_scaling, emb_dim, track_capacity, max_dets,
                           device)) {}

private:
    static mot_engine_config make(const std::string& reid_weights, bool, bool, float det_thresh, int max_age, int max_obs,
                                  int min_hits, float iou_threshold, bool per_class, int, const std::string& asso_func,
                                  bool is_obb, int delta_t, float inertia, float w_assoc, float alpha_fixed, float aw_param,
                                  bool embedding_off, bool cmc_off, bool aw_off, float q_xy, float q_s, int emb_dim,
                                  int track_capacity, int max_dets, int device) {
        only_iou_aabb(asso_func, per_class, is_obb);
        if (!cmc_off) throw std::invalid_argument("camera-motion compensation is outside the accelerated hot path (cmc_off must be true)");
        if (!reid_weights.empty() && !embedding_off)
            throw std::invalid_argument("ReID inference is outside the accelerated hot path: pass embeddings to update()");
        mot_engine_config c;
        throw_on(mot_engine_default_config(MOT_TRACKER_DEEPOCSORT, &c));
        c.n_streams = 1; c.track_capacity = track_capacity; c.max_dets = max_dets; c.device = device;
        c.det_thresh = det_thresh; c.max_age = max_age; c.max_obs = max_obs; c.min_hits = min_hits;
        c.iou_threshold = iou_threshold; c.delta_t = delta_t; c.inertia = inertia; c.w_association_emb = w_assoc;
        c.alpha_fixed_emb = alpha_fixed; c.aw_param = aw_param; c.embedding_off = embedding_off ? 1 : 0;
        c.aw_off = aw_off ? 1 : 0; c.q_xy_scaling = q_xy; c.q_s_scaling = q_s; c.emb_dim = embedding_off ? 0 : emb_dim;
        return c;
    }
};

// motcpp::trackers::BoostTrackTracker (include/motcpp/trackers/boosttrack.hpp:95-124).  ECC camera-motion compensation and
// ReID are image processing outside the association hot path: use_ecc must be false (the reference's default is true; its
// ECC yields the identity warp on a static image) and with_reid false; use_sb (a std::pow in the confidence boost,
// boosttrack.cpp:395-412) is not built.  use_duo_boost, s_sim_corr and use_rich_s are accepted and have no effect, as in
// the reference (duo_confidence_boost returns its input, boosttrack.cpp:428-432; the other two are never read).
class BoostTrackTracker : public BaseTracker {
public:
    BoostTrackTracker(const std::string& reid_weights = "", bool use_half = false, bool use_gpu = false, float det_thresh = 0.6f,
                      int max_age = 60, int max_obs = 50, int min_hits = 3, float iou_threshold = 0.3f, bool per_class = false,
                      int nr_classes = 80, const std::string& asso_func = "iou", bool is_obb = false, bool use_ecc = false,
                      int min_box_area = 10, float aspect_ratio_thresh = 1.6f, const std::string& cmc_method = "ecc",
                      float lambda_iou = 0.5f, float lambda_mhd = 0.25f, float lambda_shape = 0.25f, bool use_dlo_boost = true,
                      bool use_duo_boost = true, float dlo_boost_coef = 0.65f, bool s_sim_corr = false, bool use_rich_s = false,
                      bool use_sb = false, bool use_vt = false, bool with_reid = false, int track_capacity = 0,
                      int max_dets = 0, int device = 0)
        : BaseTracker(make(reid_weights, use_half, use_gpu, det_thresh, max_age, max_obs, min_hits, iou_threshold, per_class,
                           nr_classes, asso_func, is_obb, use_ecc, min_box_area, aspect_ratio_thresh, cmc_method, lambda_iou,
                           lambda_mhd, lambda_shape, use_dlo_boost, use_duo_boost, dlo_boost_coef, s_sim_corr, use_rich_s, use_sb,
                           use_vt, with_reid, track_capacity, max_dets, device)) {}

private:
    static mot_engine_config make(const std::string&, bool, bool, float det_thresh, int max_age, int max_obs, int min_hits,
                                  float iou_threshold, bool per_class, int, const std::string& /*asso_func: stored, never read (src/tracker.cpp:27)*/, bool is_obb,
                                  bool use_ecc, int min_box_area, float aspect_ratio_thresh, const std::string& cmc_method,
                                  float lambda_iou, float lambda_mhd, float lambda_shape, bool use_dlo_boost, bool,
                                  float dlo_boost_coef, bool, bool, bool use_sb, bool use_vt, bool with_reid,
                                  int track_capacity, int max_dets, int device) {
        if (per_class || is_obb) throw std::invalid_argument("per_class / OBB are outside the accelerated hot path");
        if (use_ecc && cmc_method == "ecc")
            throw std::invalid_argument("camera-motion compensation is outside the accelerated hot path (use_ecc must be false)");
        if (with_reid) throw std::invalid_argument("BoostTrack with ReID is outside the accelerated hot path (with_reid must be false)");
        if (use_sb) throw std::invalid_argument("use_sb (std::pow in the confidence boost) is not built");
        mot_engine_config c;
        throw_on(mot_engine_default_config(MOT_TRACKER_BOOSTTRACK, &c));
        c.n_streams = 1; c.track_capacity = track_capacity; c.max_dets = max_dets; c.device = device;
        c.det_thresh = det_thresh; c.max_age = max_age; c.max_obs = max_obs; c.min_hits = min_hits;
        c.iou_threshold = iou_threshold; c.min_box_area = min_box_area; c.aspect_ratio_thresh = aspect_ratio_thresh;
        c.lambda_iou = lambda_iou; c.lambda_mhd = lambda_mhd; c.lambda_shape = lambda_shape;
        c.use_dlo_boost = use_dlo_boost ? 1 : 0; c.dlo_boost_coef = dlo_boost_coef; c.use_sb = 0; c.use_vt = use_vt ? 1 : 0;
        return c;
    }
};

}  // namespace motcpp_b200
