// TEST INFRASTRUCTURE.  C entry points around the reference's OWN implementation of the association hot
// path, compiled IN PLACE from /root/reference (never vendored; oracle/Makefile target `ref`) against the
// stand-in <Eigen/Dense> and <opencv2/...> headers of this directory:
//
//   src/motion/kalman_filter.cpp, kalman_filters/xyah_kf.cpp, kalman_filters/xysr_kf.cpp,
//   include/motcpp/motion/kalman_filters/xywh_kf.hpp, include/motcpp/utils/{iou,ops,matching}.hpp,
//   src/utils/matching.cpp, include/motcpp/association/lap_solver.hpp, src/tracker.cpp and the state machines
//   src/trackers/{sort,bytetrack,ocsort,botsort,strongsort,deepocsort}.cpp.
//
// Everything below this comment only marshals row-major C arrays into the reference's Eigen types and calls
// the reference's functions; no algorithm is restated here.  The few definitions at the end satisfy the linker
// for classes that are OUTSIDE the path (ECC / SOF camera-motion compensation, the ONNX ReID backend):
// camera motion = identity (what the reference's ECC itself returns on its first frame / on failure),
// ReID inference unavailable (embeddings are passed in through the reference's own `embs` argument).
//
// tests/test_ref_pin.py compares oracle/liboracle.so against this library function by function and
// tracker by tracker.  Only tests/ may load it.
#include <Eigen/Dense>
#include <opencv2/opencv.hpp>

#include <motcpp/motion/kalman_filters/xyah_kf.hpp>
#include <motcpp/motion/kalman_filters/xysr_kf.hpp>
#include <motcpp/motion/kalman_filters/xywh_kf.hpp>
#include <motcpp/motion/cmc/ecc.hpp>
#include <motcpp/motion/cmc/sof.hpp>
#include <motcpp/appearance/onnx_backend.hpp>
#include <motcpp/tracker.hpp>
#include <motcpp/trackers/botsort.hpp>
#include <motcpp/trackers/bytetrack.hpp>
#include <motcpp/trackers/boosttrack.hpp>
#include <motcpp/trackers/deepocsort.hpp>
#include <motcpp/trackers/ocsort.hpp>
#include <motcpp/trackers/sort.hpp>
#include <motcpp/trackers/strongsort.hpp>
#include <motcpp/utils/iou.hpp>
#include <motcpp/utils/matching.hpp>
#include <motcpp/utils/ops.hpp>

#include <cstring>
#include <memory>
#include <string>

using Eigen::MatrixXf;
using Eigen::VectorXf;
using Eigen::Vector4f;

namespace motcpp::trackers::deepocsort_assoc {
// declared in src/trackers/deepocsort.cpp:267-292 (no header); defined there at :294
Eigen::MatrixXf compute_aw_max_metric(const Eigen::MatrixXf& emb_cost, float w_association_emb, float bottom);
}

namespace {

thread_local std::string g_err;

MatrixXf from_rows(const float* p, int n, int m, int ld = -1) {
    if (ld < 0) ld = m;
    MatrixXf a(n, m);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) a(i, j) = p[static_cast<size_t>(i) * ld + j];
    return a;
}
void to_rows(const Eigen::Ref<float>& a, float* p, int ld = -1) {
    if (ld < 0) ld = static_cast<int>(a.cols());
    for (int i = 0; i < a.rows(); ++i)
        for (int j = 0; j < a.cols(); ++j) p[static_cast<size_t>(i) * ld + j] = a(i, j);
}
VectorXf vec(const float* p, int n) {
    VectorXf v(n);
    for (int i = 0; i < n; ++i) v(i) = p[i];
    return v;
}
void put(const Eigen::Ref<float>& v, float* p) {
    for (int i = 0; i < v.size(); ++i) p[i] = v(i);
}

template <class F> int guarded(F f) {
    try {
        return f();
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1000;
    } catch (...) {
        g_err = "unknown exception";
        return -1000;
    }
}

// KalmanFilterXYSR with state injected / extracted (its members are public: xysr_kf.hpp:26-41)
motcpp::motion::KalmanFilterXYSR make_xysr(const float* x7, const float* P49, float q_xy, float q_s) {
    motcpp::motion::KalmanFilterXYSR kf(7, 4, 50);
    kf.x = vec(x7, 7);
    kf.P = from_rows(P49, 7, 7);
    kf.Q(4, 4) *= q_xy;          // what KalmanBoxTracker's ctor does, src/trackers/ocsort.cpp:77-79
    kf.Q(5, 5) *= q_xy;
    kf.Q(6, 6) *= q_s;
    return kf;
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }
// 1 = Eigen-like evaluation order, 0 = textbook (see ref_shim/Eigen/Dense)
int ref_order_mode() {
#ifdef REFSHIM_ORDER_TEXTBOOK
    return 0;
#else
    return 1;
#endif
}

// ---------------- include/motcpp/utils/ops.hpp --------------------------------------------------------
#define REF_OP4(name)                                                         \
    void ref_##name(const float* in, float* out) {                            \
        Vector4f v(in[0], in[1], in[2], in[3]);                               \
        put(motcpp::utils::name(v), out);                                     \
    }
REF_OP4(xyxy2xywh)
REF_OP4(xywh2xyxy)
REF_OP4(xywh2tlwh)
REF_OP4(tlwh2xywh)
REF_OP4(tlwh2xyxy)
REF_OP4(xyxy2tlwh)
REF_OP4(tlwh2xyah)
REF_OP4(xyah2tlwh)
REF_OP4(xywh2xyah)
REF_OP4(xyah2xywh)
REF_OP4(xyxy2xysr)
REF_OP4(xysr2xyxy)
#undef REF_OP4

// ---------------- KalmanFilterXYAH (src/motion/kalman_filter.cpp, kalman_filters/xyah_kf.cpp) ---------
void ref_kf_xyah_initiate(const float* z4, float* mean8, float* cov64) {
    motcpp::motion::KalmanFilterXYAH kf;
    auto [m, c] = kf.initiate(vec(z4, 4));
    put(m, mean8);
    to_rows(c, cov64);
}
void ref_kf_xyah_predict(float* mean8, float* cov64) {
    motcpp::motion::KalmanFilterXYAH kf;
    auto [m, c] = kf.predict(vec(mean8, 8), from_rows(cov64, 8, 8));
    put(m, mean8);
    to_rows(c, cov64);
}
void ref_kf_xyah_project(const float* mean8, const float* cov64, float conf, float* pm4, float* pc16) {
    motcpp::motion::KalmanFilterXYAH kf;
    auto [m, c] = kf.project(vec(mean8, 8), from_rows(cov64, 8, 8), conf);
    put(m, pm4);
    to_rows(c, pc16);
}
int ref_kf_xyah_update(float* mean8, float* cov64, const float* z4, float conf) {
    return guarded([&] {
        motcpp::motion::KalmanFilterXYAH kf;
        auto [m, c] = kf.update(vec(mean8, 8), from_rows(cov64, 8, 8), vec(z4, 4), conf);
        put(m, mean8);
        to_rows(c, cov64);
        return 0;
    });
}
int ref_kf_xyah_gating(const float* mean8, const float* cov64, const float* meas, int m, int only_position,
                       int metric, float* out) {
    return guarded([&] {
        motcpp::motion::KalmanFilterXYAH kf;
        VectorXf d = kf.gating_distance(vec(mean8, 8), from_rows(cov64, 8, 8), from_rows(meas, m, 4),
                                        only_position != 0, metric == 0 ? "maha" : "gaussian");
        put(d, out);
        return 0;
    });
}

// ---------------- KalmanFilterXYSR (src/motion/kalman_filters/xysr_kf.cpp) ----------------------------
// initial state of a SORT / OC-SORT track: x = [xysr, 0,0,0], P = the ctor's P (xysr_kf.cpp:40-43)
void ref_kf_xysr_init(const float* z4, float* x7, float* P49) {
    motcpp::motion::KalmanFilterXYSR kf(7, 4, 50);
    for (int i = 0; i < 7; ++i) x7[i] = i < 4 ? z4[i] : 0.0f;
    to_rows(kf.P, P49);
}
void ref_kf_xysr_predict(float* x7, float* P49, float q_xy_scale, float q_s_scale) {
    auto kf = make_xysr(x7, P49, q_xy_scale, q_s_scale);
    kf.predict();
    put(kf.x, x7);
    to_rows(kf.P, P49);
}
int ref_kf_xysr_update(float* x7, float* P49, const float* z4) {
    return guarded([&] {
        auto kf = make_xysr(x7, P49, 1.0f, 1.0f);
        kf.update(vec(z4, 4));
        put(kf.x, x7);
        to_rows(kf.P, P49);
        return 0;
    });
}
void ref_kf_xysr_affine(float* x7, float* P49, const float* m2, const float* t2) {
    auto kf = make_xysr(x7, P49, 1.0f, 1.0f);
    Eigen::Matrix2f m;
    m(0, 0) = m2[0]; m(0, 1) = m2[1]; m(1, 0) = m2[2]; m(1, 1) = m2[3];
    Eigen::Vector2f t(t2[0], t2[1]);
    kf.apply_affine_correction(m, t);
    put(kf.x, x7);
    to_rows(kf.P, P49);
}

// ---------------- KalmanFilterXYWH (include/motcpp/motion/kalman_filters/xywh_kf.hpp) ------------------
void ref_kf_xywh_initiate(const float* z4, float* mean8, float* cov64) {
    motcpp::KalmanFilterXYWH kf;
    auto [m, c] = kf.initiate(Vector4f(z4[0], z4[1], z4[2], z4[3]));
    put(m, mean8);
    to_rows(c, cov64);
}
void ref_kf_xywh_predict(float* mean8, float* cov64) {
    motcpp::KalmanFilterXYWH kf;
    auto [m, c] = kf.predict(vec(mean8, 8), from_rows(cov64, 8, 8));
    put(m, mean8);
    to_rows(c, cov64);
}
void ref_kf_xywh_update(float* mean8, float* cov64, const float* z4) {
    motcpp::KalmanFilterXYWH kf;
    auto [m, c] = kf.update(vec(mean8, 8), from_rows(cov64, 8, 8), Vector4f(z4[0], z4[1], z4[2], z4[3]));
    put(m, mean8);
    to_rows(c, cov64);
}
void ref_kf_xywh_gating(const float* mean8, const float* cov64, const float* meas, int m, int only_position,
                        float* out) {
    motcpp::KalmanFilterXYWH kf;
    VectorXf d = kf.gating_distance(vec(mean8, 8), from_rows(cov64, 8, 8), from_rows(meas, m, 4), only_position != 0);
    put(d, out);
}

// ---------------- cost build (include/motcpp/utils/iou.hpp, src/utils/matching.cpp) --------------------
void ref_iou_batch(const float* a, int n, const float* b, int m, float* out) {
    to_rows(motcpp::utils::iou_batch(from_rows(a, n, 4), from_rows(b, m, 4)), out);
}
void ref_iou_distance(const float* a, int n, const float* b, int m, float* out) {
    to_rows(motcpp::utils::iou_distance(from_rows(a, n, 4), from_rows(b, m, 4)), out);
}
void ref_fuse_score(float* cost, int n, int m, const float* det_conf) {
    to_rows(motcpp::utils::fuse_score(from_rows(cost, n, m), vec(det_conf, m)), cost);
}
int ref_embedding_distance(const float* t, int n, const float* d, int m, int dim, int metric, float* out) {
    return guarded([&] {
        to_rows(motcpp::utils::embedding_distance(from_rows(t, n, dim), from_rows(d, m, dim),
                                                  metric == 0 ? "cosine" : "euclidean"), out);
        return 0;
    });
}
// AssociationFunction by name ("iou", "hmiou", "giou", "ciou", "diou", "centroid"), iou.hpp:371-411.
// The non-"iou" variants are only shape-consistent when the second set has ONE row... and the first too
// (SURVEY trap 11); the stand-in Eigen throws on the mismatched shapes a release Eigen build would read past.
int ref_asso_func(const char* mode, const float* a, int n, const float* b, int m, int w, int h, float* out) {
    return guarded([&] {
        motcpp::utils::AssociationFunction f(w, h, mode);
        to_rows(f(from_rows(a, n, 4), from_rows(b, m, 4)), out);
        return 0;
    });
}
int ref_linear_assignment(const float* cost, int n, int m, int ld, float thresh, int* row2col, int* col2row) {
    return guarded([&] {
        for (int i = 0; i < n; ++i) row2col[i] = -1;
        for (int j = 0; j < m; ++j) col2row[j] = -1;
        auto r = motcpp::utils::linear_assignment(from_rows(cost, n, m, ld), thresh);
        for (const auto& mm : r.matches) {
            row2col[mm[0]] = mm[1];
            col2row[mm[1]] = mm[0];
        }
        return static_cast<int>(r.matches.size());
    });
}
int ref_aw_max_metric(const float* emb, int n, int m, int ld, float w_assoc, float bottom, float* out, int ld_out) {
    return guarded([&] {
        to_rows(motcpp::trackers::deepocsort_assoc::compute_aw_max_metric(from_rows(emb, n, m, ld), w_assoc, bottom),
                out, ld_out);
        return 0;
    });
}

// ---------------- tracker front-ends ------------------------------------------------------------------
// kind: "sort" | "bytetrack" | "ocsort" | "botsort" | "strongsort" | "deepocsort"; p = the numeric ctor
// arguments in the order documented per kind below (same order as the oracle's orc_*_create).
// the asso_func constructor argument of the trackers created next ("iou" unless set; OC-SORT and DeepOC-SORT use it)
static std::string g_asso_func = "iou";
void ref_set_asso_func(const char* name) { g_asso_func = name ? name : "iou"; }

void* ref_tracker_create(const char* kind, const float* p, int np) {
    g_err.clear();
    try {
        const std::string k(kind);
        auto need = [&](int n) { if (np != n) throw std::invalid_argument("ref_tracker_create: wrong parameter count for " + k); };
        motcpp::BaseTracker* t = nullptr;
        if (k == "sort") {            // det_thresh, max_age, max_obs, min_hits, iou_threshold
            need(5);
            t = new motcpp::trackers::Sort(p[0], (int)p[1], (int)p[2], (int)p[3], p[4]);
        } else if (k == "bytetrack") { // det_thresh, max_age, max_obs, min_hits, iou_threshold, min_conf, track_thresh, match_thresh, track_buffer, frame_rate
            need(10);
            t = new motcpp::trackers::ByteTrack(p[0], (int)p[1], (int)p[2], (int)p[3], p[4], false, 80, "iou", false,
                                                p[5], p[6], p[7], (int)p[8], (int)p[9]);
        } else if (k == "ocsort") {    // det_thresh, max_age, max_obs, min_hits, iou_threshold, min_conf, delta_t, inertia, use_byte, Q_xy_scaling, Q_s_scaling
            need(11);
            t = new motcpp::trackers::OCSort(p[0], (int)p[1], (int)p[2], (int)p[3], p[4], false, 80, g_asso_func, false,
                                             p[5], (int)p[6], p[7], p[8] != 0.0f, p[9], p[10]);
        } else if (k == "botsort") {   // track_high, track_low, new_track, track_buffer, match_thresh, proximity, appearance, frame_rate, fuse_first_associate, with_reid
            need(10);
            t = new motcpp::trackers::BotSort("", false, false, 0.3f, 30, 50, 3, 0.3f, false, 80, "iou", false,
                                              p[0], p[1], p[2], (int)p[3], p[4], p[5], p[6], "none", (int)p[7],
                                              p[8] != 0.0f, p[9] != 0.0f);
        } else if (k == "strongsort") { // max_age, min_conf, max_cos_dist, max_iou_dist, n_init, nn_budget, mc_lambda, ema_alpha
            need(8);
            t = new motcpp::trackers::StrongSORT("", false, false, 0.3f, (int)p[0], 50, 3, 0.3f, false, 80, "iou", false,
                                                 p[1], p[2], p[3], (int)p[4], (int)p[5], p[6], p[7]);
        } else if (k == "deepocsort") { // det_thresh, max_age, max_obs, min_hits, iou_threshold, delta_t, inertia, w_association_emb, alpha_fixed_emb, aw_param, embedding_off, aw_off, Q_xy_scaling, Q_s_scaling   (cmc_off = true)
            need(14);
            t = new motcpp::trackers::DeepOCSort("", false, false, p[0], (int)p[1], (int)p[2], (int)p[3], p[4], false, 80,
                                                 "iou", false, (int)p[5], p[6], p[7], p[8], p[9], p[10] != 0.0f, true,
                                                 p[11] != 0.0f, p[12], p[13]);
        } else if (k == "boosttrack") { // det_thresh, max_age, max_obs, min_hits, iou_threshold, min_box_area, aspect_ratio_thresh, lambda_iou, lambda_mhd, lambda_shape, use_dlo_boost, dlo_boost_coef, use_vt   (use_ecc = false, with_reid = false, use_sb = false)
            need(13);
            t = new motcpp::trackers::BoostTrackTracker("", false, false, p[0], (int)p[1], (int)p[2], (int)p[3], p[4], false, 80,
                                                        "iou", false, /*use_ecc=*/false, (int)p[5], p[6], "none", p[7], p[8], p[9],
                                                        p[10] != 0.0f, true, p[11], false, false, /*use_sb=*/false, p[12] != 0.0f,
                                                        /*with_reid=*/false);
        } else {
            throw std::invalid_argument("ref_tracker_create: unknown kind " + k);
        }
        return t;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void ref_tracker_destroy(void* h) { delete static_cast<motcpp::BaseTracker*>(h); }
int ref_tracker_reset(void* h) {
    return guarded([&] { static_cast<motcpp::BaseTracker*>(h)->reset(); return 0; });
}
// dets (n x 6) row-major, embs (n x dim) or NULL; img is a shape-only 1080 x 1920 handle.
// out: rows of 8 floats; returns the number of rows, -needed when out_cap is too small, -1000 on an exception.
int ref_tracker_update(void* h, const float* dets, int n, const float* embs, int dim, float* out, int out_cap) {
    return guarded([&] {
        cv::Mat img(1080, 1920, cv::CV_8UC3);
        MatrixXf d = n > 0 ? from_rows(dets, n, 6) : MatrixXf(0, 6);
        MatrixXf e = (embs && dim > 0 && n > 0) ? from_rows(embs, n, dim) : MatrixXf();
        MatrixXf r = static_cast<motcpp::BaseTracker*>(h)->update(d, img, e);
        const int rows = static_cast<int>(r.rows());
        if (rows > out_cap) return -rows;
        if (rows > 0 && r.cols() != 8) throw std::runtime_error("tracker returned a matrix that is not (M, 8)");
        to_rows(r, out);
        return rows;
    });
}

}  // extern "C"

// ======================================================================================================
// Link-time definitions for reference classes OUTSIDE the hot path (their sources need real OpenCV /
// ONNX Runtime: src/motion/cmc/{cmc,ecc,sof}.cpp, src/appearance/{reid,onnx}_backend.cpp).
// ======================================================================================================
namespace motcpp::motion {

cv::Mat CMC::preprocess(const cv::Mat& img, float, bool) { return img; }

ECC::ECC(int warp_mode, float eps, int max_iter, float scale, bool align, bool grayscale)
    : warp_mode_(warp_mode), eps_(eps), max_iter_(max_iter), scale_(scale), align_(align), grayscale_(grayscale) {}
// camera motion = identity
Eigen::Matrix<float, 2, 3> ECC::apply(const cv::Mat&, const Eigen::MatrixXf&) {
    Eigen::Matrix<float, 2, 3> w;
    w.setZero();
    w(0, 0) = 1.0f;
    w(1, 1) = 1.0f;
    return w;
}

const cv::Size SOF::win_size_(21, 21);
const cv::TermCriteria SOF::term_criteria_(3, 30, 0.01);
SOF::SOF(float scale) : scale_(scale), initialized_(false) {}
Eigen::Matrix<float, 2, 3> SOF::apply(const cv::Mat&, const Eigen::MatrixXf&) {
    Eigen::Matrix<float, 2, 3> w;
    w.setZero();
    w(0, 0) = 1.0f;
    w(1, 1) = 1.0f;
    return w;
}

}  // namespace motcpp::motion

namespace motcpp::appearance {

Eigen::MatrixXf ReIDBackend::get_crops(const Eigen::MatrixXf&, const cv::Mat&) {
    throw std::runtime_error("ref_shim: ReID inference is outside the hot path - pass embeddings");
}
Eigen::MatrixXf ReIDBackend::normalize_features(const Eigen::MatrixXf& f) { return f; }
std::pair<int, int> ReIDBackend::determine_input_shape(const std::string&) { return {256, 128}; }
std::pair<Eigen::Vector3f, Eigen::Vector3f> ReIDBackend::determine_normalization(const std::string&) {
    return {Eigen::Vector3f(0.485f, 0.456f, 0.406f), Eigen::Vector3f(0.229f, 0.224f, 0.225f)};
}

ONNXBackend::ONNXBackend(const std::string& model_path, const std::string& model_name, bool use_half, bool use_gpu)
    : model_path_(model_path), model_name_(model_name), use_gpu_(use_gpu) {
    use_half_ = use_half;
    input_shape_ = {256, 128};
}
ONNXBackend::~ONNXBackend() = default;
Eigen::MatrixXf ONNXBackend::get_features(const Eigen::MatrixXf&, const cv::Mat&) {
    throw std::runtime_error("ref_shim: ReID inference is outside the hot path - pass embeddings");
}
void ONNXBackend::warmup() {}

}  // namespace motcpp::appearance
