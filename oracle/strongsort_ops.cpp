// TEST INFRASTRUCTURE (oracle): the cost builders of the StrongSORT association and the XYSR
// affine (camera-motion) correction (8a9 / 8f-2), restated from
//   NearestNeighborDistanceMetric::distance / nn_cosine_distance / cosine_distance  src/trackers/strongsort.cpp:240-334
//   linear_assignment::gate_cost_matrix                                            src/trackers/strongsort.cpp:451-492
//   iou_matching::iou / iou_cost                                                   src/trackers/strongsort.cpp:502-585
//   linear_assignment::min_cost_matching (threshold clamp)                         src/trackers/strongsort.cpp:372-377
//   KalmanFilterXYSR::apply_affine_correction                                      src/motion/kalman_filters/xysr_kf.cpp:114-141
//   deepocsort_assoc::compute_aw_max_metric (DeepOC-SORT adaptive embedding weights) src/trackers/deepocsort.cpp:294-345
// Pinning: compute_aw_max_metric and the affine correction are compared with the reference's compiled code
// (tests/test_ref_pin.py), the StrongSORT builders through the whole compiled strongsort.cpp; hand-derived KATs in
// tests/test_strongsort_ops.py.  Eigen's GEMM / .norm() summation order is unspecified; sums here are sequential.
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <vector>

extern "C" {

// strongsort.cpp:240-334, metric "cosine".  samples (n_samples x dim): every target's gallery, seg[s] = the target row
// sample s belongs to; feats (m x dim) raw detection features.  out (n_targets x m): min over the target's samples of
// 1 - (s/|s|) . (f/|f|) (rows / features with norm <= 1e-10 stay unnormalised, :318-329); no samples -> 1e5 (:271).
void orc_nn_cosine_distance(const float* samples, const int* seg, int n_samples, int n_targets, const float* feats,
                            int m, int dim, float* out) {
    for (size_t k = 0; k < (size_t)n_targets * m; ++k) out[k] = 1e5f;
    std::vector<float> fn((size_t)m * dim), sn(dim);
    auto normalise = [&](const float* src, float* dst) {
        float acc = 0.0f;
        for (int k = 0; k < dim; ++k) acc = acc + src[k] * src[k];
        const float nrm = std::sqrt(acc);
        for (int k = 0; k < dim; ++k) dst[k] = (nrm > 1e-10f) ? src[k] / nrm : src[k];
    };
    for (int j = 0; j < m; ++j) normalise(feats + (size_t)j * dim, fn.data() + (size_t)j * dim);
    std::vector<char> seen(n_targets, 0);
    for (int s = 0; s < n_samples; ++s) {
        const int t = seg[s];
        if (t < 0 || t >= n_targets) continue;
        normalise(samples + (size_t)s * dim, sn.data());
        for (int j = 0; j < m; ++j) {
            const float* f = fn.data() + (size_t)j * dim;
            float acc = 0.0f;
            for (int k = 0; k < dim; ++k) acc = acc + sn[k] * f[k];
            const float d = 1.0f - acc;                                        // :333
            float& o = out[(size_t)t * m + j];
            if (!seen[t] || d < o) o = d;                                      // colwise().minCoeff() (:295)
        }
        seen[t] = 1;
    }
}

// strongsort.cpp:451-492 in place: recs = n_tracks XYAH records [mean 8 | cov 64], meas = (n_meas x 4) xyah rows.
void orc_gate_cost_matrix(float* cost, int ld, const float* recs, int n_tracks, const float* meas, int n_meas,
                          float mc_lambda, float gated_cost, int only_position) {
    const float gating_threshold = 9.4877f;                                    // :461
    std::vector<float> gd(n_meas);
    for (int r = 0; r < n_tracks; ++r) {
        orc_kf_xyah_gating(recs + (size_t)r * 72, recs + (size_t)r * 72 + 8, meas, n_meas, only_position, 0, gd.data());
        float* row = cost + (size_t)r * ld;
        for (int j = 0; j < n_meas; ++j) {
            float c = row[j];
            if (gd[j] > gating_threshold) c = gated_cost;                      // :477-481
            row[j] = mc_lambda * c + (1.0f - mc_lambda) * gd[j];               // :484-487
        }
    }
}

// strongsort.cpp:502-585: tracks / candidates as tlwh rows; tsu (nullable) = time_since_update per track, rows with
// tsu > 1 are INFTY_COST = 1e5 (:567-570).  out (n x m) = 1 - iou.
void orc_iou_cost_tlwh(const float* trk, const int* tsu, int n, const float* det, int m, float* out) {
    for (int i = 0; i < n; ++i) {
        float* row = out + (size_t)i * m;
        if (tsu && tsu[i] > 1) { for (int j = 0; j < m; ++j) row[j] = 1e5f; continue; }
        const float* b = trk + 4 * i;
        const float bx2 = b[0] + b[2], by2 = b[1] + b[3];
        const float area_b = b[2] * b[3];
        for (int j = 0; j < m; ++j) {
            const float* c = det + 4 * j;
            const float cx2 = c[0] + c[2], cy2 = c[1] + c[3];
            const float tlx = std::max(b[0], c[0]), tly = std::max(b[1], c[1]);
            const float brx = std::min(bx2, cx2), bry = std::min(by2, cy2);
            const float w = std::max(0.0f, brx - tlx), h = std::max(0.0f, bry - tly);
            const float inter = w * h;
            const float area_c = c[2] * c[3];
            const float uni = area_b + area_c - inter;
            const float iou = (uni > 1e-6f) ? (inter / uni) : 0.0f;            // :533
            row[j] = 1.0f - iou;                                               // :576
        }
    }
}

// strongsort.cpp:372-377: entries above max_distance become max_distance + 1e-5
void orc_clamp_cost(float* cost, int n, int m, int ld, float max_distance) {
    const float cap = max_distance + 1e-5f;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j)
            if (cost[(size_t)i * ld + j] > max_distance) cost[(size_t)i * ld + j] = cap;
}

// xysr_kf.cpp:114-141: x7 / P49 (row-major 7x7) in place; m2 row-major 2x2, t2.
void orc_kf_xysr_affine(float* x7, float* P49, const float* m2, const float* t2) {
    auto mv = [&](float a, float b, float& o0, float& o1) {
        o0 = m2[0] * a + m2[1] * b;
        o1 = m2[2] * a + m2[3] * b;
    };
    float c0, c1;
    mv(x7[0], x7[1], c0, c1);
    x7[0] = c0 + t2[0]; x7[1] = c1 + t2[1];
    mv(x7[4], x7[5], c0, c1);
    x7[4] = c0; x7[5] = c1;
    auto block = [&](int r0, int q0, float o[4]) {                             // m * B * m^T, B = P[r0:r0+2, q0:q0+2]
        const float b00 = P49[r0 * 7 + q0], b01 = P49[r0 * 7 + q0 + 1];
        const float b10 = P49[(r0 + 1) * 7 + q0], b11 = P49[(r0 + 1) * 7 + q0 + 1];
        const float a00 = m2[0] * b00 + m2[1] * b10, a01 = m2[0] * b01 + m2[1] * b11;
        const float a10 = m2[2] * b00 + m2[3] * b10, a11 = m2[2] * b01 + m2[3] * b11;
        o[0] = a00 * m2[0] + a01 * m2[1]; o[1] = a00 * m2[2] + a01 * m2[3];
        o[2] = a10 * m2[0] + a11 * m2[1]; o[3] = a10 * m2[2] + a11 * m2[3];
    };
    float pp[4], vv[4], pv[4];
    block(0, 0, pp);
    block(4, 4, vv);
    block(0, 4, pv);
    P49[0] = pp[0]; P49[1] = pp[1]; P49[7] = pp[2]; P49[8] = pp[3];
    P49[4 * 7 + 4] = vv[0]; P49[4 * 7 + 5] = vv[1]; P49[5 * 7 + 4] = vv[2]; P49[5 * 7 + 5] = vv[3];
    P49[4] = pv[0]; P49[5] = pv[1]; P49[7 + 4] = pv[2]; P49[7 + 5] = pv[3];
    P49[4 * 7 + 0] = pv[0]; P49[5 * 7 + 0] = pv[1]; P49[4 * 7 + 1] = pv[2]; P49[5 * 7 + 1] = pv[3];   // transpose (:140)
}

// deepocsort_assoc::compute_aw_max_metric (src/trackers/deepocsort.cpp:294-345): adaptive weighting of the embedding
// cost by how distinctive each row's / column's best match is.  emb (n x m, ld) -> out (n x m, ld_out).
void orc_aw_max_metric(const float* emb, int n, int m, int ld, float w_assoc, float bottom, float* out, int ld_out) {
    std::vector<float> w((size_t)n * m, w_assoc);
    auto weight = [&](float mx, float second) { return 1.0f - std::max((second / mx) - bottom, 0.0f) / (1.0f - bottom); };
    if (m >= 2)
        for (int i = 0; i < n; ++i) {
            float mx = -INFINITY, se = -INFINITY;                             // two largest values of the row (:303-313)
            for (int j = 0; j < m; ++j) {
                const float v = emb[(size_t)i * ld + j];
                if (v > mx) { se = mx; mx = v; } else if (v > se) se = v;
            }
            for (int j = 0; j < m; ++j) {
                if (mx == 0.0f) w[(size_t)i * m + j] = 0.0f;                   // setZero (:315-316)
                else w[(size_t)i * m + j] = w[(size_t)i * m + j] * weight(mx, se);
            }
        }
    if (n >= 2)
        for (int j = 0; j < m; ++j) {
            float mx = -INFINITY, se = -INFINITY;
            for (int i = 0; i < n; ++i) {
                const float v = emb[(size_t)i * ld + j];
                if (v > mx) { se = mx; mx = v; } else if (v > se) se = v;
            }
            for (int i = 0; i < n; ++i) {
                if (mx == 0.0f) w[(size_t)i * m + j] = 0.0f;
                else w[(size_t)i * m + j] = w[(size_t)i * m + j] * weight(mx, se);
            }
        }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) out[(size_t)i * ld_out + j] = w[(size_t)i * m + j] * emb[(size_t)i * ld + j];   // cwiseProduct (:344)
}

}  // extern "C"
