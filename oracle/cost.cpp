// TEST INFRASTRUCTURE (oracle): box conversions and cost-matrix builders.
//   conversions : include/motcpp/utils/ops.hpp:15-211
//   iou_batch   : include/motcpp/utils/iou.hpp:63-100
//   iou_distance: src/utils/matching.cpp:62-65
//   fuse_score  : src/utils/matching.cpp:130-143
//   embedding_distance: src/utils/matching.cpp:67-107
#include "oracle.h"

#include <algorithm>
#include <cmath>

extern "C" {

void orc_xyxy2xywh(const float* in, float* out) {          // ops.hpp:15-22
    const float w = in[2] - in[0], h = in[3] - in[1];
    out[0] = in[0] + w * 0.5f; out[1] = in[1] + h * 0.5f; out[2] = w; out[3] = h;
}
void orc_xywh2xyxy(const float* in, float* out) {          // ops.hpp:27-34
    const float xc = in[0], yc = in[1], w = in[2], h = in[3];
    out[0] = xc - w * 0.5f; out[1] = yc - h * 0.5f; out[2] = xc + w * 0.5f; out[3] = yc + h * 0.5f;
}
void orc_xywh2tlwh(const float* in, float* out) {          // ops.hpp:39-44
    const float xc = in[0], yc = in[1], w = in[2], h = in[3];
    out[0] = xc - w * 0.5f; out[1] = yc - h * 0.5f; out[2] = w; out[3] = h;
}
void orc_tlwh2xyah(const float* in, float* out) {          // ops.hpp:78-85
    const float t = in[0], l = in[1], w = in[2], h = in[3];
    out[0] = t + w * 0.5f; out[1] = l + h * 0.5f; out[2] = (h > 0.0f) ? (w / h) : 0.0f; out[3] = h;
}
void orc_xyah2xywh(const float* in, float* out) {          // ops.hpp:108-112
    out[0] = in[0]; out[1] = in[1]; out[2] = in[2] * in[3]; out[3] = in[3];
}
void orc_xyxy2xysr(const float* in, float* out) {          // ops.hpp:188-197
    const float w = in[2] - in[0], h = in[3] - in[1];
    out[0] = in[0] + w * 0.5f; out[1] = in[1] + h * 0.5f; out[2] = w * h;
    out[3] = (h > 1e-6f) ? (w / h) : 0.0f;
}
void orc_xysr2xyxy(const float* in, float* out) {          // ops.hpp:202-211
    const float xc = in[0], yc = in[1], s = in[2], r = in[3];
    const float w = std::sqrt(s * r);
    const float h = s / w;
    out[0] = xc - w * 0.5f; out[1] = yc - h * 0.5f; out[2] = xc + w * 0.5f; out[3] = yc + h * 0.5f;
}

void orc_iou_batch(const float* a, int n, const float* b, int m, float* out) {
    for (int i = 0; i < n; ++i) {
        const float* p = a + 4 * i;
        const float area1 = (p[2] - p[0]) * (p[3] - p[1]);
        for (int j = 0; j < m; ++j) {
            const float* q = b + 4 * j;
            const float area2 = (q[2] - q[0]) * (q[3] - q[1]);
            const float xx1 = std::max(p[0], q[0]);
            const float yy1 = std::max(p[1], q[1]);
            const float xx2 = std::min(p[2], q[2]);
            const float yy2 = std::min(p[3], q[3]);
            const float w = std::max(0.0f, xx2 - xx1);
            const float h = std::max(0.0f, yy2 - yy1);
            const float inter = w * h;
            const float uni = area1 + area2 - inter;
            out[(size_t)i * m + j] = (uni > 0.0f) ? (inter / uni) : 0.0f;
        }
    }
}

void orc_iou_distance(const float* a, int n, const float* b, int m, float* out) {
    orc_iou_batch(a, n, b, m, out);
    for (size_t k = 0; k < (size_t)n * m; ++k) out[k] = 1.0f - out[k];
}

void orc_fuse_score(float* cost, int n, int m, const float* det_conf) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) {
            const float sim = 1.0f - cost[(size_t)i * m + j];
            const float fused = sim * det_conf[j];
            cost[(size_t)i * m + j] = 1.0f - fused;
        }
}

namespace {
float dot_seq(const float* x, const float* y, int dim) {
    float acc = x[0] * y[0];
    for (int k = 1; k < dim; ++k) acc = acc + x[k] * y[k];
    return acc;
}
}  // namespace

// Eigen's .norm()/.dot() use an unspecified vectorised summation order; this restatement sums
// sequentially, so cosine costs agree with the stock reference to ~1e-6 abs, not bit-for-bit.
void orc_embedding_distance(const float* t, int n, const float* d, int m, int dim, int metric, float* out) {
    for (int i = 0; i < n; ++i) {
        const float* tf = t + (size_t)i * dim;
        const float tn = std::sqrt(dot_seq(tf, tf, dim));
        for (int j = 0; j < m; ++j) {
            const float* df = d + (size_t)j * dim;
            if (metric == 0) {
                const float dn = std::sqrt(dot_seq(df, df, dim));
                const float sim = dot_seq(tf, df, dim) / (tn * dn + 1e-10f);
                out[(size_t)i * m + j] = std::max(0.0f, 1.0f - sim);
            } else {
                float acc = 0.0f;
                for (int k = 0; k < dim; ++k) { const float e = tf[k] - df[k]; acc = acc + e * e; }
                out[(size_t)i * m + j] = std::sqrt(acc);
            }
        }
    }
}

}  // extern "C"
