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

// hmiou_batch / giou_batch / diou_batch / centroid_batch (include/motcpp/utils/iou.hpp:119-146, :151-187, :258-293,
// :298-330), evaluated PAIR-WISE: element (i, j) from box a[i] and box b[j].  That is what the reference's expressions
// mean and what they compute whenever b has ONE row (its `replicate(N, 1)` of a column of b only lines up for M == 1;
// SURVEY 8 trap 11), which is also all its tests exercise (tests/test_iou.cpp:74-115).  kind: 3 hmiou, 4 giou, 5 diou,
// 6 centroid (frame_w / frame_h only matter there), 7 ciou (:197-253).
// ciou's arc tangent: the reference applies Eigen's array .atan() (iou.hpp:238-239), whose implementation depends on the
// (unpinned) Eigen version and the build's vector width.  The contract here, as for acos (oracle/ocsort.cpp), is the
// CORRECTLY ROUNDED fp32 arc tangent: fdlibm's atan evaluated in fp64 by a fixed sequence of IEEE operations and rounded
// once; csrc/cost_device.cuh repeats the sequence operation for operation.
float orc_atanf(float xf) {
    static const double atanhi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01,
                                     1.57079632679489655800e+00};
    static const double atanlo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17,
                                     6.12323399573676603587e-17};
    static const double aT[11] = {3.33333333333329318027e-01, -1.99999999998764832476e-01, 1.42857142725034663711e-01,
                                  -1.11111104054623557880e-01, 9.09088713343650656196e-02, -7.69187620504482999495e-02,
                                  6.66107313738753120669e-02, -5.83357013379057348645e-02, 4.97687799461593236017e-02,
                                  -3.65315727442169155270e-02, 1.62858201153657823623e-02};
    double x = (double)xf;
    if (!(x == x)) return xf;
    const bool neg = std::signbit(xf);
    const double ax = std::fabs(x);
    if (ax >= 7.378697629483821e+19) {                       // 2^66
        const double r = atanhi[3] + atanlo[3];
        return (float)(neg ? -r : r);
    }
    int id;
    if (ax < 0.4375) {
        if (ax < 1.862645149230957e-09) return xf;           // 2^-29
        id = -1;
    } else {
        x = ax;
        if (ax < 1.1875) {
            if (ax < 0.6875) { id = 0; x = (2.0 * x - 1.0) / (2.0 + x); }
            else { id = 1; x = (x - 1.0) / (x + 1.0); }
        } else {
            if (ax < 2.4375) { id = 2; x = (x - 1.5) / (1.0 + 1.5 * x); }
            else { id = 3; x = -1.0 / x; }
        }
    }
    const double z = x * x, w = z * z;
    const double s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    const double s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return (float)(x - x * (s1 + s2));
    const double r = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return (float)(neg ? -r : r);
}

void orc_iou_variant(const float* a, int n, const float* b, int m, int kind, int frame_w, int frame_h, float* out) {
    const float norm = static_cast<float>(std::sqrt(frame_w * frame_w + frame_h * frame_h));            // :325
    for (int i = 0; i < n; ++i) {
        const float* p = a + 4 * i;
        for (int j = 0; j < m; ++j) {
            const float* q = b + 4 * j;
            float iou1;
            orc_iou_batch(p, 1, q, 1, &iou1);
            float r = 0.0f;
            if (kind == 3) {                                                                            // :128-145
                const float ih = std::max(std::min(p[3], q[3]) - std::max(p[1], q[1]), 0.0f);
                const float uh = std::max(std::max(p[3], q[3]) - std::min(p[1], q[1]), 1e-10f);
                r = iou1 * (ih / uh);
            } else if (kind == 4) {                                                                     // :162-186
                const float wc = std::max(p[2], q[2]) - std::min(p[0], q[0]);
                const float hc = std::max(p[3], q[3]) - std::min(p[1], q[1]);
                const float enc = wc * hc;
                const float a1 = (p[2] - p[0]) * (p[3] - p[1]), a2 = (q[2] - q[0]) * (q[3] - q[1]);
                const float inter = iou1 * (a1 + a2) / (iou1 + 1e-10f);
                const float uni = a1 + a2 - inter;
                const float g = iou1 - (enc - uni) / (enc + 1e-10f);
                r = (g + 1.0f) / 2.0f;
            } else if (kind == 5) {                                                                     // :269-292
                const float cx1 = (p[0] + p[2]) / 2.0f, cy1 = (p[1] + p[3]) / 2.0f;
                const float cx2 = (q[0] + q[2]) / 2.0f, cy2 = (q[1] + q[3]) / 2.0f;
                const float dx = cx1 - cx2, dy = cy1 - cy2;
                const float inner = dx * dx + dy * dy;
                const float ox = std::max(p[2], q[2]) - std::min(p[0], q[0]);
                const float oy = std::max(p[3], q[3]) - std::min(p[1], q[1]);
                const float outer = ox * ox + oy * oy;
                const float d = iou1 - inner / (outer + 1e-10f);
                r = (d + 1.0f) / 2.0f;
            } else if (kind == 7) {                                                                     // :197-253
                const float eps = 1e-7f;
                const float cx1 = (p[0] + p[2]) / 2.0f, cy1 = (p[1] + p[3]) / 2.0f;
                const float cx2 = (q[0] + q[2]) / 2.0f, cy2 = (q[1] + q[3]) / 2.0f;
                const float dx = cx1 - cx2, dy = cy1 - cy2;
                const float inner = dx * dx + dy * dy;
                const float ox = std::max(p[2], q[2]) - std::min(p[0], q[0]);
                const float oy = std::max(p[3], q[3]) - std::min(p[1], q[1]);
                const float outer = (ox * ox + oy * oy) + eps;
                const float w1 = p[2] - p[0], h1 = p[3] - p[1], w2 = q[2] - q[0], h2 = q[3] - q[1];
                const float ad = orc_atanf(w2 / (h2 + eps)) - orc_atanf(w1 / (h1 + eps));
                const float pi_squared = static_cast<float>(M_PI * M_PI);
                const float v = (4.0f / pi_squared) * (ad * ad);
                const float S = 1.0f - iou1;
                const float alpha = v / ((S + v) + eps);
                const float c = (iou1 - inner / outer) + alpha * v;
                r = (c + 1.0f) / 2.0f;
            } else {                                                                                    // :308-329
                const float dx = (p[0] + p[2]) / 2.0f - (q[0] + q[2]) / 2.0f;
                const float dy = (p[1] + p[3]) / 2.0f - (q[1] + q[3]) / 2.0f;
                const float dist = std::sqrt(dx * dx + dy * dy);
                r = 1.0f - dist / norm;
            }
            out[(size_t)i * m + j] = r;
        }
    }
}

}  // extern "C"
