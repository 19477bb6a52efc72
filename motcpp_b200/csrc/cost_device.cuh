// cost_device.cuh - box conversions and IoU-family cost functors (exact fp32, reference op order).
//   conversions  reference include/motcpp/utils/ops.hpp:15-211
//   iou_batch    reference include/motcpp/utils/iou.hpp:63-100
//   iou_distance reference src/utils/matching.cpp:62-65
//   fuse_score   reference src/utils/matching.cpp:130-143
#pragma once
#include "simt.cuh"

namespace mot {

// xyxy -> xywh (ops.hpp:15-22)
__device__ __forceinline__ float4 xyxy2xywh(float4 b) {
    const float w = xsub(b.z, b.x), h = xsub(b.w, b.y);
    return make_float4(xadd(b.x, xmul(w, 0.5f)), xadd(b.y, xmul(h, 0.5f)), w, h);
}
// xywh -> xyxy (ops.hpp:27-34)
__device__ __forceinline__ float4 xywh2xyxy(float4 b) {
    const float hw = xmul(b.z, 0.5f), hh = xmul(b.w, 0.5f);
    return make_float4(xsub(b.x, hw), xsub(b.y, hh), xadd(b.x, hw), xadd(b.y, hh));
}
// xywh -> tlwh -> xyah (ops.hpp:39-44, 78-85), the chain STrack's ctor runs (bytetrack.cpp:26-29)
__device__ __forceinline__ float4 xywh2xyah_via_tlwh(float4 b) {
    const float t = xsub(b.x, xmul(b.z, 0.5f)), l = xsub(b.y, xmul(b.w, 0.5f));
    const float a = (b.w > 0.0f) ? xdiv(b.z, b.w) : 0.0f;
    return make_float4(xadd(t, xmul(b.z, 0.5f)), xadd(l, xmul(b.w, 0.5f)), a, b.w);
}
// xyah -> xywh -> xyxy (ops.hpp:108-112, 27-34): STrack::xyxy() (bytetrack.cpp:118-128)
__device__ __forceinline__ float4 xyah2xyxy(float x, float y, float a, float h) {
    return xywh2xyxy(make_float4(x, y, xmul(a, h), h));
}
// xyxy -> xysr (ops.hpp:188-197)
__device__ __forceinline__ float4 xyxy2xysr(float4 b) {
    const float w = xsub(b.z, b.x), h = xsub(b.w, b.y);
    return make_float4(xadd(b.x, xmul(w, 0.5f)), xadd(b.y, xmul(h, 0.5f)), xmul(w, h),
                       (h > 1e-6f) ? xdiv(w, h) : 0.0f);
}
// xysr -> xyxy (ops.hpp:202-211)
__device__ __forceinline__ float4 xysr2xyxy(float x, float y, float s, float r) {
    const float w = xsqrt(xmul(s, r));
    const float h = xdiv(s, w);
    const float hw = xmul(w, 0.5f), hh = xmul(h, 0.5f);
    return make_float4(xsub(x, hw), xsub(y, hh), xadd(x, hw), xadd(y, hh));
}

__device__ __forceinline__ float box_area(float4 b) { return xmul(xsub(b.z, b.x), xsub(b.w, b.y)); }

// iou.hpp:81-98 for one pair; area_a precomputed
__device__ __forceinline__ float iou_pair(float4 a, float area_a, float4 b) {
    const float area_b = box_area(b);
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.0f, xsub(xx2, xx1));
    const float h = fmaxf(0.0f, xsub(yy2, yy1));
    const float inter = xmul(w, h);
    const float uni = xsub(xadd(area_a, area_b), inter);
    // 0 / uni is exactly +0; answering it directly keeps zero dividends (disjoint boxes, the vast
    // majority of an N x M matrix) off the IEEE divider's slow path.  Same bits as the division.
    if (!(uni > 0.0f)) return 0.0f;
    const bool z = (inter == 0.0f);
    const float q = xdiv(z ? 1.0f : inter, uni);
    return z ? 0.0f : q;
}

// true => the boxes share no interior => IoU is exactly 0
__device__ __forceinline__ bool boxes_disjoint(float4 a, float4 b) {
    return !((fminf(a.z, b.z) > fmaxf(a.x, b.x)) && (fminf(a.w, b.w) > fmaxf(a.y, b.y)));
}

// Tie-break infinitesimal of the OC-SORT assignments (see OcmCost::pair_bias): prefers the higher column, then
// couples rows and columns so that two rows on two bit-identical columns also have a single optimum.
__device__ __forceinline__ double twin_bias(int i, int j) { return -(double)(64 * j + ((i * j) & 63)) * 0x1p-50; }

// Cost functor for block_lap(): rows = row_box[i], columns = det_box[col_map[j]];
// cost = 1 - IoU, optionally fused with the detection score: 1 - (1 - d) * conf.
// `prune` must only be set when thresh < 1 (a disjoint pair costs exactly 1).
struct IouCost {
    static constexpr bool kWarpPerRow = false;
    static constexpr bool kGrid = true;      // rows/columns are boxes: block_lap may index the columns spatially
    const float4* row_box;
    const float4* det_box;
    const float* det_conf;
    const unsigned short* col_map;
    bool fuse;
    bool prune;
    // an IoU no candidate can be below (cost <= thresh needs iou >= 1 - thresh; with fuse, as long as every confidence is
    // <= 1): block_lap's grid walk then only looks at the corner window such a pair can sit in.  0: unknown.
    static constexpr bool kIouFloor = true;
    static constexpr bool kSmallFast = true;   // block_lap_solve's one-warp path for a handful of trivial components
    float iou_floor = 0.0f;
    struct Row { float4 b; float area; };
    __device__ __forceinline__ Row row(int i) const {
        Row r;
        r.b = row_box[i];
        r.area = box_area(r.b);
        return r;
    }
    __device__ __forceinline__ float4 col_box(int j) const { return det_box[col_map[j]]; }
    __device__ __forceinline__ bool reject(const Row& r, int j) const {
        return prune && boxes_disjoint(r.b, det_box[col_map[j]]);
    }
    __device__ __forceinline__ float cost(const Row& r, int j) const {
        const int d = col_map[j];
        float dist = xsub(1.0f, iou_pair(r.b, r.area, det_box[d]));
        if (fuse) {
            const float sim = xsub(1.0f, dist);
            dist = xsub(1.0f, xmul(sim, det_conf[d]));
        }
        return dist;
    }
    __device__ __forceinline__ float pair(int i, int j) const { return cost(row(i), j); }
    __device__ __forceinline__ bool is_candidate(const Row& r, int, int j, float thresh) const { return cost(r, j) <= thresh; }
    __device__ __forceinline__ double pair_bias(int, int) const { return 0.0; }
};

// hmiou / giou / diou / centroid for ONE pair (include/motcpp/utils/iou.hpp:119-330, pair-wise: see kernels_cost.cuh)
enum : int { kVarHmIou = 3, kVarGIoU = 4, kVarDIoU = 5, kVarCentroid = 6, kVarCIoU = 7 };

// correctly rounded fp32 arc tangent: fdlibm's atan in fp64, a fixed sequence of IEEE operations (no FMA), rounded once.
// oracle/cost.cpp: orc_atanf repeats it operation for operation (ciou, iou.hpp:238-239; see the note there).
__device__ __forceinline__ float atanf_cr(float xf) {
    const double hi0 = 4.63647609000806093515e-01, hi1 = 7.85398163397448278999e-01, hi2 = 9.82793723247329054082e-01,
                 hi3 = 1.57079632679489655800e+00;
    const double lo0 = 2.26987774529616870924e-17, lo1 = 3.06161699786838301793e-17, lo2 = 1.39033110312309984516e-17,
                 lo3 = 6.12323399573676603587e-17;
    const double aT0 = 3.33333333333329318027e-01, aT1 = -1.99999999998764832476e-01, aT2 = 1.42857142725034663711e-01,
                 aT3 = -1.11111104054623557880e-01, aT4 = 9.09088713343650656196e-02, aT5 = -7.69187620504482999495e-02,
                 aT6 = 6.66107313738753120669e-02, aT7 = -5.83357013379057348645e-02, aT8 = 4.97687799461593236017e-02,
                 aT9 = -3.65315727442169155270e-02, aT10 = 1.62858201153657823623e-02;
    double x = (double)xf;
    if (!(x == x)) return xf;
    const bool neg = (__float_as_uint(xf) >> 31) != 0;
    const double ax = fabs(x);
    if (ax >= 7.378697629483821e+19) {
        const double r = __dadd_rn(hi3, lo3);
        return __double2float_rn(neg ? -r : r);
    }
    int id;
    double hi = 0.0, lo = 0.0;
    if (ax < 0.4375) {
        if (ax < 1.862645149230957e-09) return xf;
        id = -1;
    } else {
        x = ax;
        if (ax < 1.1875) {
            if (ax < 0.6875) { id = 0; hi = hi0; lo = lo0; x = __ddiv_rn(__dsub_rn(__dmul_rn(2.0, x), 1.0), __dadd_rn(2.0, x)); }
            else { id = 1; hi = hi1; lo = lo1; x = __ddiv_rn(__dsub_rn(x, 1.0), __dadd_rn(x, 1.0)); }
        } else {
            if (ax < 2.4375) { id = 2; hi = hi2; lo = lo2; x = __ddiv_rn(__dsub_rn(x, 1.5), __dadd_rn(1.0, __dmul_rn(1.5, x))); }
            else { id = 3; hi = hi3; lo = lo3; x = __ddiv_rn(-1.0, x); }
        }
    }
    const double z = __dmul_rn(x, x), w = __dmul_rn(z, z);
    double s1 = __dadd_rn(aT8, __dmul_rn(w, aT10));
    s1 = __dadd_rn(aT6, __dmul_rn(w, s1));
    s1 = __dadd_rn(aT4, __dmul_rn(w, s1));
    s1 = __dadd_rn(aT2, __dmul_rn(w, s1));
    s1 = __dmul_rn(z, __dadd_rn(aT0, __dmul_rn(w, s1)));
    double s2 = __dadd_rn(aT7, __dmul_rn(w, aT9));
    s2 = __dadd_rn(aT5, __dmul_rn(w, s2));
    s2 = __dadd_rn(aT3, __dmul_rn(w, s2));
    s2 = __dmul_rn(w, __dadd_rn(aT1, __dmul_rn(w, s2)));
    const double t = __dmul_rn(x, __dadd_rn(s1, s2));
    if (id < 0) return __double2float_rn(__dsub_rn(x, t));
    const double r = __dsub_rn(hi, __dsub_rn(__dsub_rn(t, lo), x));
    return __double2float_rn(neg ? -r : r);
}
__device__ __forceinline__ float iou_variant_pair(int kind, float4 p, float area_p, float4 q, float norm) {
    if (kind == kVarCentroid) {
        const float dx = xsub(xdiv(xadd(p.x, p.z), 2.0f), xdiv(xadd(q.x, q.z), 2.0f));
        const float dy = xsub(xdiv(xadd(p.y, p.w), 2.0f), xdiv(xadd(q.y, q.w), 2.0f));
        return xsub(1.0f, xdiv(xsqrt(xadd(xmul(dx, dx), xmul(dy, dy))), norm));
    }
    const float iou = iou_pair(p, area_p, q);
    if (kind == kVarHmIou) {
        const float ih = fmaxf(xsub(fminf(p.w, q.w), fmaxf(p.y, q.y)), 0.0f);
        const float uh = fmaxf(xsub(fmaxf(p.w, q.w), fminf(p.y, q.y)), 1e-10f);
        return xmul(iou, xdiv(ih, uh));
    }
    const float ox = xsub(fmaxf(p.z, q.z), fminf(p.x, q.x)), oy = xsub(fmaxf(p.w, q.w), fminf(p.y, q.y));
    if (kind == kVarGIoU) {
        const float enc = xmul(ox, oy);
        const float a12 = xadd(area_p, box_area(q));
        const float inter = xdiv(xmul(iou, a12), xadd(iou, 1e-10f));
        const float uni = xsub(a12, inter);
        const float g = xsub(iou, xdiv(xsub(enc, uni), xadd(enc, 1e-10f)));
        return xdiv(xadd(g, 1.0f), 2.0f);
    }
    const float dx = xsub(xdiv(xadd(p.x, p.z), 2.0f), xdiv(xadd(q.x, q.z), 2.0f));
    const float dy = xsub(xdiv(xadd(p.y, p.w), 2.0f), xdiv(xadd(q.y, q.w), 2.0f));
    const float inner = xadd(xmul(dx, dx), xmul(dy, dy));
    const float outer = xadd(xmul(ox, ox), xmul(oy, oy));
    if (kind == kVarCIoU) {                                                  // iou.hpp:197-253
        const float eps = 1e-7f;
        const float w1 = xsub(p.z, p.x), h1 = xsub(p.w, p.y), w2 = xsub(q.z, q.x), h2 = xsub(q.w, q.y);
        const float ad = xsub(atanf_cr(xdiv(w2, xadd(h2, eps))), atanf_cr(xdiv(w1, xadd(h1, eps))));
        const float v = xmul(0.40528473f /* 4.0f / float(pi * pi) */, xmul(ad, ad));
        const float S = xsub(1.0f, iou);
        const float alpha = xdiv(v, xadd(xadd(S, v), eps));
        const float c = xadd(xsub(iou, xdiv(inner, xadd(outer, eps))), xmul(alpha, v));
        return xdiv(xadd(c, 1.0f), 2.0f);
    }
    return xdiv(xadd(xsub(iou, xdiv(inner, xadd(outer, 1e-10f))), 1.0f), 2.0f);
}

}  // namespace mot
