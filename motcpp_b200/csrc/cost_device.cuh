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
enum : int { kVarHmIou = 3, kVarGIoU = 4, kVarDIoU = 5, kVarCentroid = 6 };
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
    return xdiv(xadd(xsub(iou, xdiv(inner, xadd(outer, 1e-10f))), 1.0f), 2.0f);
}

}  // namespace mot
