// ocm_device.cuh - OC-SORT's observation-centric-momentum association cost (exact fp32, reference op
// order) and the plain -IoU cost of its BYTE / last-observation re-match passes.
//   ocsort_assoc::associate   reference src/trackers/ocsort.cpp:610-737 (cost :617-700)
//   speed_direction           reference src/trackers/ocsort.cpp:159-170
//   re-match cost             reference src/trackers/ocsort.cpp:443-449, :504-510
//
// acosf: the reference calls std::acos(float), i.e. the libm of whatever box it is built on (glibc
// 2.39's differs from the correctly rounded value by 1 ulp for ~8 % of arguments).  The contract here
// is the CORRECTLY ROUNDED fp32 arc cosine: acos evaluated in fp64 by a fixed sequence of IEEE
// operations (fdlibm's rational approximation, no FMA) and rounded once to fp32.  oracle/ocsort.cpp
// repeats the sequence operation for operation, so both sides agree bit for bit.
#pragma once
#include "cost_device.cuh"

namespace mot {

__device__ __forceinline__ double acos_R(double z) {
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
                 pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
                 pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05;
    const double qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
                 qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    double p = __dmul_rn(z, pS5);
    p = __dmul_rn(z, __dadd_rn(pS4, p));
    p = __dmul_rn(z, __dadd_rn(pS3, p));
    p = __dmul_rn(z, __dadd_rn(pS2, p));
    p = __dmul_rn(z, __dadd_rn(pS1, p));
    p = __dmul_rn(z, __dadd_rn(pS0, p));
    double q = __dmul_rn(z, qS4);
    q = __dmul_rn(z, __dadd_rn(qS3, q));
    q = __dmul_rn(z, __dadd_rn(qS2, q));
    q = __dmul_rn(z, __dadd_rn(qS1, q));
    q = __dadd_rn(1.0, q);
    return __ddiv_rn(p, q);
}

// correctly rounded acosf (see the file header); argument already clamped to [-1, 1] by the caller
__device__ __forceinline__ float acosf_cr(float xf) {
    const double pi = 3.14159265358979311600e+00, pio2 = 1.57079632679489655800e+00;
    const double x = (double)xf;
    if (!(x == x)) return xf;
    if (x >= 1.0) return 0.0f;
    if (x <= -1.0) return __double2float_rn(pi);
    double r;
    if (x >= -0.5 && x <= 0.5) {
        r = __dsub_rn(pio2, __dadd_rn(x, __dmul_rn(x, acos_R(__dmul_rn(x, x)))));
    } else if (x > 0.5) {
        const double z = __dmul_rn(__dsub_rn(1.0, x), 0.5);
        const double s = __dsqrt_rn(z);
        r = __dmul_rn(2.0, __dadd_rn(s, __dmul_rn(s, acos_R(z))));
    } else {
        const double z = __dmul_rn(__dadd_rn(1.0, x), 0.5);
        const double s = __dsqrt_rn(z);
        r = __dsub_rn(pi, __dmul_rn(2.0, __dadd_rn(s, __dmul_rn(s, acos_R(z)))));
    }
    return __double2float_rn(r);
}

__device__ __forceinline__ float box_sum4(float4 b) { return xadd(xadd(xadd(b.x, b.y), b.z), b.w); }

// unit (dy, dx) from the centre of `from` to the centre of `to` (ocsort.cpp:159-170)
__device__ __forceinline__ float2 speed_direction(float4 from, float4 to) {
    const float cx1 = xdiv(xadd(from.x, from.z), 2.0f), cy1 = xdiv(xadd(from.y, from.w), 2.0f);
    const float cx2 = xdiv(xadd(to.x, to.z), 2.0f), cy2 = xdiv(xadd(to.y, to.w), 2.0f);
    const float dy = xsub(cy2, cy1), dx = xsub(cx2, cx1);
    const float norm = xadd(xsqrt(xadd(xmul(dy, dy), xmul(dx, dx))), 1e-6f);
    return make_float2(xdiv(dy, norm), xdiv(dx, norm));
}

// per-track momentum terms gathered once per frame: centre of k_previous_obs, velocity, validity
struct OcmTrack {
    float cx, cy, vy, vx;
};

// angle cost of one (detection, track) pair before the detection-score factor (ocsort.cpp:626-667)
__device__ __forceinline__ float ocm_angle_cost(float det_cx, float det_cy, float4 t /* cx, cy, vy, vx */, float valid,
                                                float inertia) {
    const float kPi = 3.14159265358979323846f;
    const float dx = xsub(det_cx, t.x), dy = xsub(det_cy, t.y);
    const float norm = xadd(xsqrt(xadd(xmul(dx, dx), xmul(dy, dy))), 1e-6f);
    const float Y = xdiv(dy, norm), X = xdiv(dx, norm);
    float c = xadd(xmul(t.w, X), xmul(t.z, Y));
    c = fminf(fmaxf(c, -1.0f), 1.0f);
    const float ang = xdiv(xsub(kPi / 2.0f, fabsf(acosf_cr(c))), kPi);
    return xmul(xmul(valid, ang), inertia);
}

// The tracker's AssociationFunction (iou.hpp:371-411) for one pair.  kind 0 = iou_batch; kind 6 = centroid_batch
// (1 - centre distance / frame diagonal, :298-330), the one variant whose expression is defined for every N x M.
// Centroid similarities are non-zero for every pair, so the callers switch the disjoint-box pruning off with it.
// A compile-time choice: as a run-time field it cost the default "iou" path 10 % (registers spilled in the solvers' inner loops).
template <int ASSO>
__device__ __forceinline__ float asso_pair(float norm, float4 a, float area_a, float4 b) {
    if constexpr (ASSO == kVarCentroid) return iou_variant_pair(kVarCentroid, a, area_a, b, norm);
    else return iou_pair(a, area_a, b);
}

// Cost functor for block_lap(): rows = high-confidence detections (row_map -> detection index), columns =
// tracks in list order.  cost = -(iou + angle cost * score).  is_candidate() additionally tallies the
// reference's "trivial one-to-one" test (ocsort.cpp:676-689): which rows / columns see more than one
// pair with iou > iou_threshold.
template <int ASSO>
struct OcmCostT {
    static constexpr bool kWarpPerRow = false;
    static constexpr bool kGrid = true;
    const float4* det_box;
    const float* det_conf;
    const unsigned short* row_map;
    const float4* trk_box;            // predicted boxes (shared memory)
    const float4* ocm;                // {cx, cy, vy, vx} of k_previous_obs / velocity per track
    const unsigned char* valid;       // bit 0: previous_obs(4) >= 0, bit 1: the track has a live twin
    float inertia, iou_thr;
    bool prune;                       // a disjoint pair can neither be a candidate nor count as iou > thr
    unsigned* row_bits;               // [ceil(n/32)] rows that have one pair with iou > thr
    unsigned* col_bits;               // [ceil(m/32)]
    unsigned short* row_hit;          // [n] column of (the last) such pair
    int* flags;                       // [0] any pair with iou > thr, [1] a row or column saw two
    float asso_norm;                  // frame diagonal (ASSO = centroid; "iou" below means the AssociationFunction's value)
    struct Row { float4 b; float area, cx, cy, score; };
    __device__ __forceinline__ Row row(int i) const {
        Row r;
        const int d = row_map[i];
        r.b = det_box[d];
        r.area = box_area(r.b);
        r.cx = xdiv(xadd(r.b.x, r.b.z), 2.0f);
        r.cy = xdiv(xadd(r.b.y, r.b.w), 2.0f);
        r.score = det_conf[d];
        return r;
    }
    __device__ __forceinline__ float4 col_box(int j) const { return trk_box[j]; }
    __device__ __forceinline__ bool reject(const Row& r, int j) const { return prune && boxes_disjoint(r.b, trk_box[j]); }
    __device__ __forceinline__ float iou(const Row& r, int j) const { return asso_pair<ASSO>(asso_norm, r.b, r.area, trk_box[j]); }
    __device__ __forceinline__ float cost_from_iou(const Row& r, int j, float v) const {
        const float va = (valid[j] & 1) ? 1.0f : 0.0f;
        if (va == 0.0f) return -v;                     // 0 * angle * inertia * score adds exactly +-0
        const float ac = xmul(ocm_angle_cost(r.cx, r.cy, ocm[j], va, inertia), r.score);
        return -xadd(v, ac);
    }
    __device__ __forceinline__ float cost(const Row& r, int j) const { return cost_from_iou(r, j, iou(r, j)); }
    __device__ __forceinline__ float pair(int i, int j) const { return cost(row(i), j); }
    // The reference spawns bit-identical "twin" tracks (parity trap 8), so exactly tied optima are systematic
    // here.  Its LAPJV resolves them by scan order over the whole dense matrix; this solver adds an
    // infinitesimal that prefers the HIGHER column (what LAPJV's right-to-left column reduction yields in most
    // cases) - DESIGN.md "Ties".
    __device__ __forceinline__ double pair_bias(int i, int j) const { return twin_bias(i, j); }
    // the "trivial one-to-one" tallies for a pair with IoU v
    __device__ __forceinline__ void tally(int i, int j, float v) const {
        if (v > iou_thr) {
            const unsigned rb = 1u << (i & 31), cb = 1u << (j & 31);
            const unsigned ro = atomicOr(&row_bits[i >> 5], rb), co = atomicOr(&col_bits[j >> 5], cb);
            if (!(ro & rb)) {                          // the row's first such pair (its only one whenever row_hit is read)
                row_hit[i] = (unsigned short)j;
                if (*(volatile int*)&flags[0] == 0) atomicOr(&flags[0], 1);
            }
            // (same-address atomics serialise: only the first such pair of a frame pays one)
            if (((ro & rb) || (co & cb)) && *(volatile int*)&flags[1] == 0) atomicOr(&flags[1], 1);
        }
    }
    __device__ __forceinline__ bool is_candidate(const Row& r, int i, int j, float thresh) const {
        const float v = iou(r, j);
        tally(i, j, v);
        return cost_from_iou(r, j, v) <= thresh;
    }
};

using OcmCost = OcmCostT<0>;

// DeepOC-SORT's first association (deepocsort.cpp:348-504): OcmCost plus the appearance term,
//   cost = -((iou + angle cost * score) + w(i, j) * emb(i, j)),  emb = detection . track embedding where iou > 0, else 0 (:421-423),
//   w = w_assoc_emb x adaptive row weight x adaptive column weight (compute_aw_max_metric :294-345) or plain w_assoc_emb (aw_off).
// The products and weights are prepared per frame by deep_embedding_terms() (ocsort_kernel.cuh).
struct DeepOcmCost {
    static constexpr bool kWarpPerRow = false;
    static constexpr bool kGrid = true;
    OcmCost base;
    const float* dense;               // [n][m] products, valid only where iou > 0
    const float* row_w;               // [n] w_assoc_emb x row weight, 0 for an all-zero row (mode 1)
    const float* col_w;               // [m] column weight (mode 1, n >= 2)
    const unsigned char* col_z;       // [m] column maximum is 0 (mode 1, n >= 2)
    int m;
    int mode;                         // 0: no appearance term (embedding_off), 1: adaptive weights, 2: w_assoc_emb only (aw_off)
    float w_assoc;
    bool cols_weighted;               // n >= 2 (:324)
    bool prune;
    struct Row : OcmCost::Row { int i; float rw; };
    __device__ __forceinline__ Row row(int i) const {
        Row r;
        static_cast<OcmCost::Row&>(r) = base.row(i);
        r.i = i;
        r.rw = (mode == 1) ? row_w[i] : w_assoc;
        return r;
    }
    __device__ __forceinline__ float4 col_box(int j) const { return base.col_box(j); }
    __device__ __forceinline__ bool reject(const Row& r, int j) const { return base.reject(r, j); }
    __device__ __forceinline__ float iou(const Row& r, int j) const { return base.iou(r, j); }
    __device__ __forceinline__ float emb_term(const Row& r, int j, float v) const {
        if (mode == 0 || v <= 0.0f) return 0.0f;
        float w = r.rw;
        if (mode == 1 && cols_weighted) w = col_z[j] ? 0.0f : xmul(w, col_w[j]);
        return xmul(w, dense[(size_t)r.i * m + j]);
    }
    __device__ __forceinline__ float cost_from_iou(const Row& r, int j, float v) const {
        float t = v;                                   // + (0 * angle * inertia * score) adds exactly +-0
        if (base.valid[j] & 1) t = xadd(v, xmul(ocm_angle_cost(r.cx, r.cy, base.ocm[j], 1.0f, base.inertia), r.score));
        return -xadd(t, emb_term(r, j, v));
    }
    __device__ __forceinline__ float cost(const Row& r, int j) const { return cost_from_iou(r, j, iou(r, j)); }
    __device__ __forceinline__ float pair(int i, int j) const { return cost(row(i), j); }
    __device__ __forceinline__ double pair_bias(int i, int j) const { return twin_bias(i, j); }
    __device__ __forceinline__ bool is_candidate(const Row& r, int i, int j, float thresh) const {
        const float v = iou(r, j);
        base.tally(i, j, v);
        return cost_from_iou(r, j, v) <= thresh;
    }
};

// Cost functor of the BYTE pass and the last-observation re-match: cost = -iou(detection, box of a track);
// rows / columns are LIST positions (the lists may hold an index twice, see ocsort_kernel.cuh).
template <int ASSO>
struct NegIouCostT {
    static constexpr bool kWarpPerRow = false;
    static constexpr bool kGrid = true;
    const float4* det_box;
    const unsigned short* row_map;    // row -> detection index
    const float4* trk_box;
    const unsigned short* col_map;    // column -> track position
    float iou_thr;
    bool prune;
    int* flags;                       // [0] any pair with iou > thr  (the reference's max_iou > threshold gate)
    float asso_norm;
    struct Row { float4 b; float area; };
    __device__ __forceinline__ Row row(int i) const {
        Row r;
        r.b = det_box[row_map[i]];
        r.area = box_area(r.b);
        return r;
    }
    __device__ __forceinline__ float4 col_box(int j) const { return trk_box[col_map[j]]; }
    __device__ __forceinline__ bool reject(const Row& r, int j) const { return prune && boxes_disjoint(r.b, col_box(j)); }
    __device__ __forceinline__ float cost(const Row& r, int j) const { return -asso_pair<ASSO>(asso_norm, r.b, r.area, col_box(j)); }
    __device__ __forceinline__ float pair(int i, int j) const { return cost(row(i), j); }
    __device__ __forceinline__ double pair_bias(int i, int j) const { return twin_bias(i, j); }   // twins, as OcmCost
    __device__ __forceinline__ bool is_candidate(const Row& r, int, int j, float thresh) const {
        const float v = asso_pair<ASSO>(asso_norm, r.b, r.area, col_box(j));
        if (v > iou_thr && *(volatile int*)&flags[0] == 0) atomicOr(&flags[0], 1);
        return -v <= thresh;
    }
};

}  // namespace mot
