// kernels_cost.cuh - standalone N x M cost-matrix kernels behind mot_cost_*().
//   iou_batch / iou_distance / fuse_score  reference include/motcpp/utils/iou.hpp:63-100,
//                                          src/utils/matching.cpp:62-65,130-143
//   OC-SORT observation-centric momentum cost   reference src/trackers/ocsort.cpp:628-679
// One thread per (track, 4 detections): the detection block is staged in shared memory as float4,
// track boxes sit in registers, the fp32 output is written with coalesced float4 stores - the
// output (4 B per pair) is the only HBM stream that matters.
#pragma once
#include "cost_device.cuh"
#include "ocm_device.cuh"
#include "gate_device.cuh"

namespace mot {

enum : int { kCostIou = 0, kCostIouDistance = 1, kCostIouDistanceFused = 2 };

constexpr int kCostTileCols = 512;     // detections staged per tile (8 KB of boxes + 2 KB of scores)
constexpr int kCostTileRows = 8;       // track rows per CTA pass (256 threads = 8 rows x 32 column-quads x 4)

// out is row-major (n x m) with leading dimension ld (ld % 4 == 0 and 16-byte aligned => float4 stores)
__global__ void __launch_bounds__(256) iou_cost_kernel(const float* __restrict__ a, int n, const float* __restrict__ b,
                                                       int m, const float* __restrict__ conf, float* __restrict__ out,
                                                       int ld, int mode) {
    __shared__ float4 s_box[kCostTileCols];
    __shared__ float s_conf[kCostTileCols];
    const int tid = (int)threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;                 // ty: row inside the 8-row group
    const int col_tiles = (m + kCostTileCols - 1) / kCostTileCols;
    const int row_groups = (n + kCostTileRows - 1) / kCostTileRows;
    const bool vec_ok = ((ld & 3) == 0) && ((((size_t)out) & 15) == 0);
    for (int ct = (int)blockIdx.y; ct < col_tiles; ct += (int)gridDim.y) {
        const int c0 = ct * kCostTileCols;
        const int cn = min(kCostTileCols, m - c0);
        __syncthreads();
        for (int k = tid; k < cn; k += 256) {
            s_box[k] = *reinterpret_cast<const float4*>(b + (size_t)(c0 + k) * 4);
            s_conf[k] = (mode == kCostIouDistanceFused) ? conf[c0 + k] : 1.0f;
        }
        __syncthreads();
        for (int rg = (int)blockIdx.x; rg < row_groups; rg += (int)gridDim.x) {
            const int i = rg * kCostTileRows + ty;
            if (i >= n) continue;
            const float4 ra = *reinterpret_cast<const float4*>(a + (size_t)i * 4);
            const float area = box_area(ra);
            float* orow = out + (size_t)i * ld + c0;
            for (int q = tx * 4; q < cn; q += 128) {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = q + e;
                    float r = 0.0f;
                    if (j < cn) {
                        r = iou_pair(ra, area, s_box[j]);
                        if (mode != kCostIou) r = xsub(1.0f, r);
                        if (mode == kCostIouDistanceFused) r = xsub(1.0f, xmul(xsub(1.0f, r), s_conf[j]));
                    }
                    v[e] = r;
                }
                if (vec_ok && q + 3 < cn) {
                    *reinterpret_cast<float4*>(orow + q) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (q + e < cn) orow[q + e] = v[e];
                }
            }
        }
    }
}

// OC-SORT association cost -(iou + angle cost * score) for every (detection, track) pair
// (ocsort_assoc::associate, ocsort.cpp:617-700).  Rows = detections [x1,y1,x2,y2,score], columns = tracks:
// predicted box trks4, velocity vel2 = (dy, dx), k_previous_obs prev5 = [x1,y1,x2,y2,conf] (conf < 0 = none).
// out_cost / out_iou (nullable) are (n_dets x n_trks) row-major with leading dimension ld.  One thread per
// (detection, 4 tracks); the track tile sits in shared memory.  Issue-bound (one fp64 acos per pair), the 4 B
// per pair written is the only HBM stream.
__global__ void __launch_bounds__(256) ocm_cost_kernel(const float* __restrict__ dets5, int n_dets,
                                                       const float* __restrict__ trks4, const float* __restrict__ vel2,
                                                       const float* __restrict__ prev5, int n_trks, float inertia,
                                                       float* __restrict__ out_cost, float* __restrict__ out_iou, int ld) {
    __shared__ float4 s_box[kCostTileCols];
    __shared__ float4 s_ocm[kCostTileCols];
    __shared__ float s_valid[kCostTileCols];
    const int tid = (int)threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;
    const int col_tiles = (n_trks + kCostTileCols - 1) / kCostTileCols;
    const int row_groups = (n_dets + kCostTileRows - 1) / kCostTileRows;
    const bool vec_ok = ((ld & 3) == 0) && ((((size_t)out_cost) & 15) == 0) && ((((size_t)out_iou) & 15) == 0);
    for (int ct = (int)blockIdx.y; ct < col_tiles; ct += (int)gridDim.y) {
        const int c0 = ct * kCostTileCols;
        const int cn = min(kCostTileCols, n_trks - c0);
        __syncthreads();
        for (int k = tid; k < cn; k += 256) {
            const float* p = prev5 + (size_t)(c0 + k) * 5;
            s_box[k] = *reinterpret_cast<const float4*>(trks4 + (size_t)(c0 + k) * 4);
            s_ocm[k] = make_float4(xdiv(xadd(p[0], p[2]), 2.0f), xdiv(xadd(p[1], p[3]), 2.0f), vel2[(size_t)(c0 + k) * 2],
                                   vel2[(size_t)(c0 + k) * 2 + 1]);
            s_valid[k] = (p[4] >= 0.0f) ? 1.0f : 0.0f;
        }
        __syncthreads();
        for (int rg = (int)blockIdx.x; rg < row_groups; rg += (int)gridDim.x) {
            const int i = rg * kCostTileRows + ty;
            if (i >= n_dets) continue;
            const float* d = dets5 + (size_t)i * 5;
            const float4 rb = make_float4(d[0], d[1], d[2], d[3]);
            const float score = d[4];
            const float area = box_area(rb);
            const float cx = xdiv(xadd(rb.x, rb.z), 2.0f), cy = xdiv(xadd(rb.y, rb.w), 2.0f);
            float* crow = out_cost + (size_t)i * ld + c0;
            float* irow = out_iou ? out_iou + (size_t)i * ld + c0 : nullptr;
            for (int q = tx * 4; q < cn; q += 128) {
                float vc[4], vi[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = q + e;
                    vc[e] = 0.0f; vi[e] = 0.0f;
                    if (j < cn) {
                        const float iou = iou_pair(rb, area, s_box[j]);
                        const float ac = xmul(ocm_angle_cost(cx, cy, s_ocm[j], s_valid[j], inertia), score);
                        vi[e] = iou;
                        vc[e] = -xadd(iou, ac);
                    }
                }
                if (vec_ok && q + 3 < cn) {
                    *reinterpret_cast<float4*>(crow + q) = make_float4(vc[0], vc[1], vc[2], vc[3]);
                    if (irow) *reinterpret_cast<float4*>(irow + q) = make_float4(vi[0], vi[1], vi[2], vi[3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (q + e < cn) { crow[q + e] = vc[e]; if (irow) irow[q + e] = vi[e]; }
                }
            }
        }
    }
}

// linear_assignment::gate_cost_matrix (reference src/trackers/strongsort.cpp:451-492) IN PLACE on cost (n_tracks x n_meas,
// leading dimension ld): Mahalanobis gate (entries with gating distance > 9.4877 become gated_cost) followed by the
// motion blend mc_lambda * cost + (1 - mc_lambda) * gating distance.  recs = XYAH records (72 floats per track),
// meas4 = (n_meas x 4) xyah rows.  One warp per track row: lane 0's projection + Cholesky factor is broadcast, the 32
// lanes then sweep the row with coalesced read-modify-writes.  8 B per pair of HBM traffic, ~45 exact-fp32 operations
// (7 IEEE divisions) per pair: issue-bound, like iou_cost_kernel.
__global__ void __launch_bounds__(256) gate_cost_kernel(float* __restrict__ cost, int ld, const float* __restrict__ recs,
                                                        int n_tracks, const float* __restrict__ meas4, int n_meas,
                                                        float mc_lambda, float gated_cost, int only_position) {
    const int warps_per_cta = (int)blockDim.x >> 5;
    const int lane = lane_id();
    for (int i = (int)blockIdx.x * warps_per_cta + warp_id(); i < n_tracks; i += (int)gridDim.x * warps_per_cta) {
        const GateRow g = gate_prepare(recs + (size_t)i * kRecFloats);     // same values in every lane (loads broadcast)
        float* row = cost + (size_t)i * ld;
        for (int j = lane; j < n_meas; j += 32) {
            const float4 z = *reinterpret_cast<const float4*>(meas4 + (size_t)j * 4);
            row[j] = gate_blend(row[j], gate_distance(g, z, only_position != 0), mc_lambda, gated_cost);
        }
    }
}

// iou_matching::iou_cost (reference src/trackers/strongsort.cpp:502-585): 1 - IoU of tlwh boxes with the reference's
// `union > 1e-6` guard; rows whose time_since_update > 1 are INFTY_COST (:567-570).  Same tiling as iou_cost_kernel.
__global__ void __launch_bounds__(256) iou_tlwh_cost_kernel(const float* __restrict__ trk, const int* __restrict__ tsu, int n,
                                                            const float* __restrict__ det, int m, float* __restrict__ out,
                                                            int ld) {
    __shared__ float4 s_box[kCostTileCols];
    const int tid = (int)threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;
    const int col_tiles = (m + kCostTileCols - 1) / kCostTileCols;
    const int row_groups = (n + kCostTileRows - 1) / kCostTileRows;
    const bool vec_ok = ((ld & 3) == 0) && ((((size_t)out) & 15) == 0);
    for (int ct = (int)blockIdx.y; ct < col_tiles; ct += (int)gridDim.y) {
        const int c0 = ct * kCostTileCols;
        const int cn = min(kCostTileCols, m - c0);
        __syncthreads();
        for (int k = tid; k < cn; k += 256) s_box[k] = *reinterpret_cast<const float4*>(det + (size_t)(c0 + k) * 4);
        __syncthreads();
        for (int rg = (int)blockIdx.x; rg < row_groups; rg += (int)gridDim.x) {
            const int i = rg * kCostTileRows + ty;
            if (i >= n) continue;
            const float4 rb = *reinterpret_cast<const float4*>(trk + (size_t)i * 4);
            const bool stale = tsu != nullptr && tsu[i] > 1;
            float* orow = out + (size_t)i * ld + c0;
            for (int q = tx * 4; q < cn; q += 128) {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = q + e;
                    v[e] = 0.0f;
                    if (j < cn) v[e] = stale ? kInftyCost : xsub(1.0f, iou_tlwh_pair(rb, s_box[j]));
                }
                if (vec_ok && q + 3 < cn) {
                    *reinterpret_cast<float4*>(orow + q) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (q + e < cn) orow[q + e] = v[e];
                }
            }
        }
    }
}

}  // namespace mot
