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

// ---- deepocsort_assoc::compute_aw_max_metric (reference src/trackers/deepocsort.cpp:294-345): the embedding cost is
// re-weighted by how distinctive the best entry of its row and of its column is: weight = 1 - max(second / max - bottom, 0)
// / (1 - bottom), zero when the maximum is zero.  Three streaming passes over the (n x m) matrix: row top-2 (one warp
// per row, coalesced), column top-2 (32 columns per CTA, 8 row phases merged through shared memory), element-wise product.
struct Top2 { float mx, se; };
__device__ __forceinline__ void top2_push(Top2& t, float v) { if (v > t.mx) { t.se = t.mx; t.mx = v; } else if (v > t.se) t.se = v; }
__device__ __forceinline__ void top2_merge(Top2& a, const Top2& b) { top2_push(a, b.mx); top2_push(a, b.se); }
// weight as the reference writes it; encodes "row / column is all-zero-max" as a NaN-free flag value < 0 is impossible, so
// the zero case is carried separately: returns the weight, *zero = max == 0
__device__ __forceinline__ float aw_weight(const Top2& t, float bottom, bool* zero) {
    *zero = (t.mx == 0.0f);
    return xsub(1.0f, xdiv(fmaxf(xsub(xdiv(t.se, t.mx), bottom), 0.0f), xsub(1.0f, bottom)));
}

__global__ void __launch_bounds__(256) aw_row_top2_kernel(const float* __restrict__ emb, int n, int m, int ld, float bottom,
                                                          float* __restrict__ row_w, unsigned char* __restrict__ row_zero) {
    const int lane = lane_id(), wpc = (int)blockDim.x >> 5;
    for (int i = (int)blockIdx.x * wpc + warp_id(); i < n; i += (int)gridDim.x * wpc) {
        Top2 t{-INFINITY, -INFINITY};
        const float* row = emb + (size_t)i * ld;
        if ((ld & 3) == 0 && ((((size_t)emb) & 15) == 0)) {                       // 16 B per lane and iteration, four iterations in flight
            const float4* r4 = reinterpret_cast<const float4*>(row);
            const int m4 = m >> 2;
#pragma unroll 4
            for (int q = lane; q < m4; q += 32) { const float4 v = r4[q]; top2_push(t, v.x); top2_push(t, v.y); top2_push(t, v.z); top2_push(t, v.w); }
            for (int j = (m4 << 2) + lane; j < m; j += 32) top2_push(t, row[j]);
        } else {
            for (int j = lane; j < m; j += 32) top2_push(t, row[j]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Top2 b{__shfl_xor_sync(kFullMask, t.mx, o), __shfl_xor_sync(kFullMask, t.se, o)};
            top2_merge(t, b);
        }
        if (lane == 0) { bool z; row_w[i] = aw_weight(t, bottom, &z); row_zero[i] = z ? 1 : 0; }
    }
}

__global__ void __launch_bounds__(256) aw_col_top2_kernel(const float* __restrict__ emb, int n, int m, int ld, float bottom,
                                                          float* __restrict__ col_w, unsigned char* __restrict__ col_zero) {
    __shared__ Top2 part[8][32];
    const int tx = (int)threadIdx.x & 31, ty = (int)threadIdx.x >> 5;
    for (int c0 = (int)blockIdx.x * 32; c0 < m; c0 += (int)gridDim.x * 32) {
        const int j = c0 + tx;
        Top2 t{-INFINITY, -INFINITY};
        if (j < m) {
            int i = ty;
            for (; i + 56 < n; i += 64) {                                          // eight independent row loads in flight
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = emb[(size_t)(i + 8 * u) * ld + j];
#pragma unroll
                for (int u = 0; u < 8; ++u) top2_push(t, v[u]);
            }
            for (; i < n; i += 8) top2_push(t, emb[(size_t)i * ld + j]);
        }
        part[ty][tx] = t;
        __syncthreads();
        if (ty == 0 && j < m) {
            for (int k = 1; k < 8; ++k) top2_merge(t, part[k][tx]);
            bool z;
            col_w[j] = aw_weight(t, bottom, &z);
            col_zero[j] = z ? 1 : 0;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) aw_apply_kernel(const float* __restrict__ emb, int n, int m, int ld, float w_assoc,
                                                       const float* __restrict__ row_w, const unsigned char* __restrict__ row_zero,
                                                       const float* __restrict__ col_w, const unsigned char* __restrict__ col_zero,
                                                       float* __restrict__ out, int ld_out) {
    const long long total = (long long)n * m;
    const bool rows_on = m >= 2, cols_on = n >= 2;                             // fewer than two entries: no weighting (:311, :333)
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(k / m), j = (int)(k - (long long)i * m);
        float w = w_assoc;
        if (rows_on) w = row_zero[i] ? 0.0f : xmul(w, row_w[i]);
        if (cols_on) w = col_zero[j] ? 0.0f : xmul(w, col_w[j]);
        out[(size_t)i * ld_out + j] = xmul(w, emb[(size_t)i * ld + j]);
    }
}

// hmiou_batch / giou_batch / diou_batch / centroid_batch (reference include/motcpp/utils/iou.hpp:119-330),
// evaluated pair-wise - the reference's own expressions only line up when the second set has one row (trap 11), which is
// the domain on which parity is defined.  Same tiling as iou_cost_kernel.
__global__ void __launch_bounds__(256) iou_variant_kernel(const float* __restrict__ a, int n, const float* __restrict__ b, int m,
                                                          int kind, float norm, float* __restrict__ out, int ld) {
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
        for (int k = tid; k < cn; k += 256) s_box[k] = *reinterpret_cast<const float4*>(b + (size_t)(c0 + k) * 4);
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
                for (int e = 0; e < 4; ++e) v[e] = (q + e < cn) ? iou_variant_pair(kind, ra, area, s_box[q + e], norm) : 0.0f;
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
