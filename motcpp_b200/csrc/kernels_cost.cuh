// kernels_cost.cuh - standalone N x M cost-matrix kernels behind mot_cost_*().
//   iou_batch / iou_distance / fuse_score  reference include/motcpp/utils/iou.hpp:63-100,
//                                          src/utils/matching.cpp:62-65,130-143
//   OC-SORT observation-centric momentum cost   reference src/trackers/ocsort.cpp:628-679
// One thread per (track, 4 detections): the detection block is staged in shared memory as float4,
// track boxes sit in registers, the fp32 output is written with coalesced float4 stores - the
// output (4 B per pair) is the only HBM stream that matters.
#pragma once
#include "cost_device.cuh"

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

}  // namespace mot
