// kernels_lap.cuh - standalone batched linear-assignment kernel behind mot_lap() / mot_lap_batch().
// One CTA per problem; see lap_device.cuh for the algorithm and the reference lines it replaces.
#pragma once
#include "lap_device.cuh"
#include "jv_device.cuh"
#include "jv_block_device.cuh"

namespace mot {

struct LapBatchArgs {
    const float* cost;      // problem p at cost + p * stride_cost, row-major (n x m), leading dim ld
    long long stride_cost;
    const int* n_rows;      // per-problem sizes (nullptr => n for all)
    const int* n_cols;
    int n, m, ld;
    float thresh;
    int* row2col;           // problem p at row2col + p * n_max
    int* col2row;           // problem p at col2row + p * m_max
    unsigned char* gscratch;// problem p at gscratch + p * lap_gscratch_bytes(n_max, m_max)
    int n_max, m_max, e_cap;
    int n_problems;
};

__global__ void __launch_bounds__(256) lap_dense_kernel(LapBatchArgs a) {
    MOT_DYNAMIC_SMEM(smem);
    LapWorkspace ws;
    lap_carve(smem, a.n_max, a.m_max, a.e_cap, ws);
    for (int p = (int)blockIdx.x; p < a.n_problems; p += (int)gridDim.x) {
        lap_carve_gscratch(a.gscratch + (size_t)p * lap_gscratch_bytes(a.n_max, a.m_max), a.n_max, a.m_max, ws);
        const int n = a.n_rows ? a.n_rows[p] : a.n;
        const int m = a.n_cols ? a.n_cols[p] : a.m;
        MatrixCost cost{a.cost + (size_t)p * a.stride_cost, a.ld};
        block_lap(ws, n, m, a.n_max, a.m_max, a.thresh, cost);
        int* r2c = a.row2col + (size_t)p * a.n_max;
        int* c2r = a.col2row + (size_t)p * a.m_max;
        for (int i = (int)threadIdx.x; i < n; i += (int)blockDim.x) r2c[i] = (int)ws.row2col[i];
        for (int j = (int)threadIdx.x; j < m; j += (int)blockDim.x) c2r[j] = (int)ws.col2row[j];
        __syncthreads();
    }
}

// The reference's dense LAPJV itself (jv_device.cuh), one warp per problem, rows + columns <= kLapJvMax: reproduces
// utils::linear_assignment INCLUDING its scan-order tie-breaking (what the sparse solver above does not promise).
constexpr int kLapJvMax = 384;

__global__ void __launch_bounds__(32) lap_jv_kernel(const float* __restrict__ cost, long long stride_cost, int n_problems, int n, int m,
                                                    int ld, float thresh, int* __restrict__ row2col, int* __restrict__ col2row) {
    MOT_DYNAMIC_SMEM(work);                                   // jv_work_bytes(n + m + 1)
    const JvWork w = jv_carve(work, n + m + 1);
    for (int p = (int)blockIdx.x; p < n_problems; p += (int)gridDim.x) {
        warp_dense_lapjv(JvCost{cost + (size_t)p * stride_cost, n, m, ld, (double)thresh / 2.0}, n + m, w);
        for (int i = (int)threadIdx.x; i < n; i += 32) { const int j = w.x[i]; row2col[(size_t)p * n + i] = j < m ? j : -1; }
        for (int j = (int)threadIdx.x; j < m; j += 32) { const int i = w.y[j]; col2row[(size_t)p * m + j] = i < n ? i : -1; }
        __syncwarp();
    }
}


// The same algorithm for ANY size, one CTA per problem (jv_block_device.cuh): work arrays in global scratch
// (problem slot = blockIdx.x), the scan-order permutation in dynamic shared memory.
constexpr int kLapJvBlockThreads = 512;
__global__ void __launch_bounds__(kLapJvBlockThreads) lap_jv_block_kernel(const float* __restrict__ cost, long long stride_cost,
                                                                          int n_problems, int n, int m, int ld, float thresh,
                                                                          int* __restrict__ row2col, int* __restrict__ col2row,
                                                                          unsigned char* __restrict__ gscratch, int all_shared) {
    MOT_DYNAMIC_SMEM(smem);
    __shared__ BlockScratch bs;
    const int N = n + m;
    const JvBlockWork w = jv_block_carve(all_shared ? nullptr : gscratch + (size_t)blockIdx.x * jv_block_gbytes(N), smem, N, all_shared != 0);
    for (int p = (int)blockIdx.x; p < n_problems; p += (int)gridDim.x) {
        block_dense_lapjv(JvCost{cost + (size_t)p * stride_cost, n, m, ld, (double)thresh / 2.0}, N, w, &bs);
        for (int i = (int)threadIdx.x; i < n; i += (int)blockDim.x) { const int j = w.x[i]; row2col[(size_t)p * n + i] = j < m ? j : -1; }
        for (int j = (int)threadIdx.x; j < m; j += (int)blockDim.x) { const int i = w.y[j]; col2row[(size_t)p * m + j] = i < n ? i : -1; }
        __syncthreads();
    }
}

}  // namespace mot
