// kernels_pack.cuh - compaction of the padded per-frame output blocks [frames][ld_out][8] into one contiguous run of
// valid rows, so that the host-buffer path (mot_engine_update_host_packed) moves only what BaseTracker::update would
// have returned (reference: the (M, 8) matrix built at the end of every update(), e.g. src/trackers/bytetrack.cpp:596-620).
// HBM-bound, trivial next to the step kernel: 32 B read + 32 B written per valid row.
#pragma once
#include "simt.cuh"

namespace mot {

// off[0..n] = exclusive prefix sums of min(n_out[f], ld_out) over the n frames of a chunk; one CTA.
static __global__ void __launch_bounds__(1024) pack_scan_kernel(const int* __restrict__ n_out, int n, int ld_out, int* __restrict__ off) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int f = base + tid;
        const int v = f < n ? min(n_out[f], ld_out) : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        const int ws = warp_sum[lane];
        const int before = __reduce_add_sync(kFullMask, lane < warp ? ws : 0);
        const int total = __reduce_add_sync(kFullMask, ws);
        const int c = carry;
        if (f < n) off[f] = c + before + incl - v;
        __syncthreads();
        if (tid == 0) carry = c + total;
        __syncthreads();
    }
    if (tid == 0) off[n] = carry;
}

// one CTA per frame block (grid-stride): rows [0, n_out[f]) of padded block f -> packed + off[f] * 8
static __global__ void __launch_bounds__(256) pack_rows_kernel(const float4* __restrict__ padded, const int* __restrict__ n_out,
                                                               const int* __restrict__ off, int n, int ld_out,
                                                               float4* __restrict__ packed) {
    for (int f = (int)blockIdx.x; f < n; f += (int)gridDim.x) {
        const int q = 2 * min(n_out[f], ld_out);                      // float4s: 8 floats per row
        const float4* src = padded + (size_t)f * ld_out * 2;
        float4* dst = packed + (size_t)off[f] * 2;
        for (int k = (int)threadIdx.x; k < q; k += (int)blockDim.x) dst[k] = src[k];
    }
}

}  // namespace mot
