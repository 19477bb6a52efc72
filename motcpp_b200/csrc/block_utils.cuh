// block_utils.cuh - block-wide stable compaction and exclusive scan built from warp ballots /
// shuffles.  All functions must be called by every thread of the block (they synchronise).
#pragma once
#include "simt.cuh"

namespace mot {

// Scratch the block-wide primitives need: one int per warp plus a running base.
struct BlockScratch {
    int warp_sum[32];
    int base;
    int pad[3];
};

// Stable compaction of the indices k in [0, n) for which pred(k) is true.  emit(k, pos) is called
// by the thread that owns k with pos = rank of k among the kept indices (ascending k), offset by
// `start`.  Returns start + number kept (same value in every thread).
template <class Pred, class Emit>
__device__ __forceinline__ int block_compact(int n, int start, BlockScratch* bs, Pred pred, Emit emit) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    __syncthreads();                       // protect bs against a previous use
    if (tid == 0) bs->base = start;
    __syncthreads();
    for (int tile = 0; tile < n; tile += nt) {
        const int k = tile + tid;
        const bool keep = (k < n) && pred(k);
        const unsigned ballot = __ballot_sync(kFullMask, keep);
        const int within = __popc(ballot & ((1u << lane) - 1u));
        if (lane == 0) bs->warp_sum[warp] = __popc(ballot);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < nwarps; ++w) {
            const int s = bs->warp_sum[w];
            if (w < warp) before += s;
            total += s;
        }
        const int base = bs->base;
        if (keep) emit(k, base + before + within);
        __syncthreads();
        if (tid == 0) bs->base = base + total;
        __syncthreads();
    }
    return bs->base;
}

// In-place exclusive scan of data[0..n) (shared or global memory); writes the grand total to
// data[n] when write_total is set (the array must then hold n + 1 entries).  Returns the total.
__device__ __forceinline__ int block_exclusive_scan(int* data, int n, BlockScratch* bs, bool write_total) {
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    __syncthreads();
    if (tid == 0) bs->base = 0;
    __syncthreads();
    for (int tile = 0; tile < n; tile += nt) {
        const int k = tile + tid;
        const int v = (k < n) ? data[k] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) bs->warp_sum[warp] = incl;
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < nwarps; ++w) {
            const int s = bs->warp_sum[w];
            if (w < warp) before += s;
            total += s;
        }
        const int base = bs->base;
        if (k < n) data[k] = base + before + incl - v;
        __syncthreads();
        if (tid == 0) bs->base = base + total;
        __syncthreads();
    }
    const int total = bs->base;
    if (write_total && tid == 0) data[n] = total;
    __syncthreads();
    return total;
}

}  // namespace mot
