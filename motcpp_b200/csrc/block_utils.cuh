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

// Stable compaction of the indices k in [0, n) for which pred(k) is true.  Every thread takes a
// contiguous run of ceil(n / blockDim) indices (at most 32), so one block-wide scan serves any n
// up to 32 * blockDim.  emit(k, pos) is called by the thread that owns k, with pos = start + rank
// of k among the kept indices (ascending k).  Returns start + number kept, in every thread; the
// emitted data is visible to the whole block on return.  pred must be side-effect free.
template <class Pred, class Emit>
__device__ __forceinline__ int block_compact(int n, int start, BlockScratch* bs, Pred pred, Emit emit) {
    if (n <= 0) return start;
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int items = (n + nt - 1) / nt;
    const int lo = tid * items;
    unsigned keep = 0;
    for (int q = 0; q < items; ++q) {
        const int k = lo + q;
        if (k < n && pred(k)) keep |= (1u << q);
    }
    const int cnt = __popc(keep);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();                       // bs may still be read by a previous call
    if (lane == 31) bs->warp_sum[warp] = incl;
    __syncthreads();
    // every warp sums the per-warp counts itself: lane w holds warp w's count, two redux.sync do the rest
    const int ws = (lane < nwarps) ? bs->warp_sum[lane] : 0;
    const int total = __reduce_add_sync(kFullMask, ws);
    const int before = __reduce_add_sync(kFullMask, (lane < warp) ? ws : 0);
    int pos = start + before + incl - cnt;
    for (int q = 0; q < items; ++q)
        if ((keep >> q) & 1u) emit(lo + q, pos++);
    __syncthreads();
    return start + total;
}

// Two stable compactions of the SAME index range in one pass (one scan of two packed 16-bit counters, three barriers
// instead of six): indices with pred0 go through emit0 from position start0, indices with pred1 through emit1 from start1.
// The predicates need not be exclusive.  end0 / end1 receive start + number kept.  n <= 32 * blockDim and < 65536.
template <class Pred0, class Pred1, class Emit0, class Emit1>
__device__ __forceinline__ void block_compact2(int n, int start0, int start1, BlockScratch* bs, Pred0 pred0, Pred1 pred1,
                                               Emit0 emit0, Emit1 emit1, int& end0, int& end1) {
    end0 = start0; end1 = start1;
    if (n <= 0) return;
    const int nt = (int)blockDim.x, tid = (int)threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int items = (n + nt - 1) / nt;
    const int lo = tid * items;
    unsigned keep0 = 0, keep1 = 0;
    for (int q = 0; q < items; ++q) {
        const int k = lo + q;
        if (k < n) {
            if (pred0(k)) keep0 |= (1u << q);
            if (pred1(k)) keep1 |= (1u << q);
        }
    }
    const int cnt = __popc(keep0) | (__popc(keep1) << 16);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();                       // bs may still be read by a previous call
    if (lane == 31) bs->warp_sum[warp] = incl;
    __syncthreads();
    const int ws = (lane < nwarps) ? bs->warp_sum[lane] : 0;
    const int total = __reduce_add_sync(kFullMask, ws);
    const int before = __reduce_add_sync(kFullMask, (lane < warp) ? ws : 0);
    const int excl = before + incl - cnt;
    int pos0 = start0 + (excl & 0xffff), pos1 = start1 + (excl >> 16);
    for (int q = 0; q < items; ++q) {
        if ((keep0 >> q) & 1u) emit0(lo + q, pos0++);
        if ((keep1 >> q) & 1u) emit1(lo + q, pos1++);
    }
    __syncthreads();
    end0 = start0 + (total & 0xffff);
    end1 = start1 + (total >> 16);
}

// Exclusive scan of one value per thread; *total receives the block-wide sum.  Two barriers; all threads must call.
__device__ __forceinline__ int block_exclusive_scan_value(int v, BlockScratch* bs, int* total) {
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (int)blockDim.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();                       // bs may still be read by a previous call
    if (lane == 31) bs->warp_sum[warp] = incl;
    __syncthreads();
    const int ws = (lane < nwarps) ? bs->warp_sum[lane] : 0;
    *total = __reduce_add_sync(kFullMask, ws);
    return __reduce_add_sync(kFullMask, (lane < warp) ? ws : 0) + incl - v;
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
        const int ws = (lane < nwarps) ? bs->warp_sum[lane] : 0;
        const int total = __reduce_add_sync(kFullMask, ws);
        const int before = __reduce_add_sync(kFullMask, (lane < warp) ? ws : 0);
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
