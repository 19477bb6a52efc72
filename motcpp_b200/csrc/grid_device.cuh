// grid_device.cuh - a uniform 32 x 16 grid over a set of boxes in shared memory, rebuilt per use.
// The association costs of the reference (iou_batch, include/motcpp/utils/iou.hpp:63-100) are
// evaluated for ALL N x M pairs; a pair whose boxes are disjoint has IoU exactly 0, cost exactly 1,
// and can never be an assignment candidate (thresh < 1) nor a duplicate (distance < 0.15).  The
// grid only decides which pairs are *looked at*: every pair of overlapping boxes is still visited
// exactly once (in the cell holding the top-left corner of their intersection), so results are
// identical to the dense evaluation.
#pragma once
#include "block_utils.cuh"

namespace mot {

constexpr int kGridX = 32;
constexpr int kGridY = 16;
constexpr int kGridCells = kGridX * kGridY;

struct BoxGrid {
    int* cell;                 // [kGridCells + 1] start offset of every cell in items[] (exclusive scan)
    int* cursor;               // [kGridCells] build-time fill cursors
    unsigned short* items;     // [cap] column indices grouped by cell
    float* red;                // [4 * 32] block-reduction scratch
    int cap;
    float x0, y0, sx, sy;
    int valid;                 // 0: grid unusable for this set (too many entries) -> visit all pairs

    __device__ __forceinline__ int cx(float x) const {
        return (int)fminf(fmaxf(xmul(xsub(x, x0), sx), 0.0f), (float)(kGridX - 1));
    }
    __device__ __forceinline__ int cy(float y) const {
        return (int)fminf(fmaxf(xmul(xsub(y, y0), sy), 0.0f), (float)(kGridY - 1));
    }
};

MOT_HD constexpr size_t grid_smem_bytes(int cap) {
    return ((sizeof(int) * (kGridCells + 1) + 15) & ~(size_t)15) + sizeof(int) * kGridCells +
           ((sizeof(unsigned short) * (size_t)cap + 15) & ~(size_t)15) + sizeof(float) * 4 * 32;
}

__device__ __forceinline__ unsigned char* grid_carve(unsigned char* p, int cap, BoxGrid& g) {
    g.cell = (int*)p;               p += (sizeof(int) * (kGridCells + 1) + 15) & ~(size_t)15;
    g.cursor = (int*)p;             p += sizeof(int) * kGridCells;
    g.items = (unsigned short*)p;   p += (sizeof(unsigned short) * (size_t)cap + 15) & ~(size_t)15;
    g.red = (float*)p;              p += sizeof(float) * 4 * 32;
    g.cap = cap;
    g.valid = 0;
    return p;
}

__device__ __forceinline__ bool box_finite(float4 b) {
    const float s = (b.x - b.x) + (b.y - b.y) + (b.z - b.z) + (b.w - b.w);     // NaN/inf -> NaN
    return s == 0.0f;
}

// Build the grid over boxes box_of(0..n).  All threads of the block must call.  Non-finite boxes
// are left out (their IoU is NaN: never a candidate).
template <class BoxOf>
__device__ __forceinline__ void grid_build(BoxGrid& g, int n, BlockScratch* bs, BoxOf box_of) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    float lo_x = 3.0e38f, lo_y = 3.0e38f, hi_x = -3.0e38f, hi_y = -3.0e38f;
    for (int j = tid; j < n; j += nt) {
        const float4 b = box_of(j);
        if (!box_finite(b)) continue;
        lo_x = fminf(lo_x, b.x); lo_y = fminf(lo_y, b.y);
        hi_x = fmaxf(hi_x, b.z); hi_y = fmaxf(hi_y, b.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo_x = fminf(lo_x, __shfl_xor_sync(kFullMask, lo_x, o));
        lo_y = fminf(lo_y, __shfl_xor_sync(kFullMask, lo_y, o));
        hi_x = fmaxf(hi_x, __shfl_xor_sync(kFullMask, hi_x, o));
        hi_y = fmaxf(hi_y, __shfl_xor_sync(kFullMask, hi_y, o));
    }
    __syncthreads();
    if (lane == 0) { g.red[warp] = lo_x; g.red[32 + warp] = lo_y; g.red[64 + warp] = hi_x; g.red[96 + warp] = hi_y; }
    for (int c = tid; c <= kGridCells; c += nt) g.cell[c] = 0;
    for (int c = tid; c < kGridCells; c += nt) g.cursor[c] = 0;
    __syncthreads();
    for (int w = 0; w < nwarps; ++w) {
        lo_x = fminf(lo_x, g.red[w]); lo_y = fminf(lo_y, g.red[32 + w]);
        hi_x = fmaxf(hi_x, g.red[64 + w]); hi_y = fmaxf(hi_y, g.red[96 + w]);
    }
    const float ex = hi_x - lo_x, ey = hi_y - lo_y;
    g.x0 = lo_x; g.y0 = lo_y;
    g.sx = (ex > 0.0f) ? (float)kGridX / ex : 0.0f;
    g.sy = (ey > 0.0f) ? (float)kGridY / ey : 0.0f;
    if (!(g.sx == g.sx) || !(g.sy == g.sy)) { g.sx = 0.0f; g.sy = 0.0f; }
    // counts
    for (int j = tid; j < n; j += nt) {
        const float4 b = box_of(j);
        if (!box_finite(b)) continue;
        const int cx0 = g.cx(b.x), cx1 = g.cx(b.z), cy0 = g.cy(b.y), cy1 = g.cy(b.w);
        for (int yy = cy0; yy <= cy1; ++yy)
            for (int xx = cx0; xx <= cx1; ++xx) atomicAdd(&g.cell[yy * kGridX + xx], 1);
    }
    const int total = block_exclusive_scan(g.cell, kGridCells, bs, true);
    g.valid = (total <= g.cap) ? 1 : 0;
    if (g.valid) {
        for (int j = tid; j < n; j += nt) {
            const float4 b = box_of(j);
            if (!box_finite(b)) continue;
            const int cx0 = g.cx(b.x), cx1 = g.cx(b.z), cy0 = g.cy(b.y), cy1 = g.cy(b.w);
            for (int yy = cy0; yy <= cy1; ++yy)
                for (int xx = cx0; xx <= cx1; ++xx) {
                    const int c = yy * kGridX + xx;
                    g.items[g.cell[c] + atomicAdd(&g.cursor[c], 1)] = (unsigned short)j;
                }
        }
    }
    __syncthreads();
}

// Visit every column j whose box overlaps row box `a` (interior intersection), exactly once.
// visit(j, box_j) is called for those pairs only.
template <class BoxOf, class Visit>
__device__ __forceinline__ void grid_query(const BoxGrid& g, float4 a, BoxOf box_of, Visit visit) {
    const int cx0 = g.cx(a.x), cx1 = g.cx(a.z), cy0 = g.cy(a.y), cy1 = g.cy(a.w);
    for (int yy = cy0; yy <= cy1; ++yy)
        for (int xx = cx0; xx <= cx1; ++xx) {
            const int c = yy * kGridX + xx;
            const int e1 = g.cell[c + 1];
            for (int e = g.cell[c]; e < e1; ++e) {
                const int j = g.items[e];
                const float4 b = box_of(j);
                const float ix = fmaxf(a.x, b.x), iy = fmaxf(a.y, b.y);
                if (!((fminf(a.z, b.z) > ix) && (fminf(a.w, b.w) > iy))) continue;      // disjoint
                if (g.cx(ix) != xx || g.cy(iy) != yy) continue;                       // counted in another cell
                visit(j, b);
            }
        }
}

}  // namespace mot
