// grid_device.cuh - a uniform 32 x 16 grid over the TOP-LEFT CORNERS of a set of boxes, rebuilt in
// shared memory per use.
// The association costs of the reference (iou_batch, include/motcpp/utils/iou.hpp:63-100) are
// evaluated for ALL N x M pairs; a pair whose boxes are disjoint has IoU exactly 0, cost exactly 1,
// and can never be an assignment candidate (thresh < 1) nor a duplicate (distance < 0.15).  The
// grid only decides which pairs are *looked at*: each column box sits in exactly one cell, a row
// box scans the cells that can hold the corner of a box overlapping it (its own extent dilated by
// the largest column box), so every overlapping pair is visited exactly once and the results are
// identical to the dense evaluation.
// Outliers: one runaway column box (a coasting Kalman track whose height grew for 30 frames) would dilate EVERY row's
// window.  Callers that know what a sane box is (the detections' own extent) build and query with kBig = true and pass big_w / big_h: column boxes beyond
// them are kept out of the cells and of max_w / max_h, in a short list every query walks in full (stored from the end of
// items[]); column boxes that miss `roi`, the hull of all row boxes, are dropped (a track that coasted off the canvas
// would otherwise stretch the cells).  Which pairs are visited is unchanged - only where they are found.
#pragma once
#include "block_utils.cuh"

namespace mot {

constexpr int kGridX = 32;
constexpr int kGridY = 16;
constexpr int kGridCells = kGridX * kGridY;

struct BoxGrid {
    int* cell;                 // [kGridCells + 1] start offset of every cell in items[] (exclusive scan)
    int* cursor;               // [kGridCells] build-time fill cursors
    unsigned short* items;     // [cap] column indices grouped by cell (one entry per finite box)
    float* red;                // [6 * 32] block-reduction scratch, then one int: number of listed big boxes
    int cap;
    int n_big;                 // big boxes: items[cap - 1], items[cap - 2], ...
    float x0, y0, sx, sy;      // cell = clamp((corner - origin) * scale)
    float max_w, max_h;        // largest column box

    __device__ __forceinline__ int cx(float x) const {
        return (int)fminf(fmaxf(xmul(xsub(x, x0), sx), 0.0f), (float)(kGridX - 1));
    }
    __device__ __forceinline__ int cy(float y) const {
        return (int)fminf(fmaxf(xmul(xsub(y, y0), sy), 0.0f), (float)(kGridY - 1));
    }
};

MOT_HD constexpr size_t grid_smem_bytes(int cap) {
    return ((sizeof(int) * (kGridCells + 1) + 15) & ~(size_t)15) + sizeof(int) * kGridCells +
           ((sizeof(unsigned short) * (size_t)cap + 15) & ~(size_t)15) + sizeof(float) * (6 * 32 + 4);
}

__device__ __forceinline__ unsigned char* grid_carve(unsigned char* p, int cap, BoxGrid& g) {
    g.cell = (int*)p;               p += (sizeof(int) * (kGridCells + 1) + 15) & ~(size_t)15;
    g.cursor = (int*)p;             p += sizeof(int) * kGridCells;
    g.items = (unsigned short*)p;   p += (sizeof(unsigned short) * (size_t)cap + 15) & ~(size_t)15;
    g.red = (float*)p;              p += sizeof(float) * (6 * 32 + 4);
    g.cap = cap;
    return p;
}

__device__ __forceinline__ bool box_finite(float4 b) {
    const float s = (b.x - b.x) + (b.y - b.y) + (b.z - b.z) + (b.w - b.w);     // NaN/inf -> NaN
    return s == 0.0f;
}

// Build the grid over boxes box_of(0..n), n <= g.cap.  All threads of the block must call.
// Non-finite boxes are left out (their IoU is NaN: never a candidate).
__device__ __forceinline__ bool box_big(float4 b, float big_w, float big_h) { return (b.z - b.x > big_w) || (b.w - b.y > big_h); }
// no row box inside `roi` can have an interior intersection with b (the same comparisons grid_query makes, on the hull)
__device__ __forceinline__ bool box_outside(float4 b, float4 roi) { return !(fminf(roi.z, b.z) > fmaxf(roi.x, b.x)) || !(fminf(roi.w, b.w) > fmaxf(roi.y, b.y)); }

template <bool kBig = false, class BoxOf>
__device__ __forceinline__ void grid_build(BoxGrid& g, int n, BlockScratch* bs, BoxOf box_of, float big_w = 3.0e38f,
                                           float big_h = 3.0e38f, float4 roi = make_float4(-3.0e38f, -3.0e38f, 3.0e38f, 3.0e38f)) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    float lo_x = 3.0e38f, lo_y = 3.0e38f, hi_x = -3.0e38f, hi_y = -3.0e38f, mw = 0.0f, mh = 0.0f;
    for (int j = tid; j < n; j += nt) {
        const float4 b = box_of(j);
        if (!box_finite(b) || (kBig && (box_big(b, big_w, big_h) || box_outside(b, roi)))) continue;
        lo_x = fminf(lo_x, b.x); lo_y = fminf(lo_y, b.y);
        hi_x = fmaxf(hi_x, b.x); hi_y = fmaxf(hi_y, b.y);
        mw = fmaxf(mw, b.z - b.x); mh = fmaxf(mh, b.w - b.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo_x = fminf(lo_x, __shfl_xor_sync(kFullMask, lo_x, o));
        lo_y = fminf(lo_y, __shfl_xor_sync(kFullMask, lo_y, o));
        hi_x = fmaxf(hi_x, __shfl_xor_sync(kFullMask, hi_x, o));
        hi_y = fmaxf(hi_y, __shfl_xor_sync(kFullMask, hi_y, o));
        mw = fmaxf(mw, __shfl_xor_sync(kFullMask, mw, o));
        mh = fmaxf(mh, __shfl_xor_sync(kFullMask, mh, o));
    }
    __syncthreads();
    if (lane == 0) {
        g.red[warp] = lo_x; g.red[32 + warp] = lo_y; g.red[64 + warp] = hi_x; g.red[96 + warp] = hi_y;
        g.red[128 + warp] = mw; g.red[160 + warp] = mh;
    }
    for (int c = tid; c <= kGridCells; c += nt) g.cell[c] = 0;
    for (int c = tid; c < kGridCells; c += nt) g.cursor[c] = 0;
    int* big_count = reinterpret_cast<int*>(g.red + 6 * 32);
    if (tid == 0) *big_count = 0;
    __syncthreads();
    for (int w = 0; w < nwarps; ++w) {
        lo_x = fminf(lo_x, g.red[w]); lo_y = fminf(lo_y, g.red[32 + w]);
        hi_x = fmaxf(hi_x, g.red[64 + w]); hi_y = fmaxf(hi_y, g.red[96 + w]);
        mw = fmaxf(mw, g.red[128 + w]); mh = fmaxf(mh, g.red[160 + w]);
    }
    const float ex = hi_x - lo_x, ey = hi_y - lo_y;
    g.x0 = lo_x; g.y0 = lo_y;
    g.sx = (ex > 0.0f) ? (float)kGridX / ex : 0.0f;
    g.sy = (ey > 0.0f) ? (float)kGridY / ey : 0.0f;
    if (!(g.sx == g.sx) || !(g.sy == g.sy)) { g.sx = 0.0f; g.sy = 0.0f; }
    g.max_w = mw; g.max_h = mh;
    for (int j = tid; j < n; j += nt) {
        const float4 b = box_of(j);
        if (!box_finite(b) || (kBig && box_outside(b, roi))) continue;
        if (kBig && box_big(b, big_w, big_h)) g.items[g.cap - 1 - atomicAdd(big_count, 1)] = (unsigned short)j;   // n <= cap: never meets the cell items
        else atomicAdd(&g.cell[g.cy(b.y) * kGridX + g.cx(b.x)], 1);
    }
    block_exclusive_scan(g.cell, kGridCells, bs, true);
    for (int j = tid; j < n; j += nt) {
        const float4 b = box_of(j);
        if (!box_finite(b) || (kBig && (box_big(b, big_w, big_h) || box_outside(b, roi)))) continue;
        const int c = g.cy(b.y) * kGridX + g.cx(b.x);
        g.items[g.cell[c] + atomicAdd(&g.cursor[c], 1)] = (unsigned short)j;
    }
    __syncthreads();
    g.n_big = kBig ? *big_count : 0;
}

// The cells a query walks: columns cx0..cx1 of rows cy0..cy1; t > 0 marks an IoU-floor window (the big list is then
// pruned by area as well).
struct GridWindow { int cx0, cx1, cy0, cy1; float t; };

// Every column box with a non-empty interior intersection with row box `a` has its corner in this window:
// a column box b can only overlap a if  a.x1 - max_w < b.x1 < a.x2  (and likewise in y); the scan
// range is widened by a relative 1e-6 so fp32 rounding in the widths can never drop a pair.
__device__ __forceinline__ GridWindow grid_window(const BoxGrid& g, float4 a) {
    const float mx = g.max_w + (fabsf(a.x) + g.max_w) * 1e-6f, my = g.max_h + (fabsf(a.y) + g.max_h) * 1e-6f;
    return GridWindow{g.cx(a.x - mx), g.cx(a.z), g.cy(a.y - my), g.cy(a.w), 0.0f};
}

// For callers that only care about pairs with IoU > t (0 < t < 1): such a column box b has its corner
// within  a.x1 - (1 - t) max_w < b.x1 < a.x1 + (1 - t) w_a  (and likewise in y):
//   IoU > t  =>  intersection > t * area_a and > t * area_b  =>  intersection width > t * w_a and > t * w_b;
//   b.x1 >= a.x1: width <= a.x2 - b.x1, so b.x1 - a.x1 < (1 - t) w_a;   b.x1 < a.x1: width <= w_b - (a.x1 - b.x1), so
//   a.x1 - b.x1 < (1 - t) w_b <= (1 - t) max_w.
// The window is widened by a relative 1e-5; pairs outside it provably have IoU <= t, pairs inside are still judged
// exactly by the caller, so results are identical to the dense evaluation.
__device__ __forceinline__ GridWindow grid_window_iou_above(const BoxGrid& g, float4 a, float t) {
    const float u = 1.0f - t;
    const float wa = a.z - a.x, ha = a.w - a.y;
    const float lx = u * g.max_w + (fabsf(a.x) + g.max_w) * 1e-5f, ly = u * g.max_h + (fabsf(a.y) + g.max_h) * 1e-5f;
    const float rx = u * wa + (fabsf(a.x) + fabsf(wa)) * 1e-5f, ry = u * ha + (fabsf(a.y) + fabsf(ha)) * 1e-5f;
    return GridWindow{g.cx(a.x - lx), g.cx(a.x + rx), g.cy(a.y - ly), g.cy(a.y + ry), t};
}

// Visit every column of the window (and of the big list) that overlaps `a`, once.
template <bool kBig = false, class BoxOf, class Visit>
__device__ __forceinline__ void grid_walk(const BoxGrid& g, const GridWindow& w, float4 a, BoxOf box_of, Visit visit) {
    for (int yy = w.cy0; yy <= w.cy1; ++yy) {
        const int e1 = g.cell[yy * kGridX + w.cx1 + 1];
        for (int e = g.cell[yy * kGridX + w.cx0]; e < e1; ++e) {
            const int j = g.items[e];
            const float4 b = box_of(j);
            if ((fminf(a.z, b.z) > fmaxf(a.x, b.x)) && (fminf(a.w, b.w) > fmaxf(a.y, b.y))) visit(j, b);
        }
    }
    if (kBig) {
        // big boxes: IoU > t also needs area_b < area_a / t (the intersection is at most area_a, the union at least area_b)
        const float area_cap = (a.z - a.x) * (a.w - a.y) * (1.0f + 1e-4f);
        for (int e = 0; e < g.n_big; ++e) {
            const int j = g.items[g.cap - 1 - e];
            const float4 b = box_of(j);
            if (w.t > 0.0f && (b.z - b.x) * (b.w - b.y) * w.t > area_cap) continue;
            if ((fminf(a.z, b.z) > fmaxf(a.x, b.x)) && (fminf(a.w, b.w) > fmaxf(a.y, b.y))) visit(j, b);
        }
    }
}

template <bool kBig = false, class BoxOf, class Visit>
__device__ __forceinline__ void grid_query(const BoxGrid& g, float4 a, BoxOf box_of, Visit visit) {
    grid_walk<kBig>(g, grid_window(g, a), a, box_of, visit);
}
template <bool kBig = false, class BoxOf, class Visit>
__device__ __forceinline__ void grid_query_iou_above(const BoxGrid& g, float4 a, float t, BoxOf box_of, Visit visit) {
    grid_walk<kBig>(g, grid_window_iou_above(g, a, t), a, box_of, visit);
}

}  // namespace mot
