// lap_device.cuh - exact linear assignment for one (rows x cols) problem, solved by one CTA.
//
// Replaces utils::linear_assignment + LAPSolver (reference src/utils/matching.cpp:14-60,
// include/motcpp/association/lap_solver.hpp:36-332).  The reference pads the cost matrix to
// (n+m)^2 with thresh/2 and runs dense Jonker-Volgenant in fp64; that is equivalent to choosing
// the partial matching that minimises sum(c_ij - thresh), in which a pair with c_ij > thresh can
// never appear.  So instead of a dense O((n+m)^3) solve this kernel
//   1. extracts the candidate pairs c_ij <= thresh (one thread per row, columns staged in smem),
//   2. labels the connected components of the candidate graph (min-label propagation, smem atomics),
//   3. gives every component to a warp, which runs a shortest-augmenting-path (Hungarian) solve in
//      fp64 with one LANE PER COLUMN: duals, distances and predecessors live in registers and the
//      per-step minimum is a warp-shuffle reduction.  A component with more than 32 columns+rows
//      falls back to the same algorithm striding over a global-memory scratch area.
// The result equals the reference's whenever the optimum is unique (always, for continuous costs);
// ties are broken by lowest column / lowest row, not by JV's scan order (DESIGN.md "Ties").
#pragma once
#include "block_utils.cuh"
#include "grid_device.cuh"

namespace mot {

constexpr int kLapNone = 0x7fffffff;

struct LapWorkspace {
    // shared memory
    int* row_label;            // [n_max]
    int* col_label;            // [m_max]
    int* scratch_a;            // [max(e_cap, n_max + 1)] edges, later packed per-label offsets
    int* scratch_b;            // [n_max] packed per-label fill cursors
    unsigned short* comp_rows; // [n_max] row ids grouped by component
    unsigned short* comp_cols; // [m_max]
    unsigned short* comp_list; // [n_max] component roots
    short* row2col;            // [n_max] result
    short* col2row;            // [m_max] result
    int* ctl;                  // [8] counters: 0 edges, 1 overflow, 2 changed, 3 n_trivial, 4 next_warp, 5 n_team, 6 n_warp
    BlockScratch* bs;
    int e_cap;
    BoxGrid grid;              // spatial index over the columns (box costs only)
    int* pairs;                // overlapping (row, col) pairs of one row chunk; aliases scratch_b..comp_list
    int p_cap;
    // global-memory scratch for components too large for one warp's registers
    double* g_u;               // [n_max]
    double* g_v;               // [m_max + n_max]
    double* g_minv;            // [m_max + n_max]
    int* g_way;                // [m_max + n_max]
    int* g_prow;               // [m_max + n_max]
    unsigned char* g_flags;    // [m_max + 2 * n_max] used[] then in_tree[]
    PhaseClock* clk;           // optional cycle accounting (nullptr = off): slots clk_base .. clk_base + 3
    int clk_base;
};

__device__ __forceinline__ double lap_inf() { return 1.0e300; }

// cost functors that carry big_w / big_h (and say so with kBigList) have outlier column boxes listed apart from the grid
template <class C, class = void> struct lap_has_big_list { static constexpr bool value = false; };
template <class C> struct lap_has_big_list<C, decltype((void)C::kBigList)> { static constexpr bool value = true; };
template <class C, class = void> struct lap_has_small_path { static constexpr bool value = false; };
template <class C> struct lap_has_small_path<C, decltype((void)C::kSmallFast)> { static constexpr bool value = true; };
template <class C, class = void> struct lap_has_iou_floor { static constexpr bool value = false; };
template <class C> struct lap_has_iou_floor<C, decltype((void)C::kIouFloor)> { static constexpr bool value = true; };

// Ascending sort of a short segment held in shared memory, by one warp.
__device__ __forceinline__ void warp_sort_u16(unsigned short* seg, int n) {
    const int lane = lane_id();
    if (n <= 1) return;
    if (n <= 32) {
        const int v = lane < n ? (int)seg[lane] : 0x7fffffff;
        int rank = 0;
        for (int k = 0; k < n; ++k) {
            const int o = __shfl_sync(kFullMask, v, k);
            rank += (o < v) ? 1 : 0;               // ids are distinct
        }
        __syncwarp();
        if (lane < n) seg[rank] = (unsigned short)v;
        __syncwarp();
        return;
    }
    // odd-even transposition sort for the rare big component
    for (int pass = 0; pass < n; ++pass) {
        for (int k = (pass & 1) + 2 * lane; k + 1 < n; k += 64) {
            const unsigned short a = seg[k], b = seg[k + 1];
            if (a > b) { seg[k] = b; seg[k + 1] = a; }
        }
        __syncwarp();
    }
}

// Ascending sort of n <= W distinct ids by a team of W consecutive lanes: every lane ranks its element against the others.
template <int W>
__device__ __forceinline__ void team_sort_u16(unsigned mask, int tl, unsigned short* seg, int n) {
    if (n <= 1) return;
    const int v = tl < n ? (int)seg[tl] : 0x7fffffff;
    int rank = 0;
#pragma unroll
    for (int k = 0; k < W; ++k) {
        const int o = __shfl_sync(mask, v, k, W);
        rank += (k < n && o < v) ? 1 : 0;
    }
    __syncwarp(mask);
    if (tl < n) seg[rank] = (unsigned short)v;
    __syncwarp(mask);
}

// (min value, lowest lane on ties) across a team of W consecutive lanes
template <int W>
__device__ __forceinline__ void team_argmin(unsigned mask, double& key, int& arg) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) {
        const double ok = __shfl_xor_sync(mask, key, o, W);
        const int oa = __shfl_xor_sync(mask, arg, o, W);
        if (ok < key || (ok == key && oa < arg)) { key = ok; arg = oa; }
    }
}
__device__ __forceinline__ void warp_argmin(double& key, int& arg) { team_argmin<32>(kFullMask, key, arg); }

// Component with r rows and c columns, r + c <= W, solved by a team of W consecutive lanes
// (`mask` names them, `tl` is the lane's index inside the team).  Team lane l < c owns real column
// cols[l]; lane c + k owns the private "stay unmatched" column of row k (cost 0, reachable only
// from row k).  Shortest augmenting paths with duals, all state in registers:
//   u  (row duals)    in lane k for row k          v    (column duals) per lane
//   minv / way / used (per search)                 prow (row matched to this column, -1 = free)
template <int W, class Cost>
__device__ __forceinline__ void team_hungarian(unsigned mask, int tl, const unsigned short* rows, int r,
                                               const unsigned short* cols, int c, float thresh, const Cost& cost,
                                               short* row2col, short* col2row) {
    const bool is_real = tl < c;
    const bool is_col = tl < c + r;
    const int dummy_of = tl - c;
    const int col_id = is_real ? (int)cols[tl] : -1;
    const int row_id = tl < r ? (int)rows[tl] : -1;
    const double INF = lap_inf();
    const double Ld = (double)thresh;
    double v = 0.0, u = 0.0;
    int prow = -1;
    for (int s = 0; s < r; ++s) {
        double minv = INF;
        int way = -1;
        bool used = false;
        bool in_tree = (tl == s);
        int i0 = s, j0 = -1;
        for (;;) {
            const int i0_id = __shfl_sync(mask, row_id, i0, W);
            const double u0 = __shfl_sync(mask, u, i0, W);
            double cur = INF;
            if (is_real) {
                const float cf = cost.pair(i0_id, col_id);
                if (cf <= thresh) cur = (((double)cf - Ld) + cost.pair_bias(i0_id, col_id)) - u0 - v;
            } else if (is_col && dummy_of == i0) {
                cur = 0.0 - u0 - v;
            }
            if (is_col && !used && cur < minv) { minv = cur; way = j0; }
            double key = (is_col && !used) ? minv : INF;
            int arg = tl;
            team_argmin<W>(mask, key, arg);
            const double delta = key;
            if (is_col) { if (used) v -= delta; else minv -= delta; }
            if (in_tree) u += delta;
            j0 = arg;
            if (tl == j0) used = true;
            i0 = __shfl_sync(mask, prow, j0, W);
            if (i0 < 0) break;
            if (tl == i0) in_tree = true;
        }
        int j = j0;
        while (j >= 0) {
            const int jp = __shfl_sync(mask, way, j, W);
            const int from = __shfl_sync(mask, prow, jp < 0 ? 0 : jp, W);
            if (tl == j) prow = (jp < 0) ? s : from;
            j = jp;
        }
    }
    const int matched_row = (is_real && prow >= 0) ? prow : 0;
    const int matched_row_id = __shfl_sync(mask, row_id, matched_row, W);
    if (is_real && prow >= 0) {
        col2row[col_id] = (short)matched_row_id;
        row2col[matched_row_id] = (short)col_id;
    }
}

struct LapGlobalScratch {      // by-value view of the global-memory scratch (keeps LapWorkspace out of local memory)
    double* g_u; double* g_v; double* g_minv; int* g_way; int* g_prow; unsigned char* g_flags;
};

template <class Cost>
__device__ __noinline__ void warp_hungarian_big(const LapGlobalScratch ws, int m_max, int n_max,
                                                const unsigned short* rows, int r, int pr,
                                                const unsigned short* cols, int c, int pc, float thresh,
                                                const Cost cost, short* row2col, short* col2row) {
    const int lane = lane_id();
    const double INF = lap_inf();
    const double Ld = (double)thresh;
    const int total = c + r;
    unsigned char* used = ws.g_flags;
    unsigned char* in_tree = ws.g_flags + m_max + n_max;
    // position of column slot k (0 <= k < total) in the global arrays
    auto cpos = [&](int k) { return k < c ? pc + k : m_max + pr + (k - c); };
    for (int k = lane; k < total; k += 32) { ws.g_v[cpos(k)] = 0.0; ws.g_prow[cpos(k)] = -1; }
    for (int k = lane; k < r; k += 32) ws.g_u[pr + k] = 0.0;
    __syncwarp();
    for (int s = 0; s < r; ++s) {
        for (int k = lane; k < total; k += 32) { ws.g_minv[cpos(k)] = INF; ws.g_way[cpos(k)] = -1; used[cpos(k)] = 0; }
        for (int k = lane; k < r; k += 32) in_tree[pr + k] = (k == s) ? 1 : 0;
        __syncwarp();
        int i0 = s, j0 = -1;
        for (;;) {
            const int i0_id = (int)rows[i0];
            const double u0 = ws.g_u[pr + i0];
            double key = INF;
            int arg = 0x7fffffff;
            for (int k = lane; k < total; k += 32) {
                const int p = cpos(k);
                if (used[p]) continue;
                double cur = INF;
                if (k < c) {
                    const float cf = cost.pair(i0_id, (int)cols[k]);
                    if (cf <= thresh) cur = (((double)cf - Ld) + cost.pair_bias(i0_id, (int)cols[k])) - u0 - ws.g_v[p];
                } else if (k - c == i0) {
                    cur = 0.0 - u0 - ws.g_v[p];
                }
                double mv = ws.g_minv[p];
                if (cur < mv) { mv = cur; ws.g_minv[p] = cur; ws.g_way[p] = j0; }
                if (mv < key) { key = mv; arg = k; }       // ascending k per lane => lowest k on ties
            }
            warp_argmin(key, arg);
            const double delta = key;
            for (int k = lane; k < total; k += 32) {
                const int p = cpos(k);
                if (used[p]) ws.g_v[p] -= delta; else ws.g_minv[p] -= delta;
            }
            for (int k = lane; k < r; k += 32)
                if (in_tree[pr + k]) ws.g_u[pr + k] += delta;
            __syncwarp();
            j0 = arg;
            if (lane == 0) used[cpos(j0)] = 1;
            i0 = ws.g_prow[cpos(j0)];
            __syncwarp();
            if (i0 < 0) break;
            if (lane == 0) in_tree[pr + i0] = 1;
            __syncwarp();
        }
        if (lane == 0) {
            int j = j0;
            while (j >= 0) {
                const int jp = ws.g_way[cpos(j)];
                ws.g_prow[cpos(j)] = (jp < 0) ? s : ws.g_prow[cpos(jp)];
                j = jp;
            }
        }
        __syncwarp();
    }
    for (int k = lane; k < c; k += 32) {
        const int pw = ws.g_prow[pc + k];
        if (pw >= 0) {
            const int rid = (int)rows[pw], cid = (int)cols[k];
            col2row[cid] = (short)rid;
            row2col[rid] = (short)cid;
        }
    }
    __syncwarp();
}

// Cost functor contract:
//   struct Cost {
//     struct Row { ... };                                   // whatever a row needs in registers
//     __device__ Row  row(int i) const;                     // load row i once
//     __device__ bool reject(const Row&, int j) const;      // cheap test: true => cost(i,j) > thresh for sure
//     __device__ float cost(const Row&, int j) const;       // exact fp32 cost
//     __device__ float pair(int i, int j) const;            // == cost(row(i), j)
//     __device__ double pair_bias(int i, int j) const;      // 0.0, or an infinitesimal (a multiple of 2^-50 below
//                          2^-32) added to cost(i,j) in fp64: decides between exactly tied optima only
//     __device__ bool is_candidate(const Row&, int i, int j, float thresh) const;   // cost(row, j) <= thresh; called
//                          exactly ONCE per examined pair (step 1 only), so a functor may tally side statistics there
//   };
// On return (all threads) ws.row2col[0..n) / ws.col2row[0..m) hold the assignment (-1 = unmatched).
// Result / label initialisation shared by block_lap and block_lap_solve's external callers.  All threads must call.
__device__ __forceinline__ void block_lap_begin(LapWorkspace& ws, int n, int m) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
    __syncthreads();
    for (int i = tid; i < n; i += nt) { ws.row2col[i] = -1; ws.row_label[i] = kLapNone; }
    for (int j = tid; j < m; j += nt) { ws.col2row[j] = -1; ws.col_label[j] = kLapNone; }
    if (tid < 8) ws.ctl[tid] = 0;
    __syncthreads();
}

// Steps 2-4 of block_lap: expects the candidate edges (row << 16 | col) in ws.scratch_a[0 .. ws.ctl[0]) (ws.ctl[1] != 0 =
// buffer overflow) and the labels / results initialised by block_lap_begin.  A caller that finds its candidates by other
// means (the StrongSORT appearance stage) calls block_lap_begin, pushes edges, then this.  All threads must call.
template <class Cost>
__device__ void block_lap_solve(LapWorkspace& ws, int n, int m, int n_max, int m_max, float thresh, const Cost& cost) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
    const int lane = tid & 31;
    __syncthreads();
    const int n_edges = min(ws.ctl[0], ws.e_cap);
    const bool overflow = ws.ctl[1] != 0;
    if (n_edges == 0 && !overflow) {                 // no candidate at all: everything stays unmatched (results are already -1)
        if (ws.clk) { ws.clk->tick(ws.clk_base + 1); ws.clk->tick(ws.clk_base + 2); ws.clk->tick(ws.clk_base + 3); }
        return;
    }

    // ---- fast path: a handful of candidates whose components are all trivial (one row or one column) - what the second and
    //      the unconfirmed association of a ByteTrack frame look like (1-9 edges on the C2 workload, where the general
    //      machinery below cost ~20 k cycles per call in barriers, scans and queue hand-outs).  Warp 0 holds one edge per
    //      lane, finds the components as lane masks (neighbours = same row or same column, closed by OR-ing the members'
    //      masks until nothing changes) and, if every component is trivial, picks each component's best candidate by the
    //      same order as the general path (cost, then row, then column).  Anything else falls through, untouched.
    // Opt-in per cost functor (kSmallFast): in the OC-SORT kernels, already at their register limit, the extra code cost 8 %.
    if (lap_has_small_path<Cost>::value && !overflow && n_edges <= 32) {
        if (tid < 32) {
            const bool have = lane < n_edges;
            const int pk = have ? ws.scratch_a[lane] : 0;
            const int ei = have ? (pk >> 16) : -1 - lane, ej = have ? (pk & 0xffff) : -1 - lane;   // idle lanes: unique keys
            const unsigned rm = __match_any_sync(kFullMask, ei), cm = __match_any_sync(kFullMask, ej);
            unsigned comp = rm | cm;
            for (;;) {
                unsigned acc = comp;
                for (int l = 0; l < 32; ++l) {
                    const unsigned other = __shfl_sync(kFullMask, comp, l);
                    if ((comp >> l) & 1u) acc |= other;
                }
                const bool grew = acc != comp;
                comp = acc;
                if (!__any_sync(kFullMask, grew)) break;
            }
            const bool trivial = ((comp & ~rm) == 0u) || ((comp & ~cm) == 0u);
            const bool all_trivial = __all_sync(kFullMask, !have || trivial);
            if (all_trivial) {
                float cf0 = 0.0f;
                double cf = 0.0;
                if (have) { cf0 = cost.pair(ei, ej); cf = (double)cf0 + cost.pair_bias(ei, ej); }
                const bool valid = have && (cf0 <= thresh);
                bool beaten = false;
                for (int l = 0; l < 32; ++l) {
                    const double ocf = __shfl_sync(kFullMask, cf, l);
                    const int oi = __shfl_sync(kFullMask, ei, l), oj = __shfl_sync(kFullMask, ej, l);
                    const bool ovalid = __shfl_sync(kFullMask, valid ? 1 : 0, l) != 0;
                    if (l != lane && ((comp >> l) & 1u) && ovalid &&
                        (ocf < cf || (ocf == cf && (oi < ei || (oi == ei && oj < ej))))) beaten = true;
                }
                if (valid && !beaten) { ws.row2col[ei] = (short)ej; ws.col2row[ej] = (short)ei; }
            }
            if (lane == 0) ws.ctl[2] = all_trivial ? 1 : 0;
        }
        __syncthreads();
        if (ws.ctl[2] != 0) {
            if (ws.clk) { ws.clk->tick(ws.clk_base + 1); ws.clk->tick(ws.clk_base + 2); ws.clk->tick(ws.clk_base + 3); }
            return;
        }
        __syncthreads();                              // everyone has read the flag before the general path reuses ctl[2]
    }

    // ---- 2. connected components by min-label propagation
    if (overflow) {
        // too many candidates for the edge buffer: treat everything as one component (still exact)
        for (int i = tid; i < n; i += nt) ws.row_label[i] = 0;
        for (int j = tid; j < m; j += nt) ws.col_label[j] = 0;
        __syncthreads();
    } else {
        for (int e = tid; e < n_edges; e += nt) { const int i = ws.scratch_a[e] >> 16; ws.row_label[i] = i; }
        __syncthreads();
        for (;;) {
            for (int e = tid; e < n_edges; e += nt) {
                const int pk = ws.scratch_a[e];
                atomicMin(&ws.col_label[pk & 0xffff], ws.row_label[pk >> 16]);
            }
            __syncthreads();
            if (tid == 0) ws.ctl[2] = 0;
            __syncthreads();
            for (int e = tid; e < n_edges; e += nt) {
                const int pk = ws.scratch_a[e];
                const int cl = ws.col_label[pk & 0xffff];
                const int old = atomicMin(&ws.row_label[pk >> 16], cl);
                if (cl < old) ws.ctl[2] = 1;
            }
            __syncthreads();
            if (ws.ctl[2] == 0) break;
        }
    }

    if (ws.clk) ws.clk->tick(ws.clk_base + 1);
    // ---- 3. group rows / columns by component: counts (packed rows | cols << 16) -> offsets -> fill
    int* off = ws.scratch_a;       // [n + 1] packed exclusive offsets per label
    int* cur = ws.scratch_b;       // [n] packed fill cursors per label
    for (int i = tid; i <= n; i += nt) off[i] = 0;
    for (int i = tid; i < n; i += nt) cur[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nt) { const int l = ws.row_label[i]; if (l != kLapNone) atomicAdd(&off[l], 1); }
    for (int j = tid; j < m; j += nt) { const int l = ws.col_label[j]; if (l != kLapNone) atomicAdd(&off[l], 1 << 16); }
    __syncthreads();
    block_exclusive_scan(off, n, ws.bs, true);
    for (int i = tid; i < n; i += nt) {
        const int l = ws.row_label[i];
        if (l != kLapNone) { const int k = atomicAdd(&cur[l], 1) & 0xffff; ws.comp_rows[(off[l] & 0xffff) + k] = (unsigned short)i; }
    }
    for (int j = tid; j < m; j += nt) {
        const int l = ws.col_label[j];
        if (l != kLapNone) { const int k = atomicAdd(&cur[l], 1 << 16) >> 16; ws.comp_cols[(off[l] >> 16) + k] = (unsigned short)j; }
    }
    __syncthreads();

    if (ws.clk) ws.clk->tick(ws.clk_base + 2);
    // ---- 4. solve every component with the cheapest exact method that fits it
    //   trivial (one row or one column): a thread takes the best candidate directly
    //   r + c <= 8 : an 8-lane team (four components per warp at a time)
    //   r + c <= 32: a warp          larger: a warp over global scratch
    unsigned short* list_triv = ws.comp_list;
    unsigned short* list_team = (unsigned short*)ws.scratch_b;           // the fill cursors are dead now
    unsigned short* list_warp = list_team + n;
    for (int i0 = 0; i0 < n; i0 += nt) {                       // lock-step, one aggregated counter update per class and warp
        const int i = i0 + tid;
        int cls = -1;
        if (i < n) {
            const int o0 = off[i], o1 = off[i + 1];
            const int r = (o1 & 0xffff) - (o0 & 0xffff), c = (o1 >> 16) - (o0 >> 16);
            if (r != 0) cls = (r == 1 || c == 1) ? 0 : ((r + c <= 8) ? 1 : 2);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const unsigned hit = __ballot_sync(kFullMask, cls == k);
            if (!hit) continue;
            int* ctr = &ws.ctl[k == 0 ? 3 : (k == 1 ? 5 : 6)];
            int b0 = 0;
            if (lane == __ffs((int)hit) - 1) b0 = atomicAdd(ctr, __popc(hit));
            b0 = __shfl_sync(kFullMask, b0, __ffs((int)hit) - 1);
            if (cls == k) {
                unsigned short* dst = (k == 0) ? list_triv : ((k == 1) ? list_team : list_warp);
                dst[b0 + __popc(hit & ((1u << lane) - 1u))] = (unsigned short)i;
            }
        }
    }
    if (tid == 0) ws.ctl[2] = 0;                               // (the label-propagation flag is dead) cursor of the team queue
    __syncthreads();
    const int n_triv = ws.ctl[3], n_team = ws.ctl[5], n_warp = ws.ctl[6];
    // Longest work first: the warp-class components are the long poles of this phase (measured on the C2 workload: 33 k of
    // its 59 k cycles, against 18 k for the teams and 5 k for the trivial ones), so every warp pulls from their queue
    // before anything else and the short components fill in behind them; all three classes are handed out dynamically or
    // per thread, so no warp waits for another until the final barrier.  Components are disjoint: any order is the same result.
    for (;;) {
        int k = 0;
        if (lane == 0) k = atomicAdd(&ws.ctl[4], 1);
        k = __shfl_sync(kFullMask, k, 0);
        if (k >= n_warp) break;
        const int root = (int)list_warp[k];
        const int o0 = off[root], o1 = off[root + 1];
        const int pr = o0 & 0xffff, r = (o1 & 0xffff) - pr;
        const int pc = o0 >> 16, c = (o1 >> 16) - pc;
        unsigned short* rows = ws.comp_rows + pr;
        unsigned short* cols = ws.comp_cols + pc;
        warp_sort_u16(rows, r);
        warp_sort_u16(cols, c);
        if (r + c <= 32) team_hungarian<32>(kFullMask, lane, rows, r, cols, c, thresh, cost, ws.row2col, ws.col2row);
        else warp_hungarian_big(LapGlobalScratch{ws.g_u, ws.g_v, ws.g_minv, ws.g_way, ws.g_prow, ws.g_flags}, m_max, n_max,
                                rows, r, pr, cols, c, pc, thresh, cost, ws.row2col, ws.col2row);
    }
    {
        const int tl = lane & 7;
        const unsigned tmask = 0xffu << (lane & 24);
        for (;;) {                                             // four team components per warp and pull
            int k4 = 0;
            if (lane == 0) k4 = atomicAdd(&ws.ctl[2], 4);
            k4 = __shfl_sync(kFullMask, k4, 0);
            if (k4 >= n_team) break;
            const int k = k4 + (lane >> 3);
            if (k < n_team) {
                const int root = (int)list_team[k];
                const int o0 = off[root], o1 = off[root + 1];
                const int pr = o0 & 0xffff, r = (o1 & 0xffff) - pr;
                const int pc = o0 >> 16, c = (o1 >> 16) - pc;
                unsigned short* rows = ws.comp_rows + pr;
                unsigned short* cols = ws.comp_cols + pc;
                team_sort_u16<8>(tmask, tl, rows, r);          // the fill order is arbitrary; ties go to the lowest index
                team_sort_u16<8>(tmask, tl, cols, c);
                team_hungarian<8>(tmask, tl, rows, r, cols, c, thresh, cost, ws.row2col, ws.col2row);
            }
            __syncwarp();
        }
    }
    for (int k = tid; k < n_triv; k += nt) {
        const int root = (int)list_triv[k];
        const int o0 = off[root], o1 = off[root + 1];
        const int pr = o0 & 0xffff, r = (o1 & 0xffff) - pr;
        const int pc = o0 >> 16, c = (o1 >> 16) - pc;
        double best = 0.0;
        int bi = -1, bj = -1;
        for (int a = 0; a < r; ++a)
            for (int b = 0; b < c; ++b) {
                const int i = (int)ws.comp_rows[pr + a], j = (int)ws.comp_cols[pc + b];
                const float cf0 = cost.pair(i, j);
                if (!(cf0 <= thresh)) continue;
                const double cf = (double)cf0 + cost.pair_bias(i, j);
                if (bi < 0 || cf < best || (cf == best && (i < bi || (i == bi && j < bj)))) { best = cf; bi = i; bj = j; }
            }
        if (bi >= 0) { ws.row2col[bi] = (short)bj; ws.col2row[bj] = (short)bi; }
    }
    __syncthreads();
    if (ws.clk) ws.clk->tick(ws.clk_base + 3);
}

template <class Cost>
__device__ void block_lap(LapWorkspace& ws, int n, int m, int n_max, int m_max, float thresh, const Cost& cost) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    __syncthreads();
    for (int i = tid; i < n; i += nt) { ws.row2col[i] = -1; ws.row_label[i] = kLapNone; }
    for (int j = tid; j < m; j += nt) { ws.col2row[j] = -1; ws.col_label[j] = kLapNone; }
    if (tid < 8) ws.ctl[tid] = 0;
    __syncthreads();
    if (n == 0 || m == 0) return;

    // ---- 1. candidate pairs.  Costs that live in shared memory / registers are scanned one row per
    //         thread; a dense matrix in global memory is scanned one row per warp so loads coalesce.
    if (Cost::kWarpPerRow) {
        for (int i = warp; i < n; i += nwarps) {
            const typename Cost::Row rw = cost.row(i);
            for (int j0 = 0; j0 < m; j0 += 32) {
                const int j = j0 + lane;
                const bool cand = (j < m) && cost.is_candidate(rw, i, j, thresh);
                const unsigned ballot = __ballot_sync(kFullMask, cand);
                if (ballot == 0) continue;
                int e0 = 0;
                if (lane == 0) e0 = atomicAdd(&ws.ctl[0], __popc(ballot));
                e0 = __shfl_sync(kFullMask, e0, 0);
                if (cand) {
                    const int e = e0 + __popc(ballot & ((1u << lane) - 1u));
                    if (e < ws.e_cap) ws.scratch_a[e] = (i << 16) | j;
                    else ws.ctl[1] = 1;
                }
            }
        }
    } else {
        bool use_grid = false;
        if constexpr (Cost::kGrid) {
            // box costs: index the columns so that only overlapping pairs are looked at
            if (cost.prune && (long long)n * m >= 65536 && m <= ws.grid.cap) {
                if (ws.clk && ws.clk_base == 3) ws.clk->tick(16 + 3);
                if constexpr (lap_has_big_list<Cost>::value)       // outlier column boxes kept out of the cells (grid_device.cuh)
                    grid_build<true>(ws.grid, m, ws.bs, [&](int j) { return cost.col_box(j); }, cost.big_w, cost.big_h, cost.roi);
                else
                    grid_build(ws.grid, m, ws.bs, [&](int j) { return cost.col_box(j); });
                if (ws.clk && ws.clk_base == 3) ws.clk->tick(16 + 0);
                use_grid = true;
            }
        }
        auto push_edge = [&](int i, int j) {
            const int e = atomicAdd(&ws.ctl[0], 1);
            if (e < ws.e_cap) ws.scratch_a[e] = (i << 16) | j;
            else ws.ctl[1] = 1;
        };
        auto scan_row = [&](int i) {                       // every column, exact cost on the spot
            const typename Cost::Row rw = cost.row(i);
            for (int j = 0; j < m; ++j) {
                if (cost.reject(rw, j)) continue;
                if (cost.is_candidate(rw, i, j, thresh)) push_edge(i, j);
            }
        };
        if (use_grid) {
            if constexpr (Cost::kGrid) {
                // one row per thread and pass: (1) collect the overlapping pairs of the rows through the grid, (2) judge them
                // densely, one pair per thread (no divergence).  Measured alternatives, all slower on the C2 workload: judging
                // inside the walk (+16 %), count / scan / write instead of the shared counter (+20 %), a lock-step warp
                // walk with one aggregated atomic per step (+70 %): the walk is a chain of dependent shared-memory loads.
                // `step` rows per pass, halved (down to one warp's worth) whenever their pairs do not fit the pair buffer.
                int step = nt;
                for (int base = 0; base < n;) {
                    const int i = base + tid;
                    const bool mine = tid < step && i < n;
                    if (tid == 0) ws.ctl[7] = 0;
                    __syncthreads();
                    constexpr bool kBigList = lap_has_big_list<Cost>::value;
                    auto window_of = [&](const float4 b) {
                        if constexpr (kBigList || lap_has_iou_floor<Cost>::value) {
                            // such functors know an IoU below which no pair is a candidate: a tighter corner window
                            if (cost.iou_floor > 0.01f) return grid_window_iou_above(ws.grid, b, cost.iou_floor);
                        }
                        return grid_window(ws.grid, b);
                    };
                    if (mine) {
                        // (handing rows with crowded windows to whole warps was measured: no gain on C2, the walk is issue-bound)
                        const typename Cost::Row rw = cost.row(i);
                        grid_walk<kBigList>(ws.grid, window_of(rw.b), rw.b, [&](int j) { return cost.col_box(j); }, [&](int j, float4) {
                            const int q = atomicAdd(&ws.ctl[7], 1);
                            if (q < ws.p_cap) ws.pairs[q] = (i << 16) | j;
                        });
                    }
                    __syncthreads();
                    if (ws.clk && ws.clk_base == 3) ws.clk->tick(16 + 1);
                    const int n_pairs = ws.ctl[7];
                    if (n_pairs > ws.p_cap && step > 32) {          // uniform decision: retry this base with fewer rows
                        step >>= 1;
                        __syncthreads();                            // everyone has read n_pairs before thread 0 clears it
                        continue;
                    }
                    if (n_pairs <= ws.p_cap) {
                        for (int q0 = 0; q0 < n_pairs; q0 += nt) {       // lock-step: one aggregated edge push per warp and pass
                            const int q = q0 + tid;
                            const int pk = (q < n_pairs) ? ws.pairs[q] : 0;
                            const int pi = pk >> 16, pj = pk & 0xffff;
                            const bool cand = (q < n_pairs) && cost.is_candidate(cost.row(pi), pi, pj, thresh);
                            const unsigned hit = __ballot_sync(kFullMask, cand);
                            if (hit) {
                                int e0 = 0;
                                if (lane == __ffs((int)hit) - 1) e0 = atomicAdd(&ws.ctl[0], __popc(hit));
                                e0 = __shfl_sync(kFullMask, e0, __ffs((int)hit) - 1);
                                if (cand) {
                                    const int e = e0 + __popc(hit & ((1u << lane) - 1u));
                                    if (e < ws.e_cap) ws.scratch_a[e] = (pi << 16) | pj;
                                    else ws.ctl[1] = 1;
                                }
                            }
                        }
                    } else if (mine) {
                        scan_row(i);                           // 32 rows still overflow the buffer: every column of those rows
                    }
                    __syncthreads();
                    if (ws.clk && ws.clk_base == 3) ws.clk->tick(16 + 2);
                    base += step;
                }
            }
        } else {
            // no grid: one WARP per row, lanes across the columns (row in registers, coalesced column reads, one
            // aggregated edge push per 32 columns).  Up to ~64 k pairs this beats building and walking the grid - the walk is
            // a chain of dependent shared-memory loads whose latency does not shrink with the problem (measured on the
            // 192 x 192 unconfirmed-track association of a C2 frame: 30 k cycles through the grid).
            for (int i = warp; i < n; i += nwarps) {
                const typename Cost::Row rw = cost.row(i);
                for (int j0 = 0; j0 < m; j0 += 32) {
                    const int j = j0 + lane;
                    const bool cand = (j < m) && !cost.reject(rw, j) && cost.is_candidate(rw, i, j, thresh);
                    const unsigned hit = __ballot_sync(kFullMask, cand);
                    if (hit == 0) continue;
                    int e0 = 0;
                    if (lane == __ffs((int)hit) - 1) e0 = atomicAdd(&ws.ctl[0], __popc(hit));
                    e0 = __shfl_sync(kFullMask, e0, __ffs((int)hit) - 1);
                    if (cand) {
                        const int e = e0 + __popc(hit & ((1u << lane) - 1u));
                        if (e < ws.e_cap) ws.scratch_a[e] = (i << 16) | j;
                        else ws.ctl[1] = 1;
                    }
                }
            }
        }
    }
    if (ws.clk) { __syncthreads(); ws.clk->tick(ws.clk_base); }
    block_lap_solve(ws, n, m, n_max, m_max, thresh, cost);
}

}  // namespace mot

namespace mot {

// ---- shared-memory carve-up for LapWorkspace (sizes in bytes, 16-byte aligned pieces)
MOT_HD constexpr size_t lap_align16(size_t x) { return (x + 15) & ~(size_t)15; }

MOT_HD constexpr size_t lap_smem_bytes(int n_max, int m_max, int e_cap) {
    const int a = e_cap > n_max + 1 ? e_cap : n_max + 1;
    size_t b = 0;
    b += lap_align16(sizeof(int) * (size_t)n_max);              // row_label
    b += lap_align16(sizeof(int) * (size_t)m_max);              // col_label
    b += lap_align16(sizeof(int) * (size_t)a);                  // scratch_a
    b += lap_align16(sizeof(int) * (size_t)n_max);              // scratch_b
    b += lap_align16(sizeof(unsigned short) * (size_t)n_max);   // comp_rows
    b += lap_align16(sizeof(unsigned short) * (size_t)m_max);   // comp_cols
    b += lap_align16(sizeof(unsigned short) * (size_t)n_max);   // comp_list
    b += lap_align16(sizeof(short) * (size_t)n_max);            // row2col
    b += lap_align16(sizeof(short) * (size_t)m_max);            // col2row
    b += lap_align16(sizeof(int) * 8);                          // ctl
    b += lap_align16(sizeof(BlockScratch));
    b += lap_align16(grid_smem_bytes(n_max > m_max ? n_max : m_max));
    return b;
}

// bytes of [row_label, row2col): dead between two block_lap calls, reused by the CTA-wide dense LAPJV (jv_block_device.cuh)
MOT_HD constexpr size_t lap_idle_bytes(int n_max, int m_max, int e_cap) {
    const int a = e_cap > n_max + 1 ? e_cap : n_max + 1;
    return lap_align16(sizeof(int) * (size_t)n_max) + lap_align16(sizeof(int) * (size_t)m_max) + lap_align16(sizeof(int) * (size_t)a) +
           lap_align16(sizeof(int) * (size_t)n_max) + lap_align16(sizeof(unsigned short) * (size_t)n_max) +
           lap_align16(sizeof(unsigned short) * (size_t)m_max) + lap_align16(sizeof(unsigned short) * (size_t)n_max) +
           lap_align16(grid_smem_bytes(n_max > m_max ? n_max : m_max));
}

__device__ __forceinline__ unsigned char* lap_carve(unsigned char* p, int n_max, int m_max, int e_cap, LapWorkspace& ws) {
    const int a = e_cap > n_max + 1 ? e_cap : n_max + 1;
    ws.row_label = (int*)p;            p += lap_align16(sizeof(int) * (size_t)n_max);
    ws.col_label = (int*)p;            p += lap_align16(sizeof(int) * (size_t)m_max);
    ws.scratch_a = (int*)p;            p += lap_align16(sizeof(int) * (size_t)a);
    ws.scratch_b = (int*)p;            p += lap_align16(sizeof(int) * (size_t)n_max);
    ws.comp_rows = (unsigned short*)p; p += lap_align16(sizeof(unsigned short) * (size_t)n_max);
    ws.comp_cols = (unsigned short*)p; p += lap_align16(sizeof(unsigned short) * (size_t)m_max);
    ws.comp_list = (unsigned short*)p; p += lap_align16(sizeof(unsigned short) * (size_t)n_max);
    // the grid sits inside the idle span [row_label, row2col): it is rebuilt by every block_lap call, so between calls its
    // bytes serve the dense LAPJV's work arrays / the DeepOC-SORT summation tiles like the arrays above
    unsigned char* const grid_at = p;
    grid_carve(p, n_max > m_max ? n_max : m_max, ws.grid);  p += lap_align16(grid_smem_bytes(n_max > m_max ? n_max : m_max));
    ws.row2col = (short*)p;            p += lap_align16(sizeof(short) * (size_t)n_max);
    ws.col2row = (short*)p;            p += lap_align16(sizeof(short) * (size_t)m_max);
    ws.ctl = (int*)p;                  p += lap_align16(sizeof(int) * 8);
    ws.bs = (BlockScratch*)p;          p += lap_align16(sizeof(BlockScratch));
    ws.e_cap = e_cap;
    ws.clk = nullptr;
    ws.clk_base = 0;
    ws.pairs = ws.scratch_b;           // overlapping-pair buffer of the candidate search: up to the grid it is collected through
    ws.p_cap = (int)((grid_at - (unsigned char*)ws.scratch_b) / sizeof(int));
    return p;
}

// global scratch (bytes) for the large-component fallback of one problem
MOT_HD constexpr size_t lap_gscratch_bytes(int n_max, int m_max) {
    const size_t cols = (size_t)n_max + (size_t)m_max;
    size_t b = 0;
    b += lap_align16(sizeof(double) * (size_t)n_max);   // u
    b += lap_align16(sizeof(double) * cols);            // v
    b += lap_align16(sizeof(double) * cols);            // minv
    b += lap_align16(sizeof(int) * cols);               // way
    b += lap_align16(sizeof(int) * cols);               // prow
    b += lap_align16(cols + (size_t)n_max);             // flags
    return b;
}

__device__ __forceinline__ void lap_carve_gscratch(unsigned char* g, int n_max, int m_max, LapWorkspace& ws) {
    const size_t cols = (size_t)n_max + (size_t)m_max;
    ws.g_u = (double*)g;        g += lap_align16(sizeof(double) * (size_t)n_max);
    ws.g_v = (double*)g;        g += lap_align16(sizeof(double) * cols);
    ws.g_minv = (double*)g;     g += lap_align16(sizeof(double) * cols);
    ws.g_way = (int*)g;         g += lap_align16(sizeof(int) * cols);
    ws.g_prow = (int*)g;        g += lap_align16(sizeof(int) * cols);
    ws.g_flags = (unsigned char*)g;
}

// Dense-matrix cost: the standalone mot_lap() entry point (row-major fp32, leading dimension ld).
struct MatrixCost {
    static constexpr bool kWarpPerRow = true;
    static constexpr bool kGrid = false;
    const float* c;
    int ld;
    struct Row { const float* p; };
    __device__ __forceinline__ Row row(int i) const { return Row{c + (size_t)i * ld}; }
    __device__ __forceinline__ bool reject(const Row&, int) const { return false; }
    __device__ __forceinline__ float cost(const Row& r, int j) const { return r.p[j]; }
    __device__ __forceinline__ float pair(int i, int j) const { return c[(size_t)i * ld + j]; }
    __device__ __forceinline__ bool is_candidate(const Row& r, int, int j, float thresh) const { return r.p[j] <= thresh; }
    __device__ __forceinline__ double pair_bias(int, int) const { return 0.0; }
};

}  // namespace mot
