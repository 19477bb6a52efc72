// jv_block_device.cuh - the reference's dense LAPJV (include/motcpp/association/lap_solver.hpp:36-231, :251-332),
// step for step, run by a WHOLE CTA on the extended (n+m) x (n+m) matrix - the any-size sibling of the one-warp
// solver in jv_device.cuh (which keeps its state in shared memory and serves rows + columns <= kJvMax).
//
// Why it exists: frames whose optimum is not unique (OC-SORT's bit-identical twin tracks, StrongSORT's duplicated
// rows).  There the reference's answer is whatever LAPJV's scan order yields over the full dense matrix, and only
// the same algorithm reproduces it.  Such frames are rare; every other frame goes through the sparse solver
// (lap_device.cuh).  This path trades speed for being the reference's answer at ANY problem size.
//
// Parallelisation keeps the sequential semantics exactly:
//   column reduction        one thread per column (strict '<' over ascending rows = lowest row on ties); the
//                           right-to-left claim becomes "highest column wins" = atomicMax + a per-row count
//   reduction transfer      rows in ascending order, the row minimum is a block reduction; the thread that owns
//                           column x[i] applies the update, so ONE barrier per row
//   augmenting row reduct.  per free row a lexicographic (value, index) top-2 block reduction, bookkeeping on
//                           thread 0 (two barriers per row)
//   shortest augm. paths    the level-opening scan only acts at weak prefix minima of dist[order[k]]: a block-wide
//                           prefix-min marks them in a bitmap.  On the extended matrix these "records" come by the
//                           hundred (its dummy rows / columns tie exactly), so their swaps are not replayed one by one
//                           (measured: 130 k cycles per level, 80 % of the solve) but applied as the permutation they
//                           compose to, in parallel (jv_block_apply_records); a relax step updates
//                           all TODO columns in parallel and only the (rare) columns that land exactly on the level are
//                           replayed in position order - one barrier (__syncthreads_or) per relax step otherwise.
// fp64 throughout, like the reference.  Work arrays live in shared memory when the caller has room for them
// (jv_block_sbytes_full), else in global memory (per-stream scratch, L2 resident) with only the scan order permutation
// and the small reduction scratch in shared memory.
#pragma once
#include "jv_device.cuh"

#ifndef MOT_JV_SERIAL_RECORDS
#define MOT_JV_SERIAL_RECORDS 40     // level openings with at most this many records are replayed by thread 0
#endif

namespace mot {

struct JvBlockWork {            // N_max entries each
    double* v;                  // global
    double* dist;               // global
    int* x;                     // global
    int* y;                     // global
    int* fr;                    // global: free rows
    int* pred;                  // global
    int* cnt;                   // global: columns whose minimum sits in this row (column reduction); later the event ranks
    int* tmp;                   // global: staging of the level-opening permutation
    int* order;                 // shared (or global): scan-order permutation of find_path_dense
    unsigned* bits;             // shared: 2 x ceil(N_max / 32) words (event bitmap, strict-record bitmap)
    // reduction scratch, shared
    double* pv1; double* pv2;   // [32]
    int* pi1; int* pi2;         // [32]
    int* ctl;                   // [8]
};

MOT_HD constexpr size_t jv_block_gbytes(int n_max) {          // global bytes
    return ((size_t)n_max * (2 * sizeof(double) + 6 * sizeof(int)) + 64 + 15) & ~(size_t)15;
}
MOT_HD constexpr size_t jv_block_sbytes(int n_max) {          // shared bytes (order included)
    return (sizeof(int) * (size_t)n_max + sizeof(unsigned) * 2 * (size_t)((n_max + 31) / 32) + 2 * 32 * sizeof(double) +
            2 * 32 * sizeof(int) + 8 * sizeof(int) + 32 + 15) & ~(size_t)15;
}
// Everything in shared memory: the solver is a chain of a few thousand short steps (one per augmenting-row-reduction
// row, level and relax step), each a handful of DEPENDENT reads of these arrays - out of L2 that is ~4 us per step
// (42 ms for a 30 x 622 re-match of duplicated lists), out of shared memory a fraction of it.
MOT_HD constexpr size_t jv_block_sbytes_full(int n_max) { return jv_block_sbytes(n_max) + jv_block_gbytes(n_max); }
// g: global scratch of jv_block_gbytes(n_max); s: shared scratch of jv_block_sbytes(n_max), 16-byte aligned -
// or, all_shared, of jv_block_sbytes_full(n_max) (g is then not touched)
__device__ __forceinline__ JvBlockWork jv_block_carve(unsigned char* g, unsigned char* s, int n_max, bool all_shared = false) {
    JvBlockWork w;
    if (all_shared) g = s + jv_block_sbytes(n_max);
    w.v = (double*)g;       g += sizeof(double) * (size_t)n_max;
    w.dist = (double*)g;    g += sizeof(double) * (size_t)n_max;
    w.x = (int*)g;          g += sizeof(int) * (size_t)n_max;
    w.y = (int*)g;          g += sizeof(int) * (size_t)n_max;
    w.fr = (int*)g;         g += sizeof(int) * (size_t)n_max;
    w.pred = (int*)g;       g += sizeof(int) * (size_t)n_max;
    w.cnt = (int*)g;        g += sizeof(int) * (size_t)n_max;
    w.tmp = (int*)g;
    w.pv1 = (double*)s;     s += 32 * sizeof(double);
    w.pv2 = (double*)s;     s += 32 * sizeof(double);
    w.pi1 = (int*)s;        s += 32 * sizeof(int);
    w.pi2 = (int*)s;        s += 32 * sizeof(int);
    w.ctl = (int*)s;        s += 8 * sizeof(int);
    w.bits = (unsigned*)s;  s += sizeof(unsigned) * 2 * (size_t)((n_max + 31) / 32);
    w.order = (int*)s;
    return w;
}

__device__ __forceinline__ double lap_inf_jv() { return 1.0e300; }

struct JvTop2 { double v1; int i1; double v2; int i2; };
__device__ __forceinline__ bool jv_lexless(double va, int ia, double vb, int ib) { return va < vb || (va == vb && ia < ib); }
__device__ __forceinline__ JvTop2 jv_merge(JvTop2 a, const JvTop2& b) {
    if (jv_lexless(b.v1, b.i1, a.v1, a.i1)) {            // b holds the best: the runner-up is lexmin(a.best, b.second)
        JvTop2 r = b;
        if (jv_lexless(a.v1, a.i1, r.v2, r.i2)) { r.v2 = a.v1; r.i2 = a.i1; }
        return r;
    }
    if (jv_lexless(b.v1, b.i1, a.v2, a.i2)) { a.v2 = b.v1; a.i2 = b.i1; }
    return a;
}
// Lexicographic (value, index) best and runner-up over the block; valid in thread 0 after the call.
// One barrier inside; the caller must place another before the scratch is reused.
__device__ __forceinline__ JvTop2 jv_block_top2(JvTop2 t, const JvBlockWork& w) {
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = (int)blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        JvTop2 b;
        b.v1 = __shfl_xor_sync(kFullMask, t.v1, o); b.i1 = __shfl_xor_sync(kFullMask, t.i1, o);
        b.v2 = __shfl_xor_sync(kFullMask, t.v2, o); b.i2 = __shfl_xor_sync(kFullMask, t.i2, o);
        t = jv_merge(t, b);
    }
    if (lane == 0) { w.pv1[warp] = t.v1; w.pi1[warp] = t.i1; w.pv2[warp] = t.v2; w.pi2[warp] = t.i2; }
    __syncthreads();
    if (tid == 0)
        for (int k = 1; k < nw; ++k) t = jv_merge(t, JvTop2{w.pv1[k], w.pi1[k], w.pv2[k], w.pi2[k]});
    return t;
}

// The swaps of one run of level-opening records, as ONE permutation.  Sequentially (lap_solver.hpp:127-137) the records
// at positions k_0 < k_1 < ... < k_{T-1} do   swap(order[k_t], order[h0 + t]),  t = 0 .. T-1   (k_t >= h0 + t).
//   * order[k_t] is untouched before step t and position h0 + t is untouched after it, so the front ends up as
//     order'[h0 + t] = order[k_t];
//   * what step t displaces from position h0 + t goes to k_t; if k_t lies inside the front [h0, h0 + T) it is displaced
//     again at step k_t - h0, and so on: the element that STARTS at a front position which is not itself a record hops
//     along  p -> k_{p - h0}  until it leaves the front.  The hops only go up, so in-place pointer jumping on the k
//     array finds every chain's landing position in log T rounds.
// ev: record bitmap by position; records taken from [pb, pe).  K = w.cnt (ranks -> positions), staging in w.tmp, the
// per-word rank offsets in wb.  All threads of the block must call; returns T.
__device__ __forceinline__ int jv_block_apply_records(const JvBlockWork& w, const unsigned* ev, unsigned* wb, int pb, int pe, int h0) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
    int* K = w.cnt;
    const int q0 = pb >> 5, q1 = (pe - 1) >> 5;
    auto word = [&](int q) {                              // the records of word q that lie in [pb, pe)
        unsigned b = ev[q];
        if (q == q0) b &= 0xffffffffu << (pb & 31);
        if (q == q1 && (pe & 31)) b &= 0xffffffffu >> (32 - (pe & 31));
        return b;
    };
    if (tid == 0) {
        int acc = 0;
        for (int q = q0; q <= q1; ++q) { wb[q] = (unsigned)acc; acc += __popc(word(q)); }
        w.ctl[4] = acc;
    }
    __syncthreads();
    const int T = w.ctl[4];
    __syncthreads();                                      // ctl[4] is rewritten by the next run
    if (T == 0) return 0;
    for (int p = pb + tid; p < pe; p += nt) {
        const unsigned b = word(p >> 5);
        if ((b >> (p & 31)) & 1u) K[wb[p >> 5] + __popc(b & ((1u << (p & 31)) - 1u))] = p;
    }
    __syncthreads();
    auto is_record = [&](int p) { return p >= pb && p < pe && ((ev[p >> 5] >> (p & 31)) & 1u); };
    for (int t = tid; t < T; t += nt) w.tmp[h0 + t] = w.order[K[t]];
    __syncthreads();
    const int front_end = h0 + T;
    for (;;) {                                            // K[t] <- landing position of the chain through node t
        // pointer jumping in synchronous rounds: read, barrier, write (jumping in place without the barrier converges as
        // well - every pointer only ever moves further along its own chain - but it is a data race by the letter)
        bool more = false;
        const int t0 = tid;                               // T <= N <= a few thousand: a handful of nodes per thread
        int nxt[8];
        int cnt = 0;
        for (int t = t0; t < T && cnt < 8; t += nt, ++cnt) {
            const int q = K[t];
            nxt[cnt] = (q < front_end && q != h0 + t) ? K[q - h0] : -1;
        }
        __syncthreads();
        cnt = 0;
        for (int t = t0; t < T && cnt < 8; t += nt, ++cnt)
            if (nxt[cnt] >= 0) { K[t] = nxt[cnt]; more |= nxt[cnt] < front_end; }
        for (int t = t0 + 8 * nt; t < T; t += nt) {       // beyond 8 nodes per thread (never at the built shapes): in place
            const int q = K[t];
            if (q < front_end && q != h0 + t) { const int nq = K[q - h0]; K[t] = nq; more |= nq < front_end; }
        }
        if (!__syncthreads_or(more ? 1 : 0)) break;
    }
    for (int t = tid; t < T; t += nt)
        if (!is_record(h0 + t)) w.tmp[K[t]] = w.order[h0 + t];
    __syncthreads();
    for (int t = tid; t < T; t += nt) {
        w.order[h0 + t] = w.tmp[h0 + t];
        if (!is_record(h0 + t)) w.order[K[t]] = w.tmp[K[t]];
    }
    __syncthreads();
    return T;
}

// lap_solver.hpp:115-211 (find_path_dense + the augmentation), whole block.
static __device__ __noinline__ void jv_block_augment_from(const JvCost& c, int N, const JvBlockWork& w, int start) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int words = (N + 31) >> 5;
    unsigned* ev = w.bits;                 // event bitmap (records / level hits), indexed by POSITION
    unsigned* strict = w.bits + words;     // level opening: the record is a strict new minimum
    for (int j = tid; j < N; j += nt) { w.order[j] = j; w.pred[j] = start; w.dist[j] = c.at(start, j) - w.v[j]; }
    for (int q = tid; q < 2 * words; q += nt) w.bits[q] = 0;
    __syncthreads();
    int lo = 0, hi = 0, settled = 0, sink = -1;          // uniform across the block
    while (sink < 0) {
        if (lo == hi) {
            // ---- open the next distance level: the sequential scan over k = lo+1 .. N-1 only acts where
            //      a_k = dist[order[k]] <= min(a_lo .. a_{k-1}); mark those positions with a block-wide prefix-min
            settled = lo;
            const int len = N - lo;
            const int per = (len + nt - 1) / nt;
            const int p0 = lo + tid * per, p1 = min(N, p0 + per);
            double cm = lap_inf_jv();
            for (int p = p0; p < p1; ++p) { const double a = w.dist[w.order[p]]; if (a < cm) cm = a; }
            double incl = cm;                              // inclusive prefix-min inside the warp
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double t = __shfl_up_sync(kFullMask, incl, o);
                if (lane >= o && t < incl) incl = t;
            }
            double excl = __shfl_up_sync(kFullMask, incl, 1);
            if (lane == 0) excl = lap_inf_jv();
            if (lane == 31) w.pv1[warp] = incl;
            __syncthreads();
            double before = lap_inf_jv();
            for (int k = 0; k < warp; ++k) { const double t = w.pv1[k]; if (t < before) before = t; }
            double run = before < excl ? before : excl;    // min over every position before p0
            for (int p = p0; p < p1; ++p) {
                const double a = w.dist[w.order[p]];
                if (p > lo && a <= run) {
                    atomicOr(&ev[p >> 5], 1u << (p & 31));
                    if (a < run) atomicOr(&strict[p >> 5], 1u << (p & 31));
                }
                if (a < run) run = a;
            }
            __syncthreads();
            // few records: thread 0 replays their swaps; many (the extended matrix ties by the hundred): the runs between
            // strict records are applied as permutations by the whole block
            constexpr int kSerialRecords = MOT_JV_SERIAL_RECORDS, kMaxRuns = 30;
            if (tid == 0) {
                int n_rec = 0, n_strict = 0;
                for (int q = (lo + 1) >> 5; q < words; ++q) {
                    n_rec += __popc(ev[q]);
                    unsigned sq = strict[q];
                    while (sq) {
                        const int b = __ffs((int)sq) - 1;
                        sq &= sq - 1;
                        if (n_strict < kMaxRuns) w.pi1[n_strict] = (q << 5) + b;
                        ++n_strict;
                    }
                }
                const bool serial = n_rec <= kSerialRecords || n_strict > kMaxRuns;
                if (serial) {
                    int h = lo + 1;
                    for (int q = (lo + 1) >> 5; q < words; ++q) {
                        unsigned bitsq = ev[q];
                        const unsigned sq = strict[q];
                        ev[q] = 0; strict[q] = 0;
                        while (bitsq) {
                            const int b = __ffs((int)bitsq) - 1;
                            bitsq &= bitsq - 1;
                            const int k = (q << 5) + b;
                            const int j = w.order[k];
                            if ((sq >> b) & 1u) h = lo;          // strictly smaller: the level restarts at lo (:131)
                            w.order[k] = w.order[h];
                            w.order[h++] = j;
                        }
                    }
                    w.ctl[0] = h;
                }
                w.ctl[1] = -1;
                w.ctl[2] = serial ? -1 : n_strict;
            }
            __syncthreads();
            const int n_runs = w.ctl[2];                         // strict records = run boundaries; -1: already replayed
            if (n_runs >= 0) {
                int h = lo + 1;
                for (int g = 0; g <= n_runs; ++g) {
                    const int pb = g == 0 ? lo + 1 : w.pi1[g - 1];
                    const int pe = g == n_runs ? N : w.pi1[g];
                    const int h0 = g == 0 ? lo + 1 : lo;         // a strict record restarts the level at lo (:131)
                    if (pe > pb) h = h0 + jv_block_apply_records(w, ev, strict, pb, pe, h0);
                    else h = h0;
                }
                for (int q = ((lo + 1) >> 5) + tid; q < words; q += nt) { ev[q] = 0; strict[q] = 0; }
                if (tid == 0) w.ctl[0] = h;
                __syncthreads();
            }
            hi = w.ctl[0];
            // the LAST free column of the level is the sink (:139-141)
            for (int k = lo + tid; k < hi; k += nt)
                if (w.y[w.order[k]] < 0) atomicMax(&w.ctl[1], k);
            __syncthreads();
            if (w.ctl[1] >= 0) sink = w.order[w.ctl[1]];
        }
        if (sink < 0) {
            // ---- relax from the level's columns, one after the other (:143-155)
            int slo = lo, shi = hi;
            bool hit = false;
            while (slo != shi && !hit) {
                const int jq = w.order[slo++];
                const int i = w.y[jq];
                const double level = w.dist[jq];
                const double base = c.at(i, jq) - w.v[jq] - level;
                bool mine = false;
                for (int k = shi + tid; k < N; k += nt) {
                    const int j = w.order[k];
                    const double cand = c.at(i, j) - w.v[j] - base;
                    if (cand < w.dist[j]) {
                        w.dist[j] = cand;
                        w.pred[j] = i;
                        if (cand == level) { atomicOr(&ev[k >> 5], 1u << (k & 31)); mine = true; }
                    }
                }
                if (__syncthreads_or(mine ? 1 : 0)) {
                    // some columns landed exactly on the level: replay them in position order
                    if (tid == 0) {
                        int s2 = shi, found = -1;
                        for (int q = shi >> 5; q < words; ++q) {
                            unsigned bitsq = ev[q];
                            ev[q] = 0;
                            while (bitsq && found < 0) {
                                const int b = __ffs((int)bitsq) - 1;
                                bitsq &= bitsq - 1;
                                const int k = (q << 5) + b;
                                const int j = w.order[k];
                                if (w.y[j] < 0) { found = j; break; }
                                w.order[k] = w.order[s2];
                                w.order[s2++] = j;
                            }
                        }
                        w.ctl[0] = s2;
                        w.ctl[1] = found;
                    }
                    __syncthreads();
                    shi = w.ctl[0];
                    if (w.ctl[1] >= 0) { sink = w.ctl[1]; hit = true; }
                    __syncthreads();                          // ctl is rewritten by the next event
                }
            }
            if (!hit) { lo = slo; hi = shi; }                // on a hit the caller's lo / hi stay put (:152)
        }
    }
    const double level = w.dist[w.order[lo]];
    __syncthreads();
    for (int k = tid; k < settled; k += nt) {
        const int j = w.order[k];
        w.v[j] += w.dist[j] - level;
    }
    if (tid == 0) {
        int j = sink, i;
        do {                                                 // flip the path (:203-208)
            i = w.pred[j];
            w.y[j] = i;
            const int prev = w.x[i];
            w.x[i] = j;
            j = prev;
        } while (i != start);
    }
    __syncthreads();
}

// All threads of the block.  On return x[0..N) / y[0..N) hold the square assignment (visible to the block).
static __device__ __noinline__ void block_dense_lapjv(const JvCost c, int N, const JvBlockWork w, BlockScratch* bs) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
    const int n = c.n, m = c.m;
    __syncthreads();
    // ---- column reduction (lap_solver.hpp:36-52): per column the minimum over rows, lowest row on ties.
    //      Rows n..N-1 hold one constant per column block, so only the first of them can win a strict '<'.
    for (int j = tid; j < N; j += nt) {
        double vj = kJvBig;
        int yj = 0;
        if (j < m) {
            for (int i = 0; i < n; ++i) {
                const double val = (double)c.dense[(size_t)i * c.ld + j];
                if (val < vj) { vj = val; yj = i; }
            }
            if (n < N && c.half < vj) { vj = c.half; yj = n; }
        } else {
            if (n > 0 && c.half < vj) { vj = c.half; yj = 0; }
            if (n < N && 0.0 < vj) { vj = 0.0; yj = n; }
        }
        w.v[j] = vj; w.y[j] = yj; w.x[j] = -1; w.cnt[j] = 0;
    }
    __syncthreads();
    // right-to-left claim (:53-62): the highest column keeps the row, every other column of that row is released
    for (int j = tid; j < N; j += nt) { atomicMax(&w.x[w.y[j]], j); atomicAdd(&w.cnt[w.y[j]], 1); }
    __syncthreads();
    for (int j = tid; j < N; j += nt)
        if (w.x[w.y[j]] != j) w.y[j] = -1;
    __syncthreads();
    // ---- free rows (ascending) and reduction transfer (:63-72)
    int n_free = block_compact(N, 0, bs, [&](int i) { return w.x[i] < 0; }, [&](int i, int pos) { w.fr[pos] = i; });
    // rows with exactly one column, in ascending order; `pred` is free until the augmentation phase
    const int n_sole = block_compact(N, 0, bs, [&](int i) { return w.x[i] >= 0 && w.cnt[i] == 1; },
                                     [&](int i, int pos) { w.pred[pos] = i; });
    {
        const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
        for (int r = 0; r < n_sole; ++r) {
            const int i = w.pred[r];
            const int xi = w.x[i];
            double second = kJvBig;
            for (int k = tid; k < N; k += nt) {
                if (k == xi) continue;
                const double red = c.at(i, k) - w.v[k];
                if (red < second) second = red;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(kFullMask, second, o); if (t < second) second = t; }
            double* slot = (r & 1) ? w.pv2 : w.pv1;                 // double-buffered: one barrier per row
            if (lane == 0) slot[warp] = second;
            __syncthreads();
            if (xi % nt == tid) {                                   // the owner of column xi applies the update
                double s2 = slot[0];
                for (int k = 1; k < nw; ++k) { const double t = slot[k]; if (t < s2) s2 = t; }
                w.v[xi] -= s2;
            }
        }
    }
    __syncthreads();
    // ---- augmenting row reduction, at most two sweeps (:74-113, :221-224)
    for (int sweep = 0; sweep < 2 && n_free > 0; ++sweep) {
        const unsigned total = (unsigned)n_free;
        unsigned cur = 0, rounds = 0;
        int kept = 0;
        while (cur < total) {
            ++rounds;
            const int i = w.fr[cur++];
            JvTop2 t{1.0e300, 0x7fffffff, 1.0e300, 0x7fffffff};
            for (int j = tid; j < N; j += nt) {
                const double red = c.at(i, j) - w.v[j];
                if (red < t.v1) { t.v2 = t.v1; t.i2 = t.i1; t.v1 = red; t.i1 = j; }
                else if (red < t.v2) { t.v2 = red; t.i2 = j; }
            }
            t = jv_block_top2(t, w);
            if (tid == 0) {
                const double u1 = t.v1;
                int j1 = t.i1;
                double u2 = t.v2;
                int j2 = t.i2;
                if (!(u2 < kJvBig)) { u2 = kJvBig; j2 = -1; }      // the scan only accepts a runner-up below LARGE
                int owner = w.y[j1];
                const double lowered = w.v[j1] - (u2 - u1);
                const bool strictly_lower = lowered < w.v[j1];
                if (rounds < cur * (unsigned)N) {                  // unsigned arithmetic as in :101
                    if (strictly_lower) w.v[j1] = lowered;
                    else if (owner >= 0 && j2 >= 0) { j1 = j2; owner = w.y[j2]; }
                    if (owner >= 0) {
                        if (strictly_lower) w.fr[--cur] = owner;   // re-process the displaced row now
                        else w.fr[kept++] = owner;
                    }
                } else if (owner >= 0) {
                    w.fr[kept++] = owner;
                }
                w.x[i] = j1;
                w.y[j1] = i;
                w.ctl[0] = (int)cur;
                w.ctl[1] = kept;
            }
            __syncthreads();
            cur = (unsigned)w.ctl[0];
            kept = w.ctl[1];
        }
        n_free = kept;
        __syncthreads();
    }
    // ---- shortest augmenting paths for what is still free (:195-211)
    for (int f = 0; f < n_free; ++f) jv_block_augment_from(c, N, w, w.fr[f]);
    __syncthreads();
}

}  // namespace mot
