// strongsort_kernel.cuh - StrongSORT's whole per-frame update() as one kernel, one CTA per camera stream.
// ReID inference and ECC camera-motion estimation are outside the association hot path: embeddings arrive with the
// detections and the warp is the identity (what motion::ECC::apply yields on its first call / on failure), but
// Track::camera_update's xyah -> tlbr -> xyah re-rounding is still applied every frame.
// Replaces reference src/trackers/strongsort.cpp:
//   StrongSORT::update: min_conf filter, tlwh / xyah detections (:829-940), output rows (:946-972)   -> phases A, K
//   Track::camera_update (:111-132), Tracker::predict / Track::predict (:139-145, :608-612)          -> phases B, C
//   Tracker::match: gated appearance metric over the confirmed tracks - NearestNeighborDistanceMetric::distance
//     (:240-334) + gate_cost_matrix (:451-492) + min_cost_matching (:344-416)                        -> phases D, E
//     IoU stage over unconfirmed ++ just-missed tracks (:728-763, iou_cost :538-585)                 -> phase F
//   Tracker::update: Track::update with the NSA Kalman update and the feature EMA (:147-187), mark_missed (:189-195),
//     initiate_track (:768-770), deletion, NearestNeighborDistanceMetric::partial_fit (:213-238)     -> phases G .. J
// Reference behaviours kept on purpose (oracle/strongsort.cpp q1-q3): an EMPTY index list means ALL - so with no
// confirmed track the IoU stage sees every tentative track twice, a matched tentative track whose second copy stays
// unmatched is deleted right after its update, the second copy's detection is swallowed; with no IoU candidate the IoU
// stage runs over all tracks; with every detection appearance-matched it runs over all detections and re-spawns them.
// The gallery of a track is a ring of its last `budget` per-frame smoothed features, stored already re-normalised the way
// cosine_distance re-normalises them (:316-324).  Only Mahalanobis-gate-passing pairs can be candidates, so the exact
// fp32 nearest-neighbour cosine (oracle summation order "lanes32") is evaluated for those pairs only, one warp per pair;
// the dense gallery x detections contraction is available on the tensor cores as mot_cost_nn_cosine.
// Exact ties arise only between the two copies of a duplicated row.  While rows + columns <= 384 the frame's IoU
// assignment is redone by the reference's own dense LAPJV (jv_device.cuh): the reference's answer.  Larger problems keep
// the sparse solver with + (j + 1) 2^-50 on the second copies (oracle tie_mode 2 = this policy; DESIGN.md "Ties").
#pragma once
#include "shapes.cuh"
#include "block_utils.cuh"
#include "cost_device.cuh"
#include "gate_device.cuh"
#include "kf_device.cuh"
#include "lap_device.cuh"
#include "jv_device.cuh"
#include "jv_block_device.cuh"
#include "bytetrack_kernel.cuh"      // header / error enums, BtShape
#include "botsort_kernel.cuh"        // lanes32 feature arithmetic (warp_dot, lanes32_reduce, dot4)

namespace mot {

#ifndef MOT_SS_THREADS
#define MOT_SS_THREADS 512
#endif
constexpr int kSsThreads = MOT_SS_THREADS;
constexpr int kSsWarpList = 96;             // per-warp buffer of gate-passing detections (flushed when fewer than 32 slots remain)
constexpr int kSsTableSlots = 4096;          // (row, det) -> blended appearance cost of the candidate pairs
enum : int { kSsTentative = 1, kSsConfirmed = 2, kSsDeleted = 3 };       // strongsort.hpp TrackState
constexpr unsigned char kSsHasFeat = 0x10;
enum : int { kErrTable = 16 };
// duplicated-row ties (q1) are resolved by the reference's own dense LAPJV: one warp (jv_device.cuh) while rows + columns <=
// kSsJvMax, the whole CTA over per-stream global scratch (jv_block_device.cuh) above that
#ifndef MOT_SS_JVMAX
#define MOT_SS_JVMAX 384
#endif
constexpr int kSsJvMax = MOT_SS_JVMAX;
static_assert(sizeof(unsigned long long) * kSsTableSlots >= jv_work_bytes(kSsJvMax + 1), "the LAPJV work area aliases the candidate table");
// header ints: kHdrActive = live tracks, kHdrLost = cumulative count of dense-LAPJV solves, kHdrFree, kHdrIdCounter (next_id - 1), kHdrFrame, kHdrError, then
enum : int { kSHdrRowsA = 6, kSHdrColsA = 7, kSHdrRowsB = 8, kSHdrColsB = 9, kSHdrMatchA = 10, kSHdrMatchB = 11,
             kSHdrDup = 12, kSHdrSpawn = 13, kSHdrGatePairs = 14 };

struct SsParams {
    float min_conf, max_cos_dist, max_iou_dist, mc_lambda, ema_alpha;
    int max_age, n_init, budget, dim;
};

struct SsLayout {
    int cap, d_max, dim, budget;
    size_t off_lists, off_state, off_meta, off_recs, off_feat, off_gal, off_ring, off_dfeat, off_dnorm, off_grad, off_jv, off_jvw, off_gscratch, stride;
    static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }
    static SsLayout make(int cap, int d_max, int dim, int budget) {
        SsLayout L{};
        L.cap = cap; L.d_max = d_max; L.dim = dim; L.budget = budget;
        size_t o = al(sizeof(int) * kHdrInts);
        L.off_lists = o;    o = al(o + sizeof(unsigned short) * 2 * (size_t)cap);
        L.off_state = o;    o = al(o + (size_t)cap);
        L.off_meta = o;     o = al(o + sizeof(int) * 8 * (size_t)cap);
        L.off_recs = o;     o = al(o + sizeof(float) * kRecFloats * (size_t)cap);
        L.off_feat = o;     o = al(o + sizeof(float) * (size_t)dim * (size_t)cap);
        L.off_gal = o;      o = al(o + sizeof(float) * (size_t)dim * (size_t)cap);
        L.off_ring = o;     o = al(o + sizeof(float) * (size_t)dim * (size_t)budget * (size_t)cap);
        L.off_dfeat = o;    o = al(o + sizeof(float) * (size_t)dim * (size_t)d_max);
        L.off_dnorm = o;    o = al(o + sizeof(float) * (size_t)d_max);
        L.off_grad = o;     o = al(o + sizeof(float) * (size_t)cap);
        L.off_jv = o;       o = al(o + sizeof(float) * (size_t)cap * (size_t)d_max);          // dense cost matrix of the exact-tie path
        L.off_jvw = o;      o = al(o + jv_block_gbytes(cap + d_max));                        // its LAPJV work arrays
        L.off_gscratch = o; o = al(o + lap_gscratch_bytes(cap, d_max));
        L.stride = o;
        return L;
    }
};

struct SsStream {
    int* hdr;
    unsigned short *list, *freel;      // live tracks in the reference's vector order; free slots
    unsigned char* state;              // TrackState | kSsHasFeat
    int *id, *hits, *tsu, *cls, *det_ind, *ring_n, *ring_pos;
    float* conf;
    float* recs;
    float* feat;                       // [cap][dim] features.back(): the smoothed, normalised feature
    float* gal;                        // [cap][dim] the same vector re-normalised (what cosine_distance multiplies)
    float* ring;                       // [cap][budget][dim] gallery: the last `budget` per-frame copies of gal
    float* dfeat;                      // [d_max][dim] this frame's detection features, normalised (filtered index)
    float* dnorm;                      // [d_max] |raw feature|
    float* grad;                       // [cap] gallery radius: max over the ring of |sample - gal| (slightly inflated)
    float* jv_dense;                   // [cap * d_max] dense (clamped) IoU cost matrix of the exact-tie path
    unsigned char* jv_work;            // jv_block_gbytes(cap + d_max): work arrays of the CTA-wide LAPJV
    unsigned char* gscratch;
    __device__ __forceinline__ static SsStream at(unsigned char* base, const SsLayout& L) {
        SsStream s;
        s.hdr = (int*)base;
        s.list = (unsigned short*)(base + L.off_lists);
        s.freel = s.list + L.cap;
        s.state = base + L.off_state;
        int* m = (int*)(base + L.off_meta);
        s.id = m; s.hits = m + L.cap; s.tsu = m + 2 * L.cap; s.cls = m + 3 * L.cap; s.det_ind = m + 4 * L.cap;
        s.ring_n = m + 5 * L.cap; s.ring_pos = m + 6 * L.cap; s.conf = (float*)(m + 7 * L.cap);
        s.recs = (float*)(base + L.off_recs);
        s.feat = (float*)(base + L.off_feat);
        s.gal = (float*)(base + L.off_gal);
        s.ring = (float*)(base + L.off_ring);
        s.dfeat = (float*)(base + L.off_dfeat);
        s.dnorm = (float*)(base + L.off_dnorm);
        s.grad = (float*)(base + L.off_grad);
        s.jv_dense = (float*)(base + L.off_jv);
        s.jv_work = base + L.off_jvw;
        s.gscratch = base + L.off_gscratch;
        return s;
    }
};

struct SsArgs {
    unsigned char* state;
    SsLayout L;
    const float* dets;        // [T][S][ld_dets][6]
    const int* n_dets;        // [T][S]
    const float* embs;        // [T][S][ld_dets][dim] or nullptr
    float* out;               // [T][S][ld_out][8]
    int* n_out;               // [T][S]
    int T, S, ld_dets, ld_out;
    int s_begin, s_end;
    SsParams p;
};

struct SsSmem {
    float4* det_tlwh;           // [d_max] by filtered index
    float4* det_box;            // [d_max] (x, y, x + w, y + h): the corners iou_matching::iou forms
    float4* det_z;              // [d_max] Detection::to_xyah
    float* det_conf;            // [d_max]
    unsigned short* dsel;       // [d_max] filtered index -> row of the input matrix (det_ind)
    unsigned short* cols;       // [d_max] IoU-stage columns (filtered indices)
    unsigned short* spawn;      // [d_max]
    unsigned char* det_a;       // [d_max] matched by the appearance stage
    unsigned short* rows_a;     // [cap] appearance-stage rows (track positions)
    unsigned short* rows_b;     // [cap] IoU-stage rows (track positions, possibly twice)
    unsigned short* unconf;     // [cap]
    short* match_det;           // [cap] by track position: detection (filtered index) this track is updated with, -1 none
    int* claim;                 // [cap] by track position: first IoU-stage row allowed to update it
    int* tflag;                 // [cap] by track position: bit0 matched by appearance, bit1 missed
    unsigned short* upd;        // [cap] positions to update
    unsigned short* list_new;   // [cap]
    unsigned long long* cache;  // [kSsTableSlots]
    unsigned short* wlist;      // [warps][kSsWarpList]
    float* wgd;                 // [warps][kSsWarpList]
    int* ectl;                  // [4] appearance-stage row counter
    BlockScratch* bs;
    LapWorkspace lap;
};

MOT_HD constexpr size_t ss_smem_bytes(int cap, int d_max, int e_cap) {
    size_t b = 0;
    b += 3 * lap_align16(sizeof(float4) * (size_t)d_max);
    b += lap_align16(sizeof(float) * (size_t)d_max);
    b += 3 * lap_align16(sizeof(unsigned short) * (size_t)d_max);
    b += lap_align16((size_t)d_max);
    b += 3 * lap_align16(sizeof(unsigned short) * (size_t)cap);
    b += lap_align16(sizeof(short) * (size_t)cap);
    b += 2 * lap_align16(sizeof(int) * (size_t)cap);
    b += 2 * lap_align16(sizeof(unsigned short) * (size_t)cap);
    b += lap_align16(sizeof(unsigned long long) * kSsTableSlots);
    b += lap_align16(sizeof(unsigned short) * (kSsThreads / 32) * kSsWarpList) + lap_align16(sizeof(float) * (kSsThreads / 32) * kSsWarpList) + 16;
    b += lap_align16(sizeof(BlockScratch));
    b += lap_smem_bytes(cap, d_max, e_cap);
    return b;
}

__device__ __forceinline__ void ss_carve(unsigned char* p, int cap, int d_max, int e_cap, SsSmem& s) {
    s.det_tlwh = (float4*)p;           p += lap_align16(sizeof(float4) * (size_t)d_max);
    s.det_box = (float4*)p;            p += lap_align16(sizeof(float4) * (size_t)d_max);
    s.det_z = (float4*)p;              p += lap_align16(sizeof(float4) * (size_t)d_max);
    s.det_conf = (float*)p;            p += lap_align16(sizeof(float) * (size_t)d_max);
    s.dsel = (unsigned short*)p;       p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.cols = (unsigned short*)p;       p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.spawn = (unsigned short*)p;      p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.det_a = p;                       p += lap_align16((size_t)d_max);
    s.rows_a = (unsigned short*)p;     p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.rows_b = (unsigned short*)p;     p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.unconf = (unsigned short*)p;     p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.match_det = (short*)p;           p += lap_align16(sizeof(short) * (size_t)cap);
    s.claim = (int*)p;                 p += lap_align16(sizeof(int) * (size_t)cap);
    s.tflag = (int*)p;                 p += lap_align16(sizeof(int) * (size_t)cap);
    s.upd = (unsigned short*)p;        p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.list_new = (unsigned short*)p;   p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.cache = (unsigned long long*)p;  p += lap_align16(sizeof(unsigned long long) * kSsTableSlots);
    s.wlist = (unsigned short*)p;      p += lap_align16(sizeof(unsigned short) * (kSsThreads / 32) * kSsWarpList);
    s.wgd = (float*)p;                 p += lap_align16(sizeof(float) * (kSsThreads / 32) * kSsWarpList);
    s.ectl = (int*)p;                  p += 16;
    s.bs = (BlockScratch*)p;           p += lap_align16(sizeof(BlockScratch));
    lap_carve(p, cap, d_max, e_cap, s.lap);
}

// Track::to_tlwh (strongsort.cpp:93-99) from the record's mean
__device__ __forceinline__ float4 ss_track_tlwh(const float* rec) {
    const float w = xmul(rec[2], rec[3]), h = rec[3];
    return make_float4(xsub(rec[0], xdiv(w, 2.0f)), xsub(rec[1], xdiv(h, 2.0f)), w, h);
}
__device__ __forceinline__ float4 ss_corners(float4 t) { return make_float4(t.x, t.y, xadd(t.x, t.z), xadd(t.y, t.w)); }   // to_tlbr (:101-109)

// Appearance-stage cost for block_lap_solve: every candidate pair sits in the table with its blended cost.
struct SsAppCost {
    static constexpr bool kWarpPerRow = false;
    static constexpr bool kGrid = false;
    unsigned long long* cache;
    struct Row { int i; };
    __device__ __forceinline__ static unsigned tag_of(int i, int d) { return 0x80000000u | ((unsigned)i << 16) | (unsigned)d; }
    __device__ __forceinline__ static unsigned home_of(unsigned tag) { return (tag * 2654435761u) >> 20; }
    __device__ __forceinline__ bool insert(int i, int d, float v) const {
        const unsigned tag = tag_of(i, d);
        const unsigned long long entry = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
        unsigned h = home_of(tag);
        for (int probe = 0; probe < kSsTableSlots; ++probe) {
            const unsigned long long old = atomicCAS(&cache[h], 0ull, entry);
            if (old == 0ull || (unsigned)(old >> 32) == tag) return true;
            h = (h + 1) & (kSsTableSlots - 1);
        }
        return false;
    }
    __device__ __forceinline__ float pair(int i, int d) const {
        const unsigned tag = tag_of(i, d);
        unsigned h = home_of(tag);
        for (int probe = 0; probe < kSsTableSlots; ++probe) {
            const unsigned long long hit = cache[h];
            if ((unsigned)(hit >> 32) == tag) return __uint_as_float((unsigned)hit);
            if (hit == 0ull) break;
            h = (h + 1) & (kSsTableSlots - 1);
        }
        return kInftyCost;
    }
    __device__ __forceinline__ Row row(int i) const { return Row{i}; }
    __device__ __forceinline__ bool reject(const Row&, int) const { return false; }
    __device__ __forceinline__ float cost(const Row& r, int j) const { return pair(r.i, j); }
    __device__ __forceinline__ bool is_candidate(const Row& r, int, int j, float thresh) const { return pair(r.i, j) <= thresh; }
    __device__ __forceinline__ double pair_bias(int, int) const { return 0.0; }
};

// IoU-stage cost (iou_matching::iou_cost, strongsort.cpp:538-585): rows = track positions (possibly listed twice),
// columns = filtered detection indices.  The second copies of duplicated rows carry the tie-break infinitesimal.
struct SsIouCost {
    static constexpr bool kWarpPerRow = false;
    static constexpr bool kGrid = true;
    const unsigned short* rows;        // row -> track position
    const unsigned short* list;        // position -> slot
    const float* recs;
    const int* tsu;
    const float4* det_tlwh;
    const float4* det_box;
    const unsigned short* col_map;
    int dup_first;
    bool prune;
    struct Row { float4 b; float4 t; bool stale; };
    __device__ __forceinline__ Row row(int i) const {
        const int slot = list[rows[i]];
        Row r;
        r.t = ss_track_tlwh(recs + (size_t)slot * kRecFloats);
        r.b = ss_corners(r.t);
        r.stale = tsu[slot] > 1;                                         // :567-570
        return r;
    }
    __device__ __forceinline__ float4 col_box(int j) const { return det_box[col_map[j]]; }
    __device__ __forceinline__ bool reject(const Row& r, int j) const {
        return r.stale || (prune && boxes_disjoint(r.b, det_box[col_map[j]]));
    }
    __device__ __forceinline__ float cost(const Row& r, int j) const {
        if (r.stale) return kInftyCost;
        return xsub(1.0f, iou_tlwh_pair(r.t, det_tlwh[col_map[j]]));
    }
    __device__ __forceinline__ float pair(int i, int j) const { return cost(row(i), j); }
    __device__ __forceinline__ bool is_candidate(const Row& r, int, int j, float thresh) const { return cost(r, j) <= thresh; }
    __device__ __forceinline__ double pair_bias(int i, int j) const {
        return (dup_first > 0 && i >= dup_first) ? (double)(j + 1) * 0x1p-50 : 0.0;
    }
};

// v / |v| into dst when |v| > 1e-10 (else a plain copy), all 32 lanes; returns |v|.  dim % 4 == 0.
__device__ __forceinline__ float warp_unit_or_same(const float* __restrict__ src, float* __restrict__ dst, int dim) {
    const int nq = dim >> 2, lane = lane_id();
    const float4* sv = reinterpret_cast<const float4*>(src);
    float4* dv = reinterpret_cast<float4*>(dst);
    float acc = 0.0f;
    for (int q = lane; q < nq; q += 32) { const float4 v = sv[q]; acc = dot4(acc, v, v); }
    const float nrm = xsqrt(lanes32_reduce(acc));
    const bool div = nrm > 1e-10f;
    for (int q = lane; q < nq; q += 32) {
        float4 v = sv[q];
        if (div) v = make_float4(xdiv(v.x, nrm), xdiv(v.y, nrm), xdiv(v.z, nrm), xdiv(v.w, nrm));
        dv[q] = v;
    }
    return nrm;
}

template <int CAP, int DMAX>
__device__ __forceinline__ void ss_frame(const SsArgs& a, const SsStream& st, SsSmem& sm, const float* dets, const float* embs,
                                         int n_det_in, float* out, int* n_out) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
    const int lane = tid & 31, g = lane & 7, base = lane & ~7, warp = tid >> 5, nwarps = nt >> 5;
    const int groups = nt >> 3, gid = tid >> 3;
    const int dim = a.p.dim, budget = a.p.budget;
    __syncthreads();
    const int n_trk = st.hdr[kHdrActive];
    int n_free = st.hdr[kHdrFree];
    const int id_base = st.hdr[kHdrIdCounter];
    const int frame_no = st.hdr[kHdrFrame];
    int n_in = n_det_in < 0 ? 0 : n_det_in;
    if (n_in > min(DMAX, a.ld_dets)) { n_in = min(DMAX, a.ld_dets); if (tid == 0) atomicOr(&st.hdr[kHdrError], (int)kErrTooManyDets); }

    // ---- A. detections with conf >= min_conf (:849-854): tlwh (:923-932), corners, xyah (:33-40)
    const float min_conf = a.p.min_conf;
    const int n_d = block_compact(n_in, 0, sm.bs, [&](int j) { return dets[(size_t)j * 6 + 4] >= min_conf; },
                                  [&](int j, int pos) {
                                      const float* r = dets + (size_t)j * 6;
                                      const float4 t = make_float4(r[0], r[1], xsub(r[2], r[0]), xsub(r[3], r[1]));
                                      sm.dsel[pos] = (unsigned short)j;
                                      sm.det_tlwh[pos] = t;
                                      sm.det_box[pos] = ss_corners(t);
                                      sm.det_z[pos] = tlwh2xyah_strong(t);
                                      sm.det_conf[pos] = r[4];
                                      sm.det_a[pos] = 0;
                                  });
    const bool have_feat = n_d > 0 && dim > 0 && embs != nullptr;

    // ---- B. camera_update with the identity warp (:111-132), only when there are detections (:873-878)
    if (n_d > 0) {
        for (int k = tid; k < n_trk; k += nt) {
            float* rec = st.recs + (size_t)st.list[k] * kRecFloats;
            const float4 c = ss_corners(ss_track_tlwh(rec));
            const float x1 = xadd(c.x, 0.0f), y1 = xadd(c.y, 0.0f), x2 = xadd(c.z, 0.0f), y2 = xadd(c.w, 0.0f);   // W p, W = I
            const float w = xsub(x2, x1), h = xsub(y2, y1);
            rec[0] = xadd(x1, xdiv(w, 2.0f)); rec[1] = xadd(y1, xdiv(h, 2.0f)); rec[2] = xdiv(w, h); rec[3] = h;
        }
    }
    __syncthreads();

    // ---- C. predict every track (:608-612): Kalman predict, time_since_update += 1
    {
        const int rounds = (n_trk + groups - 1) / groups;
        for (int it = 0; it < rounds; ++it) {
            const int r = it * groups + gid;
            const bool live = r < n_trk;
            const int slot = live ? (int)st.list[r] : 0;
            float* rec = st.recs + (size_t)slot * kRecFloats;
            KfRow s;
            if (live) kf_load_row(rec, g, s);
            else { s.m = 1.0f; for (int j = 0; j < 8; ++j) s.p[j] = (j == g) ? 1.0f : 0.0f; }
            kf_xyah_predict(s, g, base, false);
            if (live) { kf_store_row(rec, g, s); if (g == 0) st.tsu[slot] += 1; }
        }
    }
    // detection features, normalised once per frame (cosine_distance :325-329, Track ctor :84-90, Track::update :158-164)
    if (have_feat)
        for (int k = warp; k < n_d; k += nwarps) {
            const float nrm = warp_unit_or_same(embs + (size_t)sm.dsel[k] * dim, st.dfeat + (size_t)k * dim, dim);
            if (lane == 0) st.dnorm[k] = nrm;
        }
    for (int k = tid; k < n_trk; k += nt) { sm.match_det[k] = -1; sm.tflag[k] = 0; sm.claim[k] = 0x7fffffff; }
    __syncthreads();

    // ---- D. confirmed / unconfirmed split (:705-713)
    const int n_c = block_compact(n_trk, 0, sm.bs, [&](int k) { return (st.state[st.list[k]] & 0x0f) == kSsConfirmed; },
                                  [&](int k, int pos) { sm.rows_a[pos] = (unsigned short)k; });
    const int n_u = block_compact(n_trk, 0, sm.bs, [&](int k) { return (st.state[st.list[k]] & 0x0f) != kSsConfirmed; },
                                  [&](int k, int pos) { sm.unconf[pos] = (unsigned short)k; });

    // ---- E. appearance stage (:715-726).  With no confirmed track the reference runs it over ALL tracks, none of which
    //         has a gallery: every cost is 1e5, nothing matches, everything comes back unmatched (q1).
    int n_match_a = 0, gate_pairs = 0;
    const bool stage_a = n_d > 0 && n_c > 0;
    if (stage_a) {
        for (int h = tid; h < kSsTableSlots; h += nt) sm.cache[h] = 0ull;
        if (tid == 0) sm.ectl[0] = 0;
        block_lap_begin(sm.lap, n_c, n_d);
        const SsAppCost app{sm.cache};
        if (have_feat) {
            // one warp per confirmed track: gate every detection (lanes), then the whole warp evaluates the nearest-
            // neighbour cosine of each gate-passing detection against the track's gallery ring
            const float thr = a.p.max_cos_dist, lam = a.p.mc_lambda;
            // rows are handed out dynamically (a row's cost depends on how many detections pass its gate)
            unsigned short* wl = sm.wlist + warp * kSsWarpList;               // this warp's gate-passing detections ...
            float* wg = sm.wgd + warp * kSsWarpList;                          // ... and their gating distances
            for (;;) {
                int r = 0;
                if (lane == 0) r = atomicAdd(&sm.ectl[0], 1);
                r = __shfl_sync(kFullMask, r, 0);
                if (r >= n_c) break;
                const int slot = st.list[sm.rows_a[r]];
                const int ns = st.ring_n[slot];
                if (ns == 0) continue;                                           // no samples: the row is 1e5 (:271)
                const GateRow gr = gate_prepare(st.recs + (size_t)slot * kRecFloats);
                const float* ring = st.ring + (size_t)slot * budget * dim;
                // Every gallery row lies within grad[slot] of the track's current re-normalised feature gal[slot], so
                // min_s (1 - a_s . b) >= (1 - gal . b) - grad |b|: ONE dot product proves most gate-passing pairs to be
                // non-candidates (different identities sit near cosine distance 1) and the ring is only walked for the rest.
                // The bound is evaluated with slack far above fp32 round-off, so it never changes a result.
                const float* center = st.gal + (size_t)slot * dim;
                const float rad = xadd(xmul(st.grad[slot], 1.001f), 1e-6f);
                const float slack = xadd(1e-4f, xmul(1e-6f, (float)dim));
                const bool can_prune = lam > 1e-3f && ns > 2;
                int nl = 0;                                                      // warp-uniform fill of wl / wg
                // the collected pairs: (1) the pruning bound, one pair per LANE (32 independent dot products in flight);
                // (2) the survivors' exact nearest-neighbour cosine, one gallery row per lane
                auto flush = [&]() {
                    __syncwarp();
                    for (int b0 = 0; b0 < nl; b0 += 32) {
                        const int k = b0 + lane;
                        bool keep = k < nl;
                        if (keep && can_prune) {
                            const float lower = xsub(xsub(xsub(1.0f, thread_dot_lanes32(center, st.dfeat + (size_t)wl[k] * dim, dim)), rad), slack);   // <= min_s c_s
                            const float tau = xadd(xdiv(xsub(thr, xmul(xsub(1.0f, lam), wg[k])), lam), slack);  // candidates have c <= tau
                            keep = !(lower > tau);
                        }
                        unsigned todo = __ballot_sync(kFullMask, keep);
                        while (todo) {
                            const int l = __ffs(todo) - 1;
                            todo &= todo - 1;
                            const int d = wl[b0 + l];
                            const float gdl = wg[b0 + l];
                            const float* f = st.dfeat + (size_t)d * dim;
                            float best = 3.0e38f;
                            for (int sidx = lane; sidx < ns; sidx += 32) {
                                const float c = xsub(1.0f, thread_dot_lanes32(ring + (size_t)sidx * dim, f, dim));   // :333
                                best = (c < best) ? c : best;                                              // minCoeff (:295)
                            }
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) { const float t = __shfl_xor_sync(kFullMask, best, o); best = (t < best) ? t : best; }
                            const float blended = gate_blend(best, gdl, lam, kInftyCost);                 // :484-487
                            if (lane == 0 && blended <= thr) {
                                if (app.insert(r, d, blended)) {
                                    const int e = atomicAdd(&sm.lap.ctl[0], 1);
                                    if (e < sm.lap.e_cap) sm.lap.scratch_a[e] = (r << 16) | d;
                                    else sm.lap.ctl[1] = 1;
                                } else {
                                    atomicOr(&st.hdr[kHdrError], (int)kErrTable);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    nl = 0;
                };
                for (int j0 = 0; j0 < n_d; j0 += 32) {
                    const int j = j0 + lane;
                    float gd = 0.0f;
                    bool pass = false;
                    if (j < n_d) { gd = gate_distance(gr, sm.det_z[j], false); pass = !(gd > kGatingThreshold); }
                    const unsigned m = __ballot_sync(kFullMask, pass);
                    if (m == 0u) continue;
                    if (pass) { const int pos = nl + __popc(m & ((1u << lane) - 1u)); wl[pos] = (unsigned short)j; wg[pos] = gd; }
                    nl += __popc(m);
                    if (lane == 0) atomicAdd(&sm.lap.ctl[7], __popc(m));
                    if (nl > kSsWarpList - 32) flush();
                }
                flush();
            }
        }
        __syncthreads();
        gate_pairs = sm.lap.ctl[7];
        __syncthreads();
        block_lap_solve(sm.lap, n_c, n_d, CAP, DMAX, a.p.max_cos_dist, app);
        for (int r = tid; r < n_c; r += nt) {
            const int c = sm.lap.row2col[r];
            if (c >= 0) { sm.match_det[sm.rows_a[r]] = (short)c; sm.tflag[sm.rows_a[r]] = 1; sm.det_a[c] = 1; }
        }
        __syncthreads();
        n_match_a = block_compact(n_c, 0, sm.bs, [&](int r) { return sm.lap.row2col[r] >= 0; }, [&](int, int) {});
    }

    // ---- F. IoU stage (:728-763).  Rows: unconfirmed ++ appearance-unmatched tracks with time_since_update == 1
    //         (every track again when nothing is confirmed); none => ALL tracks.  Columns: the appearance stage's
    //         leftover detections; none => ALL detections.
    int n_rb = 0, n_cb = 0, dup_first = 0, n_match_b = 0, n_spawn = 0;
    if (n_d > 0) {
        for (int k = tid; k < n_u; k += nt) sm.rows_b[k] = sm.unconf[k];
        n_rb = n_u;
        if (n_trk > 0) {
            // unmatched_tracks_a: the appearance rows left unmatched - all tracks in list order when n_c == 0
            const int n_a = (n_c > 0) ? n_c : n_trk;
            auto pos_of = [&](int r) { return (n_c > 0) ? (int)sm.rows_a[r] : r; };
            auto unmatched = [&](int r) { return (n_c > 0) ? (sm.tflag[sm.rows_a[r]] & 1) == 0 : true; };
            const int room = CAP - n_rb;
            const int want = block_compact(n_a, 0, sm.bs, [&](int r) { return unmatched(r) && st.tsu[st.list[pos_of(r)]] == 1; },
                                           [&](int r, int pos) { if (pos < room) sm.rows_b[n_rb + pos] = (unsigned short)pos_of(r); });
            for (int r = tid; r < n_a; r += nt)                                  // unmatched_tracks_a with tsu != 1 (:736-740)
                if (unmatched(r) && st.tsu[st.list[pos_of(r)]] != 1) sm.tflag[pos_of(r)] |= 2;
            if (want > room) { if (tid == 0) atomicOr(&st.hdr[kHdrError], (int)kErrCapacity); }
            if (n_c == 0 && n_u > 0 && want > 0) dup_first = n_u;
            n_rb += (want < room) ? want : room;
        }
        __syncthreads();
        if (n_rb == 0) {                                                         // "empty means all" (:358-361)
            for (int k = tid; k < n_trk; k += nt) sm.rows_b[k] = (unsigned short)k;
            n_rb = n_trk;
        }
        n_cb = block_compact(n_d, 0, sm.bs, [&](int j) { return sm.det_a[j] == 0; }, [&](int j, int pos) { sm.cols[pos] = (unsigned short)j; });
        if (n_cb == 0) {                                                         // :362-365
            for (int j = tid; j < n_d; j += nt) sm.cols[j] = (unsigned short)j;
            n_cb = n_d;
        }
        __syncthreads();
        if (n_rb > 0) {
            const SsIouCost cost{sm.rows_b, st.list, st.recs, st.tsu, sm.det_tlwh, sm.det_box, sm.cols, dup_first, a.p.max_iou_dist < 1.0f};
            block_lap(sm.lap, n_rb, n_cb, CAP, DMAX, a.p.max_iou_dist, cost);
            if (dup_first > 0) {
                // duplicated rows: the optimum is not unique, and the reference's answer is whatever its dense LAPJV yields
                // on the clamped matrix (min_cost_matching :372-379), non-candidate entries included
                const float maxd = a.p.max_iou_dist, capv = xadd(maxd, 1e-5f);
                for (int e = tid; e < n_rb * n_cb; e += nt) {
                    const float c = cost.pair(e / n_cb, e - (e / n_cb) * n_cb);
                    st.jv_dense[e] = (c > maxd) ? capv : c;
                }
                if (tid == 0) st.hdr[kHdrLost] += 1;
                __syncthreads();
                const JvCost jc{st.jv_dense, n_rb, n_cb, n_cb, (double)maxd / 2.0};
                const int* jx;
                const int* jy;
                if (n_rb + n_cb <= kSsJvMax) {
                    const JvWork w = jv_carve(reinterpret_cast<unsigned char*>(sm.cache), kSsJvMax + 1);
                    if (tid < 32) warp_dense_lapjv(jc, n_rb + n_cb, w);
                    jx = w.x; jy = w.y;
                } else {
                    // [row_label, row2col) of the sparse solver is idle here
                    const size_t idle_bytes = (size_t)((unsigned char*)sm.lap.row2col - (unsigned char*)sm.lap.row_label);
                    const JvBlockWork w = jv_block_carve(st.jv_work, (unsigned char*)sm.lap.row_label, n_rb + n_cb,
                                                         jv_block_sbytes_full(n_rb + n_cb) <= idle_bytes);
                    block_dense_lapjv(jc, n_rb + n_cb, w, sm.bs);
                    jx = w.x; jy = w.y;
                }
                __syncthreads();
                for (int i = tid; i < n_rb; i += nt) {                          // lap_solver.hpp:326-331, then :389-399
                    const int j = jx[i];
                    sm.lap.row2col[i] = (short)((j < n_cb && st.jv_dense[i * n_cb + j] <= maxd) ? j : -1);
                }
                for (int j = tid; j < n_cb; j += nt) {
                    const int i = jy[j];
                    sm.lap.col2row[j] = (short)((i < n_rb && st.jv_dense[i * n_cb + j] <= maxd) ? i : -1);
                }
                __syncthreads();
            }
            // combine (:743-759): an IoU match counts unless its track or its detection was matched by appearance, or an
            // earlier row of the same track already took one; the tracks of unmatched rows are missed (:761-765)
            for (int r = tid; r < n_rb; r += nt) {
                const int pos = sm.rows_b[r];
                const int c = sm.lap.row2col[r];
                if (c < 0) atomicOr(&sm.tflag[pos], 2);
                else if ((sm.tflag[pos] & 1) == 0 && sm.det_a[sm.cols[c]] == 0) atomicMin(&sm.claim[pos], r);
            }
            __syncthreads();
            for (int r = tid; r < n_rb; r += nt) {
                const int pos = sm.rows_b[r];
                const int c = sm.lap.row2col[r];
                if (c >= 0 && (sm.tflag[pos] & 1) == 0 && sm.det_a[sm.cols[c]] == 0 && sm.claim[pos] == r) sm.match_det[pos] = (short)sm.cols[c];
            }
            n_spawn = block_compact(n_cb, 0, sm.bs, [&](int c) { return sm.lap.col2row[c] < 0; },
                                    [&](int c, int pos) { sm.spawn[pos] = sm.cols[c]; });
        } else {                                                                 // no track at all: every column is left over (:367-369)
            for (int k = tid; k < n_cb; k += nt) sm.spawn[k] = sm.cols[k];
            n_spawn = n_cb;
        }
    } else {
        for (int k = tid; k < n_trk; k += nt) sm.tflag[k] |= 2;                  // no detection: every track is missed
    }
    __syncthreads();

    // ---- G. Track::update for every matched track (:147-187)
    const int n_upd = block_compact(n_trk, 0, sm.bs, [&](int k) { return sm.match_det[k] >= 0; },
                                    [&](int k, int pos) { sm.upd[pos] = (unsigned short)k; });
    n_match_b = n_upd - n_match_a;
    {
        const int rounds = (n_upd + groups - 1) / groups;
        for (int it = 0; it < rounds; ++it) {
            const int k = it * groups + gid;
            const bool live = k < n_upd;
            const int pos = live ? (int)sm.upd[k] : 0;
            const int slot = live ? (int)st.list[pos] : 0;
            const int d = live ? (int)sm.match_det[pos] : 0;
            float* rec = st.recs + (size_t)slot * kRecFloats;
            KfRow s;
            if (live) kf_load_row(rec, g, s);
            else { s.m = 1.0f; for (int j = 0; j < 8; ++j) s.p[j] = (j == g) ? 1.0f : 0.0f; }
            float z[4] = {0.0f, 0.0f, 1.0f, 1.0f};
            float conf = 0.0f;
            if (live) { const float4 q = sm.det_z[d]; z[0] = q.x; z[1] = q.y; z[2] = q.z; z[3] = q.w; conf = sm.det_conf[d]; }
            const bool ok = kf_xyah_update(s, g, base, z, conf);                 // NSA: R scaled by (1 - conf) (:153)
            if (live) {
                if (ok) kf_store_row(rec, g, s);
                if (g == 0) {
                    if (!ok) atomicOr(&st.hdr[kHdrError], (int)kErrKalman);
                    const int orig = sm.dsel[d];
                    st.conf[slot] = conf;
                    st.cls[slot] = (int)dets[(size_t)orig * 6 + 5];
                    st.det_ind[slot] = orig;
                    const int hits = st.hits[slot] + 1;
                    st.hits[slot] = hits;
                    st.tsu[slot] = 0;
                    if ((st.state[slot] & 0x0f) == kSsTentative && hits >= a.p.n_init)
                        st.state[slot] = (unsigned char)((st.state[slot] & 0xf0) | kSsConfirmed);
                }
            }
        }
    }
    __syncthreads();
    if (have_feat) {
        // feature EMA (:158-182), one warp per track; gal = the result re-normalised (cosine_distance :316-324)
        const float alpha = a.p.ema_alpha, beta = xsub(1.0f, alpha);
        const int nq = dim >> 2;
        for (int k = warp; k < n_upd; k += nwarps) {
            const int pos = sm.upd[k], slot = st.list[pos], d = sm.match_det[pos];
            if (st.dnorm[d] < 1e-10f) continue;                                   // zero-norm feature: skipped (:160-161)
            float4* fv = reinterpret_cast<float4*>(st.feat + (size_t)slot * dim);
            float4* gv = reinterpret_cast<float4*>(st.gal + (size_t)slot * dim);
            const float4* dv = reinterpret_cast<const float4*>(st.dfeat + (size_t)d * dim);
            const bool has = (st.state[slot] & kSsHasFeat) != 0;
            bool changed = true;
            if (has) {
                float acc = 0.0f;
                for (int q = lane; q < nq; q += 32) {
                    const float4 o = fv[q], f = dv[q];
                    const float4 v = make_float4(xadd(xmul(alpha, o.x), xmul(beta, f.x)), xadd(xmul(alpha, o.y), xmul(beta, f.y)),
                                                 xadd(xmul(alpha, o.z), xmul(beta, f.z)), xadd(xmul(alpha, o.w), xmul(beta, f.w)));
                    gv[q] = v;                                                    // parked in gal until its norm is known
                    acc = dot4(acc, v, v);
                }
                const float sn = xsqrt(lanes32_reduce(acc));
                changed = sn > 1e-10f;                                            // else the old feature stays (:173-176)
                if (changed)
                    for (int q = lane; q < nq; q += 32) {                         // every lane re-reads what it wrote
                        const float4 v = gv[q];
                        fv[q] = make_float4(xdiv(v.x, sn), xdiv(v.y, sn), xdiv(v.z, sn), xdiv(v.w, sn));
                    }
            } else {
                for (int q = lane; q < nq; q += 32) fv[q] = dv[q];
                __syncwarp();                                                     // every lane has read `has`
                if (lane == 0) st.state[slot] |= kSsHasFeat;
            }
            (void)changed;
            __syncwarp();
            warp_unit_or_same(st.feat + (size_t)slot * dim, st.gal + (size_t)slot * dim, dim);
        }
    }
    __syncthreads();

    // ---- H. mark_missed (:189-195), after the updates: a tentative track that was both updated and missed is deleted
    for (int k = tid; k < n_trk; k += nt)
        if (sm.tflag[k] & 2) {
            const int slot = st.list[k];
            const int stt = st.state[slot] & 0x0f;
            if (stt == kSsTentative || st.tsu[slot] > a.p.max_age) st.state[slot] = (unsigned char)((st.state[slot] & 0xf0) | kSsDeleted);
        }

    // ---- I. initiate_track for the IoU stage's leftover detections (:629-631, :768-770), ids in list order
    int n_new = n_spawn;
    if (n_new > n_free) { n_new = n_free; if (tid == 0) atomicOr(&st.hdr[kHdrError], (int)kErrCapacity); }
    for (int k = gid; k < n_new; k += groups) {
        const int d = sm.spawn[k];
        const int slot = st.freel[n_free - 1 - k];
        const float4 q = sm.det_z[d];
        const float z[4] = {q.x, q.y, q.z, q.w};
        KfRow s;
        kf_xyah_initiate(s, g, z);
        kf_store_row(st.recs + (size_t)slot * kRecFloats, g, s);
        if (g == 0) {
            const int orig = sm.dsel[d];
            st.id[slot] = id_base + 1 + k;
            st.hits[slot] = 1; st.tsu[slot] = 0;
            st.conf[slot] = sm.det_conf[d];
            st.cls[slot] = (int)dets[(size_t)orig * 6 + 5];
            st.det_ind[slot] = orig;
            st.ring_n[slot] = 0; st.ring_pos[slot] = 0;
            st.state[slot] = (unsigned char)(kSsTentative | ((have_feat && st.dnorm[d] > 1e-10f) ? kSsHasFeat : 0));   // :84-90
        }
    }
    if (have_feat)
        for (int k = warp; k < n_new; k += nwarps) {
            const int d = sm.spawn[k], slot = st.freel[n_free - 1 - k];
            if (!(st.dnorm[d] > 1e-10f)) continue;
            const float4* src = reinterpret_cast<const float4*>(st.dfeat + (size_t)d * dim);
            float4* dst = reinterpret_cast<float4*>(st.feat + (size_t)slot * dim);
            for (int q = lane; q < (dim >> 2); q += 32) dst[q] = src[q];
            __syncwarp();
            warp_unit_or_same(st.feat + (size_t)slot * dim, st.gal + (size_t)slot * dim, dim);
        }
    __syncthreads();

    // ---- J. drop the deleted tracks (:634-638), append the new ones, feed the galleries (partial_fit :213-238)
    int n_keep = block_compact(n_trk, 0, sm.bs, [&](int k) { return (st.state[st.list[k]] & 0x0f) != kSsDeleted; },
                               [&](int k, int pos) { sm.list_new[pos] = st.list[k]; });
    const int n_free_after = block_compact(n_trk, n_free - n_new, sm.bs, [&](int k) { return (st.state[st.list[k]] & 0x0f) == kSsDeleted; },
                                           [&](int k, int pos) { sm.upd[pos - (n_free - n_new)] = st.list[k]; });
    // (the freed slots are parked in upd[]: freel[n_free - n_new ..) still holds this frame's new slots)
    for (int k = tid; k < n_new; k += nt) sm.list_new[n_keep + k] = st.freel[n_free - 1 - k];
    __syncthreads();
    const int n_freed = n_free_after - (n_free - n_new);
    for (int k = tid; k < n_freed; k += nt) st.freel[n_free - n_new + k] = sm.upd[k];
    n_keep += n_new;
    for (int k = tid; k < n_keep; k += nt) st.list[k] = sm.list_new[k];
    __syncthreads();
    if (dim > 0)
        for (int k = warp; k < n_keep; k += nwarps) {
            const int slot = st.list[k];
            if ((st.state[slot] & 0x0f) != kSsConfirmed || (st.state[slot] & kSsHasFeat) == 0) continue;
            const int pos = st.ring_pos[slot];
            const float4* src = reinterpret_cast<const float4*>(st.gal + (size_t)slot * dim);
            float4* dst = reinterpret_cast<float4*>(st.ring + ((size_t)slot * budget + pos) * dim);
            for (int q = lane; q < (dim >> 2); q += 32) dst[q] = src[q];
            const int ns_old = st.ring_n[slot];
            __syncwarp();                                                         // every lane has read ring_pos / ring_n
            const int ns_new = (ns_old < budget) ? ns_old + 1 : budget;
            if (lane == 0) {
                st.ring_pos[slot] = (pos + 1 == budget) ? 0 : pos + 1;
                st.ring_n[slot] = ns_new;
            }
            // Gallery radius for next frame's pruning bound, max_s |ring[s] - gal|.  Exact every fourth frame (one ring row per
            // lane: the whole ring is read); in between the triangle inequality carries it: the centre moved by
            // |gal - previous gal| (the previous gal is the ring row appended last frame), so every old row is at most that much
            // farther away, and the new row is at distance 0.
            const auto dist2 = [&](const float4* rv) {
                float d2 = 0.0f;
                for (int q = 0; q < (dim >> 2); ++q) {
                    const float4 x = rv[q], c = src[q];
                    const float ex = x.x - c.x, ey = x.y - c.y, ez = x.z - c.z, ew = x.w - c.w;
                    d2 += ex * ex + ey * ey + ez * ez + ew * ew;
                }
                return d2;
            };
            if (ns_new <= 2 || ((frame_no + slot) & 3) == 0) {
                float worst = 0.0f;
                for (int sidx = lane; sidx < ns_new; sidx += 32)
                    worst = fmaxf(worst, dist2(reinterpret_cast<const float4*>(st.ring + ((size_t)slot * budget + sidx) * dim)));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) worst = fmaxf(worst, __shfl_xor_sync(kFullMask, worst, o));
                if (lane == 0) st.grad[slot] = sqrtf(worst);
            } else if (lane == 0) {
                const int prev = (pos == 0) ? budget - 1 : pos - 1;
                st.grad[slot] = st.grad[slot] + sqrtf(dist2(reinterpret_cast<const float4*>(st.ring + ((size_t)slot * budget + prev) * dim))) * 1.0001f + 1e-6f;
            }
        }

    // ---- K. output rows: confirmed tracks updated this frame (:946-972)
    const int n_rows = (n_d == 0) ? 0
        : block_compact(n_keep, 0, sm.bs, [&](int k) { const int slot = st.list[k]; return (st.state[slot] & 0x0f) == kSsConfirmed && st.tsu[slot] < 1; },
                        [&](int k, int pos) {
                            if (pos >= a.ld_out) return;
                            const int slot = st.list[k];
                            const float4 b = ss_corners(ss_track_tlwh(st.recs + (size_t)slot * kRecFloats));
                            float* w = out + (size_t)pos * 8;
                            *reinterpret_cast<float4*>(w) = b;
                            *reinterpret_cast<float4*>(w + 4) = make_float4((float)st.id[slot], st.conf[slot], (float)st.cls[slot], (float)st.det_ind[slot]);
                        });
    if (tid == 0) {
        if (n_rows > a.ld_out) atomicOr(&st.hdr[kHdrError], (int)kErrOutput);
        *n_out = n_rows < a.ld_out ? n_rows : a.ld_out;
        st.hdr[kHdrActive] = n_keep;
        st.hdr[kHdrFree] = n_free_after;
        st.hdr[kHdrIdCounter] = id_base + n_new;
        st.hdr[kHdrFrame] = st.hdr[kHdrFrame] + 1;
        st.hdr[kSHdrRowsA] = (n_d > 0 && n_trk > 0) ? (n_c > 0 ? n_c : n_trk) : 0; st.hdr[kSHdrColsA] = n_d;
        st.hdr[kSHdrRowsB] = n_rb; st.hdr[kSHdrColsB] = n_cb;
        st.hdr[kSHdrMatchA] = n_match_a; st.hdr[kSHdrMatchB] = n_match_b;
        st.hdr[kSHdrDup] = dup_first; st.hdr[kSHdrSpawn] = n_new; st.hdr[kSHdrGatePairs] = gate_pairs;
    }
    __syncthreads();
}

template <int CAP, int DMAX, int ECAP>
__global__ void __launch_bounds__(kSsThreads) strongsort_step_kernel(SsArgs a) {
    MOT_DYNAMIC_SMEM(smem);
    SsSmem sm;
    static_assert(lap_idle_bytes(CAP, DMAX, ECAP) >= jv_block_sbytes(CAP + DMAX), "the CTA-wide LAPJV's shared scratch must fit the sparse solver's idle arrays");
    ss_carve(smem, CAP, DMAX, ECAP, sm);
    for (int s = a.s_begin + (int)blockIdx.x; s < a.s_end; s += (int)gridDim.x) {
        SsStream st = SsStream::at(a.state + (size_t)s * a.L.stride, a.L);
        lap_carve_gscratch(st.gscratch, CAP, DMAX, sm.lap);
        for (int t = 0; t < a.T; ++t) {
            const size_t fs = (size_t)t * a.S + s;
            ss_frame<CAP, DMAX>(a, st, sm, a.dets + fs * (size_t)a.ld_dets * 6,
                                a.embs ? a.embs + fs * (size_t)a.ld_dets * a.p.dim : nullptr, a.n_dets[fs],
                                a.out + fs * (size_t)a.ld_out * 8, a.n_out + fs);
        }
    }
}

// Tracker::reset (strongsort.cpp:772-778): tracks and galleries cleared, next_id = 1
static __global__ void strongsort_reset_kernel(unsigned char* state, SsLayout L, int S) {
    for (int s = (int)blockIdx.x; s < S; s += (int)gridDim.x) {
        SsStream st = SsStream::at(state + (size_t)s * L.stride, L);
        for (int k = (int)threadIdx.x; k < L.cap; k += (int)blockDim.x) {
            st.freel[k] = (unsigned short)(L.cap - 1 - k);
            st.state[k] = (unsigned char)kSsDeleted;
        }
        if (threadIdx.x == 0) {
            for (int k = 0; k < kHdrInts; ++k) st.hdr[k] = 0;
            st.hdr[kHdrFree] = L.cap;
        }
        __syncthreads();
    }
}

// shapes the StrongSORT kernel is built for (track capacity, detections per frame, candidate-edge buffer)
constexpr BtShape kSsShapes[] = {{256, 64, 1024}, {1536, 512, 4096}};
constexpr int kNumSsShapes = sizeof(kSsShapes) / sizeof(kSsShapes[0]);

}  // namespace mot
