// botsort_kernel.cuh - BoT-SORT's whole per-frame update() as one kernel, one CTA per camera stream
// (cmc_method = "none": camera-motion compensation is image processing, outside the association hot
// path; ReID embeddings arrive with the detections).
// Replaces reference src/trackers/botsort.cpp:260-764 and BotSTrack (:15-193):
//   empty frame: immediate return, frame counter untouched (:267-269)              -> early exit
//   split_detections / create_detections, feature normalisation (:336-400, :39-46) -> phase A
//   unconfirmed / tracked split, pool = tracked ++ lost (:293-309)                 -> phase B
//   multi_predict IN PLACE over the pool only (:311, KalmanFilterXYWH::predict)    -> phase C
//   first association: iou_distance, proximity mask, optional fuse_score, cosine embedding distance / 2
//     with appearance threshold, element-wise min, linear_assignment(match_thresh) (:402-495) -> phases D, E
//   second association on the low-confidence detections, thresh 0.5, unmatched -> Lost (:497-562) -> phase F
//   unconfirmed tracks x leftover detections, thresh 0.7, unmatched -> Removed (:564-647)      -> phase G
//   new tracks with conf >= new_track_thresh (:649-667), lost expiry (:669-676)    -> phases H, I
//   prepare_output list algebra and output rows (:678-764)                         -> phases J, K
// Reference behaviours kept on purpose (oracle/botsort.cpp lists them): a re-found LOST track is updated and then
// dropped from both lists; unconfirmed tracks are never predicted; unmatched tracked tracks only become
// Lost when the second association actually runs; remove_duplicate_stracks is never called.
//
// The embedding term only matters where iou_distance <= proximity_thresh (everywhere else it is overwritten with
// 1, :458-460), so the kernel evaluates cosine distances for those pairs only, with the oracle's sequential
// fp32 sums: bit-identical costs.  The dense N x M x D contraction the reference computes is available as the
// tcgen05 kernel behind mot_cost_cosine (kernels_cosine.cuh).
#pragma once
#include "shapes.cuh"
#include "block_utils.cuh"
#include "cost_device.cuh"
#include "kf_device.cuh"
#include "kf_xywh_device.cuh"
#include "lap_device.cuh"
#include "bytetrack_kernel.cuh"      // state / flag / header enums shared with ByteTrack

namespace mot {

#ifndef MOT_BOT_THREADS
#define MOT_BOT_THREADS 1024
#endif
// 1024 threads x 1 CTA per SM (64 registers): the feature passes are DRAM-latency bound and want every resident warp;
// measured on the 1024 x 1024 x 512 case: 512 threads 679 us / frame, 1024 threads 577 us / frame.
constexpr int kBotThreads = MOT_BOT_THREADS;
constexpr unsigned char kFlagHasFeat = 0x20;
constexpr int kBotTableSlots = 4096;         // (pair -> embedding term) hash table entries in shared memory

enum : int { kBHdrNew = 12, kBHdrLostAfter = 13 };

struct BotParams {
    float track_high_thresh, track_low_thresh, new_track_thresh, match_thresh, proximity_thresh, appearance_thresh;
    int max_time_lost, fuse_first, with_reid, dim;
};

// per-stream slab; `dim` is a run-time size, so this layout is passed by value, not constexpr
struct BotLayout {
    int cap, d_max, dim;
    size_t off_lists, off_sflag, off_meta, off_recs, off_tnorm, off_dnorm, off_feats, off_dfeat, off_gscratch, stride;
    static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }
    static BotLayout make(int cap, int d_max, int dim) {
        BotLayout L{};
        L.cap = cap; L.d_max = d_max; L.dim = dim;
        size_t o = al(sizeof(int) * kHdrInts);
        L.off_lists = o;    o = al(o + sizeof(unsigned short) * 3 * (size_t)cap);
        L.off_sflag = o;    o = al(o + (size_t)cap);
        L.off_meta = o;     o = al(o + sizeof(int) * 7 * (size_t)cap);
        L.off_recs = o;     o = al(o + sizeof(float) * kRecFloats * (size_t)cap);
        L.off_tnorm = o;    o = al(o + sizeof(float) * (size_t)cap);
        L.off_dnorm = o;    o = al(o + sizeof(float) * 2 * (size_t)d_max);
        L.off_feats = o;    o = al(o + sizeof(float) * (size_t)dim * (size_t)cap);
        L.off_dfeat = o;    o = al(o + sizeof(float) * (size_t)dim * (size_t)d_max);
        L.off_gscratch = o; o = al(o + lap_gscratch_bytes(cap, d_max));
        L.stride = o;
        return L;
    }
};

struct BotStream {
    int* hdr;
    unsigned short *active, *lost, *freel;
    unsigned char* sflag;
    int *id, *tracklet_len, *frame_id, *start_frame, *cls, *det_ind;
    float* conf;
    float* recs;
    float* tnorm;              // [cap] |smooth_feat| of every track touched this frame
    float* dnorm;              // [d_max] |raw feature| then [d_max] |normalised feature| of every detection
    float* feats;              // [cap][dim] smooth_feat
    float* dfeat;              // [d_max][dim] normalised detection features of this frame
    unsigned char* gscratch;
    __device__ __forceinline__ static BotStream at(unsigned char* base, const BotLayout& L) {
        BotStream s;
        s.hdr = (int*)base;
        s.active = (unsigned short*)(base + L.off_lists);
        s.lost = s.active + L.cap;
        s.freel = s.lost + L.cap;
        s.sflag = base + L.off_sflag;
        int* m = (int*)(base + L.off_meta);
        s.id = m; s.tracklet_len = m + L.cap; s.frame_id = m + 2 * L.cap; s.start_frame = m + 3 * L.cap;
        s.cls = m + 4 * L.cap; s.det_ind = m + 5 * L.cap; s.conf = (float*)(m + 6 * L.cap);
        s.recs = (float*)(base + L.off_recs);
        s.tnorm = (float*)(base + L.off_tnorm);
        s.dnorm = (float*)(base + L.off_dnorm);
        s.feats = (float*)(base + L.off_feats);
        s.dfeat = (float*)(base + L.off_dfeat);
        s.gscratch = base + L.off_gscratch;
        return s;
    }
};

struct BotArgs {
    unsigned char* state;
    BotLayout L;
    const float* dets;        // [T][S][ld_dets][6]
    const int* n_dets;        // [T][S]
    const float* embs;        // [T][S][ld_dets][dim] or nullptr
    float* out;               // [T][S][ld_out][8]
    int* n_out;               // [T][S]
    int T, S, ld_dets, ld_out;
    int s_begin, s_end;
    BotParams p;
};

struct BotSmem {
    float4* det_box;            // [d_max] IoU box of every detection (xyxy -> xywh -> xyxy, as BotSTrack does)
    float* det_conf;            // [d_max]
    unsigned short* first;      // [d_max] conf > track_high_thresh
    unsigned short* second;     // [d_max] track_low_thresh < conf <= track_high_thresh
    unsigned short* udet;       // [d_max]
    unsigned short* udet2;      // [d_max]
    float4* row_box;            // [cap]
    unsigned short* pool;       // [cap] slot of every pool row (tracked ++ lost)
    unsigned short* unconf;     // [cap]
    unsigned short* sel;        // [cap]
    unsigned short* list_a;     // [cap]
    unsigned short* list_b;     // [cap]
    unsigned short* list_c;     // [cap]
    unsigned long long* cache;  // [kBotTableSlots] open-addressing table: tag (row << 16 | det, top bit set) : float bits
    BlockScratch* bs;
    LapWorkspace lap;
};

MOT_HD constexpr size_t bot_smem_bytes(int cap, int d_max, int e_cap) {
    size_t b = 0;
    b += lap_align16(sizeof(float4) * (size_t)d_max);
    b += lap_align16(sizeof(float) * (size_t)d_max);
    b += 4 * lap_align16(sizeof(unsigned short) * (size_t)d_max);
    b += lap_align16(sizeof(float4) * (size_t)cap);
    b += 6 * lap_align16(sizeof(unsigned short) * (size_t)cap);
    b += lap_align16(sizeof(unsigned long long) * kBotTableSlots);
    b += lap_align16(sizeof(BlockScratch));
    b += lap_smem_bytes(cap, d_max, e_cap);
    return b;
}

__device__ __forceinline__ void bot_carve(unsigned char* p, int cap, int d_max, int e_cap, BotSmem& s) {
    s.det_box = (float4*)p;            p += lap_align16(sizeof(float4) * (size_t)d_max);
    s.det_conf = (float*)p;            p += lap_align16(sizeof(float) * (size_t)d_max);
    s.first = (unsigned short*)p;      p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.second = (unsigned short*)p;     p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.udet = (unsigned short*)p;       p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.udet2 = (unsigned short*)p;      p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.row_box = (float4*)p;            p += lap_align16(sizeof(float4) * (size_t)cap);
    s.pool = (unsigned short*)p;       p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.unconf = (unsigned short*)p;     p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.sel = (unsigned short*)p;        p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.list_a = (unsigned short*)p;     p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.list_b = (unsigned short*)p;     p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.list_c = (unsigned short*)p;     p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.cache = (unsigned long long*)p;  p += lap_align16(sizeof(unsigned long long) * kBotTableSlots);
    s.bs = (BlockScratch*)p;           p += lap_align16(sizeof(BlockScratch));
    lap_carve(p, cap, d_max, e_cap, s.lap);
}

// ---- feature arithmetic.  Summation order ("lanes32", shared with oracle/botsort.cpp lanes_dot): the products of
// float4 number q go, in order x y z w, into partial sum q % 32 (each partial starts at +0 and takes its float4s in
// ascending q); the 32 partials are then combined by the butterfly p[l] += p[l ^ o], o = 16, 8, 4, 2, 1.  A warp
// evaluates it with coalesced float4 loads and five shuffles; one thread can evaluate the same order on its own.
__device__ __forceinline__ float lanes32_reduce(float p) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) p = xadd(p, __shfl_xor_sync(kFullMask, p, o));
    return p;
}
__device__ __forceinline__ float dot4(float acc, float4 a, float4 b) {
    acc = xadd(acc, xmul(a.x, b.x)); acc = xadd(acc, xmul(a.y, b.y));
    acc = xadd(acc, xmul(a.z, b.z)); acc = xadd(acc, xmul(a.w, b.w));
    return acc;
}
// all 32 lanes: prefetch the `dim` floats at v (the next work item of this warp) into L2
__device__ __forceinline__ void warp_prefetch_vec(const float* v, int dim) {
    for (int b = lane_id() * 128; b < dim * 4; b += 32 * 128) prefetch_l2(reinterpret_cast<const char*>(v) + b);
}
// all 32 lanes of a warp; dim % 4 == 0, both pointers 16-byte aligned
__device__ __forceinline__ float warp_dot(const float* __restrict__ x, const float* __restrict__ y, int dim) {
    const float4* xv = reinterpret_cast<const float4*>(x);
    const float4* yv = reinterpret_cast<const float4*>(y);
    const int nq = dim >> 2;
    float acc = 0.0f;
#pragma unroll 4
    for (int q = lane_id(); q < nq; q += 32) acc = dot4(acc, xv[q], yv[q]);
    return lanes32_reduce(acc);
}
// the same value computed by ONE thread (assignment-solver fallback when a pair is not in the table)
static __device__ __noinline__ float thread_dot_lanes32(const float* __restrict__ x, const float* __restrict__ y, int dim) {
    const float4* xv = reinterpret_cast<const float4*>(x);
    const float4* yv = reinterpret_cast<const float4*>(y);
    const int nq = dim >> 2;
    float part[32];
#pragma unroll
    for (int l = 0; l < 32; ++l) part[l] = 0.0f;
    for (int q0 = 0; q0 < nq; q0 += 32) {
#pragma unroll
        for (int l = 0; l < 32; ++l)
            if (q0 + l < nq) part[l] = dot4(part[l], xv[q0 + l], yv[q0 + l]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int l = 0; l < 32; ++l)
            if (l < o) part[l] = xadd(part[l], part[l + o]);
    }
    return part[0];
}

// track box from its CURRENT mean (BotSTrack::xyxy, botsort.cpp:171-181)
__device__ __forceinline__ float4 bot_track_box(const float* rec) {
    return xywh2xyxy(*reinterpret_cast<const float4*>(rec));
}

// Cost functor of the first association and of the unconfirmed-track association (botsort.cpp:438-466, :598-620)
struct BotCost {
    static constexpr bool kWarpPerRow = false;
    static constexpr bool kGrid = true;
    const float4* row_box;
    const unsigned short* row_slot;   // row -> track slot
    const float4* det_box;
    const float* det_conf;
    const unsigned short* col_map;    // column -> detection index
    const float* feats;               // [cap][dim] smooth features
    const float* tnorm;               // [cap]
    const float* dfeat;               // [d_max][dim] normalised detection features
    const float* dnorm;               // [d_max]
    const unsigned char* sflag;       // [cap]
    unsigned long long* cache;
    int dim;
    float prox, app;
    bool fuse, reid, prune;
    struct Row { float4 b; float area; int i; };
    __device__ __forceinline__ Row row(int i) const {
        Row r;
        r.b = row_box[i];
        r.area = box_area(r.b);
        r.i = i;
        return r;
    }
    __device__ __forceinline__ float4 col_box(int j) const { return det_box[col_map[j]]; }
    __device__ __forceinline__ bool reject(const Row& r, int j) const { return prune && boxes_disjoint(r.b, det_box[col_map[j]]); }
    // embedding_distance(...)/2 with the appearance threshold applied (matching.cpp:83-90, botsort.cpp:450-457)
    __device__ __forceinline__ float emb_from_dot(float dot, float tn, float dn) const {
        const float sim = xdiv(dot, xadd(xmul(tn, dn), 1e-10f));
        float e = xdiv(fmaxf(0.0f, xsub(1.0f, sim)), 2.0f);
        if (e > app) e = 1.0f;
        return e;
    }
    __device__ __forceinline__ static unsigned table_tag(int i, int d) { return 0x80000000u | ((unsigned)i << 16) | (unsigned)d; }
    __device__ __forceinline__ static unsigned table_home(unsigned tag) { return (tag * 2654435761u) >> 20; }      // 12 bits
    __device__ __forceinline__ void table_insert(int i, int d, float e) const {
        const unsigned tag = table_tag(i, d);
        const unsigned long long entry = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(e);
        unsigned h = table_home(tag);
        for (int probe = 0; probe < 64; ++probe) {
            const unsigned long long old = atomicCAS(&cache[h], 0ull, entry);
            if (old == 0ull || (unsigned)(old >> 32) == tag) return;
            h = (h + 1) & (kBotTableSlots - 1);
        }
    }
    __device__ __forceinline__ float emb_term(int i, int d) const {
        const unsigned tag = table_tag(i, d);
        unsigned h = table_home(tag);
        for (int probe = 0; probe < 64; ++probe) {
            const unsigned long long hit = cache[h];
            if ((unsigned)(hit >> 32) == tag) return __uint_as_float((unsigned)hit);
            if (hit == 0ull) break;
            h = (h + 1) & (kBotTableSlots - 1);
        }
        // not tabulated (pre-pass buffer overflow, or proximity_thresh >= 1): one thread evaluates the same sums
        const int slot = row_slot[i];
        const bool hf = dim > 0 && (sflag[slot] & kFlagHasFeat) != 0;
        const float dot = hf ? thread_dot_lanes32(feats + (size_t)slot * dim, dfeat + (size_t)d * dim, dim) : 0.0f;
        return emb_from_dot(dot, hf ? tnorm[slot] : 0.0f, dim > 0 ? dnorm[d] : 0.0f);
    }
    __device__ __forceinline__ float cost(const Row& r, int j) const {
        const int d = col_map[j];
        float dist = xsub(1.0f, iou_pair(r.b, r.area, det_box[d]));
        const bool masked = dist > prox;
        if (fuse) dist = xsub(1.0f, xmul(xsub(1.0f, dist), det_conf[d]));
        if (reid) {
            const float e = masked ? 1.0f : emb_term(r.i, d);
            dist = (e < dist) ? e : dist;                                      // cwiseMin (:465)
        }
        return dist;
    }
    __device__ __forceinline__ float pair(int i, int j) const { return cost(row(i), j); }
    __device__ __forceinline__ bool is_candidate(const Row& r, int, int j, float thresh) const { return cost(r, j) <= thresh; }
    __device__ __forceinline__ double pair_bias(int, int) const { return 0.0; }
};

// BotSTrack::update / re_activate (botsort.cpp:112-156) for a list of (track slot, detection) pairs over distinct
// tracks: KalmanFilterXYWH::update, metadata; with_feat additionally runs update_features (:158-169).
template <class SlotOf, class DetOf>
__device__ __forceinline__ void bot_update_pairs(const BotArgs& a, const BotStream& st, BotSmem& sm, const float* dets,
                                                 const float* embs, int n_pairs, int frame, bool with_feat, SlotOf slot_of,
                                                 DetOf det_of) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
    const int lane = lane_id(), g = lane & 7, base = lane & ~7;
    const int groups = nt >> 3, gid = tid >> 3;
    const int rounds = (n_pairs + groups - 1) / groups;
    for (int it = 0; it < rounds; ++it) {
        const int k = it * groups + gid;
        const bool live = k < n_pairs;
        const int slot = live ? slot_of(k) : 0;
        const int det = live ? det_of(k) : 0;
        float* rec = st.recs + (size_t)slot * kRecFloats;
        KfRow s;
        if (live) kf_load_row(rec, g, s);
        else { s.m = 1.0f; for (int j = 0; j < 8; ++j) s.p[j] = (j == g) ? 1.0f : 0.0f; }
        float z[4] = {0.0f, 0.0f, 1.0f, 1.0f};
        if (live) {
            const float* r = dets + (size_t)det * 6;
            const float4 q = xyxy2xywh(make_float4(r[0], r[1], r[2], r[3]));   // BotSTrack ctor (:26-31)
            z[0] = q.x; z[1] = q.y; z[2] = q.z; z[3] = q.w;
        }
        kf_xywh_update(s, g, base, z);
        if (live) {
            kf_store_row(rec, g, s);
            if (g == 0) {
                const int state = (int)(st.sflag[slot] & 0x0f);
                if (state == kStTracked) st.tracklet_len[slot] += 1;     // update (:138)
                else st.tracklet_len[slot] = 0;                          // re_activate (:121)
                st.sflag[slot] = (unsigned char)((st.sflag[slot] & kFlagHasFeat) | kStTracked | kFlagActivated);
                st.frame_id[slot] = frame;
                st.conf[slot] = dets[(size_t)det * 6 + 4];
                st.cls[slot] = (int)dets[(size_t)det * 6 + 5];
                st.det_ind[slot] = det;
            }
        }
    }
    const int dim = a.p.dim;
    if (!with_feat || dim <= 0 || embs == nullptr || n_pairs == 0) return;
    __syncthreads();
    // update_features (:158-169), one warp per track: smooth = alpha * smooth + (1 - alpha) * feat (the RAW detection
    // feature), its norm, the division, and the norm of the result (what embedding_distance recomputes next frame)
    const float alpha = 0.9f, beta = xsub(1.0f, alpha);
    const int nq = dim >> 2, warp = tid >> 5, nwarps = nt >> 5;
    for (int k = warp; k < n_pairs; k += nwarps) {
        const int slot = slot_of(k), det = det_of(k);
        if (k + nwarps < n_pairs) {
            warp_prefetch_vec(st.feats + (size_t)slot_of(k + nwarps) * dim, dim);
            warp_prefetch_vec(embs + (size_t)det_of(k + nwarps) * dim, dim);
        }
        float4* fv = reinterpret_cast<float4*>(st.feats + (size_t)slot * dim);
        const float4* rv = reinterpret_cast<const float4*>(embs + (size_t)det * dim);
        const bool has = (st.sflag[slot] & kFlagHasFeat) != 0;
        float acc = 0.0f, acc2 = 0.0f;
        if (nq <= 128) {
            // up to 512 floats: the whole vector stays in registers (4 float4 per lane), one read and one write
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int q = lane + 32 * j;
                v[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (q < nq) {
                    v[j] = rv[q];
                    if (has) {
                        const float4 o = fv[q];
                        v[j] = make_float4(xadd(xmul(alpha, o.x), xmul(beta, v[j].x)), xadd(xmul(alpha, o.y), xmul(beta, v[j].y)),
                                           xadd(xmul(alpha, o.z), xmul(beta, v[j].z)), xadd(xmul(alpha, o.w), xmul(beta, v[j].w)));
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (lane + 32 * j < nq) acc = dot4(acc, v[j], v[j]);
            const float nrm = xsqrt(lanes32_reduce(acc));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int q = lane + 32 * j;
                if (q < nq) {
                    if (nrm > 0.0f) v[j] = make_float4(xdiv(v[j].x, nrm), xdiv(v[j].y, nrm), xdiv(v[j].z, nrm), xdiv(v[j].w, nrm));
                    fv[q] = v[j];
                    acc2 = dot4(acc2, v[j], v[j]);
                }
            }
        } else {
#pragma unroll 4
            for (int q = lane; q < nq; q += 32) {
                float4 v = rv[q];
                if (has) {
                    const float4 o = fv[q];
                    v = make_float4(xadd(xmul(alpha, o.x), xmul(beta, v.x)), xadd(xmul(alpha, o.y), xmul(beta, v.y)),
                                    xadd(xmul(alpha, o.z), xmul(beta, v.z)), xadd(xmul(alpha, o.w), xmul(beta, v.w)));
                }
                fv[q] = v;
                acc = dot4(acc, v, v);
            }
            const float nrm = xsqrt(lanes32_reduce(acc));
#pragma unroll 4
            for (int q = lane; q < nq; q += 32) {           // every lane re-reads exactly the float4s it wrote
                float4 v = fv[q];
                if (nrm > 0.0f) { v = make_float4(xdiv(v.x, nrm), xdiv(v.y, nrm), xdiv(v.z, nrm), xdiv(v.w, nrm)); fv[q] = v; }
                acc2 = dot4(acc2, v, v);
            }
        }
        const float tn = xsqrt(lanes32_reduce(acc2));
        if (lane == 0) { st.tnorm[slot] = tn; st.sflag[slot] |= kFlagHasFeat; }
    }
    __syncthreads();
}

// Tabulates the embedding term of every (row, column) pair whose IoU distance passes the proximity gate, one warp per
// pair (coalesced feature reads), before the assignment consults them through BotCost::emb_term.
template <class Cost>
__device__ __forceinline__ void bot_emb_prepass(const BotStream& st, BotSmem& sm, const Cost& cost, int n, int m) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    for (int h = tid; h < kBotTableSlots; h += nt) sm.cache[h] = 0ull;
    if (n == 0 || m == 0 || !cost.reid || cost.dim <= 0 || !cost.prune || m > sm.lap.grid.cap) { __syncthreads(); return; }
    grid_build(sm.lap.grid, m, sm.bs, [&](int j) { return cost.col_box(j); });
    int inserted = 0;
    for (int base = 0; base < n; base += nt) {
        const int i = base + tid;
        if (tid == 0) sm.lap.ctl[7] = 0;
        __syncthreads();
        if (i < n) {
            const typename Cost::Row rw = cost.row(i);
            grid_query(sm.lap.grid, rw.b, [&](int j) { return cost.col_box(j); }, [&](int j, float4 b) {
                const float dist = xsub(1.0f, iou_pair(rw.b, rw.area, b));
                if (!(dist > cost.prox)) {
                    const int q = atomicAdd(&sm.lap.ctl[7], 1);
                    if (q < sm.lap.p_cap) sm.lap.pairs[q] = (i << 16) | j;
                }
            });
        }
        __syncthreads();
        const int n_pairs = min(sm.lap.ctl[7], sm.lap.p_cap);
        if (inserted + n_pairs <= (kBotTableSlots * 3) / 4) {
            for (int q = warp; q < n_pairs; q += nwarps) {
                const int pk = sm.lap.pairs[q];
                const int pi = pk >> 16, d = cost.col_map[pk & 0xffff];
                const int slot = cost.row_slot[pi];
                if (q + nwarps < n_pairs) {
                    const int nk = sm.lap.pairs[q + nwarps];
                    warp_prefetch_vec(st.feats + (size_t)cost.row_slot[nk >> 16] * cost.dim, cost.dim);
                    warp_prefetch_vec(st.dfeat + (size_t)cost.col_map[nk & 0xffff] * cost.dim, cost.dim);
                }
                const bool hf = (st.sflag[slot] & kFlagHasFeat) != 0;
                const float dot = hf ? warp_dot(st.feats + (size_t)slot * cost.dim, st.dfeat + (size_t)d * cost.dim, cost.dim) : 0.0f;
                if (lane == 0) cost.table_insert(pi, d, cost.emb_from_dot(dot, hf ? st.tnorm[slot] : 0.0f, cost.dnorm[d]));
            }
            inserted += n_pairs;
        }
        __syncthreads();
    }
}

template <int CAP, int DMAX>
__device__ __forceinline__ void bot_frame(const BotArgs& a, const BotStream& st, BotSmem& sm, const float* dets,
                                          const float* embs, int n_det_in, float* out, int* n_out) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
    const int lane = tid & 31, g = lane & 7, base = lane & ~7;
    const int groups = nt >> 3, gid = tid >> 3;
    const int dim = a.p.dim;
    const bool reid = a.p.with_reid != 0;
    __syncthreads();
    if (n_det_in <= 0) {                                   // botsort.cpp:267-269: nothing happens at all
        if (tid == 0) *n_out = 0;
        return;
    }
    const int frame = st.hdr[kHdrFrame] + 1;
    const int n_active = st.hdr[kHdrActive], n_lost = st.hdr[kHdrLost];
    int n_free = st.hdr[kHdrFree];
    const int id_base = st.hdr[kHdrIdCounter];
    int n_det = n_det_in;
    if (n_det > min(DMAX, a.ld_dets)) { n_det = min(DMAX, a.ld_dets); if (tid == 0) atomicOr(&st.hdr[kHdrError], (int)kErrTooManyDets); }

    // ---- A. detections: IoU boxes, confidence split, feature normalisation
    for (int j = tid; j < n_det; j += nt) {
        const float* r = dets + (size_t)j * 6;
        sm.det_box[j] = xywh2xyxy(xyxy2xywh(make_float4(r[0], r[1], r[2], r[3])));
        sm.det_conf[j] = r[4];
    }
    __syncthreads();
    const float t_hi = a.p.track_high_thresh, t_lo = a.p.track_low_thresh;
    const int n_first = block_compact(n_det, 0, sm.bs, [&](int j) { return sm.det_conf[j] > t_hi; },
                                      [&](int j, int pos) { sm.first[pos] = (unsigned short)j; });
    const int n_second = block_compact(n_det, 0, sm.bs,
                                       [&](int j) { const float c = sm.det_conf[j]; return !(c > t_hi) && c > t_lo; },
                                       [&](int j, int pos) { sm.second[pos] = (unsigned short)j; });
    const bool have_feat = reid && dim > 0 && embs != nullptr;
    if (have_feat) {
        // BotSTrack(det, feat): smooth_feat = feat / |feat| (:39-46), and the norm of THAT vector, which
        // embedding_distance recomputes for every pair (matching.cpp:86-87); one warp per detection
        const int nq = dim >> 2, warp = tid >> 5, nwarps = nt >> 5;
        for (int k = warp; k < n_first; k += nwarps) {
            const int d = sm.first[k];
            const float* raw = embs + (size_t)d * dim;
            if (k + nwarps < n_first) warp_prefetch_vec(embs + (size_t)sm.first[k + nwarps] * dim, dim);
            const float4* rv = reinterpret_cast<const float4*>(raw);
            float4* ov = reinterpret_cast<float4*>(st.dfeat + (size_t)d * dim);
            float acc = 0.0f;
            if (nq <= 128) {
                float4 v[4];
                float a0 = 0.0f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int q = lane + 32 * j;
                    v[j] = (q < nq) ? rv[q] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (lane + 32 * j < nq) a0 = dot4(a0, v[j], v[j]);
                const float nrm = xsqrt(lanes32_reduce(a0));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int q = lane + 32 * j;
                    if (q < nq) {
                        if (nrm > 0.0f) v[j] = make_float4(xdiv(v[j].x, nrm), xdiv(v[j].y, nrm), xdiv(v[j].z, nrm), xdiv(v[j].w, nrm));
                        ov[q] = v[j];
                        acc = dot4(acc, v[j], v[j]);
                    }
                }
            } else {
                const float nrm = xsqrt(warp_dot(raw, raw, dim));
#pragma unroll 4
                for (int q = lane; q < nq; q += 32) {
                    float4 v = rv[q];
                    if (nrm > 0.0f) v = make_float4(xdiv(v.x, nrm), xdiv(v.y, nrm), xdiv(v.z, nrm), xdiv(v.w, nrm));
                    ov[q] = v;
                    acc = dot4(acc, v, v);
                }
            }
            const float n2 = xsqrt(lanes32_reduce(acc));
            if (lane == 0) st.dnorm[DMAX + d] = n2;
        }
    }

    // ---- B. pool = tracked (activated) ++ lost ; unconfirmed kept aside (:293-309)
    const int n_trk = block_compact(n_active, 0, sm.bs,
                                    [&](int k) { return (st.sflag[st.active[k]] & kFlagActivated) != 0; },
                                    [&](int k, int pos) { sm.pool[pos] = st.active[k]; });
    const int n_unc = block_compact(n_active, 0, sm.bs,
                                    [&](int k) { return (st.sflag[st.active[k]] & kFlagActivated) == 0; },
                                    [&](int k, int pos) { sm.unconf[pos] = st.active[k]; });
    for (int k = tid; k < n_lost; k += nt) sm.pool[n_trk + k] = st.lost[k];
    const int n1 = n_trk + n_lost;
    __syncthreads();

    // ---- C. predict every pool track IN PLACE (:311); unconfirmed tracks are not predicted
    {
        const int rounds = (n1 + groups - 1) / groups;
        for (int it = 0; it < rounds; ++it) {
            const int r = it * groups + gid;
            const bool live = r < n1;
            const int slot = live ? (int)sm.pool[r] : 0;
            float* rec = st.recs + (size_t)slot * kRecFloats;
            KfRow s;
            if (live) kf_load_row(rec, g, s);
            else { s.m = 1.0f; for (int j = 0; j < 8; ++j) s.p[j] = (j == g) ? 1.0f : 0.0f; }
            kf_xywh_predict(s, g, base);
            if (live) kf_store_row(rec, g, s);
        }
    }
    __syncthreads();
    for (int r = tid; r < n1; r += nt) {
        const int slot = sm.pool[r];
        sm.row_box[r] = bot_track_box(st.recs + (size_t)slot * kRecFloats);
    }
    __syncthreads();       // (a track's |smooth_feat| is kept in tnorm[] from the moment the feature was last written)

    // ---- D. first association (:402-495)
    const int fdim = have_feat ? dim : 0;
    {
        BotCost cost{sm.row_box, sm.pool, sm.det_box, sm.det_conf, sm.first, st.feats, st.tnorm, st.dfeat, st.dnorm + DMAX,
                     st.sflag, sm.cache, fdim, a.p.proximity_thresh, a.p.appearance_thresh, a.p.fuse_first != 0, reid,
                     a.p.match_thresh < 1.0f && (!reid || a.p.proximity_thresh < 1.0f)};
        bot_emb_prepass(st, sm, cost, n1, n_first);
        block_lap(sm.lap, n1, n_first, CAP, DMAX, a.p.match_thresh, cost);
    }
    const int n_m1 = block_compact(n1, 0, sm.bs, [&](int r) { return sm.lap.row2col[r] >= 0; },
                                   [&](int r, int pos) { sm.sel[pos] = (unsigned short)r; });
    for (int k = tid; k < n_m1; k += nt) sm.list_c[k] = sm.first[sm.lap.row2col[sm.sel[k]]];
    const int n_udet = block_compact(n_first, 0, sm.bs, [&](int j) { return sm.lap.col2row[j] < 0; },
                                     [&](int j, int pos) { sm.udet[pos] = sm.first[j]; });
    // r_tracked: unmatched pool rows in state Tracked, i.e. the ones that came from the active list (:512-518)
    const int n2 = block_compact(n_trk, 0, sm.bs, [&](int r) { return sm.lap.row2col[r] < 0; },
                                 [&](int r, int pos) { sm.list_a[pos] = sm.pool[r]; sm.list_b[pos] = (unsigned short)r; });
    __syncthreads();

    // ---- E. update / re_activate the matched tracks
    bot_update_pairs(a, st, sm, dets, embs, n_m1, frame, have_feat, [&](int k) { return (int)sm.pool[sm.sel[k]]; },
                     [&](int k) { return (int)sm.list_c[k]; });
    __syncthreads();

    // ---- F. second association: r_tracked x low-confidence detections on the predicted boxes (:497-562)
    int n_lost_new = 0;
    if (n2 > 0 && n_second > 0) {
        // gather the r_tracked boxes to the front of row_box: list_b is increasing with list_b[i] >= i, so a chunk's
        // destinations are never the sources of a later chunk
        for (int cb = 0; cb < n2; cb += nt) {
            const int i = cb + tid;
            float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (i < n2) v = sm.row_box[sm.list_b[i]];
            __syncthreads();
            if (i < n2) sm.row_box[i] = v;
            __syncthreads();
        }
        IouCost cost{sm.row_box, sm.det_box, sm.det_conf, sm.second, false, true};
        block_lap(sm.lap, n2, n_second, CAP, DMAX, 0.5f, cost);
        const int n_m2 = block_compact(n2, 0, sm.bs, [&](int i) { return sm.lap.row2col[i] >= 0; },
                                       [&](int i, int pos) { sm.sel[pos] = (unsigned short)i; });
        for (int k = tid; k < n_m2; k += nt) sm.list_c[k] = sm.second[sm.lap.row2col[sm.sel[k]]];
        __syncthreads();
        bot_update_pairs(a, st, sm, dets, embs, n_m2, frame, false, [&](int k) { return (int)sm.list_a[sm.sel[k]]; },
                         [&](int k) { return (int)sm.list_c[k]; });
        n_lost_new = block_compact(n2, 0, sm.bs, [&](int i) { return sm.lap.row2col[i] < 0; },
                                   [&](int i, int pos) {
                                       const int slot = sm.list_a[i];
                                       sm.list_c[pos] = (unsigned short)slot;
                                       st.sflag[slot] = (unsigned char)((st.sflag[slot] & 0xf0) | kStLost);
                                   });
    }
    __syncthreads();
    // list_c[0 .. n_lost_new) holds the slots that became Lost this frame; keep it until phase J.

    // ---- G. unconfirmed tracks x leftover first-stage detections (:564-647)
    int n_final = n_udet;
    const unsigned short* final_list = sm.udet;
    if (n_unc > 0 && n_udet > 0) {
        for (int i = tid; i < n_unc; i += nt) sm.row_box[i] = bot_track_box(st.recs + (size_t)sm.unconf[i] * kRecFloats);
        __syncthreads();
        BotCost cost{sm.row_box, sm.unconf, sm.det_box, sm.det_conf, sm.udet, st.feats, st.tnorm, st.dfeat, st.dnorm + DMAX,
                     st.sflag, sm.cache, fdim, a.p.proximity_thresh, a.p.appearance_thresh, true, reid,
                     !reid || a.p.proximity_thresh < 1.0f};
        bot_emb_prepass(st, sm, cost, n_unc, n_udet);
        block_lap(sm.lap, n_unc, n_udet, CAP, DMAX, 0.7f, cost);
        const int n_m3 = block_compact(n_unc, 0, sm.bs, [&](int i) { return sm.lap.row2col[i] >= 0; },
                                       [&](int i, int pos) { sm.sel[pos] = (unsigned short)i; });
        for (int k = tid; k < n_m3; k += nt) sm.list_a[k] = sm.udet[sm.lap.row2col[sm.sel[k]]];
        n_final = block_compact(n_udet, 0, sm.bs, [&](int j) { return sm.lap.col2row[j] < 0; },
                                [&](int j, int pos) { sm.udet2[pos] = sm.udet[j]; });
        final_list = sm.udet2;
        for (int i = tid; i < n_unc; i += nt)
            if (sm.lap.row2col[i] < 0) st.sflag[sm.unconf[i]] = (unsigned char)kStRemoved;      // :637-641
        __syncthreads();
        bot_update_pairs(a, st, sm, dets, embs, n_m3, frame, have_feat, [&](int k) { return (int)sm.unconf[sm.sel[k]]; },
                         [&](int k) { return (int)sm.list_a[k]; });
    }
    __syncthreads();

    // ---- H. new tracks (:649-667): conf >= new_track_thresh, ids in list order
    const float new_thresh = a.p.new_track_thresh;
    const int n_new_want = block_compact(n_final, 0, sm.bs, [&](int k) { return !(sm.det_conf[final_list[k]] < new_thresh); },
                                         [&](int k, int pos) { sm.sel[pos] = final_list[k]; });
    int n_new = n_new_want;
    if (n_new > n_free) { n_new = n_free; if (tid == 0) atomicOr(&st.hdr[kHdrError], (int)kErrCapacity); }
    for (int k = gid; k < n_new; k += groups) {
        const int det = sm.sel[k];
        const int slot = st.freel[n_free - 1 - k];
        const float* r = dets + (size_t)det * 6;
        const float4 q = xyxy2xywh(make_float4(r[0], r[1], r[2], r[3]));
        const float z[4] = {q.x, q.y, q.z, q.w};
        KfRow s;
        kf_xywh_initiate(s, g, z);
        kf_store_row(st.recs + (size_t)slot * kRecFloats, g, s);
        if (g == 0) {
            st.id[slot] = id_base + 1 + k;
            st.sflag[slot] = (unsigned char)(kStTracked | (frame == 1 ? kFlagActivated : 0) | (have_feat ? kFlagHasFeat : 0));
            st.tracklet_len[slot] = 0;
            st.frame_id[slot] = frame;
            st.start_frame[slot] = frame;
            st.conf[slot] = r[4];
            st.cls[slot] = (int)r[5];
            st.det_ind[slot] = det;
        }
    }
    if (have_feat) {
        const int nq = dim >> 2, warp = tid >> 5, nwarps = nt >> 5;
        for (int k = warp; k < n_new; k += nwarps) {       // smooth_feat of a new track = the detection's normalised feature
            const int det = sm.sel[k], slot = st.freel[n_free - 1 - k];
            const float4* src = reinterpret_cast<const float4*>(st.dfeat + (size_t)det * dim);
            float4* dst = reinterpret_cast<float4*>(st.feats + (size_t)slot * dim);
            for (int q = lane; q < nq; q += 32) dst[q] = src[q];
            if (lane == 0) st.tnorm[slot] = st.dnorm[DMAX + det];
        }
    }
    __syncthreads();

    // ---- I. expire lost tracks (:669-676); re-found ones are Tracked by now and skipped
    for (int k = tid; k < n_lost; k += nt) {
        const int slot = st.lost[k];
        if ((st.sflag[slot] & 0x0f) == kStLost && frame - st.frame_id[slot] > a.p.max_time_lost)
            st.sflag[slot] = (unsigned char)((st.sflag[slot] & 0xf0) | kStRemoved);
    }
    __syncthreads();

    // ---- J. prepare_output (:678-744): active' = Tracked actives ++ new ; lost' = still-Lost ++ lost-this-frame ;
    //         a re-found lost track (Tracked, but in the lost list) is in neither: it vanishes
    int na = block_compact(n_active, 0, sm.bs, [&](int k) { return (st.sflag[st.active[k]] & 0x0f) == kStTracked; },
                           [&](int k, int pos) { sm.list_a[pos] = st.active[k]; });
    for (int k = tid; k < n_new; k += nt) sm.list_a[na + k] = st.freel[n_free - 1 - k];
    na += n_new;
    int nl = block_compact(n_lost, 0, sm.bs, [&](int k) { return (st.sflag[st.lost[k]] & 0x0f) == kStLost; },
                           [&](int k, int pos) { sm.list_b[pos] = st.lost[k]; });
    for (int k = tid; k < n_lost_new; k += nt) sm.list_b[nl + k] = sm.list_c[k];
    nl += n_lost_new;
    __syncthreads();
    n_free -= n_new;
    n_free = block_compact(n_active, n_free, sm.bs, [&](int k) { return (st.sflag[st.active[k]] & 0x0f) == kStRemoved; },
                           [&](int k, int pos) { st.freel[pos] = st.active[k]; });
    n_free = block_compact(n_lost, n_free, sm.bs, [&](int k) { return (st.sflag[st.lost[k]] & 0x0f) != kStLost; },
                           [&](int k, int pos) { st.freel[pos] = st.lost[k]; });
    for (int k = tid; k < na; k += nt) st.active[k] = sm.list_a[k];
    for (int k = tid; k < nl; k += nt) st.lost[k] = sm.list_b[k];
    __syncthreads();

    // ---- K. output rows: activated tracks of the new active list (:746-764)
    const int n_rows = block_compact(na, 0, sm.bs, [&](int k) { return (st.sflag[sm.list_a[k]] & kFlagActivated) != 0; },
                                     [&](int k, int pos) {
                                         if (pos >= a.ld_out) return;
                                         const int slot = sm.list_a[k];
                                         const float4 b = bot_track_box(st.recs + (size_t)slot * kRecFloats);
                                         float* w = out + (size_t)pos * 8;
                                         *reinterpret_cast<float4*>(w) = b;
                                         *reinterpret_cast<float4*>(w + 4) = make_float4((float)st.id[slot], st.conf[slot],
                                                                                         (float)st.cls[slot], (float)st.det_ind[slot]);
                                     });
    if (tid == 0) {
        if (n_rows > a.ld_out) atomicOr(&st.hdr[kHdrError], (int)kErrOutput);
        *n_out = n_rows < a.ld_out ? n_rows : a.ld_out;
        st.hdr[kHdrActive] = na;
        st.hdr[kHdrLost] = nl;
        st.hdr[kHdrFree] = n_free;
        st.hdr[kHdrIdCounter] = id_base + n_new;
        st.hdr[kHdrFrame] = frame;
        st.hdr[kHdrN1] = n1; st.hdr[kHdrM1] = n_first;
        st.hdr[kHdrN2] = (n2 > 0 && n_second > 0) ? n2 : 0; st.hdr[kHdrM2] = (n2 > 0 && n_second > 0) ? n_second : 0;
        st.hdr[kHdrN3] = (n_unc > 0 && n_udet > 0) ? n_unc : 0; st.hdr[kHdrM3] = (n_unc > 0 && n_udet > 0) ? n_udet : 0;
        st.hdr[kBHdrNew] = n_new; st.hdr[kBHdrLostAfter] = nl;
    }
    __syncthreads();
}

template <int CAP, int DMAX, int ECAP>
__global__ void __launch_bounds__(kBotThreads) botsort_step_kernel(BotArgs a) {
    MOT_DYNAMIC_SMEM(smem);
    BotSmem sm;
    bot_carve(smem, CAP, DMAX, ECAP, sm);
    for (int s = a.s_begin + (int)blockIdx.x; s < a.s_end; s += (int)gridDim.x) {
        BotStream st = BotStream::at(a.state + (size_t)s * a.L.stride, a.L);
        lap_carve_gscratch(st.gscratch, CAP, DMAX, sm.lap);
        for (int t = 0; t < a.T; ++t) {
            const size_t fs = (size_t)t * a.S + s;
            bot_frame<CAP, DMAX>(a, st, sm, a.dets + fs * (size_t)a.ld_dets * 6,
                                 a.embs ? a.embs + fs * (size_t)a.ld_dets * a.p.dim : nullptr, a.n_dets[fs],
                                 a.out + fs * (size_t)a.ld_out * 8, a.n_out + fs);
        }
    }
}

// BotSort ctor / reset(): everything cleared, ids restart at 0 (botsort.cpp:249,257)
static __global__ void botsort_reset_kernel(unsigned char* state, BotLayout L, int S) {
    for (int s = (int)blockIdx.x; s < S; s += (int)gridDim.x) {
        BotStream st = BotStream::at(state + (size_t)s * L.stride, L);
        for (int k = (int)threadIdx.x; k < L.cap; k += (int)blockDim.x) {
            st.freel[k] = (unsigned short)(L.cap - 1 - k);
            st.sflag[k] = (unsigned char)kStRemoved;
        }
        if (threadIdx.x == 0) {
            for (int k = 0; k < kHdrInts; ++k) st.hdr[k] = 0;
            st.hdr[kHdrFree] = L.cap;
        }
        __syncthreads();
    }
}

// shapes the BoT-SORT kernel is built for (track capacity, detections per frame, candidate-edge buffer)
constexpr BtShape kBotShapes[] = {{256, 64, 1024}, {1536, 512, 4096}, {2048, 1024, 4096}};
constexpr int kNumBotShapes = sizeof(kBotShapes) / sizeof(kBotShapes[0]);

}  // namespace mot
