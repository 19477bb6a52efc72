// ocsort_kernel.cuh - OC-SORT's whole per-frame update() as one kernel, one CTA per camera stream.
// Replaces reference src/trackers/ocsort.cpp:285-606 (OCSort::update), :610-737 (associate) and
// :53-156 (KalmanBoxTracker):
//   confidence split (:311-320)                                         -> phase A
//   predict every track in place, drop NaN boxes (:337-365)             -> phase B
//   no tracks: spawn from every high detection, empty output (:367-384) -> early exit
//   velocities and k_previous_obs of every track (:394-410)             -> phase C
//   first association: OCM cost, trivial 1:1 shortcut or assignment, IoU filter (:413-420, :610-737) -> phase D
//   update matched tracks (:423-430)                                    -> phase E
//   BYTE pass on low-confidence detections (:433-479, use_byte)         -> phase F
//   re-match leftovers against LAST OBSERVATIONS (:482-545)             -> phase G
//   update(None) (:548-550), new tracks (:553-561)                      -> phases H, I
//   output in reverse track order, age-out (:564-592)                   -> phase J
// Reference quirks kept on purpose (parity trap 8): an assignment pair rejected by the
// IoU filter is pushed to BOTH unmatched lists and then added again by the final sweep, so those
// lists can hold an index twice; the re-match then sees duplicated rows / columns, a track can be
// updated twice in one frame and a detection can spawn two tracks.  The lists here keep the
// duplicates, in the reference's order.
// State: [x 7 | P 7x7] fp32 padded to 64 floats, per-track observation ring (the reference's unbounded
// age -> bbox map is only ever queried for the last delta_t ages, else for its newest entry, which is
// last_observation), IDs from a per-stream counter.
#pragma once
#include "shapes.cuh"
#include "block_utils.cuh"
#include "cost_device.cuh"
#include "kf_device.cuh"
#include "lap_device.cuh"
#include "ocm_device.cuh"
#include "jv_device.cuh"
#include "jv_block_device.cuh"
#include <type_traits>

namespace mot {

#ifndef MOT_OC_THREADS
#define MOT_OC_THREADS 512
#endif
constexpr int kOcThreads = MOT_OC_THREADS;
constexpr int kOcRecFloats = 64;       // 56 used
constexpr int kOcRing = 8;             // observation ring entries per track: delta_t <= kOcRing - 1 (a track updated twice in one
                                       // frame still needs age - delta_t after its first update has stored age)
constexpr int kOcObsFloats = 8;        // last_observation[5], velocity (dy, dx), pad
// Exact reference tie-breaking: when a "twin" track (see the file header) is an assignment candidate, the frame's
// assignment is redone by the reference's own dense LAPJV on the full cost matrix - by one warp out of shared memory
// while rows + columns <= kJvMax (jv_device.cuh), by the whole CTA over per-stream global scratch above that
// (jv_block_device.cuh).  Either way the result IS the reference's, at any problem size.
#ifndef MOT_OC_JVMAX
#define MOT_OC_JVMAX 384
#endif
constexpr int kJvMax = MOT_OC_JVMAX;
constexpr unsigned short kNoTwin = 0xffff;

enum : int {
    kOHdrTracks = 0, kOHdrFree = 2, kOHdrIdCounter = 3, kOHdrFrame = 4, kOHdrError = 5,
    kOHdrNHigh = 6, kOHdrNTrk = 7, kOHdrUsedLap = 8, kOHdrMatched = 9, kOHdrLeftDets = 10, kOHdrLeftTrks = 11,
    kOHdrRematched = 12, kOHdrSpawned = 13, kOHdrExactSolves = 14      // cumulative count of dense-LAPJV re-solves
};

struct OcParams {
    float det_thresh, iou_threshold, min_conf, inertia;
    float q44, q66;                    // fl(0.01f * Q_xy_scaling), fl(0.0001f * Q_s_scaling)
    int max_age, min_hits, delta_t, use_byte;
    // DeepOC-SORT only (deepocsort.hpp:96-121)
    float w_assoc_emb, alpha_fixed_emb, aw_param;
    int aw_off, embedding_off;
    // asso_func (ocsort.hpp:93; OC-SORT only): 0 iou, 6 centroid with the frame diagonal the reference derives from img
    int asso;
    float asso_norm;
};

struct OcLayout {
    int cap, d_max;
    size_t off_lists, off_meta, off_obs, off_ring_box, off_ring_conf, off_ring_age, off_recs, off_ocm, off_valid,
        off_twin, off_jv, off_jvw, off_gscratch, stride;
    MOT_HD static constexpr size_t al(size_t x) { return (x + 255) & ~(size_t)255; }
    MOT_HD static constexpr OcLayout make(int cap, int d_max) {
        OcLayout L{};
        L.cap = cap; L.d_max = d_max;
        size_t o = al(sizeof(int) * 16);
        L.off_lists = o;     o = al(o + sizeof(unsigned short) * 2 * (size_t)cap);
        L.off_meta = o;      o = al(o + sizeof(int) * 8 * (size_t)cap);
        L.off_obs = o;       o = al(o + sizeof(float) * kOcObsFloats * (size_t)cap);
        L.off_ring_box = o;  o = al(o + sizeof(float4) * kOcRing * (size_t)cap);
        L.off_ring_conf = o; o = al(o + sizeof(float) * kOcRing * (size_t)cap);
        L.off_ring_age = o;  o = al(o + sizeof(int) * kOcRing * (size_t)cap);
        L.off_recs = o;      o = al(o + sizeof(float) * kOcRecFloats * (size_t)cap);
        L.off_ocm = o;       o = al(o + sizeof(float4) * (size_t)cap);
        L.off_valid = o;     o = al(o + (size_t)cap);
        L.off_twin = o;      o = al(o + sizeof(unsigned short) * (size_t)cap);
        L.off_jv = o;        o = al(o + sizeof(float) * (size_t)d_max * (size_t)cap);          // dense cost matrix of the exact-tie path
        L.off_jvw = o;       o = al(o + jv_block_gbytes(d_max + cap));                        // its LAPJV work arrays
        L.off_gscratch = o;  o = al(o + lap_gscratch_bytes(d_max, cap));
        L.stride = o;
        return L;
    }
};

struct OcStream {
    int* hdr;
    unsigned short *list, *freel;
    int *id, *age, *hits, *streak, *tsu, *cls, *det_ind;
    float* conf;
    float* obs;                // [cap][8]
    float4* ring_box;          // [cap][kOcRing]
    float* ring_conf;
    int* ring_age;
    float* recs;
    float4* ocm;               // per-frame scratch, indexed by track POSITION
    unsigned char* valid;
    unsigned short* twin;      // [cap] slot of the bit-identical twin spawned from the same detection, kNoTwin = none
    float* jv_dense;           // [d_max * cap] dense cost matrix of the exact-tie path
    unsigned char* jv_work;    // jv_block_gbytes(d_max + cap): work arrays of the CTA-wide LAPJV
    unsigned char* gscratch;
    __device__ __forceinline__ static OcStream at(unsigned char* base, const OcLayout& L) {
        OcStream s;
        s.hdr = (int*)base;
        s.list = (unsigned short*)(base + L.off_lists);
        s.freel = s.list + L.cap;
        int* m = (int*)(base + L.off_meta);
        s.id = m; s.age = m + L.cap; s.hits = m + 2 * L.cap; s.streak = m + 3 * L.cap; s.tsu = m + 4 * L.cap;
        s.cls = m + 5 * L.cap; s.det_ind = m + 6 * L.cap; s.conf = (float*)(m + 7 * L.cap);
        s.obs = (float*)(base + L.off_obs);
        s.ring_box = (float4*)(base + L.off_ring_box);
        s.ring_conf = (float*)(base + L.off_ring_conf);
        s.ring_age = (int*)(base + L.off_ring_age);
        s.recs = (float*)(base + L.off_recs);
        s.ocm = (float4*)(base + L.off_ocm);
        s.valid = base + L.off_valid;
        s.twin = (unsigned short*)(base + L.off_twin);
        s.jv_dense = (float*)(base + L.off_jv);
        s.jv_work = base + L.off_jvw;
        s.gscratch = base + L.off_gscratch;
        return s;
    }
};

struct OcArgs {
    unsigned char* state;
    const float* dets;        // [T][S][ld_dets][6]
    const int* n_dets;        // [T][S]
    float* out;               // [T][S][ld_out][8]
    int* n_out;               // [T][S]
    int T, S, ld_dets, ld_out;
    int s_begin, s_end;
    OcParams p;
    // DeepOC-SORT only: detection embeddings [T][S][ld_dets][dim] and the per-stream slab stride (OcLayout::stride +
    // DeepLayout::bytes; the embedding dimension is a run-time size)
    const float* embs;
    int dim;
    size_t stride;
};

// DeepOC-SORT's appearance state, appended to every stream's OC-SORT slab (at OcLayout::stride):
//   trk_emb  [cap][dim]      the tracks' unit-length embeddings, by slot (deepocsort.cpp:75-80, :143-161)
//   dense    [d_max][cap]    this frame's detection x track embedding products, written and read ONLY at the pairs whose
//                            boxes overlap - everywhere else the reference's masked matrix is exactly 0 (:421-423)
//   row_w    [d_max]         w_assoc_emb x adaptive row weight (0 when the row maximum is 0)        (:294-345)
//   col_top  [cap], row_top [d_max]   packed (largest, second largest) stored entry, order-preserving encoding
//   col_nnz  [cap], row_nnz [d_max]   number of stored entries per column / row
//   col_w    [cap], col_z [cap]   adaptive column weight / "column maximum is 0" flag
struct DeepLayout {
    int dim;
    size_t off_emb, off_dense, off_roww, off_coltop, off_colnnz, off_colw, off_colz, off_rowtop, off_rownnz, bytes;
    MOT_HD static DeepLayout make(int cap, int d_max, int dim) {
        DeepLayout D{};
        D.dim = dim;
        size_t o = 0;
        D.off_emb = o;     o = OcLayout::al(o + sizeof(float) * (size_t)cap * (size_t)(dim > 0 ? dim : 1));
        D.off_dense = o;   o = OcLayout::al(o + sizeof(float) * (size_t)d_max * (size_t)cap);
        D.off_roww = o;    o = OcLayout::al(o + sizeof(float) * (size_t)d_max);
        D.off_coltop = o;  o = OcLayout::al(o + sizeof(unsigned long long) * (size_t)cap);
        D.off_colnnz = o;  o = OcLayout::al(o + sizeof(int) * (size_t)cap);
        D.off_colw = o;    o = OcLayout::al(o + sizeof(float) * (size_t)cap);
        D.off_colz = o;    o = OcLayout::al(o + (size_t)cap);
        D.off_rowtop = o;  o = OcLayout::al(o + sizeof(unsigned long long) * (size_t)d_max);
        D.off_rownnz = o;  o = OcLayout::al(o + sizeof(int) * (size_t)d_max);
        D.bytes = o;
        return D;
    }
};

struct DeepStream {
    float* trk_emb;
    float* dense;
    float* row_w;
    unsigned long long* col_top;
    int* col_nnz;
    float* col_w;
    unsigned char* col_z;
    unsigned long long* row_top;
    int* row_nnz;
    int dim;
    __device__ __forceinline__ static DeepStream at(unsigned char* base, int cap, int d_max, int dim) {
        const DeepLayout D = DeepLayout::make(cap, d_max, dim);
        DeepStream s;
        s.trk_emb = (float*)(base + D.off_emb);
        s.dense = (float*)(base + D.off_dense);
        s.row_w = (float*)(base + D.off_roww);
        s.col_top = (unsigned long long*)(base + D.off_coltop);
        s.col_nnz = (int*)(base + D.off_colnnz);
        s.col_w = (float*)(base + D.off_colw);
        s.col_z = base + D.off_colz;
        s.row_top = (unsigned long long*)(base + D.off_rowtop);
        s.row_nnz = (int*)(base + D.off_rownnz);
        s.dim = dim;
        return s;
    }
};

struct OcSmem {
    float4* det_box;            // [d_max] raw xyxy
    float* det_conf;            // [d_max]
    unsigned short* high;       // [d_max] conf > det_thresh
    unsigned short* second;     // [d_max] min_conf < conf < det_thresh (use_byte)
    unsigned short* ud;         // [d_max] unmatched detections (detection indices, duplicates kept)
    unsigned short* ud2;        // [d_max]
    unsigned short* pair_det;   // [d_max] detection of every update pair
    unsigned char* det_flag;    // [d_max] per detection / per row: 0 none, 1 kept match, 2 filtered; later "gone"
    float4* trk_box;            // [cap] predicted boxes, later last observations
    unsigned short* list_a;     // [cap] live track slots in order
    unsigned short* ut;         // [cap] unmatched tracks (track POSITIONS, duplicates kept)
    unsigned short* ut2;        // [cap]
    unsigned short* pair_trk;   // [cap] track position of every update pair
    unsigned char* trk_flag;    // [cap]
    unsigned* row_bits;         // [d_max / 32]
    unsigned* col_bits;         // [cap / 32]
    int* flags;                 // [4]
    unsigned char* jv;          // working arrays of the dense LAPJV (aliases lap.scratch_a when that is big enough)
    BlockScratch* bs;
    LapWorkspace lap;
};

MOT_HD constexpr size_t oc_smem_bytes(int cap, int d_max, int e_cap) {
    size_t b = 0;
    b += lap_align16(sizeof(float4) * (size_t)d_max);
    b += lap_align16(sizeof(float) * (size_t)d_max);
    b += 5 * lap_align16(sizeof(unsigned short) * (size_t)d_max);
    b += lap_align16((size_t)d_max);
    b += lap_align16(sizeof(float4) * (size_t)cap);
    b += 4 * lap_align16(sizeof(unsigned short) * (size_t)cap);
    b += lap_align16((size_t)cap);
    b += lap_align16(sizeof(unsigned) * (size_t)((d_max + 31) / 32));
    b += lap_align16(sizeof(unsigned) * (size_t)((cap + 31) / 32));
    b += lap_align16(sizeof(int) * 4);
    b += lap_align16(sizeof(BlockScratch));
    b += lap_smem_bytes(d_max, cap, e_cap);
    if (sizeof(int) * (size_t)e_cap < jv_work_bytes(kJvMax + 1)) b += lap_align16(jv_work_bytes(kJvMax + 1));
    return b;
}

__device__ __forceinline__ void oc_carve(unsigned char* p, int cap, int d_max, int e_cap, OcSmem& s) {
    s.det_box = (float4*)p;             p += lap_align16(sizeof(float4) * (size_t)d_max);
    s.det_conf = (float*)p;             p += lap_align16(sizeof(float) * (size_t)d_max);
    s.high = (unsigned short*)p;        p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.second = (unsigned short*)p;      p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.ud = (unsigned short*)p;          p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.ud2 = (unsigned short*)p;         p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.pair_det = (unsigned short*)p;    p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.det_flag = p;                     p += lap_align16((size_t)d_max);
    s.trk_box = (float4*)p;             p += lap_align16(sizeof(float4) * (size_t)cap);
    s.list_a = (unsigned short*)p;      p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.ut = (unsigned short*)p;          p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.ut2 = (unsigned short*)p;         p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.pair_trk = (unsigned short*)p;    p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.trk_flag = p;                     p += lap_align16((size_t)cap);
    s.row_bits = (unsigned*)p;          p += lap_align16(sizeof(unsigned) * (size_t)((d_max + 31) / 32));
    s.col_bits = (unsigned*)p;          p += lap_align16(sizeof(unsigned) * (size_t)((cap + 31) / 32));
    s.flags = (int*)p;                  p += lap_align16(sizeof(int) * 4);
    s.bs = (BlockScratch*)p;            p += lap_align16(sizeof(BlockScratch));
    p = lap_carve(p, d_max, cap, e_cap, s.lap);
    // the candidate-edge buffer is dead once block_lap has returned, which is exactly when the dense LAPJV runs
    s.jv = (sizeof(int) * (size_t)e_cap >= jv_work_bytes(kJvMax + 1)) ? (unsigned char*)s.lap.scratch_a : p;
}

// KalmanBoxTracker::get_state / predict's return value (free convert_x_to_bbox, ocsort.cpp:172-181)
__device__ __forceinline__ float4 oc_track_box(const float* rec) { return xysr2xyxy(rec[0], rec[1], rec[2], rec[3]); }

// k_previous_obs (ocsort.cpp:24-51) for a track that has at least one observation: the oldest entry among
// ages age-k .. age-1, else the newest observation overall (= last_observation).  Returns box, conf in `c`.
__device__ __forceinline__ float4 oc_k_previous_obs(const OcStream& st, int slot, int age, int k, float& c) {
    for (int dt = k; dt >= 1; --dt) {
        const int a = age - dt;
        if (a < 1) continue;                                    // no observation is ever stored under age <= 0
        const int e = slot * kOcRing + (a & (kOcRing - 1));
        if (st.ring_age[e] == a) { c = st.ring_conf[e]; return st.ring_box[e]; }
    }
    const float* o = st.obs + (size_t)slot * kOcObsFloats;
    c = o[4];
    return make_float4(o[0], o[1], o[2], o[3]);
}

// KalmanBoxTracker::update with a real box (ocsort.cpp:89-127) for n_pairs (track position, detection)
// pairs over DISTINCT tracks, one 8-lane group per pair.
template <class TrkOf, class DetOf>
__device__ __forceinline__ void oc_update_pairs(const OcStream& st, const OcSmem& sm, const float* dets, int n_pairs,
                                                int delta_t, TrkOf trk_of, DetOf det_of) {
    const int lane = lane_id(), g = lane & 7, base = lane & ~7;
    const int groups = (int)(blockDim.x >> 3), gid = (int)(threadIdx.x >> 3);
    const int rounds = (n_pairs + groups - 1) / groups;
    for (int it = 0; it < rounds; ++it) {
        const int q = it * groups + gid;
        const bool live = q < n_pairs;
        const int slot = live ? (int)sm.list_a[trk_of(q)] : 0;
        const int det = live ? det_of(q) : 0;
        float* rec = st.recs + (size_t)slot * kOcRecFloats;
        KfRow7 s;
        kf7_load_row(rec, live ? g : 7, s);
        if (!live) { s.m = 1.0f; for (int j = 0; j < 7; ++j) s.p[j] = (j == g) ? 1.0f : 0.0f; }
        float z[4] = {0.0f, 0.0f, 1.0f, 1.0f};
        const float4 box = live ? sm.det_box[det] : make_float4(0.0f, 0.0f, 1.0f, 1.0f);
        if (live) {
            const float4 zz = xyxy2xysr(box);
            z[0] = zz.x; z[1] = zz.y; z[2] = zz.z; z[3] = zz.w;
        }
        if (live && g == 0) {
            float* o = st.obs + (size_t)slot * kOcObsFloats;
            const float conf = sm.det_conf[det];
            const int age = st.age[slot];
            const float4 last = make_float4(o[0], o[1], o[2], o[3]);
            if (box_sum4(last) >= 0.0f) {                                        // :97-108
                float pc;
                const float4 prev = oc_k_previous_obs(st, slot, age, delta_t, pc);
                const float2 v = (box_sum4(prev) >= 0.0f) ? speed_direction(prev, box) : speed_direction(last, box);
                o[5] = v.x; o[6] = v.y;
            }
            o[0] = box.x; o[1] = box.y; o[2] = box.z; o[3] = box.w; o[4] = conf;   // :111-115
            const int e = slot * kOcRing + (age & (kOcRing - 1));
            st.ring_box[e] = box; st.ring_conf[e] = conf; st.ring_age[e] = age;
            const unsigned short tw = st.twin[slot];          // an updated track stops being anybody's twin
            if (tw != kNoTwin) { st.twin[tw] = kNoTwin; st.twin[slot] = kNoTwin; }
            st.det_ind[slot] = det;
            st.conf[slot] = conf;
            st.cls[slot] = (int)dets[(size_t)det * 6 + 5];
            st.tsu[slot] = 0;
            st.hits[slot] += 1;
            st.streak[slot] += 1;
        }
        const bool ok = kf_xysr_update(s, g, base, z);
        if (live) {
            if (ok) kf7_store_row(rec, g, s);
            else if (g == 0) atomicOr(&st.hdr[kOHdrError], 8);
        }
    }
}

// Updates for the matches of a BYTE / re-match assignment, whose column list may name a track twice: the
// reference applies them in ascending row order, so a track's second update must see its first.
//   row r matched iff lap.row2col[r] >= 0; det_of_row(r), trk_of_col(c) translate list positions.
// dp != nullptr (DeepOC-SORT with embeddings): every box update is followed by the track's update_emb (:869-870).
template <class DetOfRow, class TrkOfCol>
__device__ __forceinline__ int oc_apply_matches(const OcStream& st, OcSmem& sm, const float* dets, int n_rows, int delta_t,
                                                DetOfRow det_of_row, TrkOfCol trk_of_col, const DeepStream* dp = nullptr,
                                                const float* embs = nullptr, const OcParams* prm = nullptr) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
    int* first_row = sm.lap.col_label;                 // dead between block_lap calls; one int per track position
    for (int r = tid; r < n_rows; r += nt) {
        const int c = sm.lap.row2col[r];
        if (c >= 0) first_row[trk_of_col(c)] = 0x7fffffff;
    }
    __syncthreads();
    for (int r = tid; r < n_rows; r += nt) {
        const int c = sm.lap.row2col[r];
        if (c >= 0) atomicMin(&first_row[trk_of_col(c)], r);
    }
    __syncthreads();
    int total = 0;
    for (int round = 0; round < 2; ++round) {
        const int n_pairs = block_compact(n_rows, 0, sm.bs,
                                          [&](int r) {
                                              const int c = sm.lap.row2col[r];
                                              if (c < 0) return false;
                                              return (first_row[trk_of_col(c)] == r) == (round == 0);
                                          },
                                          [&](int r, int pos) {
                                              sm.pair_det[pos] = (unsigned short)det_of_row(r);
                                              sm.pair_trk[pos] = (unsigned short)trk_of_col(sm.lap.row2col[r]);
                                          });
        if (n_pairs == 0) continue;
        total += n_pairs;
        oc_update_pairs(st, sm, dets, n_pairs, delta_t, [&](int q) { return (int)sm.pair_trk[q]; },
                        [&](int q) { return (int)sm.pair_det[q]; });
        if (dp) deep_update_embs(*dp, sm, embs, *prm, n_pairs, [&](int q) { return (int)sm.pair_trk[q]; },
                                 [&](int q) { return (int)sm.pair_det[q]; });
        __syncthreads();
    }
    return total;
}

// KalmanBoxTracker ctor (ocsort.cpp:53-87) for `n_new` detections det_of(k), appended to list_a at n_trk + k
template <class DetOf>
__device__ __forceinline__ void oc_spawn(const OcStream& st, OcSmem& sm, const float* dets, int n_new, int n_trk, int n_free,
                                         int id_base, DetOf det_of) {
    const int lane = lane_id(), g = lane & 7;
    const int groups = (int)(blockDim.x >> 3), gid = (int)(threadIdx.x >> 3);
    for (int k = gid; k < n_new; k += groups) {
        const int det = det_of(k);
        const int slot = st.freel[n_free - 1 - k];
        const float4 q = xyxy2xysr(sm.det_box[det]);
        const float z[4] = {q.x, q.y, q.z, q.w};
        KfRow7 s;
        kf_xysr_init(s, g, z);
        kf7_store_row(st.recs + (size_t)slot * kOcRecFloats, g, s);
        st.ring_age[slot * kOcRing + g] = -1;
        if (g == 0) {
            st.id[slot] = id_base + 1 + k;
            st.age[slot] = 0; st.hits[slot] = 0; st.streak[slot] = 0; st.tsu[slot] = 0;
            st.conf[slot] = sm.det_conf[det];
            st.cls[slot] = (int)dets[(size_t)det * 6 + 5];
            st.det_ind[slot] = det;
            float* o = st.obs + (size_t)slot * kOcObsFloats;
            o[0] = -1.0f; o[1] = -1.0f; o[2] = -1.0f; o[3] = -1.0f; o[4] = -1.0f; o[5] = 0.0f; o[6] = 0.0f; o[7] = 0.0f;
            sm.list_a[n_trk + k] = (unsigned short)slot;
        }
    }
}

// The reference's dense LAPJV on the full n x m cost matrix pair(i, j); overwrites lap.row2col / col2row.
// rows + columns <= kJvMax: one warp, state in shared memory (jv_device.cuh); larger: the whole CTA, state in the
// stream's global scratch, scan order + reduction scratch in the (now idle) label / edge arrays of the sparse solver.
// All threads of the block must call.
template <int CAP, int DMAX, class Pair>
__device__ __forceinline__ void oc_exact_assignment(const OcStream& st, OcSmem& sm, int n, int m, float thresh, Pair pair) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
    for (int e = tid; e < n * m; e += nt) st.jv_dense[e] = pair(e / m, e - (e / m) * m);
    if (tid == 0) st.hdr[kOHdrExactSolves] += 1;
    __syncthreads();
    const JvCost cost{st.jv_dense, n, m, m, (double)thresh / 2.0};
    if (n + m <= kJvMax) {
        const JvWork w = jv_carve(sm.jv, kJvMax + 1);
        if (tid < 32) warp_dense_lapjv(cost, n + m, w);
        __syncthreads();
        for (int i = tid; i < n; i += nt) { const int j = w.x[i]; sm.lap.row2col[i] = (short)(j < m ? j : -1); }     // lap_solver.hpp:326-331
        for (int j = tid; j < m; j += nt) { const int i = w.y[j]; sm.lap.col2row[j] = (short)(i < n ? i : -1); }
    } else {
        unsigned char* idle = (unsigned char*)sm.lap.row_label;         // [row_label, row2col) is dead between two block_lap calls
        const JvBlockWork w = jv_block_carve(st.jv_work, idle, n + m, jv_block_sbytes_full(n + m) <= (size_t)((unsigned char*)sm.lap.row2col - idle));
        block_dense_lapjv(cost, n + m, w, sm.bs);
        for (int i = tid; i < n; i += nt) { const int j = w.x[i]; sm.lap.row2col[i] = (short)(j < m ? j : -1); }
        for (int j = tid; j < m; j += nt) { const int i = w.y[j]; sm.lap.col2row[j] = (short)(i < n ? i : -1); }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ DeepOC-SORT helpers
// order-preserving float <-> unsigned (so that an integer atomic can keep a maximum); -0 sorts below +0, which the
// callers do not distinguish
__device__ __forceinline__ unsigned deep_f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float deep_ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// the reference's running (largest, second largest) scan (deepocsort.cpp:303-313): a multiset top-2, NaNs never enter
struct DeepTop2 {
    float mx, se;
    __device__ __forceinline__ void push(float v) {
        if (v > mx) { se = mx; mx = v; }
        else if (v > se) se = v;
    }
};

// 1 - max(second / max - bottom, 0) / (1 - bottom) with std::max's operand order (a NaN first operand survives) (:318-320)
__device__ __forceinline__ float deep_aw_weight(float mx, float se, float bottom) {
    const float t = xsub(xdiv(se, mx), bottom);
    const float c = (t < 0.0f) ? 0.0f : t;
    return xsub(1.0f, xdiv(c, xsub(1.0f, bottom)));
}

// Ordered sums, warp-cooperative.  The contract (oracle/deepocsort.cpp, pinned against the reference) is the ascending-index
// sum with one rounding per operation - a serial chain per sum.  A warp therefore takes up to 32 sums at once: for 32
// consecutive indices at a time the operand slices of all 32 items are staged into padded shared tiles with cp.async
// (one coalesced 128-byte row per instruction, up to 64 of them in flight per warp and no registers held - the loop is
// bound by memory latency, so bytes in flight are what counts), and lane p then forms ITS item's 32 terms and adds them in
// ascending order.  The chains of 32 items run side by side and the order of every sum is exactly the scalar loop's.
constexpr int kDeepK = 32;                                   // indices per staged slice
constexpr int kDeepRow = kDeepK + 4;                         // padded tile row, 16-byte aligned for the 16-byte copies
constexpr int kDeepTileFloats = 2 * 32 * kDeepRow;           // operand tiles A and B of one warp

__device__ __forceinline__ void deep_cp4(float* dst_shared, const float* src_global) {
#if defined(MOT_CPUSIM)
    *dst_shared = *src_global;
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst_shared)), "l"(src_global)
                 : "memory");
#endif
}
__device__ __forceinline__ void deep_cp16(float* dst_shared, const float* src_global) {
#if defined(MOT_CPUSIM)
    for (int k = 0; k < 4; ++k) dst_shared[k] = src_global[k];
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_shared)), "l"(src_global)
                 : "memory");
#endif
}
__device__ __forceinline__ void deep_cp_wait() {
#if !defined(MOT_CPUSIM)
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// [scratch_a, row2col) of the sparse solver's workspace is dead outside block_lap (and outside oc_apply_matches'
// col_label use): the tiles live there, one pair per participating warp
__device__ __forceinline__ int deep_tiles(const OcSmem& sm, float*& tiles) {
    tiles = (float*)sm.lap.scratch_a;
    const size_t bytes = (size_t)((const unsigned char*)sm.lap.row2col - (const unsigned char*)sm.lap.scratch_a);
    const int n = (int)(bytes / (sizeof(float) * kDeepTileFloats));
    const int nwarps = (int)(blockDim.x >> 5);
    return n < nwarps ? n : nwarps;
}

// stage columns k0 .. k0 + 31 of row_a(p) (and row_b(p)) for the items p < n_items into tile rows p; warp-wide
template <class RowA, class RowB>
__device__ __forceinline__ void deep_stage(float* tile, int n_items, int k0, int dim, bool has_b, RowA row_a, RowB row_b) {
    const int lane = lane_id(), kk = k0 + lane;
    float* ta = tile + lane;
    float* tb = tile + 32 * kDeepRow + lane;
    // 16-byte copies when every row is 16-byte aligned and the slice is whole: a lane moves 4 floats, 8 lanes one row slice,
    // a warp instruction four items' slices
    const bool wide = (dim & 3) == 0 && k0 + kDeepK <= dim && ((((size_t)row_a(0)) | (has_b ? (size_t)row_b(0) : 0)) & 15) == 0;
    if (wide) {
        const int sub = (lane & 7) * 4, pi = lane >> 3;
        for (int p0 = 0; p0 < n_items; p0 += 4) {
            const int q = p0 + pi;
            if (q < n_items) {
                deep_cp16(tile + q * kDeepRow + sub, row_a(q) + k0 + sub);
                if (has_b) deep_cp16(tile + 32 * kDeepRow + q * kDeepRow + sub, row_b(q) + k0 + sub);
            }
        }
    } else if (kk < dim) {
#pragma unroll 4
        for (int p = 0; p < n_items; ++p) {
            deep_cp4(ta + p * kDeepRow, row_a(p) + kk);
            if (has_b) deep_cp4(tb + p * kDeepRow, row_b(p) + kk);
        }
    }
    deep_cp_wait();
    __syncwarp();
}

// returns, in lane p < n_items (<= 32), sum_{k < dim} f(p, row_a(p)[k], row_b(p)[k]) in ascending k.  All 32 lanes must call.
template <class RowA, class RowB, class F>
__device__ __forceinline__ float deep_warp_ordered_sums(float* tile, int n_items, int dim, bool has_b, RowA row_a, RowB row_b, F f) {
    const int lane = lane_id();
    float acc = 0.0f;
    for (int k0 = 0; k0 < dim; k0 += kDeepK) {
        deep_stage(tile, n_items, k0, dim, has_b, row_a, row_b);
        if (lane < n_items) {
            const int cnt = min(kDeepK, dim - k0);
            const float* ra = tile + lane * kDeepRow;
            const float* rb = ra + 32 * kDeepRow;
            int q = 0;
            if (k0 == 0) { acc = f(lane, ra[0], has_b ? rb[0] : 0.0f); q = 1; }
            for (; q < cnt; ++q) acc = xadd(acc, f(lane, ra[q], has_b ? rb[q] : 0.0f));
        }
        __syncwarp();
    }
    return acc;
}

// emb <- normalise(blend ? alpha * emb + (1 - alpha) * det_emb : det_emb) for n_items (track slot, detection) pairs over
// DISTINCT slots: DeepOCSortKalmanBoxTracker's ctor copy (:73-80) and update_emb (:143-161) with
//   alpha = alpha_fixed + (1 - alpha_fixed) * (1 - trust),  trust = (conf - det_thresh) / (1 - det_thresh)   (:650-652)
// v <- v / ||v|| when the norm exceeds 1e-6 (:75-80, :154-158); ascending-index sum of squares.
template <class SlotOf, class DetOf>
__device__ __forceinline__ void deep_write_embs(const DeepStream& dp, const OcSmem& sm, const float* embs, const OcParams& p,
                                                int n_items, bool blend, SlotOf slot_of, DetOf det_of) {
    float* tiles;
    const int nw = deep_tiles(sm, tiles);
    const int warp = (int)(threadIdx.x >> 5), lane = lane_id();
    if (warp >= nw) return;
    float* tile = tiles + (size_t)warp * kDeepTileFloats;
    const int dim = dp.dim;
    for (int b0 = warp * 32; b0 < n_items; b0 += nw * 32) {
        const int cnt = min(32, n_items - b0);
        float alpha = 0.0f, beta = 1.0f;
        if (blend && lane < cnt) {
            const float trust = xdiv(xsub(sm.det_conf[det_of(b0 + lane)], p.det_thresh), xsub(1.0f, p.det_thresh));
            alpha = xadd(p.alpha_fixed_emb, xmul(xsub(1.0f, p.alpha_fixed_emb), xsub(1.0f, trust)));
            beta = xsub(1.0f, alpha);
        }
        // operand A: the detection's embedding; operand B (blend only): the track's current embedding
        auto row_a = [&](int q) { return embs + (size_t)det_of(b0 + q) * dim; };
        auto row_b = [&](int q) { return (const float*)(dp.trk_emb + (size_t)slot_of(b0 + q) * dim); };
        const float sq = deep_warp_ordered_sums(tile, cnt, dim, blend, row_a, row_b, [&](int, float s, float e) {
            const float v = blend ? xadd(xmul(alpha, e), xmul(beta, s)) : s;       // lane q owns item q: its own alpha / beta
            return xmul(v, v);
        });
        const float norm = xsqrt(sq);
        // second pass: the same slices again, every lane now writes column `lane` of all items (coalesced rows)
        for (int k0 = 0; k0 < dim; k0 += kDeepK) {
            deep_stage(tile, cnt, k0, dim, blend, row_a, row_b);
            const int kk = k0 + lane;
            for (int q = 0; q < cnt; ++q) {
                const float n = __shfl_sync(kFullMask, norm, q);
                const float al = __shfl_sync(kFullMask, alpha, q), be = __shfl_sync(kFullMask, beta, q);
                if (kk < dim) {
                    const float sv = tile[q * kDeepRow + lane];
                    const float v = blend ? xadd(xmul(al, tile[32 * kDeepRow + q * kDeepRow + lane]), xmul(be, sv)) : sv;
                    dp.trk_emb[(size_t)slot_of(b0 + q) * dim + kk] = (n > 1e-6f) ? xdiv(v, n) : v;
                }
            }
            __syncwarp();
        }
    }
}

template <class DetOf>
__device__ __forceinline__ void deep_spawn_embs(const DeepStream& dp, const OcSmem& sm, const float* embs, const OcParams& p,
                                                int n_new, int n_trk, DetOf det_of) {
    deep_write_embs(dp, sm, embs, p, n_new, false, [&](int k) { return (int)sm.list_a[n_trk + k]; }, det_of);
}
template <class TrkOf, class DetOf>
__device__ __forceinline__ void deep_update_embs(const DeepStream& dp, const OcSmem& sm, const float* embs, const OcParams& p,
                                                 int n_pairs, TrkOf trk_of, DetOf det_of) {
    deep_write_embs(dp, sm, embs, p, n_pairs, true, [&](int q) { return (int)sm.list_a[trk_of(q)]; }, det_of);
}

// insert v into the packed (largest, second largest) pair at *top (order-preserving encodings); NaNs never enter
__device__ __forceinline__ void deep_top2_insert(unsigned long long* top, float v) {
    if (!(v == v)) return;
    const unsigned ov = deep_f2ord(v);
    unsigned long long old = *top;
    for (;;) {
        const unsigned mx = (unsigned)(old >> 32), se = (unsigned)old;
        unsigned long long upd;
        if (ov > mx) upd = ((unsigned long long)ov << 32) | mx;
        else if (ov > se) upd = ((unsigned long long)mx << 32) | ov;
        else break;
        const unsigned long long seen = atomicCAS(top, old, upd);
        if (seen == old) break;
        old = seen;
    }
}

// (largest, second largest) of `stored` entries + `total - stored` zeros -> weight and "maximum is 0" flag (:303-321)
__device__ __forceinline__ float deep_top2_weight(unsigned long long top, int zeros, float bottom, bool& is_zero) {
    DeepTop2 t{deep_ord2f((unsigned)(top >> 32)), deep_ord2f((unsigned)top)};
    if (zeros >= 1) t.push(0.0f);
    if (zeros >= 2) t.push(0.0f);
    is_zero = (t.mx == 0.0f);
    return is_zero ? 0.0f : deep_aw_weight(t.mx, t.se, bottom);
}

// The appearance terms of the first association (deepocsort.cpp:756-766 GEMM, :420-440 mask + weights), sparse: only
// the (detection, track) pairs whose boxes overlap (iou > 0) have a non-zero entry in the reference's masked matrix, so
// only they are multiplied out; the adaptive weights' row / column top-2 scans count the remaining entries as zeros.
//   1. one thread per row walks the column grid and appends its overlapping pairs to a list (the idle dense-LAPJV
//      matrix of the stream's slab holds it: n_high x n_trk entries at most);
//   2. warps take 32 pairs at a time through deep_warp_ordered_sums, store the products in `dense` and fold them into
//      the packed row / column top-2 (one atomic per pair and side);
//   3. one thread per row / column turns its top-2 into the weight.
// All threads of the block must call.  Returns the DeepOcmCost mode (1 adaptive weights, 2 plain w_assoc_emb).
__device__ __forceinline__ int deep_embedding_terms(const DeepStream& dp, const OcStream& st, OcSmem& sm, const float* embs,
                                                    const OcParams& p, int n_high, int n_trk) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
    const int mode = p.aw_off ? 2 : 1;
    const int dim = dp.dim;
    const float ninf = __int_as_float(0xff800000);
    const unsigned long long empty = ((unsigned long long)deep_f2ord(ninf) << 32) | deep_f2ord(ninf);
    unsigned* pairs = (unsigned*)st.jv_dense;
    if (mode == 1) {
        for (int j = tid; j < n_trk; j += nt) { dp.col_top[j] = empty; dp.col_nnz[j] = 0; }
        for (int i = tid; i < n_high; i += nt) { dp.row_top[i] = empty; dp.row_nnz[i] = 0; }
    }
    grid_build(sm.lap.grid, n_trk, sm.bs, [&](int j) { return sm.trk_box[j]; });
    for (int i = tid; i < n_high; i += nt) {
        const float4 box = sm.det_box[sm.high[i]];
        const float area = box_area(box);
        grid_query(sm.lap.grid, box, [&](int j) { return sm.trk_box[j]; }, [&](int j, float4 b) {
            if (iou_pair(box, area, b) <= 0.0f) return;
            pairs[atomicAdd(&sm.flags[3], 1)] = ((unsigned)i << 16) | (unsigned)j;
        });
    }
    __syncthreads();
    const int n_pairs = sm.flags[3];
    {
        float* tiles;
        const int nw = deep_tiles(sm, tiles);
        const int warp = tid >> 5, lane = tid & 31;
        if (warp < nw) {
            float* tile = tiles + (size_t)warp * kDeepTileFloats;
            for (int b0 = warp * 32; b0 < n_pairs; b0 += nw * 32) {
                const int cnt = min(32, n_pairs - b0);
                const float val = deep_warp_ordered_sums(
                    tile, cnt, dim, true, [&](int q) { return embs + (size_t)sm.high[pairs[b0 + q] >> 16] * dim; },
                    [&](int q) { return (const float*)(dp.trk_emb + (size_t)sm.list_a[pairs[b0 + q] & 0xffffu] * dim); },
                    [&](int, float de, float te) { return xmul(de, te); });
                if (lane < cnt) {
                    const unsigned pr = pairs[b0 + lane];
                    const int i = (int)(pr >> 16), j = (int)(pr & 0xffffu);
                    dp.dense[(size_t)i * n_trk + j] = val;
                    if (mode == 1) {
                        deep_top2_insert(&dp.row_top[i], val); atomicAdd(&dp.row_nnz[i], 1);
                        deep_top2_insert(&dp.col_top[j], val); atomicAdd(&dp.col_nnz[j], 1);
                    }
                }
            }
        }
    }
    __syncthreads();
    if (tid == 0) sm.flags[3] = 0;
    if (mode == 1) {
        for (int i = tid; i < n_high; i += nt) {
            float w = p.w_assoc_emb;
            if (n_trk >= 2) {                                                    // (:302)
                bool z;
                const float rw = deep_top2_weight(dp.row_top[i], n_trk - dp.row_nnz[i], p.aw_param, z);
                w = z ? 0.0f : xmul(w, rw);
            }
            dp.row_w[i] = w;
        }
        if (n_high >= 2)                                                         // (:324)
            for (int j = tid; j < n_trk; j += nt) {
                bool z;
                dp.col_w[j] = deep_top2_weight(dp.col_top[j], n_high - dp.col_nnz[j], p.aw_param, z);
                dp.col_z[j] = z ? 1 : 0;
            }
    }
    __syncthreads();
    return mode;
}

template <int CAP, int DMAX, bool DEEP, int ASSO>
__device__ __forceinline__ void oc_frame(const OcArgs& a, const OcStream& st, OcSmem& sm, const float* dets, int n_det_in,
                                         float* out, int* n_out, const DeepStream& dp, const float* embs) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
    const int lane = tid & 31, g = lane & 7, base = lane & ~7;
    const int groups = nt >> 3, gid = tid >> 3;
    __syncthreads();
    const int frame = st.hdr[kOHdrFrame] + 1;                    // frame_count_ (:293)
    const int n_trk0 = st.hdr[kOHdrTracks];
    int n_free = st.hdr[kOHdrFree];
    const int id_base = st.hdr[kOHdrIdCounter];
    const float thr = a.p.iou_threshold;
    const int delta_t = a.p.delta_t;
    int n_det = n_det_in;
    if (n_det > min(DMAX, a.ld_dets)) { n_det = min(DMAX, a.ld_dets); if (tid == 0) atomicOr(&st.hdr[kOHdrError], 2); }

    // ---- A. detections and the confidence split (:311-320)
    float max_abs_score = 0.0f;
    for (int j = tid; j < n_det; j += nt) {
        const float* r = dets + (size_t)j * 6;
        sm.det_box[j] = make_float4(r[0], r[1], r[2], r[3]);
        sm.det_conf[j] = r[4];
        max_abs_score = fmaxf(max_abs_score, fabsf(r[4]));
    }
    if (tid < 4) sm.flags[tid] = 0;
    __syncthreads();
    const float dth = a.p.det_thresh, lo = a.p.min_conf;
    const int n_high = block_compact(n_det, 0, sm.bs, [&](int j) { return sm.det_conf[j] > dth; },
                                     [&](int j, int pos) { sm.high[pos] = (unsigned short)j; });
    int n_second = 0;
    if (a.p.use_byte)
        n_second = block_compact(n_det, 0, sm.bs, [&](int j) { const float c = sm.det_conf[j]; return c > lo && c < dth; },
                                 [&](int j, int pos) { sm.second[pos] = (unsigned short)j; });
    // a disjoint pair has iou 0: it can be an assignment candidate only if the angle cost alone reaches the
    // threshold (|angle cost| <= inertia * score / 2) and it never satisfies iou > thr when thr >= 0
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) max_abs_score = fmaxf(max_abs_score, __shfl_xor_sync(kFullMask, max_abs_score, o));
        if (lane == 0) atomicMax(&sm.flags[3], __float_as_int(max_abs_score));   // non-negative floats order as ints
    }
    __syncthreads();
    const float score_bound = __int_as_float(sm.flags[3]);
    // (a centroid similarity is non-zero for disjoint boxes too: no pruning with it)
    const bool prune_first = ASSO == 0 && thr >= 0.0f && (0.5f * fabsf(a.p.inertia) * score_bound * 1.0001f + 1e-7f) < thr;
    const bool prune_rest = ASSO == 0 && thr > 0.0f;
    const bool use_emb = DEEP && !a.p.embedding_off;
    __syncthreads();
    if (tid == 0) sm.flags[3] = 0;

    // ---- B. predict every track in place (KalmanBoxTracker::predict :134-151); a NaN box removes the track
    {
        const int rounds = (n_trk0 + groups - 1) / groups;
        for (int it = 0; it < rounds; ++it) {
            const int k = it * groups + gid;
            const bool live = k < n_trk0;
            const int slot = live ? (int)st.list[k] : 0;
            float* rec = st.recs + (size_t)slot * kOcRecFloats;
            KfRow7 s;
            kf7_load_row(rec, live ? g : 7, s);
            const float x6 = __shfl_sync(kFullMask, s.m, base + 6), x2 = __shfl_sync(kFullMask, s.m, base + 2);
            if (g == 6 && xadd(x6, x2) <= 0.0f) s.m = 0.0f;                  // :135-137
            kf_xysr_predict(s, g, base, a.p.q44, a.p.q66);
            const float x0 = __shfl_sync(kFullMask, s.m, base + 0), x1 = __shfl_sync(kFullMask, s.m, base + 1);
            const float xs = __shfl_sync(kFullMask, s.m, base + 2), xr = __shfl_sync(kFullMask, s.m, base + 3);
            if (live) {
                kf7_store_row(rec, g, s);
                if (g == 0) {
                    st.age[slot] += 1;
                    if (st.tsu[slot] > 0) st.streak[slot] = 0;
                    st.tsu[slot] += 1;
                    const float4 b = xysr2xyxy(x0, x1, xs, xr);
                    sm.trk_flag[k] = (b.x != b.x || b.y != b.y || b.z != b.z || b.w != b.w) ? 1 : 0;
                }
            }
        }
    }
    __syncthreads();
    const int n_trk = block_compact(n_trk0, 0, sm.bs, [&](int k) { return sm.trk_flag[k] == 0; },
                                    [&](int k, int pos) { sm.list_a[pos] = st.list[k]; });
    n_free = block_compact(n_trk0, n_free, sm.bs, [&](int k) { return sm.trk_flag[k] != 0; },
                           [&](int k, int pos) {
                               const int slot = st.list[k];
                               const unsigned short tw = st.twin[slot];
                               if (tw != kNoTwin) { st.twin[tw] = kNoTwin; st.twin[slot] = kNoTwin; }
                               st.freel[pos] = (unsigned short)slot;
                           });

    if (n_trk == 0) {
        // ---- no tracks: every high detection starts one, nothing is emitted (:367-384)
        int n_new = n_high;
        if (n_new > n_free) { n_new = n_free; if (tid == 0) atomicOr(&st.hdr[kOHdrError], 1); }
        oc_spawn(st, sm, dets, n_new, 0, n_free, id_base, [&](int k) { return (int)sm.high[k]; });
        __syncthreads();
        if constexpr (DEEP) {
            if (use_emb) deep_spawn_embs(dp, sm, embs, a.p, n_new, 0, [&](int k) { return (int)sm.high[k]; });
        }
        for (int k = tid; k < n_new; k += nt) { st.list[k] = sm.list_a[k]; st.twin[sm.list_a[k]] = kNoTwin; }
        if (tid == 0) {
            *n_out = 0;
            st.hdr[kOHdrTracks] = n_new;
            st.hdr[kOHdrFree] = n_free - n_new;
            st.hdr[kOHdrIdCounter] = id_base + n_new;
            st.hdr[kOHdrFrame] = frame;
            st.hdr[kOHdrNHigh] = n_high; st.hdr[kOHdrNTrk] = 0; st.hdr[kOHdrUsedLap] = 0; st.hdr[kOHdrMatched] = 0;
            st.hdr[kOHdrLeftDets] = 0; st.hdr[kOHdrLeftTrks] = 0; st.hdr[kOHdrRematched] = 0; st.hdr[kOHdrSpawned] = 0;
        }
        __syncthreads();
        return;
    }

    // ---- C. predicted boxes, velocities and k_previous_obs of every track (:394-410)
    for (int k = tid; k < n_trk; k += nt) {
        const int slot = sm.list_a[k];
        sm.trk_box[k] = oc_track_box(st.recs + (size_t)slot * kOcRecFloats);
        const float* o = st.obs + (size_t)slot * kOcObsFloats;
        float4 prev = make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
        float pc = -1.0f;
        if (st.hits[slot] > 0) prev = oc_k_previous_obs(st, slot, st.age[slot], delta_t, pc);
        st.ocm[k] = make_float4(xdiv(xadd(prev.x, prev.z), 2.0f), xdiv(xadd(prev.y, prev.w), 2.0f), o[5], o[6]);
        st.valid[k] = (unsigned char)(((pc >= 0.0f) ? 1 : 0) | ((st.twin[slot] != kNoTwin) ? 2 : 0));
        sm.trk_flag[k] = 0;
    }
    for (int i = tid; i < n_high; i += nt) sm.det_flag[i] = 0;
    for (int w = tid; w < (DMAX + 31) / 32; w += nt) sm.row_bits[w] = 0;
    for (int w = tid; w < (CAP + 31) / 32; w += nt) sm.col_bits[w] = 0;
    __syncthreads();

    // ---- D. first association (:413-420, associate :610-737); rows = high detections, columns = tracks
    static_assert(!DEEP || ASSO == 0, "DeepOC-SORT's sparse appearance terms rely on IoU (iou <= 0 masks the product)");
    const OcmCostT<ASSO> ocm_base{sm.det_box, sm.det_conf, sm.high, sm.trk_box, st.ocm, st.valid, a.p.inertia, thr, prune_first,
                                  sm.row_bits, sm.col_bits, sm.pair_det /* row_hit */, sm.flags, a.p.asso_norm};
    std::conditional_t<DEEP, DeepOcmCost, OcmCostT<ASSO>> ocm;
    if constexpr (DEEP) {
        // appearance term (deepocsort.cpp:420-440): needs detections (the reference leaves the matrix empty without, :756)
        const int emb_mode = (use_emb && n_high > 0) ? deep_embedding_terms(dp, st, sm, embs, a.p, n_high, n_trk) : 0;
        ocm = DeepOcmCost{ocm_base, dp.dense, dp.row_w, dp.col_w, dp.col_z, n_trk, emb_mode, a.p.w_assoc_emb, n_high >= 2, prune_first};
    } else {
        ocm = ocm_base;
    }
    block_lap(sm.lap, n_high, n_trk, DMAX, CAP, -thr, ocm);
    const bool trivial = sm.flags[0] != 0 && sm.flags[1] == 0;   // max row sum == 1 && max column sum == 1 (:676-680)
    // the optimum can only be non-unique if it uses a track that still has a bit-identical twin (the twin is then an
    // equally good column); bit 1 of valid[] marks such tracks
    if (!trivial)
        for (int j = tid; j < n_trk; j += nt)
            if (sm.lap.col2row[j] >= 0 && (st.valid[j] & 2)) sm.flags[2] = 1;
    __syncthreads();
    const bool exact1 = sm.flags[2] != 0;
    __syncthreads();
    if (exact1) {
        // a twin track is a candidate: the optimum may be non-unique, so redo the assignment with the reference's own
        // dense LAPJV over the full (n_high x n_trk) cost matrix (associate :690-700 -> linear_assignment)
        oc_exact_assignment<CAP, DMAX>(st, sm, n_high, n_trk, -thr, [&](int i, int j) { return ocm.pair(i, j); });
    }
    if (trivial) {
        // every pair with iou > thr is a match, nothing else is (:681-689)
        for (int j = tid; j < n_trk; j += nt) sm.lap.col2row[j] = -1;
        __syncthreads();
        for (int i = tid; i < n_high; i += nt) {
            const bool hit = (sm.row_bits[i >> 5] >> (i & 31)) & 1u;
            const int j = hit ? (int)sm.pair_det[i] : -1;
            sm.lap.row2col[i] = (short)j;
            if (hit) { sm.lap.col2row[j] = (short)i; sm.det_flag[i] = 1; sm.trk_flag[j] = 1; }
        }
    } else {
        // assignment pairs below the IoU threshold go to both unmatched lists (:702-712)
        for (int i = tid; i < n_high; i += nt) {
            const int j = sm.lap.row2col[i];
            if (j < 0) continue;
            const auto rw = ocm.row(i);
            const unsigned char f = (ocm.iou(rw, j) >= thr) ? 1 : 2;
            sm.det_flag[i] = f; sm.trk_flag[j] = f;
        }
    }
    __syncthreads();
    // unmatched lists in the reference's order: filtered pairs (ascending detection), then the sweep (:715-735)
    int n_ud = block_compact(n_high, 0, sm.bs, [&](int i) { return sm.det_flag[i] == 2; },
                             [&](int i, int pos) {
                                 if (pos < DMAX) sm.ud[pos] = sm.high[i];
                                 if (pos < CAP) sm.ut[pos] = (unsigned short)sm.lap.row2col[i];
                             });
    int n_ut = n_ud;
    int n_dup = n_ud;                       // list entries that the final sweep below adds a second time
    if constexpr (DEEP) {
        // DeepOC-SORT's associate also lists what the assignment left unmatched BEFORE the sweep (:476-481), so after
        // an assignment every unmatched detection and track sits in its list twice
        if (!trivial && n_high > 0) {
            n_ud = block_compact(n_high, n_ud, sm.bs, [&](int i) { return sm.det_flag[i] == 0; },
                                 [&](int i, int pos) { if (pos < DMAX) sm.ud[pos] = sm.high[i]; });
            n_ut = block_compact(n_trk, n_ut, sm.bs, [&](int j) { return sm.trk_flag[j] == 0; },
                                 [&](int j, int pos) { if (pos < CAP) sm.ut[pos] = (unsigned short)j; });
            n_dup = max(n_ud, n_ut);
        }
    }
    n_ud = block_compact(n_high, n_ud, sm.bs, [&](int i) { return sm.det_flag[i] != 1; },
                         [&](int i, int pos) { if (pos < DMAX) sm.ud[pos] = sm.high[i]; });
    n_ut = block_compact(n_trk, n_ut, sm.bs, [&](int j) { return sm.trk_flag[j] != 1; },
                         [&](int j, int pos) { if (pos < CAP) sm.ut[pos] = (unsigned short)j; });
    if (n_ud > DMAX || n_ut > CAP) {
        if (tid == 0) atomicOr(&st.hdr[kOHdrError], 1);
        n_ud = min(n_ud, DMAX); n_ut = min(n_ut, CAP);
    }
    const int n_match = block_compact(n_high, 0, sm.bs, [&](int i) { return sm.det_flag[i] == 1; },
                                      [&](int i, int pos) {
                                          sm.pair_det[pos] = sm.high[i];
                                          sm.pair_trk[pos] = (unsigned short)sm.lap.row2col[i];
                                      });

    // ---- E. update the matched tracks (:423-430)
    oc_update_pairs(st, sm, dets, n_match, delta_t, [&](int q) { return (int)sm.pair_trk[q]; },
                    [&](int q) { return (int)sm.pair_det[q]; });
    if constexpr (DEEP) {
        if (use_emb) deep_update_embs(dp, sm, embs, a.p, n_match, [&](int q) { return (int)sm.pair_trk[q]; },
                                      [&](int q) { return (int)sm.pair_det[q]; });
    }
    __syncthreads();
    // detection / track flags now mean "taken out of the unmatched lists"
    for (int j = tid; j < n_det; j += nt) sm.det_flag[j] = 0;
    for (int k = tid; k < n_trk; k += nt) sm.trk_flag[k] = 0;
    __syncthreads();

    // ---- F. BYTE pass: low-confidence detections x unmatched tracks, predicted boxes (:433-479)
    if (a.p.use_byte && n_second > 0 && n_ut > 0) {
        if (tid == 0) sm.flags[0] = 0;
        __syncthreads();
        if (tid == 0) sm.flags[2] = 0;
        __syncthreads();
        NegIouCostT<ASSO> cost{sm.det_box, sm.second, sm.trk_box, sm.ut, thr, prune_rest, sm.flags, a.p.asso_norm};
        block_lap(sm.lap, n_second, n_ut, DMAX, CAP, -thr, cost);
        if (sm.flags[0] != 0)
            for (int p = tid; p < n_ut; p += nt)
                if (sm.lap.col2row[p] >= 0 && (st.valid[sm.ut[p]] & 2)) sm.flags[2] = 1;
        __syncthreads();
        if (sm.flags[2] != 0) {
            __syncthreads();
            oc_exact_assignment<CAP, DMAX>(st, sm, n_second, n_ut, -thr, [&](int i, int j) { return cost.pair(i, j); });
        }
        if (sm.flags[0] != 0) {                                   // max_iou > threshold (:445-446)
            for (int r = tid; r < n_second; r += nt) {
                const int c = sm.lap.row2col[r];
                if (c >= 0) sm.trk_flag[sm.ut[c]] = 1;
            }
            oc_apply_matches(st, sm, dets, n_second, delta_t, [&](int r) { return (int)sm.second[r]; },
                             [&](int c) { return (int)sm.ut[c]; });
            const int n_keep = block_compact(n_ut, 0, sm.bs, [&](int p) { return sm.trk_flag[sm.ut[p]] == 0; },
                                             [&](int p, int pos) { sm.ut2[pos] = sm.ut[p]; });
            for (int p = tid; p < n_keep; p += nt) sm.ut[p] = sm.ut2[p];
            n_ut = n_keep;
        }
        __syncthreads();
    }

    // ---- G. re-match the leftovers on the tracks' last observations (:482-545)
    int n_rematch = 0, n_left_d = 0, n_left_t = 0;
    if (n_ud > 0 && n_ut > 0) {
        n_left_d = n_ud; n_left_t = n_ut;
        for (int p = tid; p < n_ut; p += nt) {
            const int k = sm.ut[p];
            const float* o = st.obs + (size_t)sm.list_a[k] * kOcObsFloats;
            sm.trk_box[k] = make_float4(o[0], o[1], o[2], o[3]);
        }
        if (tid == 0) { sm.flags[0] = 0; sm.flags[2] = 0; }
        __syncthreads();
        NegIouCostT<ASSO> cost{sm.det_box, sm.ud, sm.trk_box, sm.ut, thr, prune_rest, sm.flags, a.p.asso_norm};
        block_lap(sm.lap, n_ud, n_ut, DMAX, CAP, -thr, cost);
        // Exact ties here: (1) the lists hold an entry twice (pairs rejected by the IoU filter; in DeepOC-SORT also everything the assignment left unmatched) - which
        // COPY of a detection is matched decides the order in which a twice-listed track receives its two updates;
        // (2) DIFFERENT tracks with bit-identical last observations: twins that both stayed unmatched, and tracks that an
        // earlier frame updated with the same (duplicated) detection.  In both cases the reference's answer is its LAPJV's.
        if (sm.flags[0] != 0 && n_dup > 0 && tid == 0) sm.flags[2] = 1;
        if (sm.flags[0] != 0)
            for (int p = tid; p < n_ut; p += nt) {
                if (sm.lap.col2row[p] < 0) continue;
                const int k = sm.ut[p];
                const float4 b = sm.trk_box[k];
                for (int q = 0; q < n_ut; ++q) {
                    const int k2 = sm.ut[q];
                    const float4 b2 = sm.trk_box[k2];
                    if (k2 != k && b2.x == b.x && b2.y == b.y && b2.z == b.z && b2.w == b.w) { sm.flags[2] = 1; break; }
                }
            }
        __syncthreads();
        if (sm.flags[2] != 0) {
            __syncthreads();
            oc_exact_assignment<CAP, DMAX>(st, sm, n_ud, n_ut, -thr, [&](int i, int j) { return cost.pair(i, j); });
        }
        if (sm.flags[0] != 0) {                                   // max_iou > threshold (:507-509)
            for (int r = tid; r < n_ud; r += nt) {
                const int c = sm.lap.row2col[r];
                if (c >= 0) { sm.trk_flag[sm.ut[c]] = 1; sm.det_flag[sm.ud[r]] = 1; }
            }
            n_rematch = oc_apply_matches(st, sm, dets, n_ud, delta_t, [&](int r) { return (int)sm.ud[r]; },
                                         [&](int c) { return (int)sm.ut[c]; }, use_emb ? &dp : nullptr, embs, &a.p);
            const int kt = block_compact(n_ut, 0, sm.bs, [&](int p) { return sm.trk_flag[sm.ut[p]] == 0; },
                                         [&](int p, int pos) { sm.ut2[pos] = sm.ut[p]; });
            const int kd = block_compact(n_ud, 0, sm.bs, [&](int p) { return sm.det_flag[sm.ud[p]] == 0; },
                                         [&](int p, int pos) { sm.ud2[pos] = sm.ud[p]; });
            for (int p = tid; p < kt; p += nt) sm.ut[p] = sm.ut2[p];
            for (int p = tid; p < kd; p += nt) sm.ud[p] = sm.ud2[p];
            n_ut = kt; n_ud = kd;
        }
        __syncthreads();
    }

    // ---- H. update(None) for the tracks still unmatched: only det_ind changes (:548-550, :90, :128-131)
    for (int p = tid; p < n_ut; p += nt) st.det_ind[sm.list_a[sm.ut[p]]] = 0;

    // ---- I. new tracks for the detections still unmatched, list order, duplicates included (:553-561)
    int n_new = n_ud;
    if (n_new > n_free || n_trk + n_new > CAP) {
        n_new = min(n_free, CAP - n_trk);
        if (tid == 0) atomicOr(&st.hdr[kOHdrError], 1);
    }
    oc_spawn(st, sm, dets, n_new, n_trk, n_free, id_base, [&](int k) { return (int)sm.ud[k]; });
    __syncthreads();
    if constexpr (DEEP) {
        if (use_emb) deep_spawn_embs(dp, sm, embs, a.p, n_new, n_trk, [&](int k) { return (int)sm.ud[k]; });
        __syncthreads();
    }
    {
        // a detection that sits twice in the list has just spawned two bit-identical tracks: link them as twins
        int* first_pos = sm.lap.row_label;          // [DMAX] ints, idle outside block_lap
        int* last_pos = sm.lap.scratch_b;           // [DMAX] ints
        for (int k = tid; k < n_new; k += nt) { first_pos[sm.ud[k]] = 0x7fffffff; last_pos[sm.ud[k]] = -1; }
        __syncthreads();
        for (int k = tid; k < n_new; k += nt) { atomicMin(&first_pos[sm.ud[k]], k); atomicMax(&last_pos[sm.ud[k]], k); }
        __syncthreads();
        for (int k = tid; k < n_new; k += nt) {
            const int d = sm.ud[k], f = first_pos[d], l = last_pos[d];
            st.twin[sm.list_a[n_trk + k]] = (f != l) ? sm.list_a[n_trk + (k == f ? l : f)] : kNoTwin;
        }
        __syncthreads();
    }
    const int n_all = n_trk + n_new;
    n_free -= n_new;

    // ---- J. output in REVERSE track order, then age-out (:564-592)
    const int min_hits = a.p.min_hits, max_age = a.p.max_age;
    const int n_rows = block_compact(n_all, 0, sm.bs,
                                     [&](int q) {
                                         const int slot = sm.list_a[n_all - 1 - q];
                                         return st.tsu[slot] < 1 && (st.streak[slot] >= min_hits || frame <= min_hits);
                                     },
                                     [&](int q, int pos) {
                                         if (pos >= a.ld_out) return;
                                         const int slot = sm.list_a[n_all - 1 - q];
                                         const float* o = st.obs + (size_t)slot * kOcObsFloats;
                                         float4 b = make_float4(o[0], o[1], o[2], o[3]);
                                         if (box_sum4(b) < 0.0f) b = oc_track_box(st.recs + (size_t)slot * kOcRecFloats);
                                         float* w = out + (size_t)pos * 8;
                                         *reinterpret_cast<float4*>(w) = b;
                                         *reinterpret_cast<float4*>(w + 4) = make_float4((float)(st.id[slot] + (DEEP ? 0 : 1)), st.conf[slot],
                                                                                         (float)st.cls[slot], (float)st.det_ind[slot]);
                                     });
    const int n_keep = block_compact(n_all, 0, sm.bs, [&](int k) { return st.tsu[sm.list_a[k]] <= max_age; },
                                     [&](int k, int pos) { st.list[pos] = sm.list_a[k]; });
    n_free = block_compact(n_all, n_free, sm.bs, [&](int k) { return st.tsu[sm.list_a[k]] > max_age; },
                           [&](int k, int pos) {
                               const int slot = sm.list_a[k];
                               const unsigned short tw = st.twin[slot];
                               if (tw != kNoTwin) { st.twin[tw] = kNoTwin; st.twin[slot] = kNoTwin; }
                               st.freel[pos] = (unsigned short)slot;
                           });
    if (tid == 0) {
        if (n_rows > a.ld_out) atomicOr(&st.hdr[kOHdrError], 4);
        *n_out = n_rows < a.ld_out ? n_rows : a.ld_out;
        st.hdr[kOHdrTracks] = n_keep;
        st.hdr[kOHdrFree] = n_free;
        st.hdr[kOHdrIdCounter] = id_base + n_new;
        st.hdr[kOHdrFrame] = frame;
        st.hdr[kOHdrNHigh] = n_high; st.hdr[kOHdrNTrk] = n_trk; st.hdr[kOHdrUsedLap] = (!trivial && n_high > 0) ? 1 : 0;
        st.hdr[kOHdrMatched] = n_match; st.hdr[kOHdrLeftDets] = n_left_d; st.hdr[kOHdrLeftTrks] = n_left_t;
        st.hdr[kOHdrRematched] = n_rematch; st.hdr[kOHdrSpawned] = n_new;
    }
    __syncthreads();
}

template <int CAP, int DMAX, int ECAP, bool DEEP, int ASSO = 0>
__device__ __forceinline__ void oc_step_body(const OcArgs& a) {
    MOT_DYNAMIC_SMEM(smem);
    OcSmem sm;
    oc_carve(smem, CAP, DMAX, ECAP, sm);
    constexpr OcLayout L = OcLayout::make(CAP, DMAX);
    static_assert(lap_idle_bytes(DMAX, CAP, ECAP) >= jv_block_sbytes(DMAX + CAP), "the CTA-wide LAPJV's shared scratch must fit the sparse solver's idle arrays");
    static_assert(!DEEP || lap_idle_bytes(DMAX, CAP, ECAP) - lap_align16(sizeof(int) * (size_t)DMAX) - lap_align16(sizeof(int) * (size_t)CAP) >=
                               sizeof(float) * kDeepTileFloats, "DeepOC-SORT needs room for at least one pair of summation tiles");
    const size_t stride = DEEP ? a.stride : L.stride;
    for (int s = a.s_begin + (int)blockIdx.x; s < a.s_end; s += (int)gridDim.x) {
        unsigned char* base = a.state + (size_t)s * stride;
        OcStream st = OcStream::at(base, L);
        DeepStream dp{};
        if constexpr (DEEP) dp = DeepStream::at(base + L.stride, CAP, DMAX, a.dim);
        lap_carve_gscratch(st.gscratch, DMAX, CAP, sm.lap);
        for (int t = 0; t < a.T; ++t) {
            const size_t fs = (size_t)t * a.S + s;
            const float* embs = (DEEP && a.embs) ? a.embs + fs * (size_t)a.ld_dets * (size_t)a.dim : nullptr;
            oc_frame<CAP, DMAX, DEEP, ASSO>(a, st, sm, a.dets + fs * (size_t)a.ld_dets * 6, a.n_dets[fs],
                                      a.out + fs * (size_t)a.ld_out * 8, a.n_out + fs, dp, embs);
        }
    }
}

template <int CAP, int DMAX, int ECAP>
__global__ void __launch_bounds__(kOcThreads) ocsort_step_kernel(OcArgs a) { oc_step_body<CAP, DMAX, ECAP, false>(a); }

// OC-SORT with asso_func = "centroid" (iou.hpp:298-330): the same frame step over 1 - centre distance / frame diagonal
template <int CAP, int DMAX, int ECAP>
__global__ void __launch_bounds__(kOcThreads) ocsort_centroid_step_kernel(OcArgs a) { oc_step_body<CAP, DMAX, ECAP, false, kVarCentroid>(a); }

// DeepOC-SORT (reference src/trackers/deepocsort.cpp:589-944, associate :348-504): the OC-SORT frame step without the
// BYTE pass, plus the appearance term of the first association, the embedding EMA of every updated track and the
// reference's twice-listed leftovers (see oc_frame).  Camera-motion compensation is outside the hot path (cmc_off).
template <int CAP, int DMAX, int ECAP>
__global__ void __launch_bounds__(kOcThreads) deepocsort_step_kernel(OcArgs a) { oc_step_body<CAP, DMAX, ECAP, true>(a); }

static __global__ void ocsort_reset_kernel(unsigned char* state, OcLayout L, size_t stride, int S, int keep_id_counter) {
    for (int s = (int)blockIdx.x; s < S; s += (int)gridDim.x) {
        OcStream st = OcStream::at(state + (size_t)s * stride, L);
        for (int k = (int)threadIdx.x; k < L.cap; k += (int)blockDim.x) st.freel[k] = (unsigned short)(L.cap - 1 - k);
        if (threadIdx.x == 0) {
            const int idc = keep_id_counter ? st.hdr[kOHdrIdCounter] : 0;
            for (int k = 0; k < 16; ++k) st.hdr[k] = 0;
            st.hdr[kOHdrFree] = L.cap;
            st.hdr[kOHdrIdCounter] = idc;
        }
        __syncthreads();
    }
}

// The (track capacity, detections per frame, candidate-edge buffer) shapes the OC-SORT kernel is built for
struct OcShape { int cap, d_max, e_cap; };
constexpr OcShape kOcShapes[] = {{256, 64, 1024}, {1536, 512, 4096}, {3072, 2048, 4096}};
constexpr int kNumOcShapes = sizeof(kOcShapes) / sizeof(kOcShapes[0]);

}  // namespace mot
