// bytetrack_kernel.cuh - ByteTrack's whole per-frame update() as ONE kernel: one CTA per camera
// stream, tracker state resident in HBM/L2, every list operation a block-wide stable compaction.
//
// Replaces, for one stream and one frame (reference src/trackers/bytetrack.cpp:166-621):
//   confidence split (:185-213)            -> phase A
//   tracked/unconfirmed split + pool (:228-265, STrack::multi_predict :97-116) -> phases B, C
//   1st association iou_distance+fuse_score+linear_assignment (:267-365)       -> phases D, E
//   2nd association on un-predicted boxes (:367-442)                           -> phase F
//   unconfirmed association (:448-542)                                         -> phase G
//   new tracks (:546-554), lost expiry (:557-562)                              -> phases H, I
//   joint/sub list algebra (:565-578, :623-657)                                -> phase J
//   remove_duplicate_stracks (:581-585, :659-706)                              -> phase K
//   output (:589-620)                                                          -> phase L
// All reference quirks listed in SURVEY.md section 8 "parity traps" are kept (predictions are
// written back only for matched tracks, etc.).  IDs come from a per-stream counter.
#pragma once
#include "shapes.cuh"
#include "block_utils.cuh"
#include "cost_device.cuh"
#include "kf_device.cuh"
#include "lap_device.cuh"

namespace mot {

// 512 threads x 2 CTAs per SM (64 registers/thread, 32 resident warps) measured fastest on B200:
// 256x1 0.86 M frames/s, 256x2 1.23 M, 384x2 1.43 M, 512x2 1.52 M (profiles/README.md).
#ifndef MOT_BT_THREADS
#define MOT_BT_THREADS 512
#endif
#ifndef MOT_BT_MINBLOCKS
#define MOT_BT_MINBLOCKS 2
#endif
constexpr int kBtThreads = MOT_BT_THREADS;

// per-track Kalman record: the compact independent-coordinate form (kf_device.cuh), 96 B instead of 288 B - the whole
// state of 296 C2 streams (80 MB) then stays resident in the 126 MB L2 instead of cycling through HBM
constexpr int kBtRecFloats = kRecFloatsCompact;

enum : int { kStNew = 0, kStTracked = 1, kStLost = 2, kStRemoved = 3 };
constexpr unsigned char kFlagActivated = 0x10;

enum : int {              // header slots (ints) of one stream
    kHdrActive = 0, kHdrLost = 1, kHdrFree = 2, kHdrIdCounter = 3, kHdrFrame = 4, kHdrError = 5,
    kHdrN1 = 6, kHdrM1 = 7, kHdrN2 = 8, kHdrM2 = 9, kHdrN3 = 10, kHdrM3 = 11, kHdrDupA = 12, kHdrDupB = 13,
    kHdrInts = 16
};
enum : int { kErrCapacity = 1, kErrTooManyDets = 2, kErrOutput = 4, kErrKalman = 8 };

struct BtParams {
    float min_conf, track_thresh, match_thresh, det_thresh;
    int max_time_lost;
};

// ---- per-stream state layout in global memory (one contiguous slab per stream)
struct BtLayout {
    int cap, d_max;
    size_t off_lists, off_sflag, off_meta, off_recs, off_gscratch, stride;
    MOT_HD static constexpr size_t al(size_t x) { return (x + 255) & ~(size_t)255; }
    MOT_HD static constexpr BtLayout make(int cap, int d_max) {
        BtLayout L{};
        L.cap = cap; L.d_max = d_max;
        size_t o = al(sizeof(int) * kHdrInts);
        L.off_lists = o;    o = al(o + sizeof(unsigned short) * 3 * (size_t)cap);
        L.off_sflag = o;    o = al(o + (size_t)cap);
        L.off_meta = o;     o = al(o + sizeof(int) * 7 * (size_t)cap);
        L.off_recs = o;     o = al(o + sizeof(float) * kBtRecFloats * (size_t)cap);
        L.off_gscratch = o; o = al(o + lap_gscratch_bytes(cap, d_max));
        L.stride = o;
        return L;
    }
};

struct BtStream {
    int* hdr;
    unsigned short *active, *lost, *freel;
    unsigned char* sflag;
    int *id, *tracklet_len, *frame_id, *start_frame, *cls, *det_ind;
    float* conf;
    float* recs;
    unsigned char* gscratch;
    __device__ __forceinline__ static BtStream at(unsigned char* base, const BtLayout& L) {
        BtStream s;
        s.hdr = (int*)base;
        s.active = (unsigned short*)(base + L.off_lists);
        s.lost = s.active + L.cap;
        s.freel = s.lost + L.cap;
        s.sflag = base + L.off_sflag;
        int* m = (int*)(base + L.off_meta);
        s.id = m; s.tracklet_len = m + L.cap; s.frame_id = m + 2 * L.cap; s.start_frame = m + 3 * L.cap;
        s.cls = m + 4 * L.cap; s.det_ind = m + 5 * L.cap; s.conf = (float*)(m + 6 * L.cap);
        s.recs = (float*)(base + L.off_recs);
        s.gscratch = base + L.off_gscratch;
        return s;
    }
};

struct BtArgs {
    unsigned char* state;     // [S] slabs of layout.stride bytes
    const float* dets;        // [T][S][ld_dets][6]
    const int* n_dets;        // [T][S]
    float* out;               // [T][S][ld_out][8]
    int* n_out;               // [T][S]
    int T, S, ld_dets, ld_out, e_cap;
    int s_begin, s_end;       // streams handled by this launch
    BtParams p;
    unsigned long long* prof; // optional [16] per-phase cycle counters (mot_engine_profile), nullptr = off
};

// ---- shared-memory plan
struct BtSmem {
    float4* det_box;            // [d_max] IoU box of every detection (xyxy -> xywh -> xyxy, as STrack does)
    float* det_conf;            // [d_max]
    unsigned short* hi;         // [d_max] detections with conf > track_thresh
    unsigned short* lo;         // [d_max] min_conf < conf < track_thresh
    unsigned short* udet;       // [d_max] hi detections left after the 1st association (det indices)
    unsigned short* udet2;      // [d_max] ... after the unconfirmed association
    float4* row_box;            // [cap]
    unsigned short* pool;       // [cap] slot of every pool row (tracked ++ lost)
    unsigned short* unconf;     // [cap] slots of unconfirmed tracks
    unsigned short* sel;        // [cap] generic selection list (matched rows, r_tracked rows, ...)
    unsigned short* list_a;     // [cap] next active list
    unsigned short* list_b;     // [cap] next lost list
    unsigned short* list_c;     // [cap] lost-this-frame / freed slots
    unsigned char* dup_a;       // [cap]
    unsigned char* dup_b;       // [cap]
    BlockScratch* bs;
    LapWorkspace lap;
};

MOT_HD constexpr size_t bt_smem_bytes(int cap, int d_max, int e_cap) {
    size_t b = 0;
    b += lap_align16(sizeof(float4) * (size_t)d_max);
    b += lap_align16(sizeof(float) * (size_t)d_max);
    b += 4 * lap_align16(sizeof(unsigned short) * (size_t)d_max);
    b += lap_align16(sizeof(float4) * (size_t)cap);
    b += 6 * lap_align16(sizeof(unsigned short) * (size_t)cap);
    b += lap_align16(sizeof(BlockScratch));
    b += lap_smem_bytes(cap, d_max, e_cap);
    return b;
}

__device__ __forceinline__ void bt_carve(unsigned char* p, int cap, int d_max, int e_cap, BtSmem& s) {
    s.det_box = (float4*)p;            p += lap_align16(sizeof(float4) * (size_t)d_max);
    s.det_conf = (float*)p;            p += lap_align16(sizeof(float) * (size_t)d_max);
    s.hi = (unsigned short*)p;         p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.lo = (unsigned short*)p;         p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.udet = (unsigned short*)p;       p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.udet2 = (unsigned short*)p;      p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.row_box = (float4*)p;            p += lap_align16(sizeof(float4) * (size_t)cap);
    s.pool = (unsigned short*)p;       p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.unconf = (unsigned short*)p;     p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.sel = (unsigned short*)p;        p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.list_a = (unsigned short*)p;     p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.list_b = (unsigned short*)p;     p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.list_c = (unsigned short*)p;     p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.bs = (BlockScratch*)p;           p += lap_align16(sizeof(BlockScratch));
    lap_carve(p, cap, d_max, e_cap, s.lap);
    s.dup_a = (unsigned char*)s.lap.row_label;      // the assignment workspace is idle during duplicate removal
    s.dup_b = s.dup_a + cap;
}

// IoU box of a track from its CURRENT mean (STrack::xyxy, bytetrack.cpp:118-128)
__device__ __forceinline__ float4 bt_track_box(const float* rec) {
    const float4 m = *reinterpret_cast<const float4*>(rec);
    return xyah2xyxy(m.x, m.y, m.z, m.w);
}

// raw detection row -> measurement [xc, yc, a, h] (STrack ctor, bytetrack.cpp:26-29)
__device__ __forceinline__ void bt_det_xyah(const float* det_row, float (&z)[4]) {
    const float4 q = xywh2xyah_via_tlwh(xyxy2xywh(make_float4(det_row[0], det_row[1], det_row[2], det_row[3])));
    z[0] = q.x; z[1] = q.y; z[2] = q.z; z[3] = q.w;
}

// Kalman work for a list of (track slot, detection) pairs: one thread per (pair, coordinate) - four consecutive lanes
// advance the four independent (position, velocity) filters of a track (kf_device.cuh, "independent-coordinate form").
//   mode 0: predict (zeroing vh unless Tracked) then update   - 1st association
//   mode 1: predict then update                               - 2nd association (always Tracked)
//   mode 2: update only                                       - unconfirmed association
// slot_of(k) / det_of(k) give the k-th pair; the quad's lane 0 also refreshes the track's metadata
// (STrack::update / re_activate, bytetrack.cpp:51-85).
template <class SlotOf, class DetOf>
__device__ __forceinline__ void bt_kalman_pairs(const BtStream& st, const float* dets, int n_pairs, int mode,
                                                int frame, SlotOf slot_of, DetOf det_of) {
    const int lane = lane_id(), c = lane & 3, qbase = lane & ~3;
    const int quads = (int)(blockDim.x >> 2);
    const int qid = (int)(threadIdx.x >> 2);
    const int rounds = (n_pairs + quads - 1) / quads;
    for (int it = 0; it < rounds; ++it) {
        const int k = it * quads + qid;
        const bool live = k < n_pairs;
        const int slot = live ? slot_of(k) : 0;
        const int det = live ? det_of(k) : 0;
        float* rec = st.recs + (size_t)slot * kBtRecFloats;
        KfBlock s;
        if (live) kfb_load(rec, c, s);
        else { s.mc = 1.0f; s.mv = 0.0f; s.pcc = 1.0f; s.pcv = 0.0f; s.pvc = 0.0f; s.pvv = 1.0f; }
        const int state = live ? (int)(st.sflag[slot] & 0x0f) : kStTracked;
        float z[4] = {0.0f, 0.0f, 0.0f, 1.0f};
        if (live) bt_det_xyah(dets + (size_t)det * 6, z);
        const float zc = (c == 0) ? z[0] : (c == 1) ? z[1] : (c == 2) ? z[2] : z[3];
        if (mode != 2) {
            const float h0 = __shfl_sync(kFullMask, s.mc, qbase + 3);            // mean(3) before the motion step
            kfb_xyah_predict(s, c, h0, mode == 0 && state != kStTracked);
        }
        const float h1 = __shfl_sync(kFullMask, s.mc, qbase + 3);                // mean(3) of the predicted state
        const bool okc = kfb_xyah_update(s, c, h1, zc, 0.0f);
        const unsigned bad = __ballot_sync(kFullMask, !okc);
        const bool ok = ((bad >> qbase) & 0xfu) == 0;                             // all four pivots positive
        if (live) {
            if (ok) kfb_store(rec, c, s);
            if (c == 0) {
                if (!ok) atomicOr(&st.hdr[kHdrError], (int)kErrKalman);
                if (state == kStTracked) st.tracklet_len[slot] += 1;     // STrack::update
                else st.tracklet_len[slot] = 0;                          // STrack::re_activate
                st.sflag[slot] = (unsigned char)(kStTracked | kFlagActivated);
                st.frame_id[slot] = frame;
                st.conf[slot] = dets[(size_t)det * 6 + 4];
                st.cls[slot] = (int)dets[(size_t)det * 6 + 5];
                st.det_ind[slot] = det;
            }
        }
    }
}

template <int CAP, int DMAX>
__device__ __forceinline__ void bt_frame(const BtArgs& a, const BtStream& st, BtSmem& sm, const float* dets, int n_det_in,
                                         float* out, int* n_out) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
    constexpr int cap = CAP, d_max = DMAX;
    const int lap_m_max = d_max;
    __syncthreads();
    PhaseClock clk;
    clk.start(a.prof);
    const int frame = st.hdr[kHdrFrame] + 1;                  // frame_count_ == frame_id_ (:181-182)
    const int n_active = st.hdr[kHdrActive], n_lost = st.hdr[kHdrLost];
    int n_free = st.hdr[kHdrFree];
    const int id_base = st.hdr[kHdrIdCounter];
    int n_det = n_det_in;
    if (n_det > min(d_max, a.ld_dets)) { n_det = min(d_max, a.ld_dets); if (tid == 0) atomicOr(&st.hdr[kHdrError], (int)kErrTooManyDets); }

    // ---- A. detections: IoU boxes, confidence split
    int conf_above_one = 0;
    for (int j = tid; j < n_det; j += nt) {
        const float* r = dets + (size_t)j * 6;
        sm.det_box[j] = xywh2xyxy(xyxy2xywh(make_float4(r[0], r[1], r[2], r[3])));
        sm.det_conf[j] = r[4];
        conf_above_one |= !(r[4] <= 1.0f);
    }
    conf_above_one = __syncthreads_or(conf_above_one);      // fuse_score with a confidence > 1 voids the IoU floor below
    const float t_hi = a.p.track_thresh, t_lo = a.p.min_conf;
    int n_hi = 0, n_lo = 0;
    block_compact2(n_det, 0, 0, sm.bs, [&](int j) { return sm.det_conf[j] > t_hi; },
                   [&](int j) { const float c = sm.det_conf[j]; return c > t_lo && c < t_hi; },
                   [&](int j, int pos) { sm.hi[pos] = (unsigned short)j; }, [&](int j, int pos) { sm.lo[pos] = (unsigned short)j; },
                   n_hi, n_lo);

    clk.tick(0);
    // ---- B. pool = tracked (activated) ++ lost ; unconfirmed kept aside
    int n_trk = 0, n_unc = 0;
    block_compact2(n_active, 0, 0, sm.bs, [&](int k) { return (st.sflag[st.active[k]] & kFlagActivated) != 0; },
                   [&](int k) { return (st.sflag[st.active[k]] & kFlagActivated) == 0; },
                   [&](int k, int pos) { sm.pool[pos] = st.active[k]; }, [&](int k, int pos) { sm.unconf[pos] = st.active[k]; },
                   n_trk, n_unc);
    for (int k = tid; k < n_lost; k += nt) sm.pool[n_trk + k] = st.lost[k];
    const int n1 = n_trk + n_lost;
    __syncthreads();

    clk.tick(1);
    // ---- C. predicted box of every pool row (the prediction itself is redone in registers for the
    //         rows that get matched: only the mean is needed to build costs)
    for (int r = tid; r < n1; r += nt) {
        const int slot = sm.pool[r];
        const float* rec = st.recs + (size_t)slot * kBtRecFloats;
        const float4 m = *reinterpret_cast<const float4*>(rec);
        const float4 v = *reinterpret_cast<const float4*>(rec + 4);
        const float vh = ((st.sflag[slot] & 0x0f) != kStTracked) ? 0.0f : v.w;
        sm.row_box[r] = xyah2xyxy(xadd(m.x, v.x), xadd(m.y, v.y), xadd(m.z, v.z), xadd(m.w, vh));
    }
    __syncthreads();

    clk.tick(2);
    // ---- D. first association
    {
        // 1 - iou * conf <= match_thresh needs iou >= 1 - match_thresh (conf <= 1); the floor sits a relative 1e-3 under it
        const float floor1 = conf_above_one ? 0.0f : 1.0f - a.p.match_thresh * 1.001f - 1e-5f;
        IouCost cost{sm.row_box, sm.det_box, sm.det_conf, sm.hi, true, a.p.match_thresh < 1.0f, floor1};
        sm.lap.clk = a.prof ? &clk : nullptr;
        sm.lap.clk_base = 3;
        block_lap(sm.lap, n1, n_hi, cap, lap_m_max, a.p.match_thresh, cost);
        sm.lap.clk = nullptr;
    }
    // harvest what later phases need before the LAP workspace is reused
    // matched rows, and r_tracked = unmatched pool rows that came from the active list (state Tracked): one pass
    int n_m1 = 0, n2 = 0;
    block_compact2(n1, 0, 0, sm.bs, [&](int r) { return sm.lap.row2col[r] >= 0; },
                   [&](int r) { return r < n_trk && sm.lap.row2col[r] < 0; },
                   [&](int r, int pos) { sm.sel[pos] = (unsigned short)r; }, [&](int r, int pos) { sm.list_a[pos] = sm.pool[r]; },
                   n_m1, n2);
    // matched rows: stash the detection index next to the row (list_c is free until phase F)
    for (int k = tid; k < n_m1; k += nt) sm.list_c[k] = sm.hi[sm.lap.row2col[sm.sel[k]]];
    const int n_udet = block_compact(n_hi, 0, sm.bs, [&](int j) { return sm.lap.col2row[j] < 0; },
                                     [&](int j, int pos) { sm.udet[pos] = sm.hi[j]; });
    __syncthreads();

    clk.tick(7);
    // ---- E. Kalman predict + update for the matches of the first association
    bt_kalman_pairs(st, dets, n_m1, 0, frame, [&](int k) { return (int)sm.pool[sm.sel[k]]; },
                    [&](int k) { return (int)sm.list_c[k]; });
    __syncthreads();

    clk.tick(8);
    // ---- F. second association: r_tracked (slots in list_a) x low-confidence detections
    int n_lost_new = 0;
    if (n2 > 0 && n_lo > 0) {
        for (int i = tid; i < n2; i += nt) sm.row_box[i] = bt_track_box(st.recs + (size_t)sm.list_a[i] * kBtRecFloats);
        __syncthreads();
        IouCost cost{sm.row_box, sm.det_box, sm.det_conf, sm.lo, false, true};
        sm.lap.clk = a.prof ? &clk : nullptr;
        sm.lap.clk_base = 20;
        block_lap(sm.lap, n2, n_lo, cap, lap_m_max, 0.5f, cost);
        sm.lap.clk = nullptr;
        const int n_m2 = block_compact(n2, 0, sm.bs, [&](int i) { return sm.lap.row2col[i] >= 0; },
                                       [&](int i, int pos) { sm.sel[pos] = (unsigned short)i; });
        for (int k = tid; k < n_m2; k += nt) sm.list_c[k] = sm.lo[sm.lap.row2col[sm.sel[k]]];
        __syncthreads();
        bt_kalman_pairs(st, dets, n_m2, 1, frame, [&](int k) { return (int)sm.list_a[sm.sel[k]]; },
                        [&](int k) { return (int)sm.list_c[k]; });
        // unmatched -> Lost (:435-441), appended to the lost list in row order
        n_lost_new = block_compact(n2, 0, sm.bs, [&](int i) { return sm.lap.row2col[i] < 0; },
                                   [&](int i, int pos) {
                                       const int slot = sm.list_a[i];
                                       sm.list_c[pos] = (unsigned short)slot;
                                       st.sflag[slot] = (unsigned char)((st.sflag[slot] & 0xf0) | kStLost);
                                   });
    }
    __syncthreads();
    // list_c[0 .. n_lost_new) now holds the slots that became Lost this frame; keep it until phase J.

    clk.tick(9);
    // ---- G. unconfirmed tracks x leftover high detections
    int n_final = n_udet;
    const unsigned short* final_list = sm.udet;
    if (n_unc > 0 && n_udet > 0) {
        for (int i = tid; i < n_unc; i += nt) sm.row_box[i] = bt_track_box(st.recs + (size_t)sm.unconf[i] * kBtRecFloats);
        __syncthreads();
        IouCost cost{sm.row_box, sm.det_box, sm.det_conf, sm.udet, true, true};
        sm.lap.clk = a.prof ? &clk : nullptr;
        sm.lap.clk_base = 24;
        block_lap(sm.lap, n_unc, n_udet, cap, lap_m_max, 0.7f, cost);
        sm.lap.clk = nullptr;
        const int n_m3 = block_compact(n_unc, 0, sm.bs, [&](int i) { return sm.lap.row2col[i] >= 0; },
                                       [&](int i, int pos) { sm.sel[pos] = (unsigned short)i; });
        for (int k = tid; k < n_m3; k += nt) sm.list_a[k] = sm.udet[sm.lap.row2col[sm.sel[k]]];
        n_final = block_compact(n_udet, 0, sm.bs, [&](int j) { return sm.lap.col2row[j] < 0; },
                                [&](int j, int pos) { sm.udet2[pos] = sm.udet[j]; });
        final_list = sm.udet2;
        // unmatched unconfirmed tracks are removed (:533-538)
        for (int i = tid; i < n_unc; i += nt)
            if (sm.lap.row2col[i] < 0) st.sflag[sm.unconf[i]] = (unsigned char)kStRemoved;
        __syncthreads();
        bt_kalman_pairs(st, dets, n_m3, 2, frame, [&](int k) { return (int)sm.unconf[sm.sel[k]]; },
                        [&](int k) { return (int)sm.list_a[k]; });
    }
    __syncthreads();

    clk.tick(10);
    // ---- H. new tracks from the remaining high detections, IDs in list order (:546-554)
    const float det_thresh = a.p.det_thresh;
    const int n_new_want = block_compact(n_final, 0, sm.bs, [&](int k) { return sm.det_conf[final_list[k]] >= det_thresh; },
                                         [&](int k, int pos) { sm.sel[pos] = final_list[k]; });
    int n_new = n_new_want;
    if (n_new > n_free) { n_new = n_free; if (tid == 0) atomicOr(&st.hdr[kHdrError], (int)kErrCapacity); }
    {
        const int c = tid & 3;
        const int quads = nt >> 2, qid = tid >> 2;
        for (int k = qid; k < n_new; k += quads) {
            const int det = sm.sel[k];
            const int slot = st.freel[n_free - 1 - k];
            float z[4];
            bt_det_xyah(dets + (size_t)det * 6, z);
            const float zc = (c == 0) ? z[0] : (c == 1) ? z[1] : (c == 2) ? z[2] : z[3];
            KfBlock s;
            kfb_xyah_initiate(s, c, zc, z[3]);
            kfb_store(st.recs + (size_t)slot * kBtRecFloats, c, s);
            __syncwarp(0xfu << (lane_id() & ~3));                     // the quad has read sel[k]
            if (c == 0) {
                sm.sel[k] = (unsigned short)slot;                    // phase J appends the new slots from here
                st.id[slot] = id_base + 1 + k;
                st.sflag[slot] = (unsigned char)(kStTracked | (frame == 1 ? kFlagActivated : 0));
                st.tracklet_len[slot] = 0;
                st.frame_id[slot] = frame;
                st.start_frame[slot] = frame;
                st.conf[slot] = dets[(size_t)det * 6 + 4];
                st.cls[slot] = (int)dets[(size_t)det * 6 + 5];
                st.det_ind[slot] = det;
            }
        }
    }
    __syncthreads();

    clk.tick(11);
    // ---- I. expire lost tracks (:557-562); re-found ones are Tracked by now and skipped
    for (int k = tid; k < n_lost; k += nt) {
        const int slot = st.lost[k];
        if ((st.sflag[slot] & 0x0f) == kStLost && frame - st.frame_id[slot] > a.p.max_time_lost)
            st.sflag[slot] = (unsigned char)((st.sflag[slot] & 0xf0) | kStRemoved);
    }
    __syncthreads();

    // ---- J. next lists.  active' = kept active ++ new ++ re-found ; lost' = kept lost ++ lost-this-frame
    // Slots that died this frame (removed unconfirmed tracks, expired lost tracks) go back on the free stack, on top of
    // what is left after the new tracks took theirs (sel[] holds the new tracks' slots since phase H).
    int na = 0, nl = 0;
    n_free -= n_new;
    block_compact2(n_active, 0, n_free, sm.bs, [&](int k) { return (st.sflag[st.active[k]] & 0x0f) == kStTracked; },
                   [&](int k) { return (st.sflag[st.active[k]] & 0x0f) == kStRemoved; },
                   [&](int k, int pos) { sm.list_a[pos] = st.active[k]; }, [&](int k, int pos) { st.freel[pos] = st.active[k]; },
                   na, n_free);
    for (int k = tid; k < n_new; k += nt) sm.list_a[na + k] = sm.sel[k];
    na += n_new;
    block_compact2(n_lost, na, 0, sm.bs, [&](int k) { return (st.sflag[st.lost[k]] & 0x0f) == kStTracked; },
                   [&](int k) { return (st.sflag[st.lost[k]] & 0x0f) == kStLost; },
                   [&](int k, int pos) { sm.list_a[pos] = st.lost[k]; }, [&](int k, int pos) { sm.list_b[pos] = st.lost[k]; },
                   na, nl);
    for (int k = tid; k < n_lost_new; k += nt) sm.list_b[nl + k] = sm.list_c[k];
    nl += n_lost_new;
    n_free = block_compact(n_lost, n_free, sm.bs, [&](int k) { return (st.sflag[st.lost[k]] & 0x0f) == kStRemoved; },
                           [&](int k, int pos) { st.freel[pos] = st.lost[k]; });

    clk.tick(12);
    // ---- K. remove_duplicate_stracks(active', lost') (:659-706)
    __syncthreads();                       // list_a / list_b tails were written without a barrier when the lost list was empty
    for (int i = tid; i < na; i += nt) sm.dup_a[i] = 0;
    for (int j = tid; j < nl; j += nt) sm.dup_b[j] = 0;
    if (na > 0 && nl > 0) {
        for (int i = tid; i < na; i += nt) sm.row_box[i] = bt_track_box(st.recs + (size_t)sm.list_a[i] * kBtRecFloats);
        // the lost boxes share row_box's tail when they fit, else they are recomputed from global
        const bool fits = na + nl <= cap;
        if (fits)
            for (int j = tid; j < nl; j += nt) sm.row_box[na + j] = bt_track_box(st.recs + (size_t)sm.list_b[j] * kBtRecFloats);
        __syncthreads();
        auto mark = [&](int i, int j, float4 ba, float area, float4 bb) {
            const float pd = xsub(1.0f, iou_pair(ba, area, bb));
            if (pd < 0.15f) {
                const int sa = sm.list_a[i], sb = sm.list_b[j];
                const int tp = st.frame_id[sa] - st.start_frame[sa];
                const int tq = st.frame_id[sb] - st.start_frame[sb];
                if (tp > tq) sm.dup_b[j] = 1; else sm.dup_a[i] = 1;
            }
        };
        if (fits && (long long)na * nl >= 8192) {
            // lost boxes in the grid.  A duplicate needs 1 - IoU < 0.15: only boxes whose corner lies within 16 % of a box
            // size can qualify, so a row meets 0-2 of them and they are judged on the spot (collecting the pairs first -
            // count, scan, write, judge densely - walked the grid twice for pair lists this short).
            grid_build(sm.lap.grid, nl, sm.bs, [&](int j) { return sm.row_box[na + j]; });
            for (int i = tid; i < na; i += nt) {
                const float4 ba = sm.row_box[i];
                const float area = box_area(ba);
                grid_query_iou_above(sm.lap.grid, ba, 0.84f, [&](int j) { return sm.row_box[na + j]; },
                                     [&](int j, float4 bb) { mark(i, j, ba, area, bb); });
            }
        } else {
            for (int i = tid; i < na; i += nt) {
                const float4 ba = sm.row_box[i];
                const float area = box_area(ba);
                for (int j = 0; j < nl; ++j) {
                    const float4 bb = fits ? sm.row_box[na + j] : bt_track_box(st.recs + (size_t)sm.list_b[j] * kBtRecFloats);
                    if (boxes_disjoint(ba, bb)) continue;                  // distance exactly 1
                    mark(i, j, ba, area, bb);
                }
            }
        }
    }
    __syncthreads();
    int na2 = 0, nl2 = 0;
    block_compact2(na, 0, n_free, sm.bs, [&](int i) { return sm.dup_a[i] == 0; }, [&](int i) { return sm.dup_a[i] != 0; },
                   [&](int i, int pos) { st.active[pos] = sm.list_a[i]; }, [&](int i, int pos) { st.freel[pos] = sm.list_a[i]; },
                   na2, n_free);
    block_compact2(nl, 0, n_free, sm.bs, [&](int j) { return sm.dup_b[j] == 0; }, [&](int j) { return sm.dup_b[j] != 0; },
                   [&](int j, int pos) { st.lost[pos] = sm.list_b[j]; }, [&](int j, int pos) { st.freel[pos] = sm.list_b[j]; },
                   nl2, n_free);
    __syncthreads();

    clk.tick(13);
    // ---- L. output rows for activated tracks, in list order (:589-620)
    const int n_rows = block_compact(na2, 0, sm.bs, [&](int i) { return (st.sflag[st.active[i]] & kFlagActivated) != 0; },
                                     [&](int i, int pos) {
                                         if (pos >= a.ld_out) return;
                                         const int slot = st.active[i];
                                         const float4 b = bt_track_box(st.recs + (size_t)slot * kBtRecFloats);
                                         float* o = out + (size_t)pos * 8;
                                         *reinterpret_cast<float4*>(o) = b;
                                         *reinterpret_cast<float4*>(o + 4) =
                                             make_float4((float)st.id[slot], st.conf[slot], (float)st.cls[slot],
                                                         (float)st.det_ind[slot]);
                                     });
    if (tid == 0) {
        if (n_rows > a.ld_out) atomicOr(&st.hdr[kHdrError], (int)kErrOutput);
        *n_out = n_rows < a.ld_out ? n_rows : a.ld_out;
        st.hdr[kHdrActive] = na2;
        st.hdr[kHdrLost] = nl2;
        st.hdr[kHdrFree] = n_free;
        st.hdr[kHdrIdCounter] = id_base + n_new;
        st.hdr[kHdrFrame] = frame;
        st.hdr[kHdrN1] = n1; st.hdr[kHdrM1] = n_hi; st.hdr[kHdrN2] = n2; st.hdr[kHdrM2] = n_lo;
        st.hdr[kHdrN3] = n_unc; st.hdr[kHdrM3] = n_udet; st.hdr[kHdrDupA] = na; st.hdr[kHdrDupB] = nl;
    }
    __syncthreads();
    clk.tick(14);
}

// One CTA per stream; each CTA walks its streams' T frames in order (state stays hot in L1/L2).
// CAP / DMAX / ECAP are compile-time so that every shared-memory and state pointer is "base +
// constant" (no registers spent on the ~50 pointers of the carve-up).
// THREADS = kBtThreads (two CTAs per SM) when there are streams to fill the machine; kBtThreadsWide (one 1024-thread CTA
// per SM) when there are fewer streams than SMs (BASELINE configs[4]: 8 streams per GPU) - a frame's phases are then
// latency-bound per stream and twice the threads shorten the parallel ones (one pass instead of two over the pool).
constexpr int kBtThreadsWide = 1024;
template <int CAP, int DMAX, int ECAP, int THREADS = kBtThreads>
__global__ void __launch_bounds__(THREADS, (THREADS > 512) ? 1 : MOT_BT_MINBLOCKS) bytetrack_step_kernel(BtArgs a) {
    MOT_DYNAMIC_SMEM(smem);
    BtSmem sm;
    bt_carve(smem, CAP, DMAX, ECAP, sm);
    constexpr BtLayout L = BtLayout::make(CAP, DMAX);
    for (int s = a.s_begin + (int)blockIdx.x; s < a.s_end; s += (int)gridDim.x) {
        BtStream st = BtStream::at(a.state + (size_t)s * L.stride, L);
        lap_carve_gscratch(st.gscratch, CAP, DMAX, sm.lap);
        for (int t = 0; t < a.T; ++t) {
            const size_t fs = (size_t)t * a.S + s;
            if (t + 1 < a.T) {                       // the next frame's detections are fresh HBM lines: start them towards L2 now
                const char* nx = reinterpret_cast<const char*>(a.dets + (fs + a.S) * (size_t)a.ld_dets * 6);
                const int lines = (min(a.ld_dets, DMAX) * 24 + 127) >> 7;
                if ((int)threadIdx.x < lines) prefetch_l2(nx + ((size_t)threadIdx.x << 7));
                if ((int)threadIdx.x == lines) prefetch_l2(a.n_dets + fs + a.S);
            }
            bt_frame<CAP, DMAX>(a, st, sm, a.dets + fs * (size_t)a.ld_dets * 6, a.n_dets[fs],
                                a.out + fs * (size_t)a.ld_out * 8, a.n_out + fs);
        }
    }
}


// reset / first-time initialisation of the per-stream slabs
static __global__ void bytetrack_reset_kernel(unsigned char* state, BtLayout L, int S, int keep_id_counter) {
    for (int s = (int)blockIdx.x; s < S; s += (int)gridDim.x) {
        BtStream st = BtStream::at(state + (size_t)s * L.stride, L);
        for (int k = (int)threadIdx.x; k < L.cap; k += (int)blockDim.x) {
            st.freel[k] = (unsigned short)(L.cap - 1 - k);
            st.sflag[k] = (unsigned char)kStRemoved;
        }
        if (threadIdx.x == 0) {
            const int idc = keep_id_counter ? st.hdr[kHdrIdCounter] : 0;
            for (int k = 0; k < kHdrInts; ++k) st.hdr[k] = 0;
            st.hdr[kHdrFree] = L.cap;
            st.hdr[kHdrIdCounter] = idc;
        }
        __syncthreads();
    }
}

}  // namespace mot
