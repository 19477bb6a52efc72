// boosttrack_kernel.cuh - BoostTrack's whole per-frame update() as one kernel, one CTA per camera stream (SURVEY 8f-1,
// second half).  Replaces reference src/trackers/boosttrack.cpp:465-699 (BoostTrackTracker::update) with its default
// options - camera-motion compensation and ReID are image processing outside the hot path (use_ecc / with_reid off),
// use_sb (a powf) is refused by the host side:
//   predict every track (:497-513, BoostKalmanFilter::predict :56-59, BoostTrack::predict :156-163)   -> phase B
//   detection-confidence boost: conf = max(conf, max_iou * dlo_boost_coef), or the use_vt rule (:361-426) -> phase C
//   filter by det_thresh (:532-538)                                                                     -> phase C
//   cost = (1 - iou) - lambda_mhd * mh_sim (:297-358, :588-603), one linear_assignment (:614)           -> phase D
//   Kalman update of the matches (:646-654, :61-75), new tracks (:657-666)                              -> phases E, F
//   output rows in track order through filter_outputs (:669-698, :434-463), age-out (:685-689)          -> phases G, H
// Kalman state: [cx, cy, h, r | velocities] with F = [I I; 0 I], H = [I 0] and diagonal Q, R, P0: the covariance only ever
// couples a coordinate with its own velocity, so a record is the 8 means + four 2 x 2 blocks (24 floats) and every sum of
// the reference's dense Eigen expressions has one non-zero term (two in the prediction) - bit-identical to the dense
// evaluation.  The 4 x 4 innovation covariance is diagonal; the reference's LU inverse of it (:68) is the reciprocal of the
// diagonal, and the gain MULTIPLIES by that reciprocal (it does not divide).
#pragma once
#include "shapes.cuh"
#include "block_utils.cuh"
#include "cost_device.cuh"
#include "lap_device.cuh"
#include "sort_kernel.cuh"      // SortLayout / SortStream / SortArgs-style slab: list, free stack, 7 meta arrays, 64-float records

namespace mot {

#ifndef MOT_BOOST_THREADS
#define MOT_BOOST_THREADS 512
#endif
constexpr int kBoostThreads = MOT_BOOST_THREADS;
constexpr int kBoostRecFloats = kSortRecFloats;      // 24 used: x 8 | (pcc, pcv, pvc, pvv) x 4

struct BoostParams {
    float det_thresh, iou_threshold, aspect_ratio_thresh, lambda_mhd, dlo_boost_coef, min_box_area;
    int max_age, min_hits, use_dlo_boost, use_vt;
};

struct BoostArgs {
    unsigned char* state;
    const float* dets;        // [T][S][ld_dets][6]
    const int* n_dets;        // [T][S]
    float* out;               // [T][S][ld_out][8]
    int* n_out;               // [T][S]
    int T, S, ld_dets, ld_out;
    int s_begin, s_end;
    BoostParams p;
};

struct BoostSmem {
    float4* det_box;            // [d_max] raw xyxy
    float* det_conf;            // [d_max] confidence, boosted in phase C
    unsigned short* valid;      // [d_max] detections with boosted conf >= det_thresh
    float4* trk_box;            // [cap] get_state() of every predicted track
    float4* trk_mean;           // [cap] x.head(4)
    float4* trk_inv;            // [cap] 1 / covariance.diagonal().head(4)
    unsigned short* sel;        // [cap]
    unsigned short* sel2;       // [d_max]
    unsigned char* tsu1;        // [cap] min(time_since_update - 1, 255) (use_vt)
    BlockScratch* bs;
    LapWorkspace lap;
};

MOT_HD constexpr size_t boost_smem_bytes(int cap, int d_max, int e_cap) {
    size_t b = 0;
    b += lap_align16(sizeof(float4) * (size_t)d_max);
    b += lap_align16(sizeof(float) * (size_t)d_max);
    b += 2 * lap_align16(sizeof(unsigned short) * (size_t)d_max);
    b += 3 * lap_align16(sizeof(float4) * (size_t)cap);
    b += lap_align16(sizeof(unsigned short) * (size_t)cap);
    b += lap_align16((size_t)cap);
    b += lap_align16(sizeof(BlockScratch));
    b += lap_smem_bytes(d_max, cap, e_cap);
    return b;
}

__device__ __forceinline__ void boost_carve(unsigned char* p, int cap, int d_max, int e_cap, BoostSmem& s) {
    s.det_box = (float4*)p;            p += lap_align16(sizeof(float4) * (size_t)d_max);
    s.det_conf = (float*)p;            p += lap_align16(sizeof(float) * (size_t)d_max);
    s.valid = (unsigned short*)p;      p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.sel2 = (unsigned short*)p;       p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.trk_box = (float4*)p;            p += lap_align16(sizeof(float4) * (size_t)cap);
    s.trk_mean = (float4*)p;           p += lap_align16(sizeof(float4) * (size_t)cap);
    s.trk_inv = (float4*)p;            p += lap_align16(sizeof(float4) * (size_t)cap);
    s.sel = (unsigned short*)p;        p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.tsu1 = p;                        p += lap_align16((size_t)cap);
    s.bs = (BlockScratch*)p;           p += lap_align16(sizeof(BlockScratch));
    lap_carve(p, d_max, cap, e_cap, s.lap);
}

// BoostKalmanFilter::get_state (:107-115)
__device__ __forceinline__ float4 boost_state_box(float cx, float cy, float h, float r) {
    const float w = xmul(r, h);
    return make_float4(xsub(cx, xdiv(w, 2.0f)), xsub(cy, xdiv(h, 2.0f)), xadd(cx, xdiv(w, 2.0f)), xadd(cy, xdiv(h, 2.0f)));
}
// convert_bbox_to_z (:127-134)
__device__ __forceinline__ float4 boost_bbox_to_z(float4 b) {
    const float w = xsub(b.z, b.x), h = xsub(b.w, b.y);
    return make_float4(xadd(b.x, xdiv(w, 2.0f)), xadd(b.y, xdiv(h, 2.0f)), h, (h > 1e-6f) ? xdiv(w, h) : 0.0f);
}

// cost functor for block_lap(): rows = kept detections, columns = tracks (cost matrix of :588-603 without embeddings)
struct BoostCost {
    static constexpr bool kWarpPerRow = false;
    static constexpr bool kGrid = true;
    static constexpr bool kBigList = true;
    const float4* det_box;
    const unsigned short* row_map;    // row -> detection index
    const float4* trk_box;
    const float4* trk_mean;
    const float4* trk_inv;
    float lambda_mhd;
    bool prune;                       // a disjoint pair costs 1 - lambda_mhd * mh_sim >= 1 - lambda_mhd > thresh
    float big_w, big_h;               // track boxes beyond twice the largest detection: listed apart from the grid
    float4 roi;                       // hull of the frame's detections: tracks that miss it cannot be looked at by any row
    float iou_floor;                  // cost <= thresh needs IoU >= 1 - thresh - lambda_mhd (mh_sim <= 1); 0: unknown
    struct Row { float4 b; float4 z; float area; };
    __device__ __forceinline__ Row row(int i) const {
        Row r;
        r.b = det_box[row_map[i]];
        r.z = boost_bbox_to_z(r.b);
        r.area = box_area(r.b);
        return r;
    }
    __device__ __forceinline__ float4 col_box(int j) const { return trk_box[j]; }
    __device__ __forceinline__ bool reject(const Row& r, int j) const { return prune && boxes_disjoint(r.b, trk_box[j]); }
    __device__ __forceinline__ float cost(const Row& r, int j) const {
        const float4 t = trk_box[j];                                                 // get_iou_matrix (:297-329)
        const float x1 = fmaxf(r.b.x, t.x), y1 = fmaxf(r.b.y, t.y), x2 = fminf(r.b.z, t.z), y2 = fminf(r.b.w, t.w);
        const float inter = xmul(fmaxf(0.0f, xsub(x2, x1)), fmaxf(0.0f, xsub(y2, y1)));
        const float uni = xsub(xadd(r.area, box_area(t)), inter);
        const float iou = (uni > 1e-6f) ? xdiv(inter, uni) : 0.0f;
        const float4 mu = trk_mean[j], iv = trk_inv[j];                              // get_mh_dist_matrix (:331-358)
        const float d0 = xsub(r.z.x, mu.x), d1 = xsub(r.z.y, mu.y), d2 = xsub(r.z.z, mu.z), d3 = xsub(r.z.w, mu.w);
        float mh = xmul(xmul(d0, d0), iv.x);
        mh = xadd(mh, xmul(xmul(d1, d1), iv.y));
        mh = xadd(mh, xmul(xmul(d2, d2), iv.z));
        mh = xadd(mh, xmul(xmul(d3, d3), iv.w));
        const float limit = 13.2767f;                                                // :592-600
        if (mh > limit) mh = limit;
        const float sim = xdiv(xsub(limit, mh), limit);
        return xsub(xsub(1.0f, iou), xmul(lambda_mhd, sim));
    }
    __device__ __forceinline__ float pair(int i, int j) const { return cost(row(i), j); }
    __device__ __forceinline__ double pair_bias(int, int) const { return 0.0; }
    __device__ __forceinline__ bool is_candidate(const Row& r, int, int j, float thresh) const { return cost(r, j) <= thresh; }
};

template <int CAP, int DMAX>
__device__ __forceinline__ void boost_frame(const BoostArgs& a, const SortStream& st, BoostSmem& sm, const float* dets,
                                            int n_det_in, float* out, int* n_out) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
    __syncthreads();
    const int frame = st.hdr[kSHdrFrame] + 1;
    const int n_trk = st.hdr[kSHdrTracks];
    int n_free = st.hdr[kSHdrFree];
    const int id_base = st.hdr[kSHdrIdCounter];
    int n_det = n_det_in;
    if (n_det > min(DMAX, a.ld_dets)) { n_det = min(DMAX, a.ld_dets); if (tid == 0) atomicOr(&st.hdr[kSHdrError], 2); }

    // ---- A. detections
    for (int j = tid; j < n_det; j += nt) {
        const float* r = dets + (size_t)j * 6;
        sm.det_box[j] = make_float4(r[0], r[1], r[2], r[3]);
        sm.det_conf[j] = r[4];
    }

    // ---- B. predict every track in place (one thread per track: four independent (position, velocity) systems)
    for (int k = tid; k < n_trk; k += nt) {
        const int slot = st.list[k];
        float* rec = st.recs + (size_t)slot * kBoostRecFloats;
        // the record is 256-byte aligned: six 16-byte loads, five 16-byte stores (the velocities do not change)
        float4* rec4 = reinterpret_cast<float4*>(rec);
        const float4 xp = rec4[0], xv = rec4[1];
        float x[8] = {xp.x, xp.y, xp.z, xp.w, xv.x, xv.y, xv.z, xv.w};
        float p00[4];                                                                // predicted position variances
        float4 blk[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) blk[c] = rec4[2 + c];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            x[c] = xadd(x[c], x[c + 4]);                                             // x = F x
            const float pcc = blk[c].x, pcv = blk[c].y, pvc = blk[c].z, pvv = blk[c].w;
            const float fcc = xadd(pcc, pvc), fcv = xadd(pcv, pvv);                   // F P
            p00[c] = xadd(xadd(fcc, fcv), 10.0f);                                    // (F P) F^T + Q
            rec4[2 + c] = make_float4(p00[c], xadd(fcv, 0.0f), xadd(pvc, pvv), xadd(pvv, 0.01f));
        }
        rec4[0] = make_float4(x[0], x[1], x[2], x[3]);
        st.age[slot] += 1;
        if (st.tsu[slot] > 0) st.hits[slot] = 0;                                     // hit_streak (:160-162)
        st.tsu[slot] += 1;
        sm.trk_box[k] = boost_state_box(x[0], x[1], x[2], x[3]);
        sm.trk_mean[k] = make_float4(x[0], x[1], x[2], x[3]);
        sm.trk_inv[k] = make_float4(xdiv(1.0f, p00[0]), xdiv(1.0f, p00[1]), xdiv(1.0f, p00[2]), xdiv(1.0f, p00[3]));
        sm.tsu1[k] = (unsigned char)min(st.tsu[slot] - 1, 255);
    }
    __syncthreads();

    // A coasting track's height / ratio velocities can inflate its box without bound (nothing clamps them in the reference);
    // the grid keeps such boxes - wider or taller than twice the largest detection of the frame - in its overflow list.
    float big_w = 0.0f, big_h = 0.0f;
    float4 roi = make_float4(3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f);                  // hull of the detections
    {
        for (int j = tid; j < n_det; j += nt) {
            const float4 b = sm.det_box[j];
            if (!box_finite(b)) continue;
            big_w = fmaxf(big_w, b.z - b.x); big_h = fmaxf(big_h, b.w - b.y);
            roi.x = fminf(roi.x, b.x); roi.y = fminf(roi.y, b.y); roi.z = fmaxf(roi.z, b.z); roi.w = fmaxf(roi.w, b.w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            big_w = fmaxf(big_w, __shfl_xor_sync(kFullMask, big_w, o));
            big_h = fmaxf(big_h, __shfl_xor_sync(kFullMask, big_h, o));
            roi.x = fminf(roi.x, __shfl_xor_sync(kFullMask, roi.x, o));
            roi.y = fminf(roi.y, __shfl_xor_sync(kFullMask, roi.y, o));
            roi.z = fmaxf(roi.z, __shfl_xor_sync(kFullMask, roi.z, o));
            roi.w = fmaxf(roi.w, __shfl_xor_sync(kFullMask, roi.w, o));
        }
        float* red = sm.lap.grid.red;
        if ((tid & 31) == 0) {
            const int w = tid >> 5;
            red[w] = big_w; red[32 + w] = big_h; red[64 + w] = roi.x; red[96 + w] = roi.y; red[128 + w] = roi.z; red[160 + w] = roi.w;
        }
        __syncthreads();
        for (int w = 0; w < (nt >> 5); ++w) {
            big_w = fmaxf(big_w, red[w]); big_h = fmaxf(big_h, red[32 + w]);
            roi.x = fminf(roi.x, red[64 + w]); roi.y = fminf(roi.y, red[96 + w]);
            roi.z = fmaxf(roi.z, red[128 + w]); roi.w = fmaxf(roi.w, red[160 + w]);
        }
        big_w *= 2.0f; big_h *= 2.0f;
        __syncthreads();                                                             // red is reused by grid_build
    }

    // ---- C. detection-confidence boost against the predicted tracks (:361-426), then the det_thresh filter (:532-538)
    if (a.p.use_dlo_boost && n_det > 0 && n_trk > 0) {
        grid_build<true>(sm.lap.grid, n_trk, sm.bs, [&](int j) { return sm.trk_box[j]; }, big_w, big_h, roi);
        const float dth_eps = xadd(a.p.det_thresh, 1e-5f);
        for (int i = tid; i < n_det; i += nt) {
            const float4 b = sm.det_box[i];
            const float area = box_area(b);
            // Only an IoU that can change the outcome has to be found: the basic rule raises conf to mx * coef, i.e. needs
            // mx > conf / coef; the use_vt rule needs one IoU > max(0.95 - (tsu - 1), 0.8) >= 0.8 and only acts below
            // det_thresh.  t sits a relative 1e-4 under that bound; pairs the tighter window skips have IoU <= t.
            const float c0 = sm.det_conf[i];
            float t = 0.0f;
            if (!a.p.use_vt) { if (c0 > 0.0f && a.p.dlo_boost_coef > 0.0f) t = c0 / a.p.dlo_boost_coef * (1.0f - 1e-4f) - 1e-6f; }
            else t = (c0 < dth_eps) ? 0.79f : 2.0f;
            float mx = 0.0f;                       // S >= 0 and at least one track exists: the row maximum over ALL tracks
            bool boost = false;
            auto see = [&](int j, float4 tb) {
                const float v = iou_pair(b, area, tb);
                if (mx < v) mx = v;
                if (v > fmaxf(xsub(0.95f, (float)sm.tsu1[j]), 0.8f)) boost = true;
            };
            if (t >= 1.0f) {}                      // no IoU can matter
            else if (t > 0.01f) grid_query_iou_above<true>(sm.lap.grid, b, t, [&](int j) { return sm.trk_box[j]; }, see);
            else grid_query<true>(sm.lap.grid, b, [&](int j) { return sm.trk_box[j]; }, see);
            const float c = sm.det_conf[i];
            if (!a.p.use_vt) {
                const float bc = xmul(mx, a.p.dlo_boost_coef);
                sm.det_conf[i] = (c < bc) ? bc : c;
            } else if (boost) {
                sm.det_conf[i] = (c < dth_eps) ? dth_eps : c;
            }
        }
        __syncthreads();
    }
    const float dth = a.p.det_thresh;
    const int m = block_compact(n_det, 0, sm.bs, [&](int j) { return sm.det_conf[j] >= dth; },
                                [&](int j, int pos) { sm.valid[pos] = (unsigned short)j; });

    // ---- D. one association: rows = kept detections, columns = tracks (:568-633)
    {
        const float thresh = a.p.iou_threshold;
        const bool prune = a.p.lambda_mhd >= 0.0f && xsub(1.0f, a.p.lambda_mhd) > thresh * 1.0001f + 1e-6f;
        BoostCost cost{sm.det_box, sm.valid, sm.trk_box, sm.trk_mean, sm.trk_inv, a.p.lambda_mhd, prune, big_w, big_h, roi,
                       prune ? 1.0f - (thresh + a.p.lambda_mhd) * 1.001f - 1e-5f : 0.0f};
        block_lap(sm.lap, m, n_trk, DMAX, CAP, thresh, cost);
    }
    const int n_match = block_compact(m, 0, sm.bs, [&](int r) { return sm.lap.row2col[r] >= 0; },
                                      [&](int r, int pos) { sm.sel2[pos] = (unsigned short)r; });
    __syncthreads();

    // ---- E. Kalman update of the matched tracks (:646-654, BoostKalmanFilter::update :61-75, BoostTrack::update :165-181)
    for (int q = tid; q < n_match; q += nt) {
        const int r = sm.sel2[q];
        const int det = sm.valid[r];
        const int slot = st.list[sm.lap.row2col[r]];
        float4* rec4 = reinterpret_cast<float4*>(st.recs + (size_t)slot * kBoostRecFloats);
        const float4 zz = boost_bbox_to_z(sm.det_box[det]);
        const float z[4] = {zz.x, zz.y, zz.z, zz.w};
        const float R[4] = {1.0f, 1.0f, 10.0f, 0.01f};
        const float4 xp = rec4[0], xv = rec4[1];
        float x[8] = {xp.x, xp.y, xp.z, xp.w, xv.x, xv.y, xv.z, xv.w};
        float4 blk[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) blk[c] = rec4[2 + c];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float pcc = blk[c].x, pcv = blk[c].y, pvc = blk[c].z, pvv = blk[c].w;
            const float S = xadd(pcc, R[c]);
            const float Sinv = xdiv(1.0f, S);                                        // LU inverse of the diagonal S
            const float kc = xmul(pcc, Sinv), kv = xmul(pvc, Sinv);                   // K = P H^T S^-1
            const float innov = xsub(z[c], x[c]);
            x[c] = xadd(x[c], xmul(kc, innov));
            x[c + 4] = xadd(x[c + 4], xmul(kv, innov));
            const float kcs = xmul(kc, S), kvs = xmul(kv, S);                         // P - (K S) K^T
            rec4[2 + c] = make_float4(xsub(pcc, xmul(kcs, kc)), xsub(pcv, xmul(kcs, kv)), xsub(pvc, xmul(kvs, kc)),
                                      xsub(pvv, xmul(kvs, kv)));
        }
        rec4[0] = make_float4(x[0], x[1], x[2], x[3]);
        rec4[1] = make_float4(x[4], x[5], x[6], x[7]);
        st.tsu[slot] = 0;
        st.hits[slot] += 1;
        st.conf[slot] = sm.det_conf[det];
        st.cls[slot] = (int)dets[(size_t)det * 6 + 5];
        st.det_ind[slot] = det;
    }

    // ---- F. new tracks for the unmatched kept detections, ascending (:657-666, BoostKalmanFilter ctor :22-54)
    const int n_new_want = block_compact(m, 0, sm.bs, [&](int r) { return sm.lap.row2col[r] < 0; },
                                         [&](int r, int pos) { sm.sel2[pos] = sm.valid[r]; });
    int n_new = n_new_want;
    if (n_new > n_free || n_trk + n_new > CAP) { n_new = min(n_free, CAP - n_trk); if (tid == 0) atomicOr(&st.hdr[kSHdrError], 1); }
    for (int k = tid; k < n_new; k += nt) {
        const int det = sm.sel2[k];
        const int slot = st.freel[n_free - 1 - k];
        float* rec = st.recs + (size_t)slot * kBoostRecFloats;
        const float4 zz = boost_bbox_to_z(sm.det_box[det]);
        rec[0] = zz.x; rec[1] = zz.y; rec[2] = zz.z; rec[3] = zz.w;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            rec[4 + c] = 0.0f;
            rec[8 + 4 * c + 0] = 10.0f; rec[8 + 4 * c + 1] = 0.0f; rec[8 + 4 * c + 2] = 0.0f; rec[8 + 4 * c + 3] = 10000.0f;
        }
        st.id[slot] = id_base + 1 + k;
        st.hits[slot] = 0; st.tsu[slot] = 0; st.age[slot] = 0;
        st.conf[slot] = sm.det_conf[det];
        st.cls[slot] = (int)dets[(size_t)det * 6 + 5];
        st.det_ind[slot] = det;
        st.list[n_trk + k] = (unsigned short)slot;
    }
    __syncthreads();
    const int n_all = n_trk + n_new;
    n_free -= n_new;

    // ---- G. output rows in track order through filter_outputs (:669-698, :434-463)
    const int min_hits = a.p.min_hits;
    const int n_rows = block_compact(n_all, 0, sm.bs,
                                     [&](int k) {
                                         const int slot = st.list[k];
                                         if (!(st.tsu[slot] < 1 && (st.hits[slot] >= min_hits || frame <= min_hits))) return false;
                                         const float* rec = st.recs + (size_t)slot * kBoostRecFloats;
                                         const float4 b = boost_state_box(rec[0], rec[1], rec[2], rec[3]);
                                         const float w = xsub(b.z, b.x), h = xsub(b.w, b.y);
                                         return xdiv(w, xadd(h, 1e-6f)) <= a.p.aspect_ratio_thresh && xmul(w, h) > a.p.min_box_area;
                                     },
                                     [&](int k, int pos) {
                                         if (pos >= a.ld_out) return;
                                         const int slot = st.list[k];
                                         const float* rec = st.recs + (size_t)slot * kBoostRecFloats;
                                         float* o = out + (size_t)pos * 8;
                                         *reinterpret_cast<float4*>(o) = boost_state_box(rec[0], rec[1], rec[2], rec[3]);
                                         *reinterpret_cast<float4*>(o + 4) = make_float4((float)st.id[slot], st.conf[slot],
                                                                                         (float)st.cls[slot], (float)st.det_ind[slot]);
                                     });

    // ---- H. age out (:685-689): survivors keep their order (compacted through shared memory), dead slots are freed
    const int max_age = a.p.max_age;
    const int n_keep = block_compact(n_all, 0, sm.bs, [&](int k) { return st.tsu[st.list[k]] <= max_age; },
                                     [&](int k, int pos) { sm.sel[pos] = st.list[k]; });
    n_free = block_compact(n_all, n_free, sm.bs, [&](int k) { return st.tsu[st.list[k]] > max_age; },
                           [&](int k, int pos) { st.freel[pos] = st.list[k]; });
    for (int k = tid; k < n_keep; k += nt) st.list[k] = sm.sel[k];
    if (tid == 0) {
        if (n_rows > a.ld_out) atomicOr(&st.hdr[kSHdrError], 4);
        *n_out = n_rows < a.ld_out ? n_rows : a.ld_out;
        st.hdr[kSHdrTracks] = n_keep;
        st.hdr[kSHdrFree] = n_free;
        st.hdr[kSHdrIdCounter] = id_base + n_new;
        st.hdr[kSHdrFrame] = frame;
        st.hdr[kSHdrN] = m; st.hdr[kSHdrM] = n_trk;
        st.hdr[8] = n_match; st.hdr[9] = n_new;
    }
    __syncthreads();
}

template <int CAP, int DMAX, int ECAP>
__global__ void __launch_bounds__(kBoostThreads) boosttrack_step_kernel(BoostArgs a) {
    MOT_DYNAMIC_SMEM(smem);
    BoostSmem sm;
    boost_carve(smem, CAP, DMAX, ECAP, sm);
    constexpr SortLayout L = SortLayout::make(CAP, DMAX);
    for (int s = a.s_begin + (int)blockIdx.x; s < a.s_end; s += (int)gridDim.x) {
        SortStream st = SortStream::at(a.state + (size_t)s * L.stride, L);
        lap_carve_gscratch(st.gscratch, DMAX, CAP, sm.lap);
        for (int t = 0; t < a.T; ++t) {
            const size_t fs = (size_t)t * a.S + s;
            boost_frame<CAP, DMAX>(a, st, sm, a.dets + fs * (size_t)a.ld_dets * 6, a.n_dets[fs],
                                   a.out + fs * (size_t)a.ld_out * 8, a.n_out + fs);
        }
    }
}

}  // namespace mot
