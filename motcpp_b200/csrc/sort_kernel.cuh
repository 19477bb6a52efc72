// sort_kernel.cuh - SORT's whole per-frame update() as one kernel, one CTA per camera stream.
// Replaces reference src/trackers/sort.cpp:102-255 (Sort::update) and :21-76 (SortTrack):
//   confidence filter (:111-121)                         -> phase A
//   predict every track IN PLACE, drop NaN boxes (:124-150) -> phase B
//   iou_distance + linear_assignment(1 - iou_threshold) (:152-181) -> phase C
//   KalmanFilterXYSR update of matched tracks (:184-193)  -> phase D
//   spawn tracks for unmatched detections, ascending (:196-204) -> phase E
//   age-out time_since_update > max_age (:207-216)        -> phase F
//   output rows (:219-253)                                -> phase G
// State records are [x 7 | P 7x7] fp32 padded to 64 floats.  IDs come from a per-stream counter.
#pragma once
#include "shapes.cuh"
#include "block_utils.cuh"
#include "cost_device.cuh"
#include "kf_device.cuh"
#include "lap_device.cuh"

namespace mot {

#ifndef MOT_SORT_THREADS
#define MOT_SORT_THREADS 256
#endif
constexpr int kSortThreads = MOT_SORT_THREADS;
constexpr int kSortRecFloats = 64;      // 56 used

enum : int { kSHdrTracks = 0, kSHdrFree = 2, kSHdrIdCounter = 3, kSHdrFrame = 4, kSHdrError = 5, kSHdrN = 6, kSHdrM = 7 };

struct SortParams {
    float det_thresh, iou_threshold;
    int max_age, min_hits;
};

struct SortLayout {
    int cap, d_max;
    size_t off_lists, off_meta, off_recs, off_gscratch, stride;
    MOT_HD static constexpr size_t al(size_t x) { return (x + 255) & ~(size_t)255; }
    MOT_HD static constexpr SortLayout make(int cap, int d_max) {
        SortLayout L{};
        L.cap = cap; L.d_max = d_max;
        size_t o = al(sizeof(int) * 16);
        L.off_lists = o;    o = al(o + sizeof(unsigned short) * 2 * (size_t)cap);
        L.off_meta = o;     o = al(o + sizeof(int) * 7 * (size_t)cap);
        L.off_recs = o;     o = al(o + sizeof(float) * kSortRecFloats * (size_t)cap);
        L.off_gscratch = o; o = al(o + lap_gscratch_bytes(cap, d_max));
        L.stride = o;
        return L;
    }
};

struct SortStream {
    int* hdr;
    unsigned short *list, *freel;
    int *id, *hits, *tsu, *age, *cls, *det_ind;
    float* conf;
    float* recs;
    unsigned char* gscratch;
    __device__ __forceinline__ static SortStream at(unsigned char* base, const SortLayout& L) {
        SortStream s;
        s.hdr = (int*)base;
        s.list = (unsigned short*)(base + L.off_lists);
        s.freel = s.list + L.cap;
        int* m = (int*)(base + L.off_meta);
        s.id = m; s.hits = m + L.cap; s.tsu = m + 2 * L.cap; s.age = m + 3 * L.cap;
        s.cls = m + 4 * L.cap; s.det_ind = m + 5 * L.cap; s.conf = (float*)(m + 6 * L.cap);
        s.recs = (float*)(base + L.off_recs);
        s.gscratch = base + L.off_gscratch;
        return s;
    }
};

struct SortArgs {
    unsigned char* state;
    const float* dets;        // [T][S][ld_dets][6]
    const int* n_dets;        // [T][S]
    float* out;               // [T][S][ld_out][8]
    int* n_out;               // [T][S]
    int T, S, ld_dets, ld_out;
    int s_begin, s_end;
    SortParams p;
};

struct SortSmem {
    float4* det_box;            // [d_max] raw xyxy of every detection
    float* det_conf;            // [d_max]
    unsigned short* valid;      // [d_max] detections with conf >= det_thresh
    float4* row_box;            // [cap] predicted boxes
    unsigned short* list_a;     // [cap]
    unsigned short* sel;        // [cap]
    unsigned short* sel2;       // [cap]
    unsigned char* flag;        // [cap]
    BlockScratch* bs;
    LapWorkspace lap;
};

MOT_HD constexpr size_t sort_smem_bytes(int cap, int d_max, int e_cap) {
    size_t b = 0;
    b += lap_align16(sizeof(float4) * (size_t)d_max);
    b += lap_align16(sizeof(float) * (size_t)d_max);
    b += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    b += lap_align16(sizeof(float4) * (size_t)cap);
    b += 3 * lap_align16(sizeof(unsigned short) * (size_t)cap);
    b += lap_align16((size_t)cap);
    b += lap_align16(sizeof(BlockScratch));
    b += lap_smem_bytes(cap, d_max, e_cap);
    return b;
}

__device__ __forceinline__ void sort_carve(unsigned char* p, int cap, int d_max, int e_cap, SortSmem& s) {
    s.det_box = (float4*)p;            p += lap_align16(sizeof(float4) * (size_t)d_max);
    s.det_conf = (float*)p;            p += lap_align16(sizeof(float) * (size_t)d_max);
    s.valid = (unsigned short*)p;      p += lap_align16(sizeof(unsigned short) * (size_t)d_max);
    s.row_box = (float4*)p;            p += lap_align16(sizeof(float4) * (size_t)cap);
    s.list_a = (unsigned short*)p;     p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.sel = (unsigned short*)p;        p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.sel2 = (unsigned short*)p;       p += lap_align16(sizeof(unsigned short) * (size_t)cap);
    s.flag = p;                        p += lap_align16((size_t)cap);
    s.bs = (BlockScratch*)p;           p += lap_align16(sizeof(BlockScratch));
    lap_carve(p, cap, d_max, e_cap, s.lap);
}

// SortTrack::get_state (sort.cpp:72-76) from the record's first four floats
__device__ __forceinline__ float4 sort_track_box(const float* rec) {
    return xysr2xyxy(rec[0], rec[1], rec[2], rec[3]);
}

template <int CAP, int DMAX>
__device__ __forceinline__ void sort_frame(const SortArgs& a, const SortStream& st, SortSmem& sm, const float* dets,
                                           int n_det_in, float* out, int* n_out) {
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
    const int lane = tid & 31, g = lane & 7, base = lane & ~7;
    const int groups = nt >> 3, gid = tid >> 3;
    __syncthreads();
    const int frame = st.hdr[kSHdrFrame] + 1;                    // frame_count_ (:108)
    const int n_trk0 = st.hdr[kSHdrTracks];
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
    __syncthreads();
    const float dth = a.p.det_thresh;
    const int m = block_compact(n_det, 0, sm.bs, [&](int j) { return sm.det_conf[j] >= dth; },
                                [&](int j, int pos) { sm.valid[pos] = (unsigned short)j; });

    // ---- B. predict every track in place; a NaN box removes the track (:132-150)
    {
        const int rounds = (n_trk0 + groups - 1) / groups;
        for (int it = 0; it < rounds; ++it) {
            const int k = it * groups + gid;
            const bool live = k < n_trk0;
            const int slot = live ? (int)st.list[k] : 0;
            float* rec = st.recs + (size_t)slot * kSortRecFloats;
            KfRow7 s;
            kf7_load_row(rec, live ? g : 7, s);
            kf_xysr_predict(s, g, base, 0.01f, 0.0001f);
            const float x0 = __shfl_sync(kFullMask, s.m, base + 0), x1 = __shfl_sync(kFullMask, s.m, base + 1);
            const float x2 = __shfl_sync(kFullMask, s.m, base + 2), x3 = __shfl_sync(kFullMask, s.m, base + 3);
            if (live) {
                kf7_store_row(rec, g, s);
                if (g == 0) {
                    st.age[slot] += 1;
                    st.tsu[slot] += 1;
                    const float4 b = xysr2xyxy(x0, x1, x2, x3);
                    const float sum = b.x + b.y + b.z + b.w;
                    sm.flag[k] = (sum != sum) ? 1 : 0;
                }
            }
        }
    }
    __syncthreads();
    // survivors keep their order; dead slots go back on the free stack
    const int n_trk = block_compact(n_trk0, 0, sm.bs, [&](int k) { return sm.flag[k] == 0; },
                                    [&](int k, int pos) { sm.list_a[pos] = st.list[k]; });
    n_free = block_compact(n_trk0, n_free, sm.bs, [&](int k) { return sm.flag[k] != 0; },
                           [&](int k, int pos) { st.freel[pos] = st.list[k]; });
    for (int k = tid; k < n_trk; k += nt) sm.row_box[k] = sort_track_box(st.recs + (size_t)sm.list_a[k] * kSortRecFloats);
    __syncthreads();

    // ---- C. association on IoU distance
    {
        const float thresh = xsub(1.0f, a.p.iou_threshold);
        // 1 - iou <= thresh needs iou >= iou_threshold: the grid walk (taken above 64 k pairs) only looks at that corner window
        IouCost cost{sm.row_box, sm.det_box, sm.det_conf, sm.valid, false, thresh < 1.0f, 1.0f - thresh * 1.001f - 1e-5f};
        block_lap(sm.lap, n_trk, m, CAP, DMAX, thresh, cost);
    }
    const int n_match = block_compact(n_trk, 0, sm.bs, [&](int r) { return sm.lap.row2col[r] >= 0; },
                                      [&](int r, int pos) { sm.sel[pos] = (unsigned short)r; });
    const int n_new_want = block_compact(m, 0, sm.bs, [&](int j) { return sm.lap.col2row[j] < 0; },
                                         [&](int j, int pos) { sm.sel2[pos] = sm.valid[j]; });
    __syncthreads();

    // ---- D. update matched tracks (SortTrack::update, :53-70)
    {
        const int rounds = (n_match + groups - 1) / groups;
        for (int it = 0; it < rounds; ++it) {
            const int k = it * groups + gid;
            const bool live = k < n_match;
            const int r = live ? (int)sm.sel[k] : 0;
            const int slot = live ? (int)sm.list_a[r] : 0;
            const int det = live ? (int)sm.valid[sm.lap.row2col[r]] : 0;
            float* rec = st.recs + (size_t)slot * kSortRecFloats;
            KfRow7 s;
            kf7_load_row(rec, live ? g : 7, s);
            if (!live) { s.m = 1.0f; for (int j = 0; j < 7; ++j) s.p[j] = (j == g) ? 1.0f : 0.0f; }
            float z[4] = {0.0f, 0.0f, 1.0f, 1.0f};
            if (live) {
                const float4 q = xyxy2xysr(sm.det_box[det]);
                z[0] = q.x; z[1] = q.y; z[2] = q.z; z[3] = q.w;
            }
            const bool ok = kf_xysr_update(s, g, base, z);
            if (live) {
                if (ok) kf7_store_row(rec, g, s);
                if (g == 0) {
                    if (!ok) atomicOr(&st.hdr[kSHdrError], 8);
                    st.conf[slot] = sm.det_conf[det];
                    st.cls[slot] = (int)dets[(size_t)det * 6 + 5];
                    st.det_ind[slot] = det;
                    st.hits[slot] += 1;
                    st.tsu[slot] = 0;
                }
            }
        }
    }
    __syncthreads();

    // ---- E. new tracks for unmatched detections, ascending detection order (:196-204)
    int n_new = n_new_want;
    if (n_new > n_free) { n_new = n_free; if (tid == 0) atomicOr(&st.hdr[kSHdrError], 1); }
    for (int k = gid; k < n_new; k += groups) {
        const int det = sm.sel2[k];
        const int slot = st.freel[n_free - 1 - k];
        const float4 q = xyxy2xysr(sm.det_box[det]);
        const float z[4] = {q.x, q.y, q.z, q.w};
        KfRow7 s;
        kf_xysr_init(s, g, z);
        kf7_store_row(st.recs + (size_t)slot * kSortRecFloats, g, s);
        if (g == 0) {
            st.id[slot] = id_base + 1 + k;
            st.hits[slot] = 1; st.tsu[slot] = 0; st.age[slot] = 1;
            st.conf[slot] = sm.det_conf[det];
            st.cls[slot] = (int)dets[(size_t)det * 6 + 5];
            st.det_ind[slot] = det;
            sm.list_a[n_trk + k] = (unsigned short)slot;
        }
    }
    __syncthreads();
    const int n_all = n_trk + n_new;
    n_free -= n_new;

    // ---- F. age out (:207-216)
    const int max_age = a.p.max_age;
    const int n_keep = block_compact(n_all, 0, sm.bs, [&](int k) { return st.tsu[sm.list_a[k]] <= max_age; },
                                     [&](int k, int pos) { st.list[pos] = sm.list_a[k]; });
    n_free = block_compact(n_all, n_free, sm.bs, [&](int k) { return st.tsu[sm.list_a[k]] > max_age; },
                           [&](int k, int pos) { st.freel[pos] = sm.list_a[k]; });

    // ---- G. output (:219-253)
    const int min_hits = a.p.min_hits;
    const int n_rows = block_compact(n_keep, 0, sm.bs,
                                     [&](int k) {
                                         const int slot = st.list[k];
                                         return st.tsu[slot] == 0 && (st.hits[slot] >= min_hits || frame <= min_hits);
                                     },
                                     [&](int k, int pos) {
                                         if (pos >= a.ld_out) return;
                                         const int slot = st.list[k];
                                         const float4 b = sort_track_box(st.recs + (size_t)slot * kSortRecFloats);
                                         float* o = out + (size_t)pos * 8;
                                         *reinterpret_cast<float4*>(o) = b;
                                         *reinterpret_cast<float4*>(o + 4) = make_float4((float)st.id[slot], st.conf[slot],
                                                                                         (float)st.cls[slot], (float)st.det_ind[slot]);
                                     });
    if (tid == 0) {
        if (n_rows > a.ld_out) atomicOr(&st.hdr[kSHdrError], 4);
        *n_out = n_rows < a.ld_out ? n_rows : a.ld_out;
        st.hdr[kSHdrTracks] = n_keep;
        st.hdr[kSHdrFree] = n_free;
        st.hdr[kSHdrIdCounter] = id_base + n_new;
        st.hdr[kSHdrFrame] = frame;
        st.hdr[kSHdrN] = n_trk; st.hdr[kSHdrM] = m;
    }
    __syncthreads();
}

template <int CAP, int DMAX, int ECAP>
__global__ void __launch_bounds__(kSortThreads) sort_step_kernel(SortArgs a) {
    MOT_DYNAMIC_SMEM(smem);
    SortSmem sm;
    sort_carve(smem, CAP, DMAX, ECAP, sm);
    constexpr SortLayout L = SortLayout::make(CAP, DMAX);
    for (int s = a.s_begin + (int)blockIdx.x; s < a.s_end; s += (int)gridDim.x) {
        SortStream st = SortStream::at(a.state + (size_t)s * L.stride, L);
        lap_carve_gscratch(st.gscratch, CAP, DMAX, sm.lap);
        for (int t = 0; t < a.T; ++t) {
            const size_t fs = (size_t)t * a.S + s;
            sort_frame<CAP, DMAX>(a, st, sm, a.dets + fs * (size_t)a.ld_dets * 6, a.n_dets[fs],
                                  a.out + fs * (size_t)a.ld_out * 8, a.n_out + fs);
        }
    }
}

static __global__ void sort_reset_kernel(unsigned char* state, SortLayout L, int S, int keep_id_counter) {
    for (int s = (int)blockIdx.x; s < S; s += (int)gridDim.x) {
        SortStream st = SortStream::at(state + (size_t)s * L.stride, L);
        for (int k = (int)threadIdx.x; k < L.cap; k += (int)blockDim.x) st.freel[k] = (unsigned short)(L.cap - 1 - k);
        if (threadIdx.x == 0) {
            const int idc = keep_id_counter ? st.hdr[kSHdrIdCounter] : 0;
            for (int k = 0; k < 16; ++k) st.hdr[k] = 0;
            st.hdr[kSHdrFree] = L.cap;
            st.hdr[kSHdrIdCounter] = idc;
        }
        __syncthreads();
    }
}

}  // namespace mot
