// TEST INFRASTRUCTURE: builds the product's kernel sources against the SIMT emulator
// (tests/cpusim/cpusim.hpp) into tests/cpusim/libmotb200_cpusim.so.  Never shipped, never
// loaded by motcpp_b200/.
#define MOT_CPUSIM 1
#include "cpusim.hpp"
#include "../../motcpp_b200/csrc/kernels_lap.cuh"

#include <vector>

extern "C" {

// one problem, host pointers; block size is a parameter so tests can cover several shapes
int sim_lap(const float* cost, int n, int m, int ld, float thresh, int* row2col, int* col2row,
            int e_cap, int threads) {
    const int n_max = n > 0 ? n : 1, m_max = m > 0 ? m : 1;
    std::vector<unsigned char> gs(mot::lap_gscratch_bytes(n_max, m_max) + 64);
    mot::LapBatchArgs a{};
    a.cost = cost; a.stride_cost = 0; a.n_rows = nullptr; a.n_cols = nullptr;
    a.n = n; a.m = m; a.ld = ld; a.thresh = thresh;
    a.row2col = row2col; a.col2row = col2row; a.gscratch = gs.data();
    a.n_max = n_max; a.m_max = m_max; a.e_cap = e_cap; a.n_problems = 1;
    const size_t smem = mot::lap_smem_bytes(n_max, m_max, e_cap);
    cpusim::launch(dim3(1), dim3(threads), smem, [=] { mot::lap_dense_kernel(a); });
    return 0;
}

// the reference-order dense LAPJV: block = 0 -> one warp (jv_device.cuh, n + m <= kLapJvMax), 1 -> whole CTA (jv_block_device.cuh)
int sim_lap_jv(const float* cost, int n, int m, int ld, float thresh, int* row2col, int* col2row, int block, int threads) {
    if (!block) {
        if (n + m > mot::kLapJvMax) return -1;
        cpusim::launch(dim3(1), dim3(32), mot::jv_work_bytes(n + m + 1), [=] { mot::lap_jv_kernel(cost, 0, 1, n, m, ld, thresh, row2col, col2row); });
        return 0;
    }
    std::vector<unsigned char> gs(mot::jv_block_gbytes(n + m) + 64);
    unsigned char* g = gs.data();
    // block = 1: work arrays in global scratch; block = 2: everything in shared memory
    const int all_shared = block == 2;
    cpusim::launch(dim3(1), dim3(threads), all_shared ? mot::jv_block_sbytes_full(n + m) : mot::jv_block_sbytes(n + m),
                   [=] { mot::lap_jv_block_kernel(cost, 0, 1, n, m, ld, thresh, row2col, col2row, g, all_shared); });
    return 0;
}

// the corner grid with its overflow list (grid_device.cuh): visits[i * m + j] counts how often row i's query saw column j.
// t = 0: every overlapping pair; t > 0: the IoU-floor window.  big_w / big_h / use_roi as the BoostTrack kernel passes them.
int sim_grid_pairs(const float* rows, int n, const float* cols, int m, float big_w, float big_h, int use_roi, float t,
                   int* visits, int* n_big, int threads) {
    const int cap = m > 0 ? m : 1;
    const size_t smem = mot::grid_smem_bytes(cap) + sizeof(mot::BlockScratch) + 64;
    const float4* rb = reinterpret_cast<const float4*>(rows);
    const float4* cb = reinterpret_cast<const float4*>(cols);
    cpusim::launch(dim3(1), dim3(threads), smem, [=] {
        MOT_DYNAMIC_SMEM(sm);
        mot::BoxGrid g;
        unsigned char* p = mot::grid_carve(sm, cap, g);
        mot::BlockScratch* bs = reinterpret_cast<mot::BlockScratch*>(p + ((16 - ((size_t)p & 15)) & 15));
        float4 roi = make_float4(-3.0e38f, -3.0e38f, 3.0e38f, 3.0e38f);
        if (use_roi) {
            roi = make_float4(3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f);
            for (int i = 0; i < n; ++i) {
                roi.x = fminf(roi.x, rb[i].x); roi.y = fminf(roi.y, rb[i].y); roi.z = fmaxf(roi.z, rb[i].z); roi.w = fmaxf(roi.w, rb[i].w);
            }
        }
        mot::grid_build<true>(g, m, bs, [&](int j) { return cb[j]; }, big_w, big_h, roi);
        if (threadIdx.x == 0) *n_big = g.n_big;
        for (int i = (int)threadIdx.x; i < n; i += (int)blockDim.x) {
            auto see = [&](int j, float4) { visits[(size_t)i * m + j] += 1; };
            if (t > 0.0f) mot::grid_query_iou_above<true>(g, rb[i], t, [&](int j) { return cb[j]; }, see);
            else mot::grid_query<true>(g, rb[i], [&](int j) { return cb[j]; }, see);
        }
    });
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------ ByteTrack engine under the emulator
#include "../../motcpp_b200/csrc/bytetrack_kernel.cuh"

namespace {
struct SimBt {
    mot::BtLayout L;
    int S, e_cap, shape;
    mot::BtParams p;
    std::vector<unsigned char> state;
};
}  // namespace

extern "C" {

void* sim_bt_create(int S, int cap, int d_max, int e_cap, float min_conf, float track_thresh, float match_thresh,
                    int track_buffer, int frame_rate) {
    auto* h = new SimBt();
    h->shape = -1;
    for (int i = 0; i < mot::kNumBtShapes; ++i)
        if (mot::kBtShapes[i].cap >= cap && mot::kBtShapes[i].d_max >= d_max) { h->shape = i; break; }
    if (h->shape < 0) { delete h; return nullptr; }
    (void)e_cap;
    h->L = mot::BtLayout::make(mot::kBtShapes[h->shape].cap, mot::kBtShapes[h->shape].d_max);
    h->S = S; h->e_cap = mot::kBtShapes[h->shape].e_cap;
    h->p.min_conf = min_conf; h->p.track_thresh = track_thresh; h->p.match_thresh = match_thresh;
    h->p.det_thresh = track_thresh;
    h->p.max_time_lost = (int)(frame_rate / 30.0f * track_buffer);
    h->state.assign(h->L.stride * (size_t)S + 256, 0);
    unsigned char* st = h->state.data();
    const mot::BtLayout L = h->L;
    cpusim::launch(dim3(S), dim3(64), 0, [=] { mot::bytetrack_reset_kernel(st, L, S, 0); });
    return h;
}
void sim_bt_destroy(void* hv) { delete (SimBt*)hv; }

int sim_bt_update(void* hv, const float* dets, const int* n_dets, int T, int ld_dets, float* out, int* n_out,
                  int ld_out, int threads, int os_threads) {
    auto* h = (SimBt*)hv;
    mot::BtArgs a{};
    a.state = h->state.data();
    a.dets = dets; a.n_dets = n_dets; a.out = out; a.n_out = n_out;
    a.T = T; a.S = h->S; a.ld_dets = ld_dets; a.ld_out = ld_out; a.e_cap = h->e_cap; a.p = h->p; a.s_begin = 0; a.s_end = h->S;
    const size_t smem = mot::bt_smem_bytes(h->L.cap, h->L.d_max, h->e_cap);
    auto run = [&](auto tag) {
        constexpr int I = decltype(tag)::value;
        constexpr mot::BtShape sh = mot::kBtShapes[I];
        cpusim::launch(dim3(h->S), dim3(threads), smem, [=] { mot::bytetrack_step_kernel<sh.cap, sh.d_max, sh.e_cap>(a); }, os_threads);
    };
    switch (h->shape) {
        case 0: run(std::integral_constant<int, 0>{}); break;
        case 1: run(std::integral_constant<int, 1>{}); break;
        case 2: run(std::integral_constant<int, 2>{}); break;
        default: run(std::integral_constant<int, 3>{}); break;
    }
    return 0;
}

// header (16 ints) of stream s
void sim_bt_header(void* hv, int s, int* hdr16) {
    auto* h = (SimBt*)hv;
    std::memcpy(hdr16, h->state.data() + (size_t)s * h->L.stride, sizeof(int) * mot::kHdrInts);
}

// rows of [id, state, activated, frame_id, start_frame, tracklet_len, mean 8, cov 64] for list `which`
int sim_bt_dump(void* hv, int s, int which, float* outrows, int cap_rows) {
    auto* h = (SimBt*)hv;
    unsigned char* base = h->state.data() + (size_t)s * h->L.stride;
    const int* hdr = (const int*)base;
    const unsigned short* lists = (const unsigned short*)(base + h->L.off_lists);
    const unsigned short* list = lists + (which == 0 ? 0 : h->L.cap);
    const int n = which == 0 ? hdr[mot::kHdrActive] : hdr[mot::kHdrLost];
    const unsigned char* sflag = base + h->L.off_sflag;
    const int* meta = (const int*)(base + h->L.off_meta);
    const float* recs = (const float*)(base + h->L.off_recs);
    const int cap = h->L.cap;
    int k = 0;
    for (; k < n && k < cap_rows; ++k) {
        const int slot = list[k];
        float* o = outrows + 78 * k;
        o[0] = (float)meta[slot]; o[1] = (float)(sflag[slot] & 0x0f); o[2] = (sflag[slot] & 0x10) ? 1.0f : 0.0f;
        o[3] = (float)meta[2 * cap + slot]; o[4] = (float)meta[3 * cap + slot]; o[5] = (float)meta[cap + slot];
        mot::kfb_expand(recs + (size_t)slot * mot::kBtRecFloats, o + 6);
    }
    return k;
}

}  // extern "C"

// ------------------------------------------------------------------ SORT engine under the emulator
#include "../../motcpp_b200/csrc/sort_kernel.cuh"

namespace {
struct SimSort {
    mot::SortLayout L;
    int S;
    mot::SortParams p;
    std::vector<unsigned char> state;
};
}  // namespace

extern "C" {

void* sim_sort_create(int S, float det_thresh, int max_age, int min_hits, float iou_threshold) {
    auto* h = new SimSort();
    h->L = mot::SortLayout::make(256, 64);
    h->S = S;
    h->p.det_thresh = det_thresh; h->p.max_age = max_age; h->p.min_hits = min_hits; h->p.iou_threshold = iou_threshold;
    h->state.assign(h->L.stride * (size_t)S + 256, 0);
    unsigned char* st = h->state.data();
    const mot::SortLayout L = h->L;
    cpusim::launch(dim3(S), dim3(64), 0, [=] { mot::sort_reset_kernel(st, L, S, 0); });
    return h;
}
void sim_sort_destroy(void* hv) { delete (SimSort*)hv; }

int sim_sort_update(void* hv, const float* dets, const int* n_dets, int T, int ld_dets, float* out, int* n_out,
                    int ld_out, int threads) {
    auto* h = (SimSort*)hv;
    mot::SortArgs a{};
    a.state = h->state.data(); a.dets = dets; a.n_dets = n_dets; a.out = out; a.n_out = n_out;
    a.T = T; a.S = h->S; a.ld_dets = ld_dets; a.ld_out = ld_out; a.s_begin = 0; a.s_end = h->S; a.p = h->p;
    const size_t smem = mot::sort_smem_bytes(256, 64, 1024);
    cpusim::launch(dim3(h->S), dim3(threads), smem, [=] { mot::sort_step_kernel<256, 64, 1024>(a); });
    return 0;
}
void sim_sort_header(void* hv, int s, int* hdr16) {
    auto* h = (SimSort*)hv;
    std::memcpy(hdr16, h->state.data() + (size_t)s * h->L.stride, sizeof(int) * 16);
}

}  // extern "C"

// ------------------------------------------------------------------ OC-SORT engine under the emulator
#include "../../motcpp_b200/csrc/ocsort_kernel.cuh"

namespace {
struct SimOc {
    mot::OcLayout L;
    int S;
    mot::OcParams p;
    std::vector<unsigned char> state;
};
}  // namespace

extern "C" {

float sim_acosf(float x) { return mot::acosf_cr(x); }

void* sim_oc_create(int S, float det_thresh, int max_age, int min_hits, float iou_threshold, float min_conf, int delta_t,
                    float inertia, int use_byte, float q_xy, float q_s) {
    auto* h = new SimOc();
    h->L = mot::OcLayout::make(256, 64);
    h->S = S;
    h->p.det_thresh = det_thresh; h->p.max_age = max_age; h->p.min_hits = min_hits; h->p.iou_threshold = iou_threshold;
    h->p.min_conf = min_conf; h->p.delta_t = delta_t; h->p.inertia = inertia; h->p.use_byte = use_byte;
    h->p.q44 = 0.01f * q_xy; h->p.q66 = 0.0001f * q_s;
    h->state.assign(h->L.stride * (size_t)S + 256, 0);
    unsigned char* st = h->state.data();
    const mot::OcLayout L = h->L;
    cpusim::launch(dim3(S), dim3(64), 0, [=] { mot::ocsort_reset_kernel(st, L, L.stride, S, 0); });
    return h;
}
void sim_oc_set_asso(void* hv, int asso, int frame_w, int frame_h) {
    auto* h = (SimOc*)hv;
    h->p.asso = asso;
    h->p.asso_norm = static_cast<float>(std::sqrt((double)(frame_w * frame_w + frame_h * frame_h)));
}
void sim_oc_destroy(void* hv) { delete (SimOc*)hv; }

int sim_oc_update(void* hv, const float* dets, const int* n_dets, int T, int ld_dets, float* out, int* n_out, int ld_out,
                  int threads) {
    auto* h = (SimOc*)hv;
    mot::OcArgs a{};
    a.state = h->state.data(); a.dets = dets; a.n_dets = n_dets; a.out = out; a.n_out = n_out;
    a.T = T; a.S = h->S; a.ld_dets = ld_dets; a.ld_out = ld_out; a.s_begin = 0; a.s_end = h->S; a.p = h->p;
    const size_t smem = mot::oc_smem_bytes(256, 64, 1024);
    if (h->p.asso == mot::kVarCentroid) cpusim::launch(dim3(h->S), dim3(threads), smem, [=] { mot::ocsort_centroid_step_kernel<256, 64, 1024>(a); });
    else cpusim::launch(dim3(h->S), dim3(threads), smem, [=] { mot::ocsort_step_kernel<256, 64, 1024>(a); });
    return 0;
}
void sim_oc_header(void* hv, int s, int* hdr16) {
    auto* h = (SimOc*)hv;
    std::memcpy(hdr16, h->state.data() + (size_t)s * h->L.stride, sizeof(int) * 16);
}

// rows of [id, age, hits, hit_streak, time_since_update, conf, cls, det_ind, last_obs 5, velocity 2, x 7, P 49] = 71 floats
int sim_oc_dump(void* hv, int s, float* rows, int cap_rows) {
    auto* h = (SimOc*)hv;
    unsigned char* base = h->state.data() + (size_t)s * h->L.stride;
    const mot::OcLayout& L = h->L;
    const int* hdr = (const int*)base;
    const unsigned short* list = (const unsigned short*)(base + L.off_lists);
    const int* m = (const int*)(base + L.off_meta);
    const float* obs = (const float*)(base + L.off_obs);
    const float* recs = (const float*)(base + L.off_recs);
    const int n = hdr[mot::kOHdrTracks], cap = L.cap;
    int k = 0;
    for (; k < n && k < cap_rows; ++k) {
        const int slot = list[k];
        float* o = rows + 71 * k;
        o[0] = (float)m[slot]; o[1] = (float)m[cap + slot]; o[2] = (float)m[2 * cap + slot]; o[3] = (float)m[3 * cap + slot];
        o[4] = (float)m[4 * cap + slot]; o[5] = ((const float*)m)[7 * cap + slot]; o[6] = (float)m[5 * cap + slot];
        o[7] = (float)m[6 * cap + slot];
        std::memcpy(o + 8, obs + (size_t)slot * mot::kOcObsFloats, 7 * sizeof(float));
        std::memcpy(o + 15, recs + (size_t)slot * mot::kOcRecFloats, 56 * sizeof(float));
    }
    return k;
}

}  // extern "C"

// ------------------------------------------------------------------ DeepOC-SORT engine under the emulator
namespace {
struct SimDeepOc {
    mot::OcLayout L;
    mot::DeepLayout D;
    size_t stride;
    int S;
    mot::OcParams p;
    std::vector<unsigned char> state;
};
}  // namespace

extern "C" {

void* sim_deepoc_create(int S, int dim, float det_thresh, int max_age, int min_hits, float iou_threshold, int delta_t,
                        float inertia, float w_assoc_emb, float alpha_fixed_emb, float aw_param, int embedding_off, int aw_off,
                        float q_xy, float q_s) {
    auto* h = new SimDeepOc();
    h->L = mot::OcLayout::make(256, 64);
    h->D = mot::DeepLayout::make(256, 64, embedding_off ? 0 : dim);
    h->stride = h->L.stride + h->D.bytes;
    h->S = S;
    h->p = mot::OcParams{};
    h->p.det_thresh = det_thresh; h->p.max_age = max_age; h->p.min_hits = min_hits; h->p.iou_threshold = iou_threshold;
    h->p.min_conf = 0.1f; h->p.delta_t = delta_t; h->p.inertia = inertia; h->p.use_byte = 0;
    h->p.q44 = 0.01f * q_xy; h->p.q66 = 0.0001f * q_s;
    h->p.w_assoc_emb = w_assoc_emb; h->p.alpha_fixed_emb = alpha_fixed_emb; h->p.aw_param = aw_param;
    h->p.embedding_off = embedding_off; h->p.aw_off = aw_off;
    h->state.assign(h->stride * (size_t)S + 256, 0);
    unsigned char* st = h->state.data();
    const mot::OcLayout L = h->L;
    const size_t stride = h->stride;
    cpusim::launch(dim3(S), dim3(64), 0, [=] { mot::ocsort_reset_kernel(st, L, stride, S, 0); });
    return h;
}
void sim_deepoc_destroy(void* hv) { delete (SimDeepOc*)hv; }

int sim_deepoc_update(void* hv, const float* dets, const int* n_dets, const float* embs, int T, int ld_dets, float* out,
                      int* n_out, int ld_out, int threads) {
    auto* h = (SimDeepOc*)hv;
    mot::OcArgs a{};
    a.state = h->state.data(); a.dets = dets; a.n_dets = n_dets; a.out = out; a.n_out = n_out;
    a.T = T; a.S = h->S; a.ld_dets = ld_dets; a.ld_out = ld_out; a.s_begin = 0; a.s_end = h->S; a.p = h->p;
    a.embs = embs; a.dim = h->D.dim; a.stride = h->stride;
    const size_t smem = mot::oc_smem_bytes(256, 64, 1024);
    cpusim::launch(dim3(h->S), dim3(threads), smem, [=] { mot::deepocsort_step_kernel<256, 64, 1024>(a); });
    return 0;
}
void sim_deepoc_header(void* hv, int s, int* hdr16) {
    auto* h = (SimDeepOc*)hv;
    std::memcpy(hdr16, h->state.data() + (size_t)s * h->stride, sizeof(int) * 16);
}

// rows as sim_oc_dump (71 floats); embs (nullable): dim floats per row
int sim_deepoc_dump(void* hv, int s, float* rows, float* embs, int cap_rows) {
    auto* h = (SimDeepOc*)hv;
    unsigned char* base = h->state.data() + (size_t)s * h->stride;
    const mot::OcLayout& L = h->L;
    const int* hdr = (const int*)base;
    const unsigned short* list = (const unsigned short*)(base + L.off_lists);
    const int* m = (const int*)(base + L.off_meta);
    const float* obs = (const float*)(base + L.off_obs);
    const float* recs = (const float*)(base + L.off_recs);
    const float* temb = (const float*)(base + L.stride + h->D.off_emb);
    const int n = hdr[mot::kOHdrTracks], cap = L.cap;
    int k = 0;
    for (; k < n && k < cap_rows; ++k) {
        const int slot = list[k];
        float* o = rows + 71 * k;
        o[0] = (float)m[slot]; o[1] = (float)m[cap + slot]; o[2] = (float)m[2 * cap + slot]; o[3] = (float)m[3 * cap + slot];
        o[4] = (float)m[4 * cap + slot]; o[5] = ((const float*)m)[7 * cap + slot]; o[6] = (float)m[5 * cap + slot];
        o[7] = (float)m[6 * cap + slot];
        std::memcpy(o + 8, obs + (size_t)slot * mot::kOcObsFloats, 7 * sizeof(float));
        std::memcpy(o + 15, recs + (size_t)slot * mot::kOcRecFloats, 56 * sizeof(float));
        if (embs && h->D.dim > 0) std::memcpy(embs + (size_t)k * h->D.dim, temb + (size_t)slot * h->D.dim, sizeof(float) * h->D.dim);
    }
    return k;
}

}  // extern "C"

// ------------------------------------------------------------------ BoostTrack engine under the emulator
#include "../../motcpp_b200/csrc/boosttrack_kernel.cuh"
namespace {
struct SimBoost {
    mot::SortLayout L;
    int S;
    mot::BoostParams p;
    std::vector<unsigned char> state;
};
}  // namespace

extern "C" {

void* sim_boost_create(int S, float det_thresh, int max_age, int min_hits, float iou_threshold, int min_box_area,
                       float aspect_ratio_thresh, float lambda_mhd, int use_dlo_boost, float dlo_boost_coef, int use_vt) {
    auto* h = new SimBoost();
    h->L = mot::SortLayout::make(256, 64);
    h->S = S;
    h->p = mot::BoostParams{det_thresh, iou_threshold, aspect_ratio_thresh, lambda_mhd, dlo_boost_coef, (float)min_box_area,
                            max_age, min_hits, use_dlo_boost, use_vt};
    h->state.assign(h->L.stride * (size_t)S + 256, 0);
    unsigned char* st = h->state.data();
    const mot::SortLayout L = h->L;
    cpusim::launch(dim3(S), dim3(64), 0, [=] { mot::sort_reset_kernel(st, L, S, 0); });
    return h;
}
void sim_boost_destroy(void* hv) { delete (SimBoost*)hv; }
int sim_boost_update(void* hv, const float* dets, const int* n_dets, int T, int ld_dets, float* out, int* n_out, int ld_out, int threads) {
    auto* h = (SimBoost*)hv;
    mot::BoostArgs a{};
    a.state = h->state.data(); a.dets = dets; a.n_dets = n_dets; a.out = out; a.n_out = n_out;
    a.T = T; a.S = h->S; a.ld_dets = ld_dets; a.ld_out = ld_out; a.s_begin = 0; a.s_end = h->S; a.p = h->p;
    cpusim::launch(dim3(h->S), dim3(threads), mot::boost_smem_bytes(256, 64, 1024), [=] { mot::boosttrack_step_kernel<256, 64, 1024>(a); });
    return 0;
}
void sim_boost_header(void* hv, int s, int* hdr16) {
    auto* h = (SimBoost*)hv;
    std::memcpy(hdr16, h->state.data() + (size_t)s * h->L.stride, sizeof(int) * 16);
}
// rows of [id, age, streak, tsu, conf, cls, det_ind, 0, x 8, P 64]
int sim_boost_dump(void* hv, int s, float* rows, int cap_rows) {
    auto* h = (SimBoost*)hv;
    unsigned char* base = h->state.data() + (size_t)s * h->L.stride;
    const mot::SortLayout& L = h->L;
    const int* hdr = (const int*)base;
    const unsigned short* list = (const unsigned short*)(base + L.off_lists);
    const int* m = (const int*)(base + L.off_meta);
    const float* recs = (const float*)(base + L.off_recs);
    const int n = hdr[mot::kSHdrTracks], cap = L.cap;
    int k = 0;
    for (; k < n && k < cap_rows; ++k) {
        const int slot = list[k];
        float* o = rows + 80 * (size_t)k;
        std::memset(o, 0, 80 * sizeof(float));
        o[0] = (float)m[slot]; o[1] = (float)m[3 * cap + slot]; o[2] = (float)m[cap + slot]; o[3] = (float)m[2 * cap + slot];
        o[4] = ((const float*)m)[6 * cap + slot]; o[5] = (float)m[4 * cap + slot]; o[6] = (float)m[5 * cap + slot];
        const float* rec = recs + (size_t)slot * mot::kBoostRecFloats;
        std::memcpy(o + 8, rec, 8 * sizeof(float));
        for (int c = 0; c < 4; ++c) {
            const float* P = rec + 8 + 4 * c;
            o[16 + c * 8 + c] = P[0]; o[16 + c * 8 + c + 4] = P[1]; o[16 + (c + 4) * 8 + c] = P[2]; o[16 + (c + 4) * 8 + c + 4] = P[3];
        }
    }
    return k;
}

}  // extern "C"

// ------------------------------------------------------------------ BoT-SORT engine under the emulator
#include "../../motcpp_b200/csrc/botsort_kernel.cuh"
#include "../../motcpp_b200/csrc/strongsort_kernel.cuh"

namespace {
struct SimBot {
    mot::BotLayout L;
    int S;
    mot::BotParams p;
    std::vector<unsigned char> state;
};
}  // namespace

extern "C" {

void* sim_bot_create(int S, int dim, float high, float low, float new_thresh, int track_buffer, float match_thresh,
                     float prox, float app, int frame_rate, int fuse_first, int with_reid) {
    auto* h = new SimBot();
    h->L = mot::BotLayout::make(256, 64, dim);
    h->S = S;
    h->p.track_high_thresh = high; h->p.track_low_thresh = low; h->p.new_track_thresh = new_thresh;
    h->p.match_thresh = match_thresh; h->p.proximity_thresh = prox; h->p.appearance_thresh = app;
    h->p.max_time_lost = (int)(frame_rate / 30.0f * track_buffer);
    h->p.fuse_first = fuse_first; h->p.with_reid = with_reid; h->p.dim = dim;
    h->state.assign(h->L.stride * (size_t)S + 256, 0);
    unsigned char* st = h->state.data();
    const mot::BotLayout L = h->L;
    cpusim::launch(dim3(S), dim3(64), 0, [=] { mot::botsort_reset_kernel(st, L, S); });
    return h;
}
void sim_bot_destroy(void* hv) { delete (SimBot*)hv; }

int sim_bot_update(void* hv, const float* dets, const int* n_dets, const float* embs, int T, int ld_dets, float* out,
                   int* n_out, int ld_out, int threads) {
    auto* h = (SimBot*)hv;
    mot::BotArgs a{};
    a.state = h->state.data(); a.L = h->L; a.dets = dets; a.n_dets = n_dets; a.embs = embs; a.out = out; a.n_out = n_out;
    a.T = T; a.S = h->S; a.ld_dets = ld_dets; a.ld_out = ld_out; a.s_begin = 0; a.s_end = h->S; a.p = h->p;
    const size_t smem = mot::bot_smem_bytes(256, 64, 1024);
    cpusim::launch(dim3(h->S), dim3(threads), smem, [=] { mot::botsort_step_kernel<256, 64, 1024>(a); });
    return 0;
}
void sim_bot_header(void* hv, int s, int* hdr16) {
    auto* h = (SimBot*)hv;
    std::memcpy(hdr16, h->state.data() + (size_t)s * h->L.stride, sizeof(int) * 16);
}

// list `which` (0 active, 1 lost): rows of [id, state, is_activated, frame_id, start_frame, tracklet_len, conf, cls,
// det_ind, has_feat, mean 8, cov 64] = 82 floats; feats (nullable): smooth_feat rows of dim floats
int sim_bot_dump(void* hv, int s, int which, float* rows, float* feats, int cap_rows) {
    auto* h = (SimBot*)hv;
    unsigned char* base = h->state.data() + (size_t)s * h->L.stride;
    const mot::BotLayout& L = h->L;
    const int* hdr = (const int*)base;
    const unsigned short* list = (const unsigned short*)(base + L.off_lists) + (which == 0 ? 0 : L.cap);
    const int n = which == 0 ? hdr[mot::kHdrActive] : hdr[mot::kHdrLost];
    const unsigned char* sflag = base + L.off_sflag;
    const int* m = (const int*)(base + L.off_meta);
    const float* recs = (const float*)(base + L.off_recs);
    const float* ft = (const float*)(base + L.off_feats);
    const int cap = L.cap;
    int k = 0;
    for (; k < n && k < cap_rows; ++k) {
        const int slot = list[k];
        float* o = rows + 82 * k;
        o[0] = (float)m[slot]; o[1] = (float)(sflag[slot] & 0x0f); o[2] = (sflag[slot] & 0x10) ? 1.0f : 0.0f;
        o[3] = (float)m[2 * cap + slot]; o[4] = (float)m[3 * cap + slot]; o[5] = (float)m[cap + slot];
        o[6] = ((const float*)m)[6 * cap + slot]; o[7] = (float)m[4 * cap + slot]; o[8] = (float)m[5 * cap + slot];
        o[9] = (sflag[slot] & 0x20) ? 1.0f : 0.0f;
        std::memcpy(o + 10, recs + (size_t)slot * mot::kRecFloats, sizeof(float) * mot::kRecFloats);
        if (feats && L.dim > 0) std::memcpy(feats + (size_t)L.dim * k, ft + (size_t)slot * L.dim, sizeof(float) * L.dim);
    }
    return k;
}

}  // extern "C"

// ------------------------------------------------------------------ StrongSORT engine on the emulator
namespace {
struct SimSs {
    mot::SsLayout L;
    mot::SsParams p;
    int S;
    std::vector<unsigned char> state;
};
}  // namespace

extern "C" {

void* sim_ss_create(int S, int dim, int max_age, float min_conf, float max_cos, float max_iou, int n_init, int budget,
                    float mc_lambda, float ema_alpha) {
    auto* h = new SimSs();
    h->L = mot::SsLayout::make(256, 64, dim, budget);
    h->S = S;
    h->p.min_conf = min_conf; h->p.max_cos_dist = max_cos; h->p.max_iou_dist = max_iou; h->p.mc_lambda = mc_lambda;
    h->p.ema_alpha = ema_alpha; h->p.max_age = max_age; h->p.n_init = n_init; h->p.budget = budget; h->p.dim = dim;
    h->state.assign(h->L.stride * (size_t)S + 256, 0);
    unsigned char* st = h->state.data();
    const mot::SsLayout L = h->L;
    cpusim::launch(dim3(S), dim3(64), 0, [=] { mot::strongsort_reset_kernel(st, L, S); });
    return h;
}
void sim_ss_destroy(void* hv) { delete (SimSs*)hv; }

int sim_ss_update(void* hv, const float* dets, const int* n_dets, const float* embs, int T, int ld_dets, float* out,
                  int* n_out, int ld_out, int threads) {
    auto* h = (SimSs*)hv;
    mot::SsArgs a{};
    a.state = h->state.data(); a.L = h->L; a.dets = dets; a.n_dets = n_dets; a.embs = embs; a.out = out; a.n_out = n_out;
    a.T = T; a.S = h->S; a.ld_dets = ld_dets; a.ld_out = ld_out; a.s_begin = 0; a.s_end = h->S; a.p = h->p;
    const size_t smem = mot::ss_smem_bytes(256, 64, 1024);
    cpusim::launch(dim3(h->S), dim3(threads), smem, [=] { mot::strongsort_step_kernel<256, 64, 1024>(a); });
    return 0;
}
void sim_ss_header(void* hv, int s, int* hdr16) {
    auto* h = (SimSs*)hv;
    std::memcpy(hdr16, h->state.data() + (size_t)s * h->L.stride, sizeof(int) * 16);
}
// rows of [id, state, hits, 0, tsu, conf, cls, det_ind, has_feat, n_samples, mean 8, cov 64]; feats nullable
int sim_ss_dump(void* hv, int s, float* rows, float* feats, int cap_rows) {
    auto* h = (SimSs*)hv;
    unsigned char* base = h->state.data() + (size_t)s * h->L.stride;
    const mot::SsLayout& L = h->L;
    const int* hdr = (const int*)base;
    const unsigned short* list = (const unsigned short*)(base + L.off_lists);
    const unsigned char* state = base + L.off_state;
    const int* m = (const int*)(base + L.off_meta);
    const float* recs = (const float*)(base + L.off_recs);
    const float* ft = (const float*)(base + L.off_feat);
    const int cap = L.cap, n = hdr[mot::kHdrActive];
    int k = 0;
    for (; k < n && k < cap_rows; ++k) {
        const int slot = list[k];
        float* o = rows + 82 * k;
        o[0] = (float)m[slot]; o[1] = (float)(state[slot] & 0x0f); o[2] = (float)m[cap + slot]; o[3] = 0.0f;
        o[4] = (float)m[2 * cap + slot]; o[5] = ((const float*)m)[7 * cap + slot]; o[6] = (float)m[3 * cap + slot];
        o[7] = (float)m[4 * cap + slot]; o[8] = (state[slot] & mot::kSsHasFeat) ? 1.0f : 0.0f; o[9] = (float)m[5 * cap + slot];
        std::memcpy(o + 10, recs + (size_t)slot * mot::kRecFloats, sizeof(float) * mot::kRecFloats);
        if (feats && L.dim > 0) {
            if (state[slot] & mot::kSsHasFeat) std::memcpy(feats + (size_t)L.dim * k, ft + (size_t)slot * L.dim, sizeof(float) * L.dim);
            else std::memset(feats + (size_t)L.dim * k, 0, sizeof(float) * L.dim);
        }
    }
    return k;
}

}  // extern "C"
