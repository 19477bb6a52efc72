// cabi.cu - the extern "C" boundary (include/motb200.h) over the sm_100a kernels.
// Host code only launches kernels and moves bytes; there is no CPU implementation of any
// operator in this library.
#include "../../include/motb200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "engine_launch.h"
#include "bytetrack_kernel.cuh"
#include "kernels_cost.cuh"
#include "kernels_kf.cuh"
#include "kernels_lap.cuh"
#include "kernels_cosine.cuh"
#include "kernels_pack.cuh"
#include "sort_kernel.cuh"
#include "ocsort_kernel.cuh"
#include "boosttrack_kernel.cuh"
#include "botsort_kernel.cuh"
#include "strongsort_kernel.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define MOT_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fail(e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver ? MOT_ERR_NO_DEVICE : MOT_ERR_CUDA, \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

int require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(MOT_ERR_NO_DEVICE, "no CUDA device available (libmotb200 has no CPU fallback): %s",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return MOT_OK;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

constexpr int kMaxChunks = 32;

}  // namespace

struct mot_engine {
    mot_engine_config cfg;
    mot::BtLayout layout;          // ByteTrack slab layout (kind == BYTETRACK)
    mot::SortLayout sort_layout;   // SORT slab layout (kind == SORT)
    mot::OcLayout oc_layout;       // OC-SORT slab layout (kind == OCSORT)
    mot::OcParams ocp;
    mot::BoostParams boostp;       // BoostTrack (kind == BOOSTTRACK) lives in a SORT slab
    mot::DeepLayout deep_layout;   // DeepOC-SORT appearance state behind every OC-SORT slab (kind == DEEPOCSORT)
    mot::BotLayout bot_layout;     // BoT-SORT slab layout (kind == BOTSORT; feature dimension is a run-time size)
    mot::BotParams botp;
    mot::SsLayout ss_layout;       // StrongSORT slab layout (kind == STRONGSORT; dim and gallery budget are run-time sizes)
    mot::SsParams ssp;
    float* d_embs = nullptr;  size_t embs_cap = 0;
    size_t stride = 0;             // bytes per stream slab (whichever layout is live)
    int threads = 0;
    mot::BtParams bt;
    mot::SortParams sortp;
    int shape = 0;             // index into mot::kBtShapes
    int e_cap = 4096;
    size_t smem_bytes = 0;
    unsigned char* d_state = nullptr;
    int n_chunks = 1;
    cudaStream_t streams[kMaxChunks] = {};
    cudaEvent_t ev_in[kMaxChunks] = {}, ev_run[kMaxChunks] = {};   // frame-chunk pipeline of the host-buffer path
    // staging for the host-buffer path (grow-only)
    float* d_dets = nullptr;  size_t dets_cap = 0;
    int* d_ndets = nullptr;   size_t ndets_cap = 0;
    float* d_out = nullptr;   size_t out_cap = 0;
    int* d_nout = nullptr;    size_t nout_cap = 0;
    unsigned long long* d_prof = nullptr;    // per-phase cycle counters (mot_engine_profile), off by default
    // packed host path: compacted rows, per-frame offsets, per-chunk totals (pinned)
    float* d_packed = nullptr; size_t packed_cap = 0;
    int* d_off = nullptr;      size_t off_cap = 0;
    int* h_off = nullptr;      size_t h_off_cap = 0;   // pinned mirror of d_off
    cudaEvent_t ev_tot[kMaxChunks] = {};
};


// ---- helpers (C++ linkage)
static int engine_reset_impl(mot_engine* e, int keep_ids) {
    const int grid = std::min(e->cfg.n_streams, 4096);
    if (e->cfg.kind == MOT_TRACKER_SORT)
        mot::sort_reset_kernel<<<grid, 256, 0, e->streams[0]>>>(e->d_state, e->sort_layout, e->cfg.n_streams, keep_ids);
    else if (e->cfg.kind == MOT_TRACKER_BOOSTTRACK)      // BoostTrack::next_id_ restarts (boosttrack.cpp:272-277)
        mot::sort_reset_kernel<<<grid, 256, 0, e->streams[0]>>>(e->d_state, e->sort_layout, e->cfg.n_streams, 0);
    else if (e->cfg.kind == MOT_TRACKER_OCSORT || e->cfg.kind == MOT_TRACKER_DEEPOCSORT)
        mot::ocsort_reset_kernel<<<grid, 256, 0, e->streams[0]>>>(e->d_state, e->oc_layout, e->stride, e->cfg.n_streams, keep_ids);
    else if (e->cfg.kind == MOT_TRACKER_BOTSORT)
        mot::botsort_reset_kernel<<<grid, 256, 0, e->streams[0]>>>(e->d_state, e->bot_layout, e->cfg.n_streams);
    else if (e->cfg.kind == MOT_TRACKER_STRONGSORT)
        mot::strongsort_reset_kernel<<<grid, 256, 0, e->streams[0]>>>(e->d_state, e->ss_layout, e->cfg.n_streams);
    else
        mot::bytetrack_reset_kernel<<<grid, 256, 0, e->streams[0]>>>(e->d_state, e->layout, e->cfg.n_streams, keep_ids);
    MOT_CUDA(cudaGetLastError());
    MOT_CUDA(cudaStreamSynchronize(e->streams[0]));
    return MOT_OK;
}

// one launch covering streams [s0, s1) for T frames, whatever the tracker kind
static void engine_launch(mot_engine* e, int T, const float* dets, const int* nd, int ld_dets, const float* embs,
                          float* out, int* nout, int ld_out, int s0, int s1, cudaStream_t st);

static mot::BtArgs make_args(mot_engine* e, int T, const float* dets, const int* nd, int ld_dets, float* out, int* nout,
                             int ld_out, int s_begin, int s_end) {
    mot::BtArgs a{};
    a.state = e->d_state;
    a.dets = dets; a.n_dets = nd; a.out = out; a.n_out = nout;
    a.T = T; a.S = e->cfg.n_streams; a.s_begin = s_begin; a.s_end = s_end;
    a.ld_dets = ld_dets; a.ld_out = ld_out; a.e_cap = e->e_cap; a.p = e->bt;
    a.prof = e->d_prof;
    return a;
}

static void engine_launch(mot_engine* e, int T, const float* dets, const int* nd, int ld_dets, const float* embs,
                          float* out, int* nout, int ld_out, int s0, int s1, cudaStream_t st) {
    if (e->cfg.kind == MOT_TRACKER_STRONGSORT) {
        mot::SsArgs a{};
        a.state = e->d_state; a.L = e->ss_layout; a.dets = dets; a.n_dets = nd; a.embs = embs; a.out = out; a.n_out = nout;
        a.T = T; a.S = e->cfg.n_streams; a.ld_dets = ld_dets; a.ld_out = ld_out; a.s_begin = s0; a.s_end = s1;
        a.p = e->ssp;
        mot::ss_launch(e->shape, s1 - s0, e->smem_bytes, st, a);
    } else if (e->cfg.kind == MOT_TRACKER_BOTSORT) {
        mot::BotArgs a{};
        a.state = e->d_state; a.L = e->bot_layout; a.dets = dets; a.n_dets = nd; a.embs = embs; a.out = out; a.n_out = nout;
        a.T = T; a.S = e->cfg.n_streams; a.ld_dets = ld_dets; a.ld_out = ld_out; a.s_begin = s0; a.s_end = s1;
        a.p = e->botp;
        mot::bot_launch(e->shape, s1 - s0, e->smem_bytes, st, a);
    } else if (e->cfg.kind == MOT_TRACKER_BOOSTTRACK) {
        mot::BoostArgs a{};
        a.state = e->d_state; a.dets = dets; a.n_dets = nd; a.out = out; a.n_out = nout;
        a.T = T; a.S = e->cfg.n_streams; a.ld_dets = ld_dets; a.ld_out = ld_out; a.s_begin = s0; a.s_end = s1;
        a.p = e->boostp;
        mot::boost_launch(e->shape, s1 - s0, e->smem_bytes, st, a);
    } else if (e->cfg.kind == MOT_TRACKER_SORT) {
        mot::SortArgs a{};
        a.state = e->d_state; a.dets = dets; a.n_dets = nd; a.out = out; a.n_out = nout;
        a.T = T; a.S = e->cfg.n_streams; a.ld_dets = ld_dets; a.ld_out = ld_out; a.s_begin = s0; a.s_end = s1;
        a.p = e->sortp;
        mot::sort_launch(e->shape, s1 - s0, e->smem_bytes, st, a);
    } else if (e->cfg.kind == MOT_TRACKER_OCSORT || e->cfg.kind == MOT_TRACKER_DEEPOCSORT) {
        mot::OcArgs a{};
        a.state = e->d_state; a.dets = dets; a.n_dets = nd; a.out = out; a.n_out = nout;
        a.T = T; a.S = e->cfg.n_streams; a.ld_dets = ld_dets; a.ld_out = ld_out; a.s_begin = s0; a.s_end = s1;
        a.p = e->ocp;
        a.embs = embs; a.dim = e->deep_layout.dim; a.stride = e->stride;
        if (e->cfg.kind == MOT_TRACKER_DEEPOCSORT) mot::deepoc_launch(e->shape, s1 - s0, e->smem_bytes, st, a);
        else if (e->ocp.asso == mot::kVarCentroid) mot::oc_centroid_launch(e->shape, s1 - s0, e->smem_bytes, st, a);
        else mot::oc_launch(e->shape, s1 - s0, e->smem_bytes, st, a);
    } else {
        mot::BtArgs a = make_args(e, T, dets, nd, ld_dets, out, nout, ld_out, s0, s1);
        mot::bt_launch(e->shape, s1 - s0, e->smem_bytes, st, a, e->threads);
    }
}

template <class T>
static int grow(T** p, size_t* cap, size_t need) {
    if (*cap >= need) return MOT_OK;
    if (*p) MOT_CUDA(cudaFree(*p));
    *p = nullptr; *cap = 0;
    MOT_CUDA(cudaMalloc(p, need * sizeof(T)));
    *cap = need;
    return MOT_OK;
}

static int kf_grid(long long n) {
    const long long groups_per_block = 256 / 8;
    long long blocks = (n + groups_per_block - 1) / groups_per_block;
    const long long cap = (long long)sm_count() * 8;
    return (int)std::max<long long>(1, std::min(blocks, cap));
}


extern "C" {

const char* mot_last_error(void) { return g_err.c_str(); }
int mot_version(void) { return 100; }
int mot_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int mot_device_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(MOT_ERR_INVALID_ARGUMENT, "null out pointer");
    if (int rc = require_device()) return rc;
    MOT_CUDA(cudaMalloc(ptr, bytes ? bytes : 1));
    return MOT_OK;
}
int mot_device_free(void* ptr) { MOT_CUDA(cudaFree(ptr)); return MOT_OK; }
int mot_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(MOT_ERR_INVALID_ARGUMENT, "null out pointer");
    if (int rc = require_device()) return rc;
    MOT_CUDA(cudaMallocHost(ptr, bytes ? bytes : 1));
    return MOT_OK;
}
int mot_host_free(void* ptr) { MOT_CUDA(cudaFreeHost(ptr)); return MOT_OK; }
int mot_copy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
    MOT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return MOT_OK;
}
int mot_copy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
    MOT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return MOT_OK;
}
int mot_memset_device(void* dst, int value, size_t bytes, void* stream) {
    MOT_CUDA(cudaMemsetAsync(dst, value, bytes, (cudaStream_t)stream));
    return MOT_OK;
}
int mot_stream_sync(void* stream) { MOT_CUDA(cudaStreamSynchronize((cudaStream_t)stream)); return MOT_OK; }

// ------------------------------------------------------------------------------ engine
int mot_engine_default_config(int kind, mot_engine_config* c) {
    if (!c) return fail(MOT_ERR_INVALID_ARGUMENT, "null config");
    std::memset(c, 0, sizeof(*c));
    c->kind = kind;
    c->n_streams = 1;
    c->track_capacity = 0; c->max_dets = 0; c->device = 0; c->n_chunks = 0;
    // BaseTracker defaults (include/motcpp/tracker.hpp:47-55)
    c->det_thresh = 0.3f; c->max_age = 30; c->max_obs = 50; c->min_hits = 3; c->iou_threshold = 0.3f;
    // ByteTrack (bytetrack.hpp:97-110)
    c->min_conf = 0.1f; c->track_thresh = 0.45f; c->match_thresh = 0.8f; c->track_buffer = 25; c->frame_rate = 30;
    // OCSort (ocsort.hpp:88-102)
    c->delta_t = 3; c->inertia = 0.2f; c->use_byte = 0; c->q_xy_scaling = 0.01f; c->q_s_scaling = 0.0001f;
    // BotSort (botsort.hpp:108-134)
    c->track_high_thresh = 0.5f; c->track_low_thresh = 0.1f; c->new_track_thresh = 0.6f;
    c->proximity_thresh = 0.5f; c->appearance_thresh = 0.25f; c->fuse_first_associate = 0; c->with_reid = 1;
    c->emb_dim = 0;
    // StrongSORT (strongsort.hpp:287-305)
    c->max_cos_dist = 0.2f; c->max_iou_dist = 0.7f; c->n_init = 3; c->nn_budget = 100; c->mc_lambda = 0.98f; c->ema_alpha = 0.9f;
    // DeepOCSort (deepocsort.hpp:93-117)
    c->w_association_emb = 0.5f; c->alpha_fixed_emb = 0.95f; c->aw_param = 0.5f; c->embedding_off = 0; c->aw_off = 0;
    c->asso_func = 0; c->frame_width = 0; c->frame_height = 0;
    // BoostTrackTracker (boosttrack.hpp:95-124)
    c->min_box_area = 10; c->aspect_ratio_thresh = 1.6f; c->lambda_iou = 0.5f; c->lambda_mhd = 0.25f; c->lambda_shape = 0.25f;
    c->use_dlo_boost = 1; c->dlo_boost_coef = 0.65f; c->use_sb = 0; c->use_vt = 0;
    switch (kind) {
        case MOT_TRACKER_SORT: c->max_age = 1; break;             // sort.hpp:70
        case MOT_TRACKER_BYTETRACK: break;
        case MOT_TRACKER_OCSORT: c->det_thresh = 0.2f; break;
        case MOT_TRACKER_BOTSORT: c->track_buffer = 30; break;
        case MOT_TRACKER_STRONGSORT: break;
        case MOT_TRACKER_DEEPOCSORT: break;
        case MOT_TRACKER_BOOSTTRACK: c->det_thresh = 0.6f; c->max_age = 60; break;
        default: return fail(MOT_ERR_INVALID_ARGUMENT, "unknown tracker kind %d", kind);
    }
    return MOT_OK;
}

int mot_engine_create(const mot_engine_config* cfg, mot_engine** out) {
    if (!cfg || !out) return fail(MOT_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (cfg->kind < MOT_TRACKER_SORT || cfg->kind > MOT_TRACKER_BOOSTTRACK)
        return fail(MOT_ERR_INVALID_ARGUMENT, "unknown tracker kind %d", cfg->kind);
    if (cfg->kind == MOT_TRACKER_STRONGSORT && (cfg->nn_budget < 1 || cfg->nn_budget > 4096))
        return fail(MOT_ERR_UNSUPPORTED, "nn_budget %d is outside 1..4096 (gallery ring size; the reference's unlimited budget is not supported)", cfg->nn_budget);
    if ((cfg->kind == MOT_TRACKER_BOTSORT || cfg->kind == MOT_TRACKER_STRONGSORT) && (cfg->emb_dim < 0 || (cfg->emb_dim & 3)))
        return fail(MOT_ERR_INVALID_ARGUMENT, "emb_dim %d must be a non-negative multiple of 4", cfg->emb_dim);
    if (cfg->n_streams <= 0) return fail(MOT_ERR_INVALID_ARGUMENT, "n_streams must be positive");
    if (cfg->kind == MOT_TRACKER_BOOSTTRACK && cfg->use_sb)
        return fail(MOT_ERR_UNSUPPORTED, "BoostTrack's use_sb confidence boost (std::pow(iou, 1.5f)) is not built; use_dlo_boost / use_vt are");
    if (cfg->asso_func != 0) {
        if (cfg->asso_func != mot::kVarCentroid)
            return fail(MOT_ERR_UNSUPPORTED, "Invalid association mode: %d (engines take 0 \"iou\" or 6 \"centroid\"; the reference's hmiou / giou / diou / ciou are only defined for one-row box sets)", cfg->asso_func);
        if (cfg->kind != MOT_TRACKER_OCSORT)
            return fail(MOT_ERR_UNSUPPORTED, "asso_func \"centroid\" is wired into the OC-SORT engine only");
        if (cfg->frame_width <= 0 || cfg->frame_height <= 0)
            return fail(MOT_ERR_INVALID_ARGUMENT, "asso_func \"centroid\" needs frame_width / frame_height (the reference reads them from img)");
    }
    if (cfg->kind == MOT_TRACKER_DEEPOCSORT && !cfg->embedding_off && cfg->emb_dim < 1)
        return fail(MOT_ERR_INVALID_ARGUMENT, "a DeepOC-SORT engine needs emb_dim >= 1 (or embedding_off = 1)");
    if ((cfg->kind == MOT_TRACKER_OCSORT || cfg->kind == MOT_TRACKER_DEEPOCSORT) && (cfg->delta_t < 1 || cfg->delta_t > mot::kOcRing - 1))
        return fail(MOT_ERR_UNSUPPORTED, "delta_t %d is outside 1..%d (the observation ring keeps the current age and the %d before it)", cfg->delta_t, mot::kOcRing - 1, mot::kOcRing - 1);
    if (int rc = require_device()) return rc;
    int prev_device = 0;
    MOT_CUDA(cudaGetDevice(&prev_device));
    MOT_CUDA(cudaSetDevice(cfg->device));
    mot_engine* e = new mot_engine();
    e->cfg = *cfg;
    struct Guard {                       // every early return below frees the half-built engine and restores the device
        mot_engine** e; int dev;
        ~Guard() { if (*e) mot_engine_destroy(*e); cudaSetDevice(dev); }
    } guard{&e, prev_device};
    const bool is_deep = cfg->kind == MOT_TRACKER_DEEPOCSORT;
    const bool is_boost = cfg->kind == MOT_TRACKER_BOOSTTRACK;
    const bool is_sort = cfg->kind == MOT_TRACKER_SORT || is_boost /* same slab layout and shapes */, is_oc = cfg->kind == MOT_TRACKER_OCSORT || is_deep;   // same kernel text and shapes
    const bool is_bot = cfg->kind == MOT_TRACKER_BOTSORT, is_ss = cfg->kind == MOT_TRACKER_STRONGSORT;
    if (e->cfg.track_capacity <= 0) e->cfg.track_capacity = 1536;
    if (e->cfg.max_dets <= 0) e->cfg.max_dets = 512;
    // round the request up to the nearest shape the kernel is instantiated for
    e->shape = -1;
    if (is_oc) {
        for (int i = 0; i < mot::kNumOcShapes; ++i)
            if (mot::kOcShapes[i].cap >= e->cfg.track_capacity && mot::kOcShapes[i].d_max >= e->cfg.max_dets) { e->shape = i; break; }
    } else if (is_ss) {
        for (int i = 0; i < mot::kNumSsShapes; ++i)
            if (mot::kSsShapes[i].cap >= e->cfg.track_capacity && mot::kSsShapes[i].d_max >= e->cfg.max_dets) { e->shape = i; break; }
    } else if (is_bot) {
        for (int i = 0; i < mot::kNumBotShapes; ++i)
            if (mot::kBotShapes[i].cap >= e->cfg.track_capacity && mot::kBotShapes[i].d_max >= e->cfg.max_dets) { e->shape = i; break; }
    } else {
        for (int i = 0; i < mot::kNumBtShapes; ++i)
            if (mot::kBtShapes[i].cap >= e->cfg.track_capacity && mot::kBtShapes[i].d_max >= e->cfg.max_dets) { e->shape = i; break; }
    }
    if (e->shape < 0) {
        const int tc = e->cfg.track_capacity, md = e->cfg.max_dets;
        return fail(MOT_ERR_INVALID_ARGUMENT, "track_capacity %d / max_dets %d exceed the largest built shape (%s)", tc, md,
                    is_oc ? "3072 tracks / 2048 detections" : (is_bot ? "2048 tracks / 1024 detections" : (is_ss ? "1536 tracks / 512 detections" : "3072 tracks / 1024 detections")));
    }
    if (is_ss) {
        e->cfg.track_capacity = mot::kSsShapes[e->shape].cap; e->cfg.max_dets = mot::kSsShapes[e->shape].d_max; e->e_cap = mot::kSsShapes[e->shape].e_cap;
    } else {
    e->cfg.track_capacity = is_oc ? mot::kOcShapes[e->shape].cap : (is_bot ? mot::kBotShapes[e->shape].cap : mot::kBtShapes[e->shape].cap);
    e->cfg.max_dets = is_oc ? mot::kOcShapes[e->shape].d_max : (is_bot ? mot::kBotShapes[e->shape].d_max : mot::kBtShapes[e->shape].d_max);
    e->e_cap = is_oc ? mot::kOcShapes[e->shape].e_cap : (is_bot ? mot::kBotShapes[e->shape].e_cap : mot::kBtShapes[e->shape].e_cap);
    }
    // BaseTracker ctor fix-up (src/tracker.cpp:37-39)
    if (e->cfg.max_age >= e->cfg.max_obs) e->cfg.max_obs = e->cfg.max_age + 5;
    e->layout = mot::BtLayout::make(e->cfg.track_capacity, e->cfg.max_dets);
    e->sort_layout = mot::SortLayout::make(e->cfg.track_capacity, e->cfg.max_dets);
    e->oc_layout = mot::OcLayout::make(e->cfg.track_capacity, e->cfg.max_dets);
    e->bt.min_conf = cfg->min_conf;
    e->bt.track_thresh = cfg->track_thresh;
    e->bt.match_thresh = cfg->match_thresh;
    e->bt.det_thresh = cfg->track_thresh;                                        // bytetrack.cpp:145
    e->bt.max_time_lost = (int)(cfg->frame_rate / 30.0f * cfg->track_buffer);    // bytetrack.cpp:141-142
    e->sortp.det_thresh = cfg->det_thresh;
    e->sortp.iou_threshold = cfg->iou_threshold;
    e->sortp.max_age = cfg->max_age;
    e->sortp.min_hits = cfg->min_hits;
    e->boostp.det_thresh = cfg->det_thresh; e->boostp.iou_threshold = cfg->iou_threshold;
    e->boostp.aspect_ratio_thresh = cfg->aspect_ratio_thresh; e->boostp.lambda_mhd = cfg->lambda_mhd;
    e->boostp.dlo_boost_coef = cfg->dlo_boost_coef; e->boostp.min_box_area = (float)cfg->min_box_area;
    e->boostp.max_age = cfg->max_age; e->boostp.min_hits = cfg->min_hits; e->boostp.use_dlo_boost = cfg->use_dlo_boost;
    e->boostp.use_vt = cfg->use_vt;
    e->ocp.det_thresh = cfg->det_thresh;
    e->ocp.iou_threshold = cfg->iou_threshold;                                   // asso_threshold_ (ocsort.cpp:195)
    e->ocp.min_conf = cfg->min_conf;
    e->ocp.inertia = cfg->inertia;
    e->ocp.q44 = 0.01f * cfg->q_xy_scaling;                                      // xysr_kf.cpp:58-61 then ocsort.cpp:77-79, in fp32
    e->ocp.q66 = 0.0001f * cfg->q_s_scaling;
    e->ocp.max_age = cfg->max_age;
    e->ocp.min_hits = cfg->min_hits;
    e->ocp.delta_t = cfg->delta_t;
    e->ocp.use_byte = is_deep ? 0 : cfg->use_byte;                               // DeepOC-SORT has no BYTE pass
    e->ocp.w_assoc_emb = cfg->w_association_emb;
    e->ocp.alpha_fixed_emb = cfg->alpha_fixed_emb;
    e->ocp.aw_param = cfg->aw_param;
    e->ocp.aw_off = cfg->aw_off;
    e->ocp.embedding_off = cfg->embedding_off;
    e->ocp.asso = cfg->asso_func;
    e->ocp.asso_norm = cfg->asso_func ? static_cast<float>(std::sqrt((double)(cfg->frame_width * cfg->frame_width + cfg->frame_height * cfg->frame_height))) : 1.0f;   // iou.hpp:325
    e->deep_layout = mot::DeepLayout::make(e->cfg.track_capacity, e->cfg.max_dets, cfg->embedding_off ? 0 : cfg->emb_dim);
    e->bot_layout = mot::BotLayout::make(e->cfg.track_capacity, e->cfg.max_dets, cfg->emb_dim);
    e->botp.track_high_thresh = cfg->track_high_thresh;
    e->botp.track_low_thresh = cfg->track_low_thresh;
    e->botp.new_track_thresh = cfg->new_track_thresh;
    e->botp.match_thresh = cfg->match_thresh;
    e->botp.proximity_thresh = cfg->proximity_thresh;
    e->botp.appearance_thresh = cfg->appearance_thresh;
    e->botp.max_time_lost = (int)(cfg->frame_rate / 30.0f * cfg->track_buffer);  // botsort.cpp:235-236
    e->botp.fuse_first = cfg->fuse_first_associate;
    e->botp.with_reid = cfg->with_reid;
    e->botp.dim = cfg->emb_dim;
    e->ss_layout = mot::SsLayout::make(e->cfg.track_capacity, e->cfg.max_dets, cfg->emb_dim, std::max(1, cfg->nn_budget));
    e->ssp.min_conf = cfg->min_conf; e->ssp.max_cos_dist = cfg->max_cos_dist; e->ssp.max_iou_dist = cfg->max_iou_dist;
    e->ssp.mc_lambda = cfg->mc_lambda; e->ssp.ema_alpha = cfg->ema_alpha; e->ssp.max_age = cfg->max_age;
    e->ssp.n_init = cfg->n_init; e->ssp.budget = std::max(1, cfg->nn_budget); e->ssp.dim = cfg->emb_dim;
    e->stride = is_ss ? e->ss_layout.stride : is_sort ? e->sort_layout.stride : (is_oc ? e->oc_layout.stride + (is_deep ? e->deep_layout.bytes : 0) : (is_bot ? e->bot_layout.stride : e->layout.stride));
    e->threads = is_ss ? mot::kSsThreads : is_boost ? mot::kBoostThreads : is_sort ? mot::kSortThreads : (is_oc ? mot::kOcThreads : (is_bot ? mot::kBotThreads : mot::bt_threads(e->shape, cfg->n_streams, sm_count())));
    e->smem_bytes = is_ss ? mot::ss_smem_bytes(e->layout.cap, e->layout.d_max, e->e_cap) : is_boost ? mot::boost_smem_bytes(e->layout.cap, e->layout.d_max, e->e_cap) : is_sort ? mot::sort_smem_bytes(e->layout.cap, e->layout.d_max, e->e_cap)
                  : is_oc   ? mot::oc_smem_bytes(e->layout.cap, e->layout.d_max, e->e_cap)
                  : is_bot  ? mot::bot_smem_bytes(e->layout.cap, e->layout.d_max, e->e_cap)
                            : mot::bt_smem_bytes(e->layout.cap, e->layout.d_max, e->e_cap);
    int max_optin = 0;
    MOT_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device));
    if (e->smem_bytes > (size_t)max_optin) {
        const size_t need = e->smem_bytes;
        return fail(MOT_ERR_INVALID_ARGUMENT, "track_capacity/max_dets need %zu B of shared memory per CTA (limit %d)",
                    need, max_optin);
    }
    MOT_CUDA(is_ss ? mot::ss_prepare(e->shape, e->smem_bytes) : is_boost ? mot::boost_prepare(e->shape, e->smem_bytes) : is_sort ? mot::sort_prepare(e->shape, e->smem_bytes)
                     : (is_oc ? (is_deep ? mot::deepoc_prepare(e->shape, e->smem_bytes)
                                        : (cfg->asso_func == mot::kVarCentroid ? mot::oc_centroid_prepare(e->shape, e->smem_bytes) : mot::oc_prepare(e->shape, e->smem_bytes)))
                              : (is_bot ? mot::bot_prepare(e->shape, e->smem_bytes) : mot::bt_prepare(e->shape, e->smem_bytes, e->threads))));
    e->n_chunks = cfg->n_chunks > 0 ? std::min(cfg->n_chunks, kMaxChunks) : (cfg->n_streams >= 128 ? 8 : (cfg->n_streams >= 32 ? 4 : 1));
    e->n_chunks = std::min(e->n_chunks, cfg->n_streams);
    int prio_lo = 0, prio_hi = 0;
    MOT_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    for (int c = 0; c < kMaxChunks; ++c) {
        // streams[2] (copy-out + row compaction of the host paths) outranks the compute stream: its small kernels must not
        // queue behind the next chunk's step kernel, which fills every SM
        if (c == 2) MOT_CUDA(cudaStreamCreateWithPriority(&e->streams[c], cudaStreamNonBlocking, prio_hi));
        else MOT_CUDA(cudaStreamCreateWithFlags(&e->streams[c], cudaStreamNonBlocking));
        MOT_CUDA(cudaEventCreateWithFlags(&e->ev_in[c], cudaEventDisableTiming));
        MOT_CUDA(cudaEventCreateWithFlags(&e->ev_run[c], cudaEventDisableTiming));
    }
    MOT_CUDA(cudaMalloc(&e->d_state, e->stride * (size_t)cfg->n_streams));
    MOT_CUDA(cudaMemsetAsync(e->d_state, 0, e->stride * (size_t)cfg->n_streams, e->streams[0]));
    if (int rc = engine_reset_impl(e, 0)) return rc;
    *out = e;
    e = nullptr;                         // ownership passes to the caller; the guard only restores the device
    return MOT_OK;
}

int mot_engine_destroy(mot_engine* e) {
    if (!e) return MOT_OK;
    cudaSetDevice(e->cfg.device);
    for (int c = 0; c < kMaxChunks; ++c) {
        if (e->streams[c]) cudaStreamDestroy(e->streams[c]);
        if (e->ev_in[c]) cudaEventDestroy(e->ev_in[c]);
        if (e->ev_run[c]) cudaEventDestroy(e->ev_run[c]);
        if (e->ev_tot[c]) cudaEventDestroy(e->ev_tot[c]);
    }
    cudaFree(e->d_packed); cudaFree(e->d_off); cudaFree(e->d_prof);
    if (e->h_off) cudaFreeHost(e->h_off);
    cudaFree(e->d_state); cudaFree(e->d_embs); cudaFree(e->d_dets); cudaFree(e->d_ndets); cudaFree(e->d_out); cudaFree(e->d_nout);
    delete e;
    return MOT_OK;
}

int mot_engine_reset(mot_engine* e) {
    if (!e) return fail(MOT_ERR_INVALID_ARGUMENT, "null engine");
    MOT_CUDA(cudaSetDevice(e->cfg.device));
    for (int c = 0; c < kMaxChunks; ++c) MOT_CUDA(cudaStreamSynchronize(e->streams[c]));
    return engine_reset_impl(e, /*keep_ids=*/1);     // BoT-SORT ignores the flag: its ids restart (botsort.cpp:257)
}

int mot_engine_update_device_embs(mot_engine* e, int T, const float* d_dets, const int* d_n_dets, int ld_dets,
                                  const float* d_embs, float* d_out, int* d_n_out, int ld_out, void* stream) {
    if (!e || !d_dets || !d_n_dets || !d_out || !d_n_out) return fail(MOT_ERR_INVALID_ARGUMENT, "null argument");
    if (T <= 0 || ld_dets <= 0 || ld_out <= 0) return fail(MOT_ERR_INVALID_ARGUMENT, "non-positive size");
    if ((ld_out * 8 * sizeof(float)) % 16 != 0 || (((size_t)d_out) & 15)) return fail(MOT_ERR_INVALID_ARGUMENT, "out must be 16-byte aligned");
    if (d_embs && ((e->cfg.kind != MOT_TRACKER_BOTSORT && e->cfg.kind != MOT_TRACKER_STRONGSORT && e->cfg.kind != MOT_TRACKER_DEEPOCSORT) || e->cfg.emb_dim <= 0))
        return fail(MOT_ERR_INVALID_ARGUMENT, "embeddings need a BoT-SORT / StrongSORT / DeepOC-SORT engine created with emb_dim > 0");
    if (!d_embs && e->cfg.kind == MOT_TRACKER_DEEPOCSORT && !e->cfg.embedding_off)
        return fail(MOT_ERR_INVALID_ARGUMENT, "a DeepOC-SORT engine with embedding_off = 0 needs the detections' embeddings (the reference would run its ReID network here)");
    if (d_embs && (((size_t)d_embs) & 15)) return fail(MOT_ERR_INVALID_ARGUMENT, "embs must be 16-byte aligned");
    MOT_CUDA(cudaSetDevice(e->cfg.device));          // a NULL stream must mean the ENGINE's device
    const int S = e->cfg.n_streams;
    engine_launch(e, T, d_dets, d_n_dets, ld_dets, d_embs, d_out, d_n_out, ld_out, 0, S, (cudaStream_t)stream);
    MOT_CUDA(cudaGetLastError());
    return MOT_OK;
}

int mot_engine_update_device(mot_engine* e, int T, const float* d_dets, const int* d_n_dets, int ld_dets,
                             float* d_out, int* d_n_out, int ld_out, void* stream) {
    return mot_engine_update_device_embs(e, T, d_dets, d_n_dets, ld_dets, nullptr, d_out, d_n_out, ld_out, stream);
}

int mot_engine_update_host_embs(mot_engine* e, int T, const float* dets, const int* n_dets, int ld_dets, const float* embs,
                                float* out, int* n_out, int ld_out) {
    if (!e || !dets || !n_dets || !out || !n_out) return fail(MOT_ERR_INVALID_ARGUMENT, "null argument");
    if (T <= 0 || ld_dets <= 0 || ld_out <= 0) return fail(MOT_ERR_INVALID_ARGUMENT, "non-positive size");
    if (embs && ((e->cfg.kind != MOT_TRACKER_BOTSORT && e->cfg.kind != MOT_TRACKER_STRONGSORT && e->cfg.kind != MOT_TRACKER_DEEPOCSORT) || e->cfg.emb_dim <= 0))
        return fail(MOT_ERR_INVALID_ARGUMENT, "embeddings need a BoT-SORT / StrongSORT / DeepOC-SORT engine created with emb_dim > 0");
    if (!embs && e->cfg.kind == MOT_TRACKER_DEEPOCSORT && !e->cfg.embedding_off)
        return fail(MOT_ERR_INVALID_ARGUMENT, "a DeepOC-SORT engine with embedding_off = 0 needs the detections' embeddings (the reference would run its ReID network here)");
    MOT_CUDA(cudaSetDevice(e->cfg.device));
    const int S = e->cfg.n_streams;
    const size_t TS = (size_t)T * S;
    const size_t dim = embs ? (size_t)e->cfg.emb_dim : 0;
    if (int rc = grow(&e->d_dets, &e->dets_cap, TS * ld_dets * 6)) return rc;
    if (int rc = grow(&e->d_ndets, &e->ndets_cap, TS)) return rc;
    if (int rc = grow(&e->d_out, &e->out_cap, TS * ld_out * 8)) return rc;
    if (int rc = grow(&e->d_nout, &e->nout_cap, TS)) return rc;
    if (embs) if (int rc = grow(&e->d_embs, &e->embs_cap, TS * ld_dets * dim)) return rc;
    if (T >= 4) {
        // Many frames per call: pipeline over FRAME chunks.  A stream's frames are sequential on its CTA, so a launch
        // lasts T x (time per frame) whatever the number of streams; cutting T lets the copies of chunk c+1 / c-1
        // overlap the kernel of chunk c (streams[0] computes, streams[1] copies in, streams[2] copies out).
        const int C = std::min(kMaxChunks, T / 2);
        cudaStream_t s_run = e->streams[0], s_in = e->streams[1], s_out = e->streams[2];
        const size_t det_fr = (size_t)S * ld_dets * 6, out_fr = (size_t)S * ld_out * 8, emb_fr = (size_t)S * ld_dets * dim;
        for (int c = 0; c < C; ++c) {
            const int t0 = (int)((long long)T * c / C), t1 = (int)((long long)T * (c + 1) / C), nt = t1 - t0;
            MOT_CUDA(cudaMemcpyAsync(e->d_dets + t0 * det_fr, dets + t0 * det_fr, nt * det_fr * sizeof(float), cudaMemcpyHostToDevice, s_in));
            MOT_CUDA(cudaMemcpyAsync(e->d_ndets + (size_t)t0 * S, n_dets + (size_t)t0 * S, (size_t)nt * S * sizeof(int), cudaMemcpyHostToDevice, s_in));
            if (embs)
                MOT_CUDA(cudaMemcpyAsync(e->d_embs + t0 * emb_fr, embs + t0 * emb_fr, nt * emb_fr * sizeof(float), cudaMemcpyHostToDevice, s_in));
            MOT_CUDA(cudaEventRecord(e->ev_in[c], s_in));
            MOT_CUDA(cudaStreamWaitEvent(s_run, e->ev_in[c], 0));
            engine_launch(e, nt, e->d_dets + t0 * det_fr, e->d_ndets + (size_t)t0 * S, ld_dets, embs ? e->d_embs + t0 * emb_fr : nullptr,
                          e->d_out + t0 * out_fr, e->d_nout + (size_t)t0 * S, ld_out, 0, S, s_run);
            MOT_CUDA(cudaGetLastError());
            MOT_CUDA(cudaEventRecord(e->ev_run[c], s_run));
            MOT_CUDA(cudaStreamWaitEvent(s_out, e->ev_run[c], 0));
            MOT_CUDA(cudaMemcpyAsync(out + t0 * out_fr, e->d_out + t0 * out_fr, nt * out_fr * sizeof(float), cudaMemcpyDeviceToHost, s_out));
            MOT_CUDA(cudaMemcpyAsync(n_out + (size_t)t0 * S, e->d_nout + (size_t)t0 * S, (size_t)nt * S * sizeof(int), cudaMemcpyDeviceToHost, s_out));
        }
        MOT_CUDA(cudaStreamSynchronize(s_out));
        return MOT_OK;
    }
    // Few frames per call: pipeline over STREAM chunks instead (each chunk on its own CUDA stream)
    const int C = e->n_chunks;
    for (int c = 0; c < C; ++c) {
        const int s0 = (int)((long long)S * c / C), s1 = (int)((long long)S * (c + 1) / C);
        if (s1 <= s0) continue;
        cudaStream_t st = e->streams[c];
        const size_t det_row = (size_t)ld_dets * 6 * sizeof(float), out_row = (size_t)ld_out * 8 * sizeof(float);
        const size_t emb_row = (size_t)ld_dets * dim * sizeof(float);
        // [T][S][...] -> a (T x chunk) sub-block is a 2-D copy with pitch S * row
        MOT_CUDA(cudaMemcpy2DAsync(e->d_dets + (size_t)s0 * ld_dets * 6, S * det_row, dets + (size_t)s0 * ld_dets * 6,
                                   S * det_row, (s1 - s0) * det_row, T, cudaMemcpyHostToDevice, st));
        MOT_CUDA(cudaMemcpy2DAsync(e->d_ndets + s0, S * sizeof(int), n_dets + s0, S * sizeof(int),
                                   (s1 - s0) * sizeof(int), T, cudaMemcpyHostToDevice, st));
        if (embs)
            MOT_CUDA(cudaMemcpy2DAsync(e->d_embs + (size_t)s0 * ld_dets * dim, S * emb_row, embs + (size_t)s0 * ld_dets * dim,
                                       S * emb_row, (s1 - s0) * emb_row, T, cudaMemcpyHostToDevice, st));
        engine_launch(e, T, e->d_dets, e->d_ndets, ld_dets, embs ? e->d_embs : nullptr, e->d_out, e->d_nout, ld_out, s0, s1, st);
        MOT_CUDA(cudaGetLastError());
        MOT_CUDA(cudaMemcpy2DAsync(out + (size_t)s0 * ld_out * 8, S * out_row, e->d_out + (size_t)s0 * ld_out * 8,
                                   S * out_row, (s1 - s0) * out_row, T, cudaMemcpyDeviceToHost, st));
        MOT_CUDA(cudaMemcpy2DAsync(n_out + s0, S * sizeof(int), e->d_nout + s0, S * sizeof(int),
                                   (s1 - s0) * sizeof(int), T, cudaMemcpyDeviceToHost, st));
    }
    for (int c = 0; c < C; ++c) MOT_CUDA(cudaStreamSynchronize(e->streams[c]));
    return MOT_OK;
}

// Host buffers in, PACKED rows out: only the valid rows cross the bus (on the C2 workload 361 of the 512 padded rows).
int mot_engine_update_host_packed(mot_engine* e, int T, const float* dets, const int* n_dets, int ld_dets, int max_rows,
                                  float* out_rows, long long out_cap_rows, long long* offsets, int* n_out) {
    if (!e || !dets || !n_dets || !out_rows || !offsets || !n_out) return fail(MOT_ERR_INVALID_ARGUMENT, "null argument");
    if (T <= 0 || ld_dets <= 0 || max_rows <= 0 || out_cap_rows < 0) return fail(MOT_ERR_INVALID_ARGUMENT, "non-positive size");
    if (e->cfg.kind == MOT_TRACKER_BOTSORT || e->cfg.kind == MOT_TRACKER_STRONGSORT)
        if (e->cfg.emb_dim > 0) return fail(MOT_ERR_UNSUPPORTED, "the packed host path carries no embeddings; use mot_engine_update_host_embs");
    if (e->cfg.kind == MOT_TRACKER_DEEPOCSORT && !e->cfg.embedding_off)
        return fail(MOT_ERR_UNSUPPORTED, "the packed host path carries no embeddings; use mot_engine_update_host_embs");
    MOT_CUDA(cudaSetDevice(e->cfg.device));
    const int S = e->cfg.n_streams, ld_out = max_rows;
    const size_t TS = (size_t)T * S;
    // frames per pipeline chunk: 1-2 is the plateau on the C2 workload (1.79 M frames/s; 3 -> 1.75, 5 -> 1.72, 8 -> 1.65 M);
    // what is left between this path and the device-resident rate (12 %) is the compaction kernels, which cannot co-reside
    // with the frame-step kernel (its two CTAs per SM take the whole register file) and so run between its launches.
    // (Also measured: an L2 persisting window over the tracker state - no effect.)
    int chunk_frames = 2;
    if (const char* ev = std::getenv("MOT_PACKED_CHUNK_FRAMES")) chunk_frames = std::max(1, std::atoi(ev));   // measurement aid
    const int C = std::max(1, std::min(kMaxChunks, T / chunk_frames));
    if (int rc = grow(&e->d_dets, &e->dets_cap, TS * ld_dets * 6)) return rc;
    if (int rc = grow(&e->d_ndets, &e->ndets_cap, TS)) return rc;
    if (int rc = grow(&e->d_out, &e->out_cap, TS * ld_out * 8)) return rc;
    if (int rc = grow(&e->d_nout, &e->nout_cap, TS)) return rc;
    cudaStream_t s_run = e->streams[0], s_in = e->streams[1], s_out = e->streams[2], s_rows = e->streams[3];
    const size_t det_fr = (size_t)S * ld_dets * 6, out_fr = (size_t)S * ld_out * 8;
    auto t_of = [&](int c) { return (int)((long long)T * c / C); };
    // Rows are compacted into a device staging buffer and leave through the copy engine, chunk by chunk.  (Storing them
    // straight into pinned host memory from the compaction kernel was measured and is slower: SM-issued PCIe writes hold
    // the CTAs for the whole round trip - 1.15 M frames/s against 1.45 M on the C2 workload.)
    if (int rc = grow(&e->d_packed, &e->packed_cap, TS * ld_out * 8)) return rc;
    if (int rc = grow(&e->d_off, &e->off_cap, TS + (size_t)C)) return rc;
    if (e->h_off_cap < TS + (size_t)C) {
        if (e->h_off) cudaFreeHost(e->h_off);
        e->h_off = nullptr; e->h_off_cap = 0;
        MOT_CUDA(cudaHostAlloc((void**)&e->h_off, sizeof(int) * (TS + (size_t)C), cudaHostAllocDefault));
        e->h_off_cap = TS + (size_t)C;
    }
    for (int c = 0; c < C; ++c)
        if (!e->ev_tot[c]) MOT_CUDA(cudaEventCreateWithFlags(&e->ev_tot[c], cudaEventDisableTiming));
    long long base = 0;                                    // rows already handed to the caller
    auto enqueue = [&](int c) -> int {
        const int t0 = t_of(c), nt = t_of(c + 1) - t0, nf = nt * S;
        MOT_CUDA(cudaMemcpyAsync(e->d_dets + t0 * det_fr, dets + t0 * det_fr, nt * det_fr * sizeof(float), cudaMemcpyHostToDevice, s_in));
        MOT_CUDA(cudaMemcpyAsync(e->d_ndets + (size_t)t0 * S, n_dets + (size_t)t0 * S, (size_t)nf * sizeof(int), cudaMemcpyHostToDevice, s_in));
        MOT_CUDA(cudaEventRecord(e->ev_in[c], s_in));
        MOT_CUDA(cudaStreamWaitEvent(s_run, e->ev_in[c], 0));
        engine_launch(e, nt, e->d_dets + t0 * det_fr, e->d_ndets + (size_t)t0 * S, ld_dets, nullptr, e->d_out + t0 * out_fr,
                      e->d_nout + (size_t)t0 * S, ld_out, 0, S, s_run);
        MOT_CUDA(cudaGetLastError());
        MOT_CUDA(cudaEventRecord(e->ev_run[c], s_run));
        // the compaction runs on the copy-out stream: the next chunk's step kernel starts at once
        MOT_CUDA(cudaStreamWaitEvent(s_out, e->ev_run[c], 0));
        int* off = e->d_off + (size_t)t0 * S + c;          // nf + 1 ints per chunk
        mot::pack_scan_kernel<<<1, 1024, 0, s_out>>>(e->d_nout + (size_t)t0 * S, nf, ld_out, off);
        mot::pack_rows_kernel<<<std::min(nf, sm_count() * 4), 256, 0, s_out>>>((const float4*)(e->d_out + t0 * out_fr), e->d_nout + (size_t)t0 * S,
                                                                                 off, nf, ld_out, (float4*)(e->d_packed + t0 * out_fr));
        MOT_CUDA(cudaGetLastError());
        MOT_CUDA(cudaMemcpyAsync(e->h_off + (size_t)t0 * S + c, off, (size_t)(nf + 1) * sizeof(int), cudaMemcpyDeviceToHost, s_out));
        MOT_CUDA(cudaMemcpyAsync(n_out + (size_t)t0 * S, e->d_nout + (size_t)t0 * S, (size_t)nf * sizeof(int), cudaMemcpyDeviceToHost, s_out));
        MOT_CUDA(cudaEventRecord(e->ev_tot[c], s_out));
        return MOT_OK;
    };
    auto drain = [&](int c) -> int {                       // the chunk's row count is known: copy exactly that many rows
        const int t0 = t_of(c), nf = (t_of(c + 1) - t0) * S;
        MOT_CUDA(cudaEventSynchronize(e->ev_tot[c]));
        const int* off = e->h_off + (size_t)t0 * S + c;
        const long long rows = off[nf];
        if (base + rows > out_cap_rows) return fail(MOT_ERR_INVALID_ARGUMENT, "out_rows holds %lld rows, %lld needed so far", out_cap_rows, base + rows);
        if (rows > 0)     // own stream: s_out already holds the compaction of the next two chunks (the event implies pack(c) is done)
            MOT_CUDA(cudaMemcpyAsync(out_rows + base * 8, e->d_packed + t0 * out_fr, (size_t)rows * 8 * sizeof(float), cudaMemcpyDeviceToHost, s_rows));
        for (int f = 0; f < nf; ++f) offsets[(size_t)t0 * S + f] = base + off[f];
        base += rows;
        return MOT_OK;
    };
    // software pipeline: chunks c + 1 and c + 2 are queued before the host waits for the row count of chunk c
    if (int rc = enqueue(0)) return rc;
    if (C > 1) if (int rc = enqueue(1)) return rc;
    for (int c = 0; c < C; ++c) {
        if (c + 2 < C) if (int rc = enqueue(c + 2)) return rc;
        if (int rc = drain(c)) { cudaStreamSynchronize(s_run); cudaStreamSynchronize(s_out); cudaStreamSynchronize(s_rows); return rc; }
    }
    offsets[TS] = base;
    MOT_CUDA(cudaStreamSynchronize(s_out));
    MOT_CUDA(cudaStreamSynchronize(s_rows));
    return MOT_OK;
}

int mot_engine_update_host(mot_engine* e, int T, const float* dets, const int* n_dets, int ld_dets, float* out,
                           int* n_out, int ld_out) {
    return mot_engine_update_host_embs(e, T, dets, n_dets, ld_dets, nullptr, out, n_out, ld_out);
}

// StrongSORT engines: the track list as rows of [id, state, hits, 0, tsu, conf, cls, det_ind, has_feat, n_samples, mean 8, cov 64]
int mot_engine_dump_strong(mot_engine* e, int s, float* rows82, float* feats, int cap_rows, int* n_rows) {
    if (!e || !rows82 || !n_rows || s < 0 || s >= e->cfg.n_streams) return fail(MOT_ERR_INVALID_ARGUMENT, "bad argument");
    if (e->cfg.kind != MOT_TRACKER_STRONGSORT) return fail(MOT_ERR_UNSUPPORTED, "not a StrongSORT engine");
    MOT_CUDA(cudaSetDevice(e->cfg.device));
    MOT_CUDA(cudaDeviceSynchronize());
    const mot::SsLayout& L = e->ss_layout;
    std::vector<unsigned char> slab(L.off_gal);
    const unsigned char* dbase = e->d_state + (size_t)s * L.stride;
    MOT_CUDA(cudaMemcpy(slab.data(), dbase, slab.size(), cudaMemcpyDeviceToHost));
    const unsigned char* base = slab.data();
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
        float* o = rows82 + 82 * (size_t)k;
        o[0] = (float)m[slot]; o[1] = (float)(state[slot] & 0x0f); o[2] = (float)m[cap + slot]; o[3] = 0.0f;
        o[4] = (float)m[2 * cap + slot]; o[5] = ((const float*)m)[7 * cap + slot]; o[6] = (float)m[3 * cap + slot];
        o[7] = (float)m[4 * cap + slot]; o[8] = (state[slot] & mot::kSsHasFeat) ? 1.0f : 0.0f; o[9] = (float)m[5 * cap + slot];
        std::memcpy(o + 10, recs + (size_t)slot * mot::kRecFloats, sizeof(float) * mot::kRecFloats);
        if (feats && L.dim > 0) {
            if (state[slot] & mot::kSsHasFeat) std::memcpy(feats + (size_t)L.dim * k, ft + (size_t)slot * L.dim, sizeof(float) * L.dim);
            else std::memset(feats + (size_t)L.dim * k, 0, sizeof(float) * L.dim);
        }
    }
    *n_rows = k;
    return MOT_OK;
}

// BoT-SORT engines: list `which` (0 active, 1 lost) as rows of [id, state, is_activated, frame_id, start_frame,
// tracklet_len, conf, cls, det_ind, has_feat, mean 8, cov 64] (82 floats) and, when feats != NULL, the smooth features
int mot_engine_dump_bot(mot_engine* e, int s, int which, float* rows82, float* feats, int cap_rows, int* n_rows) {
    if (!e || !rows82 || !n_rows || s < 0 || s >= e->cfg.n_streams) return fail(MOT_ERR_INVALID_ARGUMENT, "bad argument");
    if (e->cfg.kind != MOT_TRACKER_BOTSORT) return fail(MOT_ERR_UNSUPPORTED, "not a BoT-SORT engine");
    MOT_CUDA(cudaSetDevice(e->cfg.device));
    MOT_CUDA(cudaDeviceSynchronize());
    const mot::BotLayout& L = e->bot_layout;
    std::vector<unsigned char> slab(L.off_tnorm);
    const unsigned char* dbase = e->d_state + (size_t)s * L.stride;
    MOT_CUDA(cudaMemcpy(slab.data(), dbase, slab.size(), cudaMemcpyDeviceToHost));
    const int* hdr = (const int*)slab.data();
    const unsigned short* list = (const unsigned short*)(slab.data() + L.off_lists) + (which == 0 ? 0 : L.cap);
    const int n = which == 0 ? hdr[mot::kHdrActive] : hdr[mot::kHdrLost];
    const unsigned char* sflag = slab.data() + L.off_sflag;
    const int* m = (const int*)(slab.data() + L.off_meta);
    const float* recs = (const float*)(slab.data() + L.off_recs);
    const int cap = L.cap;
    int k = 0;
    for (; k < n && k < cap_rows; ++k) {
        const int slot = list[k];
        float* o = rows82 + 82 * (size_t)k;
        o[0] = (float)m[slot]; o[1] = (float)(sflag[slot] & 0x0f); o[2] = (sflag[slot] & 0x10) ? 1.0f : 0.0f;
        o[3] = (float)m[2 * cap + slot]; o[4] = (float)m[3 * cap + slot]; o[5] = (float)m[cap + slot];
        o[6] = ((const float*)m)[6 * cap + slot]; o[7] = (float)m[4 * cap + slot]; o[8] = (float)m[5 * cap + slot];
        o[9] = (sflag[slot] & 0x20) ? 1.0f : 0.0f;
        std::memcpy(o + 10, recs + (size_t)slot * mot::kRecFloats, sizeof(float) * mot::kRecFloats);
        if (feats && L.dim > 0)
            MOT_CUDA(cudaMemcpy(feats + (size_t)L.dim * k, dbase + L.off_feats + sizeof(float) * (size_t)slot * L.dim,
                                sizeof(float) * L.dim, cudaMemcpyDeviceToHost));
    }
    *n_rows = k;
    return MOT_OK;
}

int mot_engine_check(mot_engine* e, int* flags) {
    if (!e) return fail(MOT_ERR_INVALID_ARGUMENT, "null engine");
    MOT_CUDA(cudaSetDevice(e->cfg.device));
    MOT_CUDA(cudaDeviceSynchronize());
    const int S = e->cfg.n_streams;
    std::vector<int> err(S);
    MOT_CUDA(cudaMemcpy2D(err.data(), sizeof(int), e->d_state + sizeof(int) * mot::kHdrError, e->stride,
                          sizeof(int), S, cudaMemcpyDeviceToHost));
    int all = 0;
    for (int s = 0; s < S; ++s) { all |= err[s]; if (flags) flags[s] = err[s]; }
    // read-and-clear: a transient condition (one crowded frame, one truncated output) is reported once, by the check
    // that follows it, instead of failing every later update() until reset()
    if (all) MOT_CUDA(cudaMemset2D(e->d_state + sizeof(int) * mot::kHdrError, e->stride, 0, sizeof(int), S));
    if (all & (mot::kErrCapacity | mot::kErrTooManyDets | mot::kErrOutput))
        return fail(MOT_ERR_CAPACITY, "engine capacity exceeded (flags 0x%x: 1 track slots, 2 detections, 4 output rows)", all);
    if (all & mot::kErrTable) return fail(MOT_ERR_CAPACITY, "StrongSORT appearance candidate table full (flag 16)");
    if (all & mot::kErrKalman) return fail(MOT_ERR_NUMERIC, "a Kalman update left the Cholesky path");
    return MOT_OK;
}

int mot_engine_profile(mot_engine* e, int enable, unsigned long long* cycles32) {
    if (!e) return fail(MOT_ERR_INVALID_ARGUMENT, "null engine");
    MOT_CUDA(cudaSetDevice(e->cfg.device));
    MOT_CUDA(cudaDeviceSynchronize());
    if (e->d_prof && cycles32) MOT_CUDA(cudaMemcpy(cycles32, e->d_prof, 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    else if (cycles32) memset(cycles32, 0, 32 * sizeof(unsigned long long));
    if (enable) {
        if (!e->d_prof) MOT_CUDA(cudaMalloc((void**)&e->d_prof, 32 * sizeof(unsigned long long)));
        MOT_CUDA(cudaMemset(e->d_prof, 0, 32 * sizeof(unsigned long long)));
    } else if (e->d_prof) {
        cudaFree(e->d_prof);
        e->d_prof = nullptr;
    }
    return MOT_OK;
}

int mot_engine_stream_header(mot_engine* e, int s, int* hdr16) {
    if (!e || !hdr16 || s < 0 || s >= e->cfg.n_streams) return fail(MOT_ERR_INVALID_ARGUMENT, "bad argument");
    MOT_CUDA(cudaSetDevice(e->cfg.device));
    MOT_CUDA(cudaDeviceSynchronize());
    MOT_CUDA(cudaMemcpy(hdr16, e->d_state + (size_t)s * e->stride, sizeof(int) * mot::kHdrInts, cudaMemcpyDeviceToHost));
    return MOT_OK;
}

int mot_engine_dump_list(mot_engine* e, int s, int which, float* rows, int cap_rows, int* n_rows) {
    if (!e || !rows || !n_rows || s < 0 || s >= e->cfg.n_streams) return fail(MOT_ERR_INVALID_ARGUMENT, "bad argument");
    if (e->cfg.kind != MOT_TRACKER_BYTETRACK && e->cfg.kind != MOT_TRACKER_OCSORT && e->cfg.kind != MOT_TRACKER_DEEPOCSORT)
        return fail(MOT_ERR_UNSUPPORTED, "list dumps exist for ByteTrack, OC-SORT and DeepOC-SORT engines only");
    MOT_CUDA(cudaSetDevice(e->cfg.device));
    MOT_CUDA(cudaDeviceSynchronize());
    if (e->cfg.kind == MOT_TRACKER_OCSORT || e->cfg.kind == MOT_TRACKER_DEEPOCSORT) {
        // rows of [id, age, hits, hit_streak, time_since_update, conf, cls, det_ind, last_obs 5, velocity 2, x 7, P 49, pad 7]
        const mot::OcLayout& L = e->oc_layout;
        std::vector<unsigned char> slab(L.off_ocm);
        MOT_CUDA(cudaMemcpy(slab.data(), e->d_state + (size_t)s * e->stride, slab.size(), cudaMemcpyDeviceToHost));
        const int* hdr = (const int*)slab.data();
        const unsigned short* list = (const unsigned short*)(slab.data() + L.off_lists);
        const int* m = (const int*)(slab.data() + L.off_meta);
        const float* obs = (const float*)(slab.data() + L.off_obs);
        const float* recs = (const float*)(slab.data() + L.off_recs);
        const int n = which == 0 ? hdr[mot::kOHdrTracks] : 0, cap = L.cap;
        int k = 0;
        for (; k < n && k < cap_rows; ++k) {
            const int slot = list[k];
            float* o = rows + 78 * (size_t)k;
            std::memset(o, 0, 78 * sizeof(float));
            o[0] = (float)m[slot]; o[1] = (float)m[cap + slot]; o[2] = (float)m[2 * cap + slot]; o[3] = (float)m[3 * cap + slot];
            o[4] = (float)m[4 * cap + slot]; o[5] = ((const float*)m)[7 * cap + slot]; o[6] = (float)m[5 * cap + slot];
            o[7] = (float)m[6 * cap + slot];
            std::memcpy(o + 8, obs + (size_t)slot * mot::kOcObsFloats, 7 * sizeof(float));
            std::memcpy(o + 15, recs + (size_t)slot * mot::kOcRecFloats, 56 * sizeof(float));
        }
        *n_rows = k;
        return MOT_OK;
    }
    std::vector<unsigned char> slab(e->layout.off_gscratch);
    MOT_CUDA(cudaMemcpy(slab.data(), e->d_state + (size_t)s * e->layout.stride, slab.size(), cudaMemcpyDeviceToHost));
    const mot::BtLayout& L = e->layout;
    const int* hdr = (const int*)slab.data();
    const unsigned short* list = (const unsigned short*)(slab.data() + L.off_lists) + (which == 0 ? 0 : L.cap);
    const int n = which == 0 ? hdr[mot::kHdrActive] : hdr[mot::kHdrLost];
    const unsigned char* sflag = slab.data() + L.off_sflag;
    const int* meta = (const int*)(slab.data() + L.off_meta);
    const float* recs = (const float*)(slab.data() + L.off_recs);
    int k = 0;
    for (; k < n && k < cap_rows; ++k) {
        const int slot = list[k];
        float* o = rows + 78 * (size_t)k;
        o[0] = (float)meta[slot]; o[1] = (float)(sflag[slot] & 0x0f); o[2] = (sflag[slot] & 0x10) ? 1.0f : 0.0f;
        o[3] = (float)meta[2 * L.cap + slot]; o[4] = (float)meta[3 * L.cap + slot]; o[5] = (float)meta[L.cap + slot];
        mot::kfb_expand(recs + (size_t)slot * mot::kBtRecFloats, o + 6);     // compact record -> [mean 8 | cov 8x8]
    }
    *n_rows = k;
    return MOT_OK;
}

int mot_engine_dump_boost(mot_engine* e, int s, float* rows, int cap_rows, int* n_rows) {
    if (!e || !rows || !n_rows || s < 0 || s >= e->cfg.n_streams) return fail(MOT_ERR_INVALID_ARGUMENT, "bad argument");
    if (e->cfg.kind != MOT_TRACKER_BOOSTTRACK) return fail(MOT_ERR_UNSUPPORTED, "not a BoostTrack engine");
    MOT_CUDA(cudaSetDevice(e->cfg.device));
    MOT_CUDA(cudaDeviceSynchronize());
    const mot::SortLayout& L = e->sort_layout;
    std::vector<unsigned char> slab(L.off_gscratch);
    MOT_CUDA(cudaMemcpy(slab.data(), e->d_state + (size_t)s * L.stride, slab.size(), cudaMemcpyDeviceToHost));
    const int* hdr = (const int*)slab.data();
    const unsigned short* list = (const unsigned short*)(slab.data() + L.off_lists);
    const int* m = (const int*)(slab.data() + L.off_meta);
    const float* recs = (const float*)(slab.data() + L.off_recs);
    const int n = hdr[mot::kSHdrTracks], cap = L.cap;
    int k = 0;
    for (; k < n && k < cap_rows; ++k) {
        const int slot = list[k];
        float* o = rows + 80 * (size_t)k;
        std::memset(o, 0, 80 * sizeof(float));
        // meta arrays of the SORT slab: id, hits (= hit_streak here), tsu, age, cls, det_ind, conf
        o[0] = (float)m[slot]; o[1] = (float)m[3 * cap + slot]; o[2] = (float)m[cap + slot]; o[3] = (float)m[2 * cap + slot];
        o[4] = ((const float*)m)[6 * cap + slot]; o[5] = (float)m[4 * cap + slot]; o[6] = (float)m[5 * cap + slot];
        const float* rec = recs + (size_t)slot * mot::kBoostRecFloats;
        std::memcpy(o + 8, rec, 8 * sizeof(float));
        for (int c = 0; c < 4; ++c) {
            const float* P = rec + 8 + 4 * c;
            o[16 + c * 8 + c] = P[0]; o[16 + c * 8 + c + 4] = P[1]; o[16 + (c + 4) * 8 + c] = P[2]; o[16 + (c + 4) * 8 + c + 4] = P[3];
        }
    }
    *n_rows = k;
    return MOT_OK;
}

int mot_engine_dump_deep_embs(mot_engine* e, int s, float* embs, int cap_rows, int* n_rows) {
    if (!e || !embs || !n_rows || s < 0 || s >= e->cfg.n_streams) return fail(MOT_ERR_INVALID_ARGUMENT, "bad argument");
    if (e->cfg.kind != MOT_TRACKER_DEEPOCSORT || e->deep_layout.dim <= 0)
        return fail(MOT_ERR_UNSUPPORTED, "not a DeepOC-SORT engine with embeddings");
    MOT_CUDA(cudaSetDevice(e->cfg.device));
    MOT_CUDA(cudaDeviceSynchronize());
    const mot::OcLayout& L = e->oc_layout;
    const int dim = e->deep_layout.dim;
    const unsigned char* base = e->d_state + (size_t)s * e->stride;
    std::vector<unsigned char> head(L.off_meta);
    MOT_CUDA(cudaMemcpy(head.data(), base, head.size(), cudaMemcpyDeviceToHost));
    const int n = ((const int*)head.data())[mot::kOHdrTracks];
    const unsigned short* list = (const unsigned short*)(head.data() + L.off_lists);
    const float* d_emb = (const float*)(base + L.stride + e->deep_layout.off_emb);
    int k = 0;
    for (; k < n && k < cap_rows; ++k)
        MOT_CUDA(cudaMemcpy(embs + (size_t)k * dim, d_emb + (size_t)list[k] * dim, sizeof(float) * dim, cudaMemcpyDeviceToHost));
    *n_rows = k;
    return MOT_OK;
}

int mot_engine_info(mot_engine* e, int* threads, int* smem, int* ctas, int* state_bytes) {
    if (!e) return fail(MOT_ERR_INVALID_ARGUMENT, "null engine");
    if (threads) *threads = e->threads;
    if (smem) *smem = (int)e->smem_bytes;
    if (ctas) *ctas = e->cfg.n_streams;
    if (state_bytes) *state_bytes = (int)e->stride;
    return MOT_OK;
}

// ------------------------------------------------------------------------------ standalone kernels
int mot_kf_initiate(int kind, float* recs, const float* z, long long n, void* stream) {
    if (!recs || !z || n < 0) return fail(MOT_ERR_INVALID_ARGUMENT, "bad argument");
    if (int rc = require_device()) return rc;
    if (n == 0) return MOT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (kind) {
        case mot::kKfXYAH: mot::kf_initiate_kernel<mot::kKfXYAH><<<kf_grid(n), 256, 0, st>>>(recs, z, n); break;
        case mot::kKfXYSR: mot::kf_initiate_kernel<mot::kKfXYSR><<<kf_grid(n), 256, 0, st>>>(recs, z, n); break;
        case mot::kKfXYWH: mot::kf_initiate_kernel<mot::kKfXYWH><<<kf_grid(n), 256, 0, st>>>(recs, z, n); break;
        default: return fail(MOT_ERR_INVALID_ARGUMENT, "unknown Kalman kind %d", kind);
    }
    MOT_CUDA(cudaGetLastError());
    return MOT_OK;
}

int mot_kf_predict(int kind, float* recs, const unsigned char* flags, long long n, float q_xy, float q_s, void* stream) {
    if (!recs || n < 0) return fail(MOT_ERR_INVALID_ARGUMENT, "bad argument");
    if (int rc = require_device()) return rc;
    if (n == 0) return MOT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const float q44 = 0.01f * q_xy, q66 = 0.0001f * q_s;     // xysr_kf.cpp:58-61 then ocsort.cpp:77-79, in fp32
    switch (kind) {
        case mot::kKfXYAH: mot::kf_predict_kernel<mot::kKfXYAH><<<kf_grid(n), 256, 0, st>>>(recs, flags, n, q44, q66); break;
        case mot::kKfXYSR: mot::kf_predict_kernel<mot::kKfXYSR><<<kf_grid(n), 256, 0, st>>>(recs, flags, n, q44, q66); break;
        case mot::kKfXYWH: mot::kf_predict_kernel<mot::kKfXYWH><<<kf_grid(n), 256, 0, st>>>(recs, flags, n, q44, q66); break;
        default: return fail(MOT_ERR_INVALID_ARGUMENT, "unknown Kalman kind %d", kind);
    }
    MOT_CUDA(cudaGetLastError());
    return MOT_OK;
}

int mot_kf_update(int kind, float* recs, const float* z, const float* conf, long long n, unsigned char* failflags,
                  void* stream) {
    if (!recs || !z || n < 0) return fail(MOT_ERR_INVALID_ARGUMENT, "bad argument");
    if (int rc = require_device()) return rc;
    if (n == 0) return MOT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (kind) {
        case mot::kKfXYAH: mot::kf_update_kernel<mot::kKfXYAH><<<kf_grid(n), 256, 0, st>>>(recs, z, conf, n, failflags); break;
        case mot::kKfXYSR: mot::kf_update_kernel<mot::kKfXYSR><<<kf_grid(n), 256, 0, st>>>(recs, z, conf, n, failflags); break;
        case mot::kKfXYWH: mot::kf_update_kernel<mot::kKfXYWH><<<kf_grid(n), 256, 0, st>>>(recs, z, conf, n, failflags); break;
        default: return fail(MOT_ERR_INVALID_ARGUMENT, "unknown Kalman kind %d", kind);
    }
    MOT_CUDA(cudaGetLastError());
    return MOT_OK;
}

int mot_kf_gating(int kind, const float* recs, int n_tracks, const float* meas, int n_meas, int only_position,
                  int metric, float* out, void* stream) {
    if (!recs || !meas || !out || n_tracks < 0 || n_meas < 0) return fail(MOT_ERR_INVALID_ARGUMENT, "bad argument");
    if (int rc = require_device()) return rc;
    if (n_tracks == 0 || n_meas == 0) return MOT_OK;
    if (kind == mot::kKfXYAH && metric != 0 && metric != 1)
        return fail(MOT_ERR_INVALID_ARGUMENT, "Invalid metric: %d", metric);     // kalman_filter.cpp:174
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)n_tracks * n_meas;
    const int blocks = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)sm_count() * 16));
    if (kind == mot::kKfXYAH)
        mot::kf_gating_kernel<mot::kKfXYAH><<<blocks, 256, 0, st>>>(recs, n_tracks, meas, n_meas, only_position, metric, out);
    else if (kind == mot::kKfXYWH)
        mot::kf_gating_kernel<mot::kKfXYWH><<<blocks, 256, 0, st>>>(recs, n_tracks, meas, n_meas, only_position, metric, out);
    else
        return fail(MOT_ERR_INVALID_ARGUMENT, "gating is defined for XYAH (0) and XYWH (2) only");
    MOT_CUDA(cudaGetLastError());
    return MOT_OK;
}

int mot_cost_iou(const float* a, int n, const float* b, int m, const float* conf, float* out, int ld, int mode,
                 void* stream) {
    if (n < 0 || m < 0 || ld < m) return fail(MOT_ERR_INVALID_ARGUMENT, "bad sizes");
    if (n == 0 || m == 0) return MOT_OK;
    if (!a || !b || !out) return fail(MOT_ERR_INVALID_ARGUMENT, "null argument");
    if ((((size_t)a) | ((size_t)b)) & 15) return fail(MOT_ERR_INVALID_ARGUMENT, "box arrays must be 16-byte aligned (they are read as float4)");
    if (mode == mot::kCostIouDistanceFused && !conf) return fail(MOT_ERR_INVALID_ARGUMENT, "fuse_score needs det confidences");
    if (mode < 0 || mode > 2) return fail(MOT_ERR_INVALID_ARGUMENT, "unknown mode %d", mode);
    if (int rc = require_device()) return rc;
    const int col_tiles = (m + mot::kCostTileCols - 1) / mot::kCostTileCols;
    const int row_groups = (n + mot::kCostTileRows - 1) / mot::kCostTileRows;
    dim3 grid((unsigned)std::min(row_groups, sm_count() * 8), (unsigned)std::min(col_tiles, 64));
    mot::iou_cost_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, n, b, m, conf, out, ld, mode);
    MOT_CUDA(cudaGetLastError());
    return MOT_OK;
}

int mot_cost_iou_variant(const float* a, int n, const float* b, int m, int kind, int frame_w, int frame_h, float* out, int ld,
                         void* stream) {
    if (n < 0 || m < 0 || ld < m) return fail(MOT_ERR_INVALID_ARGUMENT, "bad sizes");
    if (kind < mot::kVarHmIou || kind > mot::kVarCIoU)
        return fail(MOT_ERR_INVALID_ARGUMENT, "Invalid association mode: %d (3 hmiou, 4 giou, 5 diou, 6 centroid, 7 ciou)", kind);   // iou.hpp:407
    if ((((size_t)a) | ((size_t)b)) & 15) return fail(MOT_ERR_INVALID_ARGUMENT, "box arrays must be 16-byte aligned (they are read as float4)");
    if (n == 0 || m == 0) return MOT_OK;
    if (!a || !b || !out) return fail(MOT_ERR_INVALID_ARGUMENT, "null argument");
    if (int rc = require_device()) return rc;
    const float norm = static_cast<float>(std::sqrt((double)(frame_w * frame_w + frame_h * frame_h)));     // iou.hpp:325
    const int col_tiles = (m + mot::kCostTileCols - 1) / mot::kCostTileCols;
    const int row_groups = (n + mot::kCostTileRows - 1) / mot::kCostTileRows;
    dim3 grid((unsigned)std::min(row_groups, sm_count() * 8), (unsigned)std::min(col_tiles, 64));
    mot::iou_variant_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, n, b, m, kind, norm, out, ld);
    MOT_CUDA(cudaGetLastError());
    return MOT_OK;
}

int mot_cost_ocm(const float* dets5, int n_dets, const float* trks4, const float* vel2, const float* prev5, int n_trks,
                 float inertia, float* out_cost, float* out_iou, int ld, void* stream) {
    if (n_dets < 0 || n_trks < 0 || ld < n_trks) return fail(MOT_ERR_INVALID_ARGUMENT, "bad sizes");
    if (n_dets == 0 || n_trks == 0) return MOT_OK;
    if (!dets5 || !trks4 || !vel2 || !prev5 || !out_cost) return fail(MOT_ERR_INVALID_ARGUMENT, "null argument");
    if (int rc = require_device()) return rc;
    const int col_tiles = (n_trks + mot::kCostTileCols - 1) / mot::kCostTileCols;
    const int row_groups = (n_dets + mot::kCostTileRows - 1) / mot::kCostTileRows;
    dim3 grid((unsigned)std::min(row_groups, sm_count() * 8), (unsigned)std::min(col_tiles, 64));
    mot::ocm_cost_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dets5, n_dets, trks4, vel2, prev5, n_trks, inertia,
                                                                 out_cost, out_iou, ld);
    MOT_CUDA(cudaGetLastError());
    return MOT_OK;
}

int mot_cost_cosine(const float* t, int n, const float* d, int m, int dim, float* out, int ld, void* stream) {
    if (n < 0 || m < 0 || dim <= 0 || ld < m) return fail(MOT_ERR_INVALID_ARGUMENT, "bad sizes");
    if (n == 0 || m == 0) return MOT_OK;
    if (!t || !d || !out) return fail(MOT_ERR_INVALID_ARGUMENT, "null argument");
    if (int rc = require_device()) return rc;
    std::string err;
    const int rc = mot::launch_cosine(t, n, d, m, dim, out, ld, (cudaStream_t)stream, err);
    if (rc != MOT_OK) return fail(rc, "%s", err.c_str());
    return MOT_OK;
}

int mot_cost_nn_cosine(const float* samples, const int* seg, int n_samples, int n_targets, const float* feats, int m,
                       int dim, float* out, int ld, void* stream) {
    if (n_samples < 0 || n_targets < 0 || m < 0 || dim <= 0 || ld < m) return fail(MOT_ERR_INVALID_ARGUMENT, "bad sizes");
    if (n_targets == 0 || m == 0) return MOT_OK;
    if ((n_samples > 0 && (!samples || !seg)) || !feats || !out) return fail(MOT_ERR_INVALID_ARGUMENT, "null argument");
    if (int rc = require_device()) return rc;
    std::string err;
    const int rc = mot::launch_nn_cosine(samples, seg, n_samples, n_targets, feats, m, dim, out, ld, (cudaStream_t)stream, err);
    if (rc != MOT_OK) return fail(rc, "%s", err.c_str());
    return MOT_OK;
}

int mot_cost_gate(float* cost, int ld, const float* recs, int n_tracks, const float* meas4, int n_meas, float mc_lambda,
                  float gated_cost, int only_position, void* stream) {
    if (n_tracks < 0 || n_meas < 0 || ld < n_meas) return fail(MOT_ERR_INVALID_ARGUMENT, "bad sizes");
    if (n_tracks == 0 || n_meas == 0) return MOT_OK;
    if (!cost || !recs || !meas4) return fail(MOT_ERR_INVALID_ARGUMENT, "null argument");
    if (((size_t)meas4) & 15) return fail(MOT_ERR_INVALID_ARGUMENT, "meas4 must be 16-byte aligned (read as float4)");
    if (int rc = require_device()) return rc;
    const int blocks = std::max(1, std::min((n_tracks + 7) / 8, sm_count() * 8));
    mot::gate_cost_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(cost, ld, recs, n_tracks, meas4, n_meas, mc_lambda,
                                                                    gated_cost, only_position);
    MOT_CUDA(cudaGetLastError());
    return MOT_OK;
}

int mot_cost_iou_tlwh(const float* trk_tlwh, const int* tsu, int n, const float* det_tlwh, int m, float* out, int ld,
                      void* stream) {
    if (n < 0 || m < 0 || ld < m) return fail(MOT_ERR_INVALID_ARGUMENT, "bad sizes");
    if (n == 0 || m == 0) return MOT_OK;
    if (!trk_tlwh || !det_tlwh || !out) return fail(MOT_ERR_INVALID_ARGUMENT, "null argument");
    if ((((size_t)trk_tlwh) | ((size_t)det_tlwh)) & 15) return fail(MOT_ERR_INVALID_ARGUMENT, "box arrays must be 16-byte aligned (they are read as float4)");
    if (int rc = require_device()) return rc;
    const int col_tiles = (m + mot::kCostTileCols - 1) / mot::kCostTileCols;
    const int row_groups = (n + mot::kCostTileRows - 1) / mot::kCostTileRows;
    dim3 grid((unsigned)std::min(row_groups, sm_count() * 8), (unsigned)std::min(col_tiles, 64));
    mot::iou_tlwh_cost_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(trk_tlwh, tsu, n, det_tlwh, m, out, ld);
    MOT_CUDA(cudaGetLastError());
    return MOT_OK;
}

int mot_cost_aw_max_metric(const float* emb_cost, int n, int m, int ld, float w_association_emb, float bottom, float* out,
                           int ld_out, void* stream) {
    if (n < 0 || m < 0 || ld < m || ld_out < m) return fail(MOT_ERR_INVALID_ARGUMENT, "bad sizes");
    if (n == 0 || m == 0) return MOT_OK;
    if (!emb_cost || !out) return fail(MOT_ERR_INVALID_ARGUMENT, "null argument");
    if (int rc = require_device()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* ws = nullptr;                                                 // row / column weights + zero flags
    const size_t fl = sizeof(float) * ((size_t)n + m);
    MOT_CUDA(cudaMallocAsync((void**)&ws, fl + (size_t)n + m + 16, st));
    float* row_w = (float*)ws; float* col_w = row_w + n;
    unsigned char* row_z = ws + fl; unsigned char* col_z = row_z + n;
    const int cap = sm_count() * 8;
    mot::aw_row_top2_kernel<<<std::max(1, std::min((n + 7) / 8, cap)), 256, 0, st>>>(emb_cost, n, m, ld, bottom, row_w, row_z);
    mot::aw_col_top2_kernel<<<std::max(1, std::min((m + 31) / 32, cap)), 256, 0, st>>>(emb_cost, n, m, ld, bottom, col_w, col_z);
    const long long total = (long long)n * m;
    mot::aw_apply_kernel<<<(int)std::max<long long>(1, std::min<long long>((total + 255) / 256, cap)), 256, 0, st>>>(
        emb_cost, n, m, ld, w_association_emb, row_w, row_z, col_w, col_z, out, ld_out);
    MOT_CUDA(cudaGetLastError());
    MOT_CUDA(cudaFreeAsync(ws, st));
    return MOT_OK;
}

int mot_kf_xysr_affine(float* recs, long long n, const float* m2x2, const float* t2, void* stream) {
    if (n < 0) return fail(MOT_ERR_INVALID_ARGUMENT, "bad sizes");
    if (n == 0) return MOT_OK;
    if (!recs || !m2x2 || !t2) return fail(MOT_ERR_INVALID_ARGUMENT, "null argument");
    if (int rc = require_device()) return rc;
    const mot::Affine6 a{m2x2[0], m2x2[1], m2x2[2], m2x2[3], t2[0], t2[1]};      // HOST pointers: six scalars
    const int blocks = (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)sm_count() * 8));
    mot::kf_xysr_affine_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(recs, n, a);
    MOT_CUDA(cudaGetLastError());
    return MOT_OK;
}

int mot_lap_batch_device(const float* cost, long long stride_cost, int n_problems, const int* n_rows,
                         const int* n_cols, int n, int m, int ld, float thresh, int* row2col, int* col2row,
                         void* stream) {
    if (n < 0 || m < 0 || n_problems < 0 || (m > 0 && ld < m)) return fail(MOT_ERR_INVALID_ARGUMENT, "bad sizes");
    if (n > 32767 || m > 32767) return fail(MOT_ERR_INVALID_ARGUMENT, "n, m above 32767 are not supported");
    if (int rc = require_device()) return rc;
    if (n_problems == 0) return MOT_OK;
    if (!row2col || !col2row) return fail(MOT_ERR_INVALID_ARGUMENT, "null result pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0 || m == 0) {
        if (n) MOT_CUDA(cudaMemsetAsync(row2col, 0xff, sizeof(int) * (size_t)n * n_problems, st));
        if (m) MOT_CUDA(cudaMemsetAsync(col2row, 0xff, sizeof(int) * (size_t)m * n_problems, st));
        return MOT_OK;
    }
    if (!cost) return fail(MOT_ERR_INVALID_ARGUMENT, "null cost pointer");
    int dev = 0, max_optin = 0;
    MOT_CUDA(cudaGetDevice(&dev));
    MOT_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    int e_cap = 8192;
    size_t smem = mot::lap_smem_bytes(n, m, e_cap);
    while (smem > (size_t)max_optin && e_cap > 1024) { e_cap /= 2; smem = mot::lap_smem_bytes(n, m, e_cap); }
    if (smem > (size_t)max_optin)
        return fail(MOT_ERR_INVALID_ARGUMENT, "problem %d x %d needs %zu B of shared memory (limit %d)", n, m, smem, max_optin);
    MOT_CUDA(cudaFuncSetAttribute(mot::lap_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned char* gs = nullptr;
    const size_t gs_bytes = mot::lap_gscratch_bytes(n, m);
    const int grid = std::min(n_problems, sm_count() * 4);
    MOT_CUDA(cudaMallocAsync((void**)&gs, gs_bytes * (size_t)n_problems, st));
    mot::LapBatchArgs a{};
    a.cost = cost; a.stride_cost = stride_cost; a.n_rows = n_rows; a.n_cols = n_cols;
    a.n = n; a.m = m; a.ld = ld; a.thresh = thresh; a.row2col = row2col; a.col2row = col2row;
    a.gscratch = gs; a.n_max = n; a.m_max = m; a.e_cap = e_cap; a.n_problems = n_problems;
    mot::lap_dense_kernel<<<grid, 256, smem, st>>>(a);
    MOT_CUDA(cudaGetLastError());
    MOT_CUDA(cudaFreeAsync(gs, st));
    return MOT_OK;
}

int mot_lap_jv_batch_device(const float* cost, long long stride_cost, int n_problems, int n, int m, int ld, float thresh,
                            int* row2col, int* col2row, void* stream) {
    if (n <= 0 || m <= 0 || n_problems < 0 || ld < m) return fail(MOT_ERR_INVALID_ARGUMENT, "bad sizes");
    if (n + m > 32000) return fail(MOT_ERR_INVALID_ARGUMENT, "rows + columns above 32000 are not supported");
    if (int rc = require_device()) return rc;
    if (n_problems == 0) return MOT_OK;
    if (!cost || !row2col || !col2row) return fail(MOT_ERR_INVALID_ARGUMENT, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    int warp_max = mot::kLapJvMax;
    if (const char* ev = std::getenv("MOT_LAPJV_WARP_MAX")) warp_max = std::atoi(ev);      // measurement aid
    if (n + m <= warp_max) {                // one warp per problem, all state in shared memory
        const size_t wsm = mot::jv_work_bytes(n + m + 1);
        if (wsm > 48 * 1024) MOT_CUDA(cudaFuncSetAttribute(mot::lap_jv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsm));
        mot::lap_jv_kernel<<<std::min(n_problems, sm_count() * 16), 32, wsm, st>>>(cost, stride_cost, n_problems, n, m, ld, thresh,
                                                                                   row2col, col2row);
        MOT_CUDA(cudaGetLastError());
        return MOT_OK;
    }
    // any size: one CTA per problem; work arrays in shared memory when they fit, else in stream-ordered global scratch
    size_t smem = mot::jv_block_sbytes(n + m);
    int dev = 0, max_optin = 0;
    MOT_CUDA(cudaGetDevice(&dev));
    MOT_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const bool all_shared = mot::jv_block_sbytes_full(n + m) + 1024 <= (size_t)max_optin && !std::getenv("MOT_LAPJV_GLOBAL_WORK");
    if (all_shared) smem = mot::jv_block_sbytes_full(n + m);
    if (smem > (size_t)max_optin)
        return fail(MOT_ERR_INVALID_ARGUMENT, "problem %d x %d needs %zu B of shared memory (limit %d)", n, m, smem, max_optin);
    MOT_CUDA(cudaFuncSetAttribute(mot::lap_jv_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = std::min(n_problems, sm_count() * 2);
    unsigned char* gs = nullptr;
    if (!all_shared) MOT_CUDA(cudaMallocAsync((void**)&gs, mot::jv_block_gbytes(n + m) * (size_t)grid, st));
    mot::lap_jv_block_kernel<<<grid, mot::kLapJvBlockThreads, smem, st>>>(cost, stride_cost, n_problems, n, m, ld, thresh, row2col,
                                                                          col2row, gs, all_shared ? 1 : 0);
    MOT_CUDA(cudaGetLastError());
    if (gs) MOT_CUDA(cudaFreeAsync(gs, st));
    return MOT_OK;
}

int mot_lap_device(const float* cost, int n, int m, int ld, float thresh, int* row2col, int* col2row, void* stream) {
    return mot_lap_batch_device(cost, 0, 1, nullptr, nullptr, n, m, ld, thresh, row2col, col2row, stream);
}

int mot_lap_host(const float* cost, int n, int m, int ld, float thresh, int* row2col, int* col2row) {
    if (n < 0 || m < 0) return fail(MOT_ERR_INVALID_ARGUMENT, "bad sizes");
    if (int rc = require_device()) return rc;
    for (int i = 0; i < n; ++i) row2col[i] = -1;
    for (int j = 0; j < m; ++j) col2row[j] = -1;
    if (n == 0 || m == 0) return MOT_OK;                      // matching.cpp:20-28
    float* d_cost = nullptr; int *d_r = nullptr, *d_c = nullptr;
    MOT_CUDA(cudaMalloc(&d_cost, sizeof(float) * (size_t)n * ld));
    MOT_CUDA(cudaMalloc(&d_r, sizeof(int) * n));
    MOT_CUDA(cudaMalloc(&d_c, sizeof(int) * m));
    MOT_CUDA(cudaMemcpy(d_cost, cost, sizeof(float) * (size_t)n * ld, cudaMemcpyHostToDevice));
    int rc = mot_lap_device(d_cost, n, m, ld, thresh, d_r, d_c, nullptr);
    if (rc == MOT_OK) {
        MOT_CUDA(cudaMemcpy(row2col, d_r, sizeof(int) * n, cudaMemcpyDeviceToHost));
        MOT_CUDA(cudaMemcpy(col2row, d_c, sizeof(int) * m, cudaMemcpyDeviceToHost));
    }
    cudaFree(d_cost); cudaFree(d_r); cudaFree(d_c);
    return rc;
}

}  // extern "C"
