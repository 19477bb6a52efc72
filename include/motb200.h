/*
 * motb200.h - C ABI of the B200-native association engine (libmotb200.so).
 *
 * This is the drop-in boundary for motcpp's per-frame hot path.  The reference has no FFI of its
 * own: its boundary is the C++ class motcpp::BaseTracker and a few free functions.  Each entry
 * point below names the reference interface it replaces (paths relative to the motcpp tree);
 * INTEGRATION.md shows the C++ binding a motcpp maintainer would add on top of this header, and
 * the headers under include/motcpp_b200/ ship that binding (same class names, constructor arguments and
 * exceptions as the reference).
 *
 * Conventions
 *   - plain C types only; every function returns a mot_status (0 = ok); mot_last_error() gives
 *     the message of the calling thread's last failure.  No exceptions cross this boundary.
 *   - matrices are ROW-MAJOR float32 unless stated otherwise (Eigen::MatrixXf is column-major: the
 *     C++ binding transposes on the way in/out).
 *   - "device" pointers are CUDA device memory of the engine's GPU; "host" pointers are ordinary
 *     host memory (pinned memory from mot_host_alloc makes the copies asynchronous).
 *   - `stream` arguments are a cudaStream_t passed as void* (NULL = the legacy default stream).
 *   - there is no CPU fallback: every compute entry point fails with MOT_ERR_NO_DEVICE when no
 *     CUDA device is present.
 */
#ifndef MOTB200_H
#define MOTB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum mot_status {
    MOT_OK = 0,
    MOT_ERR_INVALID_ARGUMENT = 1,  /* -> std::invalid_argument in the C++ binding (src/tracker.cpp:108-125) */
    MOT_ERR_CUDA = 2,              /* -> std::runtime_error */
    MOT_ERR_NO_DEVICE = 3,
    MOT_ERR_CAPACITY = 4,          /* a stream ran out of track slots / output rows / detection slots */
    MOT_ERR_NUMERIC = 5,           /* a Kalman update hit the reference's pseudo-inverse fallback */
    MOT_ERR_UNSUPPORTED = 6
} mot_status;

const char* mot_last_error(void);
int mot_version(void);                 /* 100 * major + minor */
int mot_device_count(void);            /* 0 when no CUDA device is usable */

/* ---- memory helpers (so a caller needs no CUDA headers) ------------------------------------ */
int mot_device_alloc(void** ptr, size_t bytes);
int mot_device_free(void* ptr);
int mot_host_alloc(void** ptr, size_t bytes);            /* pinned */
int mot_host_free(void* ptr);
int mot_copy_h2d(void* dst_device, const void* src_host, size_t bytes, void* stream);
int mot_copy_d2h(void* dst_host, const void* src_device, size_t bytes, void* stream);
int mot_memset_device(void* dst_device, int value, size_t bytes, void* stream);
int mot_stream_sync(void* stream);

/* ---- tracker engine: BaseTracker::update for S independent camera streams ------------------ */
/* replaces motcpp::BaseTracker::update / reset (include/motcpp/tracker.hpp:67-74) and the concrete
 * front-ends' update(): ByteTrack src/trackers/bytetrack.cpp:166, Sort src/trackers/sort.cpp:102,
 * OCSort src/trackers/ocsort.cpp:285, BotSort src/trackers/botsort.cpp:260. */
typedef enum mot_tracker_kind {
    MOT_TRACKER_SORT = 0,
    MOT_TRACKER_BYTETRACK = 1,
    MOT_TRACKER_OCSORT = 2,
    MOT_TRACKER_BOTSORT = 3,
    MOT_TRACKER_STRONGSORT = 4,
    MOT_TRACKER_DEEPOCSORT = 5,
    MOT_TRACKER_BOOSTTRACK = 6
} mot_tracker_kind;

typedef struct mot_engine_config {
    int kind;               /* mot_tracker_kind */
    int n_streams;          /* independent trackers living on this GPU */
    int track_capacity;     /* live + lost tracks per stream (0 = default 1536) */
    int max_dets;           /* detections per stream per frame (0 = default 512) */
    int device;             /* CUDA device ordinal */
    int n_chunks;           /* host-buffer calls are pipelined over this many stream chunks (0 = auto) */
    /* BaseTracker ctor (include/motcpp/tracker.hpp:47-55) */
    float det_thresh;
    int max_age, max_obs, min_hits;
    float iou_threshold;
    /* ByteTrack ctor (include/motcpp/trackers/bytetrack.hpp:97-110) */
    float min_conf, track_thresh, match_thresh;
    int track_buffer, frame_rate;
    /* OCSort ctor (include/motcpp/trackers/ocsort.hpp:88-102) */
    int delta_t;
    float inertia;
    int use_byte;
    float q_xy_scaling, q_s_scaling;
    /* BotSort ctor (include/motcpp/trackers/botsort.hpp:108-134) */
    float track_high_thresh, track_low_thresh, new_track_thresh;
    float proximity_thresh, appearance_thresh;
    int fuse_first_associate, with_reid, emb_dim;
    /* StrongSORT ctor (include/motcpp/trackers/strongsort.hpp:287-305); min_conf, max_age and emb_dim above are shared.
     * nn_budget >= 1 is the gallery ring size per track. */
    float max_cos_dist, max_iou_dist;
    int n_init, nn_budget;
    float mc_lambda, ema_alpha;
    /* DeepOCSort ctor (include/motcpp/trackers/deepocsort.hpp:93-117); det_thresh, max_age, min_hits, iou_threshold,
     * delta_t, inertia, q_*_scaling and emb_dim above are shared.  Camera-motion compensation is not part of the hot
     * path (the engine behaves as cmc_off = true).  emb_dim may be any positive size; embedding_off = 1 needs none. */
    float w_association_emb, alpha_fixed_emb, aw_param;
    int embedding_off, aw_off;
    /* asso_func constructor argument (include/motcpp/tracker.hpp:47-55, AssociationFunction include/motcpp/utils/iou.hpp:371-411),
     * OC-SORT engines: 0 = "iou" (default), 6 = "centroid" (iou.hpp:298-330) with frame_width / frame_height = the size of
     * the frames the reference would read from `img` (src/trackers/ocsort.cpp:413).  The other variants ("hmiou", "giou",
     * "diou", "ciou") are only defined by the reference when the second box set has one row and are refused; every other
     * front-end uses plain IoU internally whatever asso_func says, as in the reference. */
    int asso_func, frame_width, frame_height;
    /* BoostTrackTracker ctor (include/motcpp/trackers/boosttrack.hpp:95-124); det_thresh, max_age, min_hits, iou_threshold
     * above are shared.  Camera-motion compensation and ReID are outside the hot path (the engine behaves as use_ecc =
     * false, with_reid = false); use_sb (a powf in the confidence boost) is refused. */
    int min_box_area;
    float aspect_ratio_thresh, lambda_iou, lambda_mhd, lambda_shape;
    int use_dlo_boost;
    float dlo_boost_coef;
    int use_sb, use_vt;
} mot_engine_config;

typedef struct mot_engine mot_engine;

/* fills cfg with the reference constructor defaults of `kind` */
int mot_engine_default_config(int kind, mot_engine_config* cfg);
int mot_engine_create(const mot_engine_config* cfg, mot_engine** out);
int mot_engine_destroy(mot_engine* e);
/* BaseTracker::reset(): clears every stream; ByteTrack/SORT/OC-SORT keep their ID counters
 * (reference bytetrack.hpp:38-40), BoT-SORT restarts at 0 (botsort.cpp:257). */
int mot_engine_reset(mot_engine* e);

/* One call = n_frames consecutive update()s for every stream, HOST buffers:
 *   dets   [n_frames][S][ld_dets][6]  rows [x1,y1,x2,y2,conf,cls]
 *   n_dets [n_frames][S]
 *   out    [n_frames][S][ld_out][8]   rows [x1,y1,x2,y2,id,conf,cls,det_ind]
 *   n_out  [n_frames][S]
 * Copies in, runs, copies out and returns when the results are in `out` (pipelined over stream
 * chunks).  n_frames = 1 is the reference's tracker->update(dets, img) for S trackers at once. */
int mot_engine_update_host(mot_engine* e, int n_frames, const float* dets, const int* n_dets, int ld_dets,
                           float* out, int* n_out, int ld_out);
/* Host buffers in, PACKED rows out - what n_frames x S calls of BaseTracker::update would have returned, back to back
 * (reference: the (M, 8) result matrix, e.g. src/trackers/bytetrack.cpp:596-620), without the padding of the call above:
 *   out_rows [out_cap_rows][8]   rows of frame f = t * S + s at out_rows[offsets[f] .. offsets[f + 1])
 *   offsets  [n_frames * S + 1]  exclusive row offsets (offsets[n_frames * S] = total rows)
 *   n_out    [n_frames][S]
 * max_rows bounds the rows of ONE frame (more are truncated and flagged, see mot_engine_check).  The rows are compacted
 * on the device, so only valid rows cross the bus (C2 workload: 361 of 512).  MOT_ERR_INVALID_ARGUMENT when out_cap_rows
 * is too small (state has advanced).  ByteTrack / SORT / OC-SORT engines, and BoT-SORT / StrongSORT without embeddings. */
int mot_engine_update_host_packed(mot_engine* e, int n_frames, const float* dets, const int* n_dets, int ld_dets, int max_rows,
                                  float* out_rows, long long out_cap_rows, long long* offsets, int* n_out);
/* Same contract with DEVICE buffers, asynchronous on `stream`; no host synchronisation. */
int mot_engine_update_device(mot_engine* e, int n_frames, const float* d_dets, const int* d_n_dets, int ld_dets,
                             float* d_out, int* d_n_out, int ld_out, void* stream);
/* BoT-SORT / StrongSORT / DeepOC-SORT engines (created with emb_dim = D > 0): the same calls with the detections' ReID embeddings,
 * embs [n_frames][S][ld_dets][D] fp32 (row j of a frame belongs to detection j; NULL = no embeddings).  Replaces the
 * `embs` argument of BaseTracker::update (include/motcpp/tracker.hpp:67-69, src/trackers/botsort.cpp:260-283). */
int mot_engine_update_host_embs(mot_engine* e, int n_frames, const float* dets, const int* n_dets, int ld_dets,
                                const float* embs, float* out, int* n_out, int ld_out);
int mot_engine_update_device_embs(mot_engine* e, int n_frames, const float* d_dets, const int* d_n_dets, int ld_dets,
                                  const float* d_embs, float* d_out, int* d_n_out, int ld_out, void* stream);
/* Per-stream error bits raised since the previous mot_engine_check / reset (0 = fine): 1 track capacity, 2 too many
 * detections, 4 output rows truncated, 8 Kalman fallback, 16 StrongSORT candidate table full.  READ-AND-CLEAR: a
 * transient condition is reported once, by the check that follows it.  Synchronises.  flags may be NULL; the return
 * value is MOT_OK or the most severe condition as a mot_status. */
int mot_engine_check(mot_engine* e, int* flags_per_stream);
/* Diagnostics (ByteTrack engines): per-phase SM cycle counters of the fused frame step, summed over streams and frames
 * since the previous call.  Returns the 32 counters accumulated so far (cycles32 may be NULL; slots 16.. are reserved for finer splits), then enables / disables
 * and clears the accounting.  Slots: 0 detections + confidence split, 1 pool lists, 2 predicted boxes, 3 first association
 * candidates, 4 components, 5 grouping, 6 exact solves, 7 harvest, 8 Kalman of the matches, 9 second association, 10
 * unconfirmed association, 11 new tracks, 12 list algebra, 13 duplicate removal, 14 output rows.  Synchronises. */
int mot_engine_profile(mot_engine* e, int enable, unsigned long long* cycles32);
/* Introspection for tests: header ints of one stream [n_active,n_lost,n_free,id_counter,frame,err,
 * n1,m1,n2,m2,n3,m3,dupA,dupB,..] (16 ints), and a dump of one list (0 active, 1 lost) as rows of
 * [id,state,is_activated,frame_id,start_frame,tracklet_len,mean 8,cov 64] (78 floats). */
int mot_engine_stream_header(mot_engine* e, int stream_index, int* hdr16);
/* BoT-SORT engines: list `which` (0 active, 1 lost) as rows of [id,state,is_activated,frame_id,start_frame,
 * tracklet_len,conf,cls,det_ind,has_feat,mean 8,cov 64] (82 floats); feats (nullable) receives the tracks'
 * smoothed features, emb_dim floats per row. */
int mot_engine_dump_bot(mot_engine* e, int stream_index, int which, float* rows82, float* feats, int cap_rows, int* n_rows);
int mot_engine_dump_list(mot_engine* e, int stream_index, int which, float* rows78, int cap_rows, int* n_rows);
/* DeepOC-SORT engines (created with embeddings on): the tracks' unit-length embeddings in track-list order (the row
 * order of mot_engine_dump_list), emb_dim floats per row. */
int mot_engine_dump_deep_embs(mot_engine* e, int stream_index, float* embs, int cap_rows, int* n_rows);
/* StrongSORT engines: the track list (reference order) as rows of [id,state,hits,0,time_since_update,conf,cls,det_ind,
 * has_feat,n_gallery_samples,mean 8,cov 64] (82 floats); feats (nullable) receives the smoothed features. */
int mot_engine_dump_strong(mot_engine* e, int stream_index, float* rows82, float* feats, int cap_rows, int* n_rows);
/* BoostTrack engines: the track list as rows of [id, age, hit_streak, time_since_update, conf, cls, det_ind, 0, x 8, P 8x8]
 * (80 floats; the covariance expanded from its four 2 x 2 blocks). */
int mot_engine_dump_boost(mot_engine* e, int stream_index, float* rows80, int cap_rows, int* n_rows);
/* launch geometry actually used (for the bench's gpu_launches / roofline bookkeeping) */
int mot_engine_info(mot_engine* e, int* threads_per_cta, int* smem_bytes, int* ctas, int* state_bytes_per_stream);

/* ---- standalone kernels (DEVICE pointers, asynchronous on `stream`) ------------------------ */
/* Kalman state records: kind 0 XYAH / 2 XYWH = 72 floats [mean 8 | cov 8x8]; kind 1 XYSR = 56
 * floats [x 7 | P 7x7].  n tracks, contiguous.
 *   initiate: KalmanFilterXYAH::initiate src/motion/kalman_filter.cpp:29-42 (+ xyah_kf.cpp:14-29),
 *             KalmanFilterXYSR ctor state xysr_kf.cpp:49-55, KalmanFilterXYWH::initiate xywh_kf.hpp:41-63
 *   predict : kalman_filter.cpp:44-58 / xysr_kf.cpp:71-77 / xywh_kf.hpp:70-94
 *   update  : kalman_filter.cpp:77-112 / xysr_kf.cpp:79-112 / xywh_kf.hpp:103-135
 *   gating  : kalman_filter.cpp:148-176 (XYAH; "maha" is the reference's d^T S^-2 d) /
 *             xywh_kf.hpp:140-177 (XYWH) */
int mot_kf_initiate(int kind, float* recs, const float* z4, long long n, void* stream);
/* flags: optional n bytes, bit0 = zero the height velocity before predicting (ByteTrack lost tracks);
 * q_xy_scaling / q_s_scaling only matter for XYSR (OC-SORT: 0.01 / 0.0001, SORT: 1 / 1). */
int mot_kf_predict(int kind, float* recs, const unsigned char* flags, long long n, float q_xy_scaling,
                   float q_s_scaling, void* stream);
/* conf: optional n floats (XYAH NSA scaling); fail: optional n bytes, 1 where the reference would
 * have taken its pseudo-inverse fallback (that record is left untouched). */
int mot_kf_update(int kind, float* recs, const float* z4, const float* conf, long long n, unsigned char* fail,
                  void* stream);
/* out (n_tracks x n_meas) row-major.  metric 0 = "maha", 1 = "gaussian" (XYAH only). */
int mot_kf_gating(int kind, const float* recs, int n_tracks, const float* meas4, int n_meas, int only_position,
                  int metric, float* out, void* stream);

/* N x M cost matrices.  a (n x 4), b (m x 4) xyxy boxes; out row-major with leading dimension ld.
 *   mode 0 iou_batch (include/motcpp/utils/iou.hpp:63-100), 1 iou_distance (src/utils/matching.cpp:62-65),
 *   2 iou_distance then fuse_score with conf[m] (matching.cpp:130-143) */
int mot_cost_iou(const float* a, int n, const float* b, int m, const float* conf, float* out, int ld, int mode,
                 void* stream);
/* hmiou_batch / giou_batch / diou_batch / centroid_batch / ciou_batch (include/motcpp/utils/iou.hpp:119-146, :151-187,
 * :258-293, :298-330, :197-253) evaluated pair-wise, kind 3 / 4 / 5 / 6 / 7 (frame_w, frame_h: centroid
 * normalisation).  The reference's own expressions only line up when b has one row (SURVEY 8 trap 11); on that domain the
 * results are identical.  ciou's arc tangent is the correctly rounded fp32 atan (Eigen's array atan is version dependent). */
int mot_cost_iou_variant(const float* a, int n, const float* b, int m, int kind, int frame_w, int frame_h, float* out, int ld,
                         void* stream);
/* OC-SORT association cost (ocsort_assoc::associate, src/trackers/ocsort.cpp:617-700): rows = detections
 * dets5 (n_dets x 5) [x1,y1,x2,y2,score], columns = tracks: predicted boxes trks4 (n_trks x 4), velocities
 * vel2 (n_trks x 2) as (dy, dx), k_previous_obs rows prev5 (n_trks x 5) [x1,y1,x2,y2,conf] (conf < 0: none).
 * out_cost = -(iou + valid * angle * inertia * score) and out_iou (nullable) = iou_batch(dets, trks), both
 * (n_dets x n_trks) row-major with leading dimension ld.  acosf is the correctly rounded one. */
int mot_cost_ocm(const float* dets5, int n_dets, const float* trks4, const float* vel2, const float* prev5, int n_trks,
                 float inertia, float* out_cost, float* out_iou, int ld, void* stream);
/* embedding_distance(metric="cosine") (src/utils/matching.cpp:67-92): max(0, 1 - t.d/(|t||d| + 1e-10)).
 * t (n x dim), d (m x dim) fp32 row-major; tcgen05 tensor-core contraction with a 3-term bf16
 * split (fp32-level accuracy), out row-major ld. */
int mot_cost_cosine(const float* t, int n, const float* d, int m, int dim, float* out, int ld, void* stream);

/* ---- StrongSORT cost builders (reference src/trackers/strongsort.cpp) ------------------------------- */
/* NearestNeighborDistanceMetric::distance, metric "cosine" (strongsort.cpp:240-334): samples (n_samples x dim) = every
 * target's gallery rows, seg[n_samples] = the target (output row) each sample belongs to, feats (m x dim) the raw
 * detection features.  out (n_targets x m, ld) = min over a target's samples of 1 - (s/|s|).(f/|f|); a target without
 * samples gets 1e5 (:271).  tcgen05 contraction (3-term bf16 split, fp32 accumulate) with the per-target minimum folded
 * into the GEMM epilogue. */
int mot_cost_nn_cosine(const float* samples, const int* seg, int n_samples, int n_targets, const float* feats, int m,
                       int dim, float* out, int ld, void* stream);
/* linear_assignment::gate_cost_matrix (strongsort.cpp:451-492) IN PLACE on cost (n_tracks x n_meas, ld): entries whose
 * gating distance (KalmanFilterXYAH::gating_distance "maha", kalman_filter.cpp:148-176) exceeds 9.4877 become gated_cost
 * (INFTY_COST = 1e5 in the reference), then cost = mc_lambda * cost + (1 - mc_lambda) * gating distance.
 * recs = XYAH records (72 floats per track), meas4 = (n_meas x 4) xyah rows (Detection::to_xyah). */
int mot_cost_gate(float* cost, int ld, const float* recs, int n_tracks, const float* meas4, int n_meas, float mc_lambda,
                  float gated_cost, int only_position, void* stream);
/* iou_matching::iou_cost (strongsort.cpp:502-585): 1 - IoU of tlwh boxes (union > 1e-6 guard); tsu (nullable) = the
 * tracks' time_since_update, rows with tsu > 1 are 1e5.  out (n x m, ld). */
int mot_cost_iou_tlwh(const float* trk_tlwh, const int* tsu, int n, const float* det_tlwh, int m, float* out, int ld,
                      void* stream);
/* deepocsort_assoc::compute_aw_max_metric (src/trackers/deepocsort.cpp:294-345), DeepOC-SORT's adaptive embedding weights
 * (SURVEY 8f-2): out = w * emb_cost with w = w_association_emb * row_weight(i) * col_weight(j), weight = 1 - max(second / max
 * - bottom, 0) / (1 - bottom) over the row's / column's two largest entries (0 when the largest is 0).  emb_cost, out (n x m). */
int mot_cost_aw_max_metric(const float* emb_cost, int n, int m, int ld, float w_association_emb, float bottom, float* out,
                           int ld_out, void* stream);
/* KalmanFilterXYSR::apply_affine_correction (src/motion/kalman_filters/xysr_kf.cpp:114-141) on n XYSR records in place:
 * the camera-motion warp of DeepOC-SORT.  m2x2 (row-major) and t2 are HOST pointers. */
int mot_kf_xysr_affine(float* recs, long long n, const float* m2x2, const float* t2, void* stream);

/* utils::linear_assignment (src/utils/matching.cpp:14-60) + LAPSolver (include/motcpp/association/
 * lap_solver.hpp:251-332): cost row-major (n x m), pairs with cost > thresh never match.
 * row2col[n] / col2row[m]: -1 = unmatched.  Batched: problem p reads cost + p*stride_cost and
 * writes row2col + p*n, col2row + p*m (n_rows/n_cols: optional per-problem sizes <= n, m). */
int mot_lap_device(const float* cost, int n, int m, int ld, float thresh, int* row2col, int* col2row, void* stream);
int mot_lap_batch_device(const float* cost, long long stride_cost, int n_problems, const int* n_rows,
                         const int* n_cols, int n, int m, int ld, float thresh, int* row2col, int* col2row,
                         void* stream);
/* The reference's dense LAPJV itself, step for step (lap_solver.hpp:36-231 on the (n+m)^2 extended matrix): one warp per
 * problem while n + m <= 384, one CTA per problem above that (any size up to n + m = 32000).  Same matches as mot_lap_*
 * whenever the optimum is unique, and the REFERENCE's choice when it is not (exact cost ties).  The OC-SORT, DeepOC-SORT and
 * StrongSORT engines run it on every frame that can tie (twin tracks, duplicated lists / rows). */
int mot_lap_jv_batch_device(const float* cost, long long stride_cost, int n_problems, int n, int m, int ld, float thresh,
                            int* row2col, int* col2row, void* stream);
/* convenience: host pointers, synchronous */
int mot_lap_host(const float* cost, int n, int m, int ld, float thresh, int* row2col, int* col2row);

#ifdef __cplusplus
}
#endif
#endif /* MOTB200_H */
