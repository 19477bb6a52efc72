/*
 * oracle.h - CPU restatement of motcpp's association hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or
 * called from the product library (motcpp_b200/).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load liboracle.so, and there
 * only as the checker / the CPU baseline.
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 * Arithmetic is plain fp32 (fp64 inside the assignment solver), evaluated in the
 * operation order documented in DESIGN.md "Arithmetic contract", compiled with
 * -ffp-contract=off so no FMA is ever formed.
 *
 * Pinning status (see oracle/README.md): every function and all six tracker state
 * machines are compared with the reference's OWN sources compiled in place (oracle/_ref/libref_lap.so,
 * libref_core_tb.so, libref_core.so; tests/test_ref_pin.py): bit-exact when the reference's Eigen sums run in
 * textbook order, ids exact and <= 2e-5 on floats in Eigen 3.4's packet order.  The reference's own KATs
 * (tests/test_matching.cpp, test_iou.cpp, test_kalman_filter.cpp, test_sort.cpp, test_trackers.cpp) are a second check.
 *
 * All matrices are ROW-MAJOR float unless stated otherwise.
 */
#ifndef MOT_ORACLE_H
#define MOT_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------- box conversions (include/motcpp/utils/ops.hpp:15-211) ---------- */
void orc_xyxy2xywh(const float* in, float* out);
void orc_xywh2xyxy(const float* in, float* out);
void orc_xywh2tlwh(const float* in, float* out);
void orc_tlwh2xyah(const float* in, float* out);
void orc_xyah2xywh(const float* in, float* out);
void orc_xyxy2xysr(const float* in, float* out);
void orc_xysr2xyxy(const float* in, float* out);

/* ---------------- Kalman XYAH (src/motion/kalman_filter.cpp, kalman_filters/xyah_kf.cpp) */
void orc_kf_xyah_initiate(const float* z4, float* mean8, float* cov64);
void orc_kf_xyah_predict(float* mean8, float* cov64);
void orc_kf_xyah_project(const float* mean8, const float* cov64, float conf, float* pm4, float* pc16);
/* returns 0 on the Cholesky path, 1 when the reference would take its pseudo-inverse fallback
 * (state left untouched by the oracle in that case). */
int orc_kf_xyah_update(float* mean8, float* cov64, const float* z4, float conf);
/* metric: 0 = "maha" (reference computes d^T S^-2 d, kalman_filter.cpp:166-172), 1 = "gaussian" */
void orc_kf_xyah_gating(const float* mean8, const float* cov64, const float* meas, int m,
                        int only_position, int metric, float* out);

/* ---------------- Kalman XYSR (src/motion/kalman_filters/xysr_kf.cpp:10-112) ----------- */
/* q_xy_scale / q_s_scale: OC-SORT multiplies Q(4,4),Q(5,5) and Q(6,6) (ocsort.cpp:77-79); SORT passes 1,1 */
void orc_kf_xysr_init(const float* z4, float* x7, float* P49);
void orc_kf_xysr_predict(float* x7, float* P49, float q_xy_scale, float q_s_scale);
int orc_kf_xysr_update(float* x7, float* P49, const float* z4);

/* ---------------- Kalman XYWH (include/motcpp/motion/kalman_filters/xywh_kf.hpp:17-185) -- */
void orc_kf_xywh_initiate(const float* z4, float* mean8, float* cov64);
void orc_kf_xywh_predict(float* mean8, float* cov64);
void orc_kf_xywh_update(float* mean8, float* cov64, const float* z4);
void orc_kf_xywh_gating(const float* mean8, const float* cov64, const float* meas, int m,
                        int only_position, float* out);

/* ---------------- cost build (include/motcpp/utils/iou.hpp:63-100, src/utils/matching.cpp) */
void orc_iou_batch(const float* a, int n, const float* b, int m, float* out);      /* (n,m) */
void orc_iou_distance(const float* a, int n, const float* b, int m, float* out);   /* 1 - iou */
void orc_fuse_score(float* cost, int n, int m, const float* det_conf);             /* in place */
/* pair-wise hmiou (kind 3) / giou (4) / diou (5) / centroid (6): include/motcpp/utils/iou.hpp:119-330 */
float orc_atanf(float x);   /* correctly rounded fp32 arc tangent (ciou contract, cost.cpp) */
void orc_iou_variant(const float* a, int n, const float* b, int m, int kind, int frame_w, int frame_h, float* out);
/* metric 0 = cosine, 1 = euclidean (matching.cpp:67-107) */
void orc_embedding_distance(const float* t, int n, const float* d, int m, int dim, int metric, float* out);

/* ---------------- linear assignment (matching.cpp:14-60, lap_solver.hpp:36-332) --------- */
/* row2col[n], col2row[m] = -1 when unmatched; returns number of matches */
int orc_linear_assignment(const float* cost, int n, int m, int ld, float thresh,
                          int* row2col, int* col2row);

/* NOT reference behaviour: same problem with cost(i,j) - (64 j + (i j mod 64)) 2^-50 in fp64, i.e. exact ties
 * resolved towards the higher column - the CUDA OC-SORT kernel's rule for bit-identical twin tracks. */
int orc_linear_assignment_biased(const float* cost, int n, int m, int ld, float thresh, int* row2col, int* col2row);

/* ---------------- StrongSORT cost builders (src/trackers/strongsort.cpp) + XYSR affine correction ------------ */
/* NearestNeighborDistanceMetric::distance, "cosine" (:240-334): samples (n_samples x dim) with seg[s] = target row,
 * feats (m x dim) raw; out (n_targets x m) = min over the target's samples of 1 - cos; no samples -> 1e5 */
void orc_nn_cosine_distance(const float* samples, const int* seg, int n_samples, int n_targets, const float* feats,
                            int m, int dim, float* out);
/* linear_assignment::gate_cost_matrix (:451-492), in place; recs = XYAH records of 72 floats, meas (n_meas x 4) xyah */
void orc_gate_cost_matrix(float* cost, int ld, const float* recs, int n_tracks, const float* meas, int n_meas,
                          float mc_lambda, float gated_cost, int only_position);
/* iou_matching::iou_cost (:502-585): tlwh boxes, tsu nullable (rows with tsu > 1 -> 1e5); out (n x m) */
void orc_iou_cost_tlwh(const float* trk, const int* tsu, int n, const float* det, int m, float* out);
/* min_cost_matching's clamp (:372-377): entries > max_distance -> max_distance + 1e-5 */
void orc_clamp_cost(float* cost, int n, int m, int ld, float max_distance);
/* KalmanFilterXYSR::apply_affine_correction (src/motion/kalman_filters/xysr_kf.cpp:114-141); m2 row-major 2x2 */
void orc_kf_xysr_affine(float* x7, float* P49, const float* m2, const float* t2);

/* deepocsort_assoc::compute_aw_max_metric (src/trackers/deepocsort.cpp:294-345) */
void orc_aw_max_metric(const float* emb, int n, int m, int ld, float w_assoc, float bottom, float* out, int ld_out);
/* NOT reference behaviour: + (j + 1) 2^-50 on rows i >= n_first (StrongSORT's duplicated track rows), see lapjv.cpp */
int orc_linear_assignment_rowdup(const float* cost, int n, int m, int ld, float thresh, int n_first, int* row2col, int* col2row);

/* ---------------- StrongSORT (src/trackers/strongsort.cpp; ECC warp = identity, embeddings passed in) -------- */
typedef struct OrcStrongSort OrcStrongSort;
/* the StrongSORT ctor arguments that reach the association (include/motcpp/trackers/strongsort.hpp:287-305) */
OrcStrongSort* orc_strongsort_create(int max_age, float min_conf, float max_cos_dist, float max_iou_dist, int n_init,
                                     int nn_budget, float mc_lambda, float ema_alpha);
void orc_strongsort_destroy(OrcStrongSort*);
void orc_strongsort_reset(OrcStrongSort*);
/* dets (n x 6), embs (n x dim) or NULL; out rows [x1,y1,x2,y2,id,conf,cls,det_ind] */
int orc_strongsort_update(OrcStrongSort*, const float* dets, int n, const float* embs, int dim, float* out, int out_cap);
/* duplicated-row ties of the IoU stage: 0 = the reference's LAPJV, 1 = orc_linear_assignment_rowdup,
 * 2 = the CUDA kernel's policy (LAPJV while rows + columns <= 384, rowdup above) */
void orc_strongsort_set_tie_mode(OrcStrongSort*, int mode);
/* [rows_a, cols_a, rows_b, cols_b, n_matches_a, n_matches_b, dup_first, n_spawned] of the last update() */
void orc_strongsort_last_sizes(const OrcStrongSort*, int* out8);
int orc_strongsort_count(const OrcStrongSort*);
/* rows of [id,state,hits,age,tsu,conf,cls,det_ind,has_feat,n_samples,mean 8,cov 64] (82 floats) in track-list order */
int orc_strongsort_dump(const OrcStrongSort*, float* rows82, float* feats, int dim, int cap_rows);

/* ---------------- trackers (state machines) --------------------------------------------- */
typedef struct OrcByteTrack OrcByteTrack;
/* arguments follow ByteTrack's ctor (include/motcpp/trackers/bytetrack.hpp:97-110); the
 * unused BaseTracker knobs (per_class, nr_classes, asso_func, is_obb) are fixed at defaults. */
OrcByteTrack* orc_bytetrack_create(float det_thresh, int max_age, int max_obs, int min_hits,
                                   float iou_threshold, float min_conf, float track_thresh,
                                   float match_thresh, int track_buffer, int frame_rate);
void orc_bytetrack_destroy(OrcByteTrack*);
void orc_bytetrack_reset(OrcByteTrack*);
/* dets: (n,6) row-major [x1,y1,x2,y2,conf,cls]; out: capacity out_cap rows of 8 floats
 * [x1,y1,x2,y2,id,conf,cls,det_ind]; returns number of rows (or -needed if out_cap too small) */
int orc_bytetrack_update(OrcByteTrack*, const float* dets, int n, float* out, int out_cap);
/* introspection for parity tests */
int orc_bytetrack_counts(const OrcByteTrack*, int* n_active, int* n_lost);
/* dump list (0 = active, 1 = lost) as rows of [id, state, is_activated, frame_id, start_frame,
 * tracklet_len, mean0..7, cov0..63] = 78 floats; returns rows written */
int orc_bytetrack_dump(const OrcByteTrack*, int which, float* out, int cap_rows);
/* sizes of the three LAP sub-problems of the last update(): [n1,m1,n2,m2,n3,m3,n_dup_a,n_dup_b] */
void orc_bytetrack_last_sizes(const OrcByteTrack*, int* sizes8);

typedef struct OrcSort OrcSort;
OrcSort* orc_sort_create(float det_thresh, int max_age, int max_obs, int min_hits, float iou_threshold);
void orc_sort_destroy(OrcSort*);
void orc_sort_reset(OrcSort*);
int orc_sort_update(OrcSort*, const float* dets, int n, float* out, int out_cap);

/* ---------------- OC-SORT (src/trackers/ocsort.cpp) -------------------------------------- */
/* correctly rounded fp32 arc cosine (fixed fp64 operation sequence, see oracle/ocsort.cpp) */
float orc_acosf(float x);
/* ocsort_assoc::associate cost (ocsort.cpp:610-700): dets5 (n_dets x 5) [xyxy,score], trks4 (n_trks x 4),
 * vel2 (n_trks x 2) (dy,dx), prev5 (n_trks x 5) k_previous_obs rows; out_cost = -(iou + angle cost),
 * out_iou (nullable) = iou_batch(dets, trks); both (n_dets x n_trks) row-major */
void orc_ocm_cost(const float* dets5, int n_dets, const float* trks4, const float* vel2, const float* prev5,
                  int n_trks, float inertia, float* out_cost, float* out_iou);
typedef struct OrcOcSort OrcOcSort;
/* arguments follow OCSort's ctor (include/motcpp/trackers/ocsort.hpp:88-102) */
OrcOcSort* orc_ocsort_create(float det_thresh, int max_age, int max_obs, int min_hits, float iou_threshold,
                             float min_conf, int delta_t, float inertia, int use_byte, float q_xy_scaling,
                             float q_s_scaling);
/* asso_func (ocsort.hpp:93): 0 "iou", 6 "centroid" + the frame size the reference reads from img (returns -1 otherwise) */
int orc_ocsort_set_asso(OrcOcSort* s, int asso, int frame_w, int frame_h);
void orc_ocsort_destroy(OrcOcSort*);
void orc_ocsort_reset(OrcOcSort*);
int orc_ocsort_update(OrcOcSort*, const float* dets, int n, float* out, int out_cap);
int orc_ocsort_count(const OrcOcSort*);
/* tests: keep / fetch the first-association cost matrix (n_high x n_trk) of the last update() */
void orc_ocsort_capture(OrcOcSort*, int on);
/* 0 (default) = the reference's LAPJV tie-breaking; 1 = orc_linear_assignment_biased; 2 = the CUDA kernel's policy:
 * the reference's LAPJV while rows + columns <= 384 (kJvMax in csrc/ocsort_kernel.cuh), the biased solver above that */
void orc_ocsort_set_tie_mode(OrcOcSort*, int mode);
int orc_ocsort_last_cost(const OrcOcSort*, float* out, int cap);
/* [n_high, n_trk, used_lap, n_first_matches, n_left_dets, n_left_trks, n_rematched, n_spawned] of the last update() */
void orc_ocsort_last_sizes(const OrcOcSort*, int* sizes8);
/* rows of [id, age, hits, hit_streak, time_since_update, conf, cls, det_ind, last_obs 5, velocity 2,
 * k_previous_obs(delta_t) 5, x 7, P 49] = 76 floats */
int orc_ocsort_dump(const OrcOcSort*, float* out, int cap_rows);

/* ---------------- DeepOC-SORT (src/trackers/deepocsort.cpp; cmc_off = true, embeddings passed in) -------------- */
typedef struct OrcDeepOcSort OrcDeepOcSort;
/* arguments follow DeepOCSort's ctor (include/motcpp/trackers/deepocsort.hpp:95-114) without the ReID / BaseTracker knobs */
OrcDeepOcSort* orc_deepocsort_create(float det_thresh, int max_age, int max_obs, int min_hits, float iou_threshold, int delta_t,
                                     float inertia, float w_association_emb, float alpha_fixed_emb, float aw_param, int embedding_off,
                                     int aw_off, float q_xy_scaling, float q_s_scaling);
void orc_deepocsort_destroy(OrcDeepOcSort*);
void orc_deepocsort_reset(OrcDeepOcSort*);
/* dets (n x 6), embs (n x dim) or NULL (embedding_off); out rows [x1,y1,x2,y2,id,conf,cls,det_ind] */
int orc_deepocsort_update(OrcDeepOcSort*, const float* dets, int n, const float* embs, int dim, float* out, int out_cap);
int orc_deepocsort_count(const OrcDeepOcSort*);
/* [n_dets, n_trks, used_lap, n_first_matches, n_left_dets, n_left_trks, n_rematched, n_spawned] of the last update() */
void orc_deepocsort_last_sizes(const OrcDeepOcSort*, int* sizes8);
/* rows of [id, age, hits, hit_streak, time_since_update, conf, cls, det_ind, last_obs 5, velocity 2, x 7, P 49] = 71 floats */
int orc_deepocsort_dump(const OrcDeepOcSort*, float* out, float* embs, int dim, int cap_rows);

/* ---------------- BoT-SORT (src/trackers/botsort.cpp; cmc_method = "none", embeddings passed in) ------ */
typedef struct OrcBotSort OrcBotSort;
/* BotSort-specific ctor arguments (include/motcpp/trackers/botsort.hpp:108-134); BaseTracker knobs are unused by
 * BotSort::update */
OrcBotSort* orc_botsort_create(float track_high_thresh, float track_low_thresh, float new_track_thresh, int track_buffer,
                               float match_thresh, float proximity_thresh, float appearance_thresh, int frame_rate,
                               int fuse_first_associate, int with_reid);
void orc_botsort_destroy(OrcBotSort*);
void orc_botsort_reset(OrcBotSort*);
/* dets (n,6); embs (n,dim) row-major or NULL; returns rows written (or -needed) */
int orc_botsort_update(OrcBotSort*, const float* dets, int n, const float* embs, int dim, float* out, int out_cap);
int orc_botsort_counts(const OrcBotSort*, int* n_active, int* n_lost);          /* returns frame_count */
/* [n1, m1, n2, m2, n3, m3, n_new, n_lost_after] of the last update() */
void orc_botsort_last_sizes(const OrcBotSort*, int* sizes8);
/* list `which` (0 active, 1 lost): rows of [id, state, is_activated, frame_id, start_frame, tracklet_len, conf, cls,
 * det_ind, has_feat, mean 8, cov 64] = 82 floats; feats (nullable): smooth_feat rows of dim floats */
int orc_botsort_dump(const OrcBotSort*, int which, float* out, float* feats, int dim, int cap_rows);

/* ---------------- BoostTrack (src/trackers/boosttrack.cpp; default options, ECC / ReID off) ---------- */
typedef struct OrcBoostTrack OrcBoostTrack;
void orc_boost_iou_dist(const float* dets4, int n, const float* trk4, int m, float* out);                  /* :297-329 */
void orc_boost_mh_dist(const float* dets4, int n, const float* mean4, const float* var4, int m, float* out); /* :331-358 */
void orc_boost_cost(const float* iou_dist, const float* mh_dist, const float* emb, int n, int m, float lambda_iou, float lambda_mhd,
                    float lambda_shape, float* out);                                                        /* :571-626 */
OrcBoostTrack* orc_boosttrack_create(float det_thresh, int max_age, int max_obs, int min_hits, float iou_threshold, int min_box_area,
                                     float aspect_ratio_thresh, float lambda_iou, float lambda_mhd, float lambda_shape,
                                     int use_dlo_boost, float dlo_boost_coef, int use_vt);
void orc_boosttrack_destroy(OrcBoostTrack* s);
void orc_boosttrack_reset(OrcBoostTrack* s);
int orc_boosttrack_update(OrcBoostTrack* s, const float* dets, int n, float* out, int out_cap);
int orc_boosttrack_count(const OrcBoostTrack* s);
void orc_boosttrack_last_sizes(const OrcBoostTrack* s, int* out4);
int orc_boosttrack_dump(const OrcBoostTrack* s, float* out80, int cap_rows);

#ifdef __cplusplus
}
#endif
#endif
