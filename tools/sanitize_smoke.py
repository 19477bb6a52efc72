"""Small runs of every engine and stand-alone kernel for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motcpp_b200 import _lib, api, synth  # noqa: E402

T = int(os.environ.get("SAN_FRAMES", "12"))
d, c = synth.stress_stream(7, n_frames=T)
if os.environ.get("SAN_ONLY") == "boosttrack":                       # the BoostTrack engine alone: small shape, use_vt, and the 1536 x 512 shape
    for kw in (dict(max_age=6), dict(max_age=4, use_vt=True, det_thresh=0.4, min_hits=1)):
        trk = api.BoostTrack(track_capacity=256, max_dets=64, **kw)
        for t in range(T):
            trk.update(d[t, :c[t]], (540, 960))
    dd = synth.bytetrack_stream(0, n_frames=6)
    eng = api.Engine(_lib.TRACKER_BOOSTTRACK, 2, 1536, 512, max_age=3)
    eng.update(np.stack([dd, dd], 1), np.full((6, 2), 512, np.int32), ld_out=1536)
    eng.check()
    print("sanitize_smoke (boosttrack) done")
    sys.exit(0)
trk = api.OCSort(track_capacity=256, max_dets=64, use_byte=True)
for t in range(T):
    trk.update(d[t, :c[t]], (540, 960))
trk = api.Sort()
for t in range(T):
    trk.update(d[t, :c[t]])
trk = api.ByteTrack(track_capacity=256, max_dets=64)
for t in range(T):
    trk.update(d[t, :c[t]], (540, 960))
d, c, e = synth.stress_stream_reid(3, n_frames=T, dim=32)
trk = api.BotSort(emb_dim=32, track_capacity=256, max_dets=64)
for t in range(T):
    trk.update(d[t, :c[t]], (540, 960), e[t, :c[t]])
dd = synth.bytetrack_stream(0, n_frames=3)
eng = api.Engine(_lib.TRACKER_BYTETRACK, 2, 1536, 512)
eng.update(np.stack([dd, dd], 1), np.full((3, 2), 512, np.int32), ld_out=512)
eng.check()
rng = np.random.default_rng(0)
api.embedding_distance(rng.normal(size=(130, 64)).astype(np.float32), rng.normal(size=(70, 64)).astype(np.float32))
a = rng.uniform(0, 500, (40, 2)); A = np.concatenate([a, a + 50], 1).astype(np.float32)
api.iou_distance(A, A[:17])
api.linear_assignment(rng.random((30, 40)).astype(np.float32), 0.5)
api.ocm_cost(np.concatenate([A, rng.random((40, 1)).astype(np.float32)], 1), A[:9], rng.normal(size=(9, 2)).astype(np.float32),
             np.concatenate([A[:9], np.ones((9, 1), np.float32)], 1), 0.2)
trk = api.StrongSort(emb_dim=32, track_capacity=256, max_dets=64, nn_budget=4, max_cos_dist=0.4)
for t in range(T):
    trk.update(d[t, :c[t]], (540, 960), e[t, :c[t]])
smp, seg = rng.normal(size=(150, 64)).astype(np.float32), np.repeat(np.arange(15), 10).astype(np.int32)
api.nn_cosine_distance(smp, seg, 15, rng.normal(size=(70, 64)).astype(np.float32))
mu = np.tile(np.array([100, 100, 0.5, 80, 0, 0, 0, 0], np.float32), (9, 1)); cv = np.tile(np.eye(8, dtype=np.float32) * 10, (9, 1, 1))
api.gate_cost_matrix(rng.random((9, 40)).astype(np.float32), mu, cv, np.tile(np.array([101, 99, 0.5, 81], np.float32), (40, 1)), 0.98)
api.iou_cost_tlwh(A[:9], A, np.ones(9, np.int32))
api.aw_max_metric(rng.random((33, 70)).astype(np.float32))
api.KalmanFilterXYSR().apply_affine_correction(rng.normal(size=(5, 7)).astype(np.float32), np.tile(np.eye(7, dtype=np.float32), (5, 1, 1)),
                                               np.eye(2, dtype=np.float32), np.zeros(2, np.float32))
for _name in ("hmiou", "giou", "diou", "centroid", "ciou"):
    api.asso_batch(_name, A, A[:17], 640, 480)
# round 2: DeepOC-SORT (appearance terms, ordered-sum tiles, embedding EMA, twice-listed leftovers -> LAPJV every frame),
# OC-SORT with the centroid association, the reference-order LAPJV kernels (one warp; CTA-wide with the parallel record
# permutations, work arrays in shared memory and - forced - in global scratch), the packed host path
trk = api.DeepOCSort(max_age=6, track_capacity=256, max_dets=64)
for t in range(T):
    trk.update(d[t, :c[t]], (540, 960), e[t, :c[t]])
trk = api.OCSort(iou_threshold=0.95, asso_func="centroid", track_capacity=256, max_dets=64)
for t in range(T):
    trk.update(d[t, :c[t]], (540, 960))
trk = api.BoostTrack(max_age=6, track_capacity=256, max_dets=64)
for t in range(T):
    trk.update(d[t, :c[t]], (540, 960))
tie = (rng.integers(0, 4, (2, 90, 120)) / 4).astype(np.float32)
api.linear_assignment_reference_order(tie, 0.8)                       # one warp (rows + columns <= 384)
tie = (rng.integers(0, 4, (2, 200, 260)) / 4).astype(np.float32)
api.linear_assignment_reference_order(tie, 0.8)                       # CTA-wide, shared work arrays
os.environ["MOT_LAPJV_GLOBAL_WORK"] = "1"
api.linear_assignment_reference_order(tie, 0.8)                       # CTA-wide, global work arrays
del os.environ["MOT_LAPJV_GLOBAL_WORK"]
streams = [synth.stress_stream(400 + s, n_frames=4, n_obj=260, canvas=(1600, 900)) for s in range(2)]   # crowded: exact solves above 384
eng = api.Engine(_lib.TRACKER_DEEPOCSORT, 2, 1536, 512, emb_dim=8, max_age=3)
dets2 = np.stack([s_[0] for s_ in streams], 1); cnt2 = np.stack([s_[1] for s_ in streams], 1).astype(np.int32)
eng.update(dets2, cnt2, ld_out=1536, embs=rng.normal(size=dets2.shape[:3] + (8,)).astype(np.float32))
eng.check()
eng = api.Engine(_lib.TRACKER_BYTETRACK, 2, 1536, 512)
eng.update_packed(np.stack([dd, dd], 1), np.full((3, 2), 512, np.int32), max_rows=512)
eng.check()
print("sanitize_smoke done")
