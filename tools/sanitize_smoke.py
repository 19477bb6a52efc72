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
print("sanitize_smoke done")
