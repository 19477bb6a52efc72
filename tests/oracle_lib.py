"""ctypes bindings for oracle/liboracle.so and oracle/_ref/libref_lap.so.

TEST INFRASTRUCTURE: imported only by tests/, bench.py's cpu_baseline / --impl reference legs
and __graft_entry__.smoke().  The product package (motcpp_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None
_REF = None

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build_oracle(force: bool = False) -> None:
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"], stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        build_oracle()
        L = C.CDLL(os.path.join(ORACLE_DIR, "liboracle.so"))
        L.orc_linear_assignment.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_float, i32p, i32p]
        L.orc_linear_assignment.restype = C.c_int
        for name in ("orc_iou_batch", "orc_iou_distance"):
            getattr(L, name).argtypes = [f32p, C.c_int, f32p, C.c_int, f32p]
            getattr(L, name).restype = None
        L.orc_fuse_score.argtypes = [f32p, C.c_int, C.c_int, f32p]
        L.orc_embedding_distance.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_int, C.c_int, f32p]
        for name in ("orc_xyxy2xywh", "orc_xywh2xyxy", "orc_xywh2tlwh", "orc_tlwh2xyah", "orc_xyah2xywh",
                     "orc_xyxy2xysr", "orc_xysr2xyxy"):
            getattr(L, name).argtypes = [f32p, f32p]
        L.orc_kf_xyah_initiate.argtypes = [f32p, f32p, f32p]
        L.orc_kf_xyah_predict.argtypes = [f32p, f32p]
        L.orc_kf_xyah_project.argtypes = [f32p, f32p, C.c_float, f32p, f32p]
        L.orc_kf_xyah_update.argtypes = [f32p, f32p, f32p, C.c_float]
        L.orc_kf_xyah_update.restype = C.c_int
        L.orc_kf_xyah_gating.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, f32p]
        L.orc_kf_xysr_init.argtypes = [f32p, f32p, f32p]
        L.orc_kf_xysr_predict.argtypes = [f32p, f32p, C.c_float, C.c_float]
        L.orc_kf_xysr_update.argtypes = [f32p, f32p, f32p]
        L.orc_kf_xysr_update.restype = C.c_int
        L.orc_kf_xywh_initiate.argtypes = [f32p, f32p, f32p]
        L.orc_kf_xywh_predict.argtypes = [f32p, f32p]
        L.orc_kf_xywh_update.argtypes = [f32p, f32p, f32p]
        L.orc_kf_xywh_gating.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, f32p]
        L.orc_bytetrack_create.argtypes = [C.c_float, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                           C.c_float, C.c_float, C.c_int, C.c_int]
        L.orc_bytetrack_create.restype = C.c_void_p
        L.orc_bytetrack_destroy.argtypes = [C.c_void_p]
        L.orc_bytetrack_reset.argtypes = [C.c_void_p]
        L.orc_bytetrack_update.argtypes = [C.c_void_p, f32p, C.c_int, f32p, C.c_int]
        L.orc_bytetrack_update.restype = C.c_int
        L.orc_bytetrack_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_bytetrack_counts.restype = C.c_int
        L.orc_bytetrack_dump.argtypes = [C.c_void_p, C.c_int, f32p, C.c_int]
        L.orc_bytetrack_dump.restype = C.c_int
        L.orc_bytetrack_last_sizes.argtypes = [C.c_void_p, i32p]
        L.orc_sort_create.argtypes = [C.c_float, C.c_int, C.c_int, C.c_int, C.c_float]
        L.orc_sort_create.restype = C.c_void_p
        L.orc_sort_destroy.argtypes = [C.c_void_p]
        L.orc_sort_reset.argtypes = [C.c_void_p]
        L.orc_sort_update.argtypes = [C.c_void_p, f32p, C.c_int, f32p, C.c_int]
        L.orc_sort_update.restype = C.c_int
        L.orc_acosf.argtypes = [C.c_float]
        L.orc_acosf.restype = C.c_float
        L.orc_ocm_cost.argtypes = [f32p, C.c_int, f32p, f32p, f32p, C.c_int, C.c_float, f32p, f32p]
        L.orc_strongsort_create.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_float, C.c_float]
        L.orc_strongsort_create.restype = C.c_void_p
        L.orc_strongsort_destroy.argtypes = [C.c_void_p]
        L.orc_strongsort_reset.argtypes = [C.c_void_p]
        L.orc_strongsort_update.argtypes = [C.c_void_p, f32p, C.c_int, C.c_void_p, C.c_int, f32p, C.c_int]
        L.orc_strongsort_set_tie_mode.argtypes = [C.c_void_p, C.c_int]
        L.orc_strongsort_last_sizes.argtypes = [C.c_void_p, i32p]
        L.orc_strongsort_count.argtypes = [C.c_void_p]
        L.orc_strongsort_dump.argtypes = [C.c_void_p, f32p, C.c_void_p, C.c_int, C.c_int]
        L.orc_nn_cosine_distance.argtypes = [f32p, i32p, C.c_int, C.c_int, f32p, C.c_int, C.c_int, f32p]
        L.orc_gate_cost_matrix.argtypes = [f32p, C.c_int, f32p, C.c_int, f32p, C.c_int, C.c_float, C.c_float, C.c_int]
        L.orc_iou_cost_tlwh.argtypes = [f32p, C.c_void_p, C.c_int, f32p, C.c_int, f32p]
        L.orc_clamp_cost.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_float]
        L.orc_kf_xysr_affine.argtypes = [f32p, f32p, f32p, f32p]
        L.orc_atanf.argtypes, L.orc_atanf.restype = [C.c_float], C.c_float
        L.orc_boost_iou_dist.argtypes = [f32p, C.c_int, f32p, C.c_int, f32p]
        L.orc_boost_mh_dist.argtypes = [f32p, C.c_int, f32p, f32p, C.c_int, f32p]
        L.orc_boost_cost.argtypes = [f32p, f32p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, f32p]
        L.orc_boosttrack_create.argtypes = [C.c_float, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float,
                                            C.c_float, C.c_int, C.c_float, C.c_int]
        L.orc_boosttrack_create.restype = C.c_void_p
        L.orc_boosttrack_destroy.argtypes = [C.c_void_p]
        L.orc_boosttrack_reset.argtypes = [C.c_void_p]
        L.orc_boosttrack_update.argtypes = [C.c_void_p, f32p, C.c_int, f32p, C.c_int]
        L.orc_boosttrack_count.argtypes = [C.c_void_p]
        L.orc_boosttrack_last_sizes.argtypes = [C.c_void_p, i32p]
        L.orc_boosttrack_dump.argtypes = [C.c_void_p, f32p, C.c_int]
        L.orc_iou_variant.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
        L.orc_aw_max_metric.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, f32p, C.c_int]
        L.orc_ocsort_create.argtypes = [C.c_float, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int,
                                        C.c_float, C.c_int, C.c_float, C.c_float]
        L.orc_ocsort_create.restype = C.c_void_p
        L.orc_ocsort_set_asso.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_ocsort_destroy.argtypes = [C.c_void_p]
        L.orc_ocsort_reset.argtypes = [C.c_void_p]
        L.orc_ocsort_update.argtypes = [C.c_void_p, f32p, C.c_int, f32p, C.c_int]
        L.orc_ocsort_update.restype = C.c_int
        L.orc_ocsort_count.argtypes = [C.c_void_p]
        L.orc_ocsort_count.restype = C.c_int
        L.orc_ocsort_set_tie_mode.argtypes = [C.c_void_p, C.c_int]
        L.orc_linear_assignment_biased.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_float, i32p, i32p]
        L.orc_linear_assignment_biased.restype = C.c_int
        L.orc_ocsort_capture.argtypes = [C.c_void_p, C.c_int]
        L.orc_ocsort_last_cost.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_ocsort_last_cost.restype = C.c_int
        L.orc_ocsort_last_sizes.argtypes = [C.c_void_p, i32p]
        L.orc_ocsort_dump.argtypes = [C.c_void_p, f32p, C.c_int]
        L.orc_ocsort_dump.restype = C.c_int
        L.orc_botsort_create.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float,
                                         C.c_int, C.c_int, C.c_int]
        L.orc_botsort_create.restype = C.c_void_p
        L.orc_botsort_destroy.argtypes = [C.c_void_p]
        L.orc_botsort_reset.argtypes = [C.c_void_p]
        L.orc_botsort_update.argtypes = [C.c_void_p, f32p, C.c_int, C.c_void_p, C.c_int, f32p, C.c_int]
        L.orc_botsort_update.restype = C.c_int
        L.orc_botsort_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_botsort_counts.restype = C.c_int
        L.orc_botsort_last_sizes.argtypes = [C.c_void_p, i32p]
        L.orc_botsort_dump.argtypes = [C.c_void_p, C.c_int, f32p, C.c_void_p, C.c_int, C.c_int]
        L.orc_botsort_dump.restype = C.c_int
        L.orc_deepocsort_create.argtypes = [C.c_float, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float,
                                            C.c_float, C.c_int, C.c_int, C.c_float, C.c_float]
        L.orc_deepocsort_create.restype = C.c_void_p
        L.orc_deepocsort_destroy.argtypes = [C.c_void_p]
        L.orc_deepocsort_reset.argtypes = [C.c_void_p]
        L.orc_deepocsort_update.argtypes = [C.c_void_p, f32p, C.c_int, C.c_void_p, C.c_int, f32p, C.c_int]
        L.orc_deepocsort_count.argtypes = [C.c_void_p]
        L.orc_deepocsort_last_sizes.argtypes = [C.c_void_p, i32p]
        L.orc_deepocsort_dump.argtypes = [C.c_void_p, f32p, C.c_void_p, C.c_int, C.c_int]
        _LIB = L
    return _LIB


def ref_lap():
    """The reference's real LAP solver (oracle/_ref/libref_lap.so) or None if it was never built."""
    global _REF
    if _REF is None:
        so = os.path.join(ORACLE_DIR, "_ref", "libref_lap.so")
        if not os.path.exists(so):
            return None
        R = C.CDLL(so)
        R.ref_linear_assignment.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_float, i32p, i32p]
        R.ref_linear_assignment.restype = C.c_int
        _REF = R
    return _REF


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ------------------------------------------------------------------ functional wrappers
def linear_assignment(cost, thresh, use_ref=False):
    cost = _f32(cost)
    n, m = cost.shape
    r2c = np.full(n, -1, np.int32)
    c2r = np.full(m, -1, np.int32)
    if n and m:
        fn = ref_lap().ref_linear_assignment if use_ref else lib().orc_linear_assignment
        fn(cost, n, m, m, float(thresh), r2c, c2r)
    return r2c, c2r


def iou_batch(a, b):
    a, b = _f32(a).reshape(-1, 4), _f32(b).reshape(-1, 4)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    if out.size:
        lib().orc_iou_batch(a, a.shape[0], b, b.shape[0], out)
    return out


def iou_distance(a, b):
    a, b = _f32(a).reshape(-1, 4), _f32(b).reshape(-1, 4)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    if out.size:
        lib().orc_iou_distance(a, a.shape[0], b, b.shape[0], out)
    return out


def fuse_score(cost, conf):
    cost = _f32(cost).copy()
    if cost.size:
        lib().orc_fuse_score(cost, cost.shape[0], cost.shape[1], _f32(conf))
    return cost


def embedding_distance(t, d, metric="cosine"):
    t, d = _f32(t), _f32(d)
    out = np.zeros((t.shape[0], d.shape[0]), np.float32)
    if out.size:
        lib().orc_embedding_distance(t, t.shape[0], d, d.shape[0], t.shape[1], 0 if metric == "cosine" else 1, out)
    return out


def convert(name, box):
    out = np.zeros(4, np.float32)
    getattr(lib(), "orc_" + name)(_f32(box), out)
    return out


class KFXYAH:
    @staticmethod
    def initiate(z):
        m, c = np.zeros(8, np.float32), np.zeros(64, np.float32)
        lib().orc_kf_xyah_initiate(_f32(z), m, c)
        return m, c.reshape(8, 8)

    @staticmethod
    def predict(mean, cov):
        m, c = _f32(mean).copy(), _f32(cov).reshape(-1).copy()
        lib().orc_kf_xyah_predict(m, c)
        return m, c.reshape(8, 8)

    @staticmethod
    def update(mean, cov, z, conf=0.0):
        m, c = _f32(mean).copy(), _f32(cov).reshape(-1).copy()
        rc = lib().orc_kf_xyah_update(m, c, _f32(z), float(conf))
        return m, c.reshape(8, 8), rc

    @staticmethod
    def gating(mean, cov, meas, only_position=False, metric="maha"):
        meas = _f32(meas).reshape(-1, 4)
        out = np.zeros(meas.shape[0], np.float32)
        lib().orc_kf_xyah_gating(_f32(mean), _f32(cov).reshape(-1), meas, meas.shape[0], int(only_position),
                                 0 if metric == "maha" else 1, out)
        return out


class KFXYSR:
    @staticmethod
    def init(z):
        x, p = np.zeros(7, np.float32), np.zeros(49, np.float32)
        lib().orc_kf_xysr_init(_f32(z), x, p)
        return x, p.reshape(7, 7)

    @staticmethod
    def predict(x, p, q_xy=1.0, q_s=1.0):
        x, p = _f32(x).copy(), _f32(p).reshape(-1).copy()
        lib().orc_kf_xysr_predict(x, p, float(q_xy), float(q_s))
        return x, p.reshape(7, 7)

    @staticmethod
    def update(x, p, z):
        x, p = _f32(x).copy(), _f32(p).reshape(-1).copy()
        rc = lib().orc_kf_xysr_update(x, p, _f32(z))
        return x, p.reshape(7, 7), rc


class KFXYWH:
    @staticmethod
    def initiate(z):
        m, c = np.zeros(8, np.float32), np.zeros(64, np.float32)
        lib().orc_kf_xywh_initiate(_f32(z), m, c)
        return m, c.reshape(8, 8)

    @staticmethod
    def predict(mean, cov):
        m, c = _f32(mean).copy(), _f32(cov).reshape(-1).copy()
        lib().orc_kf_xywh_predict(m, c)
        return m, c.reshape(8, 8)

    @staticmethod
    def update(mean, cov, z):
        m, c = _f32(mean).copy(), _f32(cov).reshape(-1).copy()
        lib().orc_kf_xywh_update(m, c, _f32(z))
        return m, c.reshape(8, 8)

    @staticmethod
    def gating(mean, cov, meas, only_position=False):
        meas = _f32(meas).reshape(-1, 4)
        out = np.zeros(meas.shape[0], np.float32)
        lib().orc_kf_xywh_gating(_f32(mean), _f32(cov).reshape(-1), meas, meas.shape[0], int(only_position), out)
        return out


class ByteTrack:
    """Oracle ByteTrack with the reference's constructor argument order (bytetrack.hpp:97-110)."""

    def __init__(self, det_thresh=0.3, max_age=30, max_obs=50, min_hits=3, iou_threshold=0.3,
                 min_conf=0.1, track_thresh=0.45, match_thresh=0.8, track_buffer=25, frame_rate=30):
        self._h = lib().orc_bytetrack_create(det_thresh, max_age, max_obs, min_hits, iou_threshold,
                                             min_conf, track_thresh, match_thresh, track_buffer, frame_rate)
        self._cap = 4096
        self._out = np.zeros((self._cap, 8), np.float32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_bytetrack_destroy(self._h)
            self._h = None

    def reset(self):
        lib().orc_bytetrack_reset(self._h)

    def update(self, dets):
        dets = _f32(dets).reshape(-1, 6)
        n = lib().orc_bytetrack_update(self._h, dets, dets.shape[0], self._out, self._cap)
        if n < 0:
            self._cap = -n * 2
            self._out = np.zeros((self._cap, 8), np.float32)
            raise RuntimeError("oracle output buffer too small; state already advanced")
        return self._out[:n].copy()

    def counts(self):
        a, b = C.c_int(), C.c_int()
        lib().orc_bytetrack_counts(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def dump(self, which):
        na, nl = self.counts()
        cap = max(1, na if which == 0 else nl)
        buf = np.zeros((cap, 78), np.float32)
        k = lib().orc_bytetrack_dump(self._h, which, buf, cap)
        return buf[:k]

    def last_sizes(self):
        s = np.zeros(8, np.int32)
        lib().orc_bytetrack_last_sizes(self._h, s)
        return s


class Sort:
    def __init__(self, det_thresh=0.3, max_age=1, max_obs=50, min_hits=3, iou_threshold=0.3):
        self._h = lib().orc_sort_create(det_thresh, max_age, max_obs, min_hits, iou_threshold)
        self._out = np.zeros((4096, 8), np.float32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_sort_destroy(self._h)
            self._h = None

    def reset(self):
        lib().orc_sort_reset(self._h)

    def update(self, dets):
        dets = _f32(dets).reshape(-1, 6)
        n = lib().orc_sort_update(self._h, dets, dets.shape[0], self._out, self._out.shape[0])
        assert n >= 0
        return self._out[:n].copy()


def nn_cosine_distance(samples, seg, n_targets, feats):
    """NearestNeighborDistanceMetric::distance "cosine" (strongsort.cpp:240-334) -> (n_targets, m)."""
    samples, feats = _f32(samples), _f32(feats)
    seg = np.ascontiguousarray(seg, np.int32)
    out = np.zeros((n_targets, feats.shape[0]), np.float32)
    if out.size:
        lib().orc_nn_cosine_distance(samples.reshape(-1), seg, samples.shape[0], n_targets, feats.reshape(-1),
                                     feats.shape[0], feats.shape[1], out)
    return out


def gate_cost_matrix(cost, means, covs, meas, mc_lambda, gated_cost=1e5, only_position=False):
    """linear_assignment::gate_cost_matrix (strongsort.cpp:451-492); returns the gated copy."""
    out = _f32(cost).copy()
    n, m = out.shape
    recs = np.ascontiguousarray(np.concatenate([_f32(means).reshape(n, 8), _f32(covs).reshape(n, 64)], axis=1))
    if out.size:
        lib().orc_gate_cost_matrix(out, m, recs.reshape(-1), n, _f32(meas).reshape(-1), m, float(mc_lambda),
                                   float(gated_cost), int(only_position))
    return out


def iou_cost_tlwh(trk, det, tsu=None):
    """iou_matching::iou_cost (strongsort.cpp:502-585)."""
    trk, det = _f32(trk).reshape(-1, 4), _f32(det).reshape(-1, 4)
    out = np.zeros((trk.shape[0], det.shape[0]), np.float32)
    t = np.ascontiguousarray(tsu, np.int32) if tsu is not None else None
    if out.size:
        lib().orc_iou_cost_tlwh(trk, t.ctypes.data if t is not None else None, trk.shape[0], det, det.shape[0], out)
    return out


def iou_variant(kind, a, b, frame_w=0, frame_h=0):
    """pair-wise hmiou (3) / giou (4) / diou (5) / centroid (6), include/motcpp/utils/iou.hpp:119-330."""
    a, b = _f32(a).reshape(-1, 4), _f32(b).reshape(-1, 4)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    if out.size:
        lib().orc_iou_variant(a, a.shape[0], b, b.shape[0], int(kind), int(frame_w), int(frame_h), out)
    return out


def aw_max_metric(emb, w_assoc=0.5, bottom=0.5):
    """deepocsort_assoc::compute_aw_max_metric (deepocsort.cpp:294-345)."""
    emb = _f32(emb)
    out = np.zeros_like(emb)
    if emb.size:
        lib().orc_aw_max_metric(emb, emb.shape[0], emb.shape[1], emb.shape[1], float(w_assoc), float(bottom), out, emb.shape[1])
    return out


def kf_xysr_affine(x7, P, m2, t2):
    """KalmanFilterXYSR::apply_affine_correction (xysr_kf.cpp:114-141) -> (x, P)."""
    x, p = _f32(x7).copy(), _f32(P).reshape(-1).copy()
    lib().orc_kf_xysr_affine(x, p, _f32(m2).reshape(-1), _f32(t2))
    return x, p.reshape(7, 7)


def ocm_cost(dets5, trks4, vel2, prev5, inertia):
    """(cost, iou), both (n_dets, n_trks): ocsort_assoc::associate's -(iou + angle cost) and iou_batch(dets, trks)."""
    dets5, trks4 = _f32(dets5).reshape(-1, 5), _f32(trks4).reshape(-1, 4)
    vel2, prev5 = _f32(vel2).reshape(-1, 2), _f32(prev5).reshape(-1, 5)
    cost = np.zeros((dets5.shape[0], trks4.shape[0]), np.float32)
    iou = np.zeros_like(cost)
    if cost.size:
        lib().orc_ocm_cost(dets5, dets5.shape[0], trks4, vel2, prev5, trks4.shape[0], float(inertia), cost, iou)
    return cost, iou


class OCSort:
    """Oracle OC-SORT with the reference's constructor argument order (ocsort.hpp:88-102)."""

    def __init__(self, det_thresh=0.2, max_age=30, max_obs=50, min_hits=3, iou_threshold=0.3, min_conf=0.1,
                 delta_t=3, inertia=0.2, use_byte=False, q_xy_scaling=0.01, q_s_scaling=0.0001, tie_mode=0,
                 asso_func="iou", frame=(1920, 1080)):
        self._h = lib().orc_ocsort_create(det_thresh, max_age, max_obs, min_hits, iou_threshold, min_conf, delta_t,
                                          inertia, int(use_byte), q_xy_scaling, q_s_scaling)
        lib().orc_ocsort_set_tie_mode(self._h, int(tie_mode))
        # asso_func: "iou" | "centroid" (the reference's other variants are undefined beyond one row); frame = (width, height)
        assert lib().orc_ocsort_set_asso(self._h, {"iou": 0, "centroid": 6}[asso_func], int(frame[0]), int(frame[1])) == 0
        self._out = np.zeros((8192, 8), np.float32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_ocsort_destroy(self._h)
            self._h = None

    def reset(self):
        lib().orc_ocsort_reset(self._h)

    def update(self, dets):
        dets = _f32(dets).reshape(-1, 6)
        n = lib().orc_ocsort_update(self._h, dets, dets.shape[0], self._out, self._out.shape[0])
        assert n >= 0
        return self._out[:n].copy()

    def dump(self):
        cap = max(1, lib().orc_ocsort_count(self._h))
        buf = np.zeros((cap, 76), np.float32)
        k = lib().orc_ocsort_dump(self._h, buf, cap)
        return buf[:k]

    def last_sizes(self):
        s = np.zeros(8, np.int32)
        lib().orc_ocsort_last_sizes(self._h, s)
        return s

    def capture(self, on=True):
        lib().orc_ocsort_capture(self._h, int(on))

    def last_cost(self):
        n = lib().orc_ocsort_last_cost(self._h, None, 0)
        sz = self.last_sizes()
        buf = np.zeros(max(n, 1), np.float32)
        lib().orc_ocsort_last_cost(self._h, buf.ctypes.data_as(C.c_void_p), n)
        return buf[:n].reshape(int(sz[0]), int(sz[1])) if n else np.zeros((0, 0), np.float32)


class BotSort:
    """Oracle BoT-SORT; BotSort-specific constructor arguments in the reference's order (botsort.hpp:121-133),
    cmc_method fixed to "none", embeddings passed to update()."""

    def __init__(self, track_high_thresh=0.5, track_low_thresh=0.1, new_track_thresh=0.6, track_buffer=30,
                 match_thresh=0.8, proximity_thresh=0.5, appearance_thresh=0.25, frame_rate=30,
                 fuse_first_associate=False, with_reid=True):
        self._h = lib().orc_botsort_create(track_high_thresh, track_low_thresh, new_track_thresh, track_buffer,
                                           match_thresh, proximity_thresh, appearance_thresh, frame_rate,
                                           int(fuse_first_associate), int(with_reid))
        self._out = np.zeros((8192, 8), np.float32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_botsort_destroy(self._h)
            self._h = None

    def reset(self):
        lib().orc_botsort_reset(self._h)

    def update(self, dets, embs=None):
        dets = _f32(dets).reshape(-1, 6)
        if embs is not None and np.size(embs):
            embs = _f32(embs).reshape(dets.shape[0], -1)
            ep, dim = embs.ctypes.data_as(C.c_void_p), embs.shape[1]
        else:
            ep, dim = None, 0
        n = lib().orc_botsort_update(self._h, dets, dets.shape[0], ep, dim, self._out, self._out.shape[0])
        assert n >= 0
        return self._out[:n].copy()

    def counts(self):
        a, b = C.c_int(), C.c_int()
        lib().orc_botsort_counts(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def dump(self, which, dim=0):
        na, nl = self.counts()
        cap = max(1, na if which == 0 else nl)
        buf = np.zeros((cap, 82), np.float32)
        feats = np.zeros((cap, max(dim, 1)), np.float32)
        k = lib().orc_botsort_dump(self._h, which, buf, feats.ctypes.data_as(C.c_void_p) if dim else None, dim, cap)
        return (buf[:k], feats[:k]) if dim else buf[:k]

    def last_sizes(self):
        s = np.zeros(8, np.int32)
        lib().orc_botsort_last_sizes(self._h, s)
        return s


class StrongSort:
    """Oracle StrongSORT: the constructor arguments that reach the association, reference names and defaults
    (include/motcpp/trackers/strongsort.hpp:287-305); ECC warp = identity, embeddings passed to update()."""

    def __init__(self, max_age=30, min_conf=0.1, max_cos_dist=0.2, max_iou_dist=0.7, n_init=3, nn_budget=100,
                 mc_lambda=0.98, ema_alpha=0.9, tie_mode=0):
        self._h = lib().orc_strongsort_create(max_age, min_conf, max_cos_dist, max_iou_dist, n_init, nn_budget,
                                              mc_lambda, ema_alpha)
        lib().orc_strongsort_set_tie_mode(self._h, tie_mode)
        self._out = np.zeros((8192, 8), np.float32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_strongsort_destroy(self._h)
            self._h = None

    def reset(self):
        lib().orc_strongsort_reset(self._h)

    def update(self, dets, embs=None):
        dets = _f32(dets).reshape(-1, 6)
        if embs is not None and np.size(embs):
            embs = _f32(embs).reshape(dets.shape[0], -1)
            ep, dim = embs.ctypes.data_as(C.c_void_p), embs.shape[1]
        else:
            ep, dim = None, 0
        n = lib().orc_strongsort_update(self._h, dets, dets.shape[0], ep, dim, self._out, self._out.shape[0])
        assert n >= 0
        return self._out[:n].copy()

    def count(self):
        return lib().orc_strongsort_count(self._h)

    def dump(self, dim=0):
        cap = max(1, self.count())
        buf = np.zeros((cap, 82), np.float32)
        feats = np.zeros((cap, max(dim, 1)), np.float32)
        k = lib().orc_strongsort_dump(self._h, buf, feats.ctypes.data_as(C.c_void_p) if dim else None, dim, cap)
        return (buf[:k], feats[:k]) if dim else buf[:k]

    def last_sizes(self):
        s = np.zeros(8, np.int32)
        lib().orc_strongsort_last_sizes(self._h, s)
        return s


class DeepOCSort:
    """Oracle DeepOC-SORT: the constructor arguments that reach the association, reference names and defaults
    (include/motcpp/trackers/deepocsort.hpp:95-114); cmc_off = True, embeddings passed to update()."""

    def __init__(self, det_thresh=0.3, max_age=30, max_obs=50, min_hits=3, iou_threshold=0.3, delta_t=3, inertia=0.2,
                 w_association_emb=0.5, alpha_fixed_emb=0.95, aw_param=0.5, embedding_off=False, aw_off=False,
                 q_xy_scaling=0.01, q_s_scaling=0.0001):
        self._h = lib().orc_deepocsort_create(det_thresh, max_age, max_obs, min_hits, iou_threshold, delta_t, inertia,
                                              w_association_emb, alpha_fixed_emb, aw_param, int(embedding_off), int(aw_off),
                                              q_xy_scaling, q_s_scaling)
        self._out = np.zeros((8192, 8), np.float32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_deepocsort_destroy(self._h)
            self._h = None

    def reset(self):
        lib().orc_deepocsort_reset(self._h)

    def update(self, dets, embs=None):
        dets = _f32(dets).reshape(-1, 6)
        if embs is not None and np.size(embs):
            embs = _f32(embs).reshape(dets.shape[0], -1)
            ep, dim = embs.ctypes.data_as(C.c_void_p), embs.shape[1]
        else:
            ep, dim = None, 0
        n = lib().orc_deepocsort_update(self._h, dets, dets.shape[0], ep, dim, self._out, self._out.shape[0])
        assert n >= 0
        return self._out[:n].copy()

    def count(self):
        return lib().orc_deepocsort_count(self._h)

    def dump(self, dim=0):
        cap = max(1, self.count())
        buf = np.zeros((cap, 71), np.float32)
        embs = np.zeros((cap, max(dim, 1)), np.float32)
        k = lib().orc_deepocsort_dump(self._h, buf, embs.ctypes.data_as(C.c_void_p) if dim else None, dim, cap)
        return (buf[:k], embs[:k]) if dim else buf[:k]

    def last_sizes(self):
        s = np.zeros(8, np.int32)
        lib().orc_deepocsort_last_sizes(self._h, s)
        return s


class BoostTrack:
    """Oracle BoostTrack: the constructor arguments that reach the association, reference names and defaults
    (include/motcpp/trackers/boosttrack.hpp:95-124); ECC and ReID off, use_sb = False."""

    def __init__(self, det_thresh=0.6, max_age=60, max_obs=50, min_hits=3, iou_threshold=0.3, min_box_area=10,
                 aspect_ratio_thresh=1.6, lambda_iou=0.5, lambda_mhd=0.25, lambda_shape=0.25, use_dlo_boost=True,
                 dlo_boost_coef=0.65, use_vt=False):
        self._h = lib().orc_boosttrack_create(det_thresh, max_age, max_obs, min_hits, iou_threshold, int(min_box_area),
                                              aspect_ratio_thresh, lambda_iou, lambda_mhd, lambda_shape, int(use_dlo_boost),
                                              dlo_boost_coef, int(use_vt))
        self._out = np.zeros((8192, 8), np.float32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_boosttrack_destroy(self._h)
            self._h = None

    def reset(self):
        lib().orc_boosttrack_reset(self._h)

    def update(self, dets, embs=None):
        dets = _f32(dets).reshape(-1, 6)
        n = lib().orc_boosttrack_update(self._h, dets, dets.shape[0], self._out, self._out.shape[0])
        assert n >= 0
        return self._out[:n].copy()

    def dump(self):
        cap = max(1, lib().orc_boosttrack_count(self._h))
        buf = np.zeros((cap, 80), np.float32)
        return buf[:lib().orc_boosttrack_dump(self._h, buf, cap)]

    def last_sizes(self):
        s = np.zeros(4, np.int32)
        lib().orc_boosttrack_last_sizes(self._h, s)
        return s


def boost_cost(dets4, trk_boxes, mean4, var4, lambda_iou=0.5, lambda_mhd=0.25, lambda_shape=0.25, emb=None):
    """BoostTrack's association cost (boosttrack.cpp:297-358, :571-626): (cost, iou_dist, mh_dist), each (n, m)"""
    d, t, mu, va = _f32(dets4).reshape(-1, 4), _f32(trk_boxes).reshape(-1, 4), _f32(mean4).reshape(-1, 4), _f32(var4).reshape(-1, 4)
    n, m = d.shape[0], t.shape[0]
    iou, mh, cost = (np.zeros((n, m), np.float32) for _ in range(3))
    if n and m:
        lib().orc_boost_iou_dist(d, n, t, m, iou)
        lib().orc_boost_mh_dist(d, n, mu, va, m, mh)
        ep = None
        if emb is not None:
            emb = _f32(emb).reshape(n, m)
            ep = emb.ctypes.data_as(C.c_void_p)
        lib().orc_boost_cost(iou, mh, ep, n, m, lambda_iou, lambda_mhd, lambda_shape, cost)
    return cost, iou, mh
