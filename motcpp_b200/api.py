"""Host-side mirror of the reference's operator interface for the association hot path, on top of
the C ABI (include/motb200.h).  Names, argument meaning and error behaviour follow motcpp:

    motcpp::trackers::ByteTrack(det_thresh, max_age, ..., frame_rate).update(dets, img, embs)
        include/motcpp/trackers/bytetrack.hpp:97-110, src/trackers/bytetrack.cpp:166
    motcpp::utils::{iou_batch, iou_distance, fuse_score, linear_assignment, embedding_distance}
        include/motcpp/utils/iou.hpp:63, include/motcpp/utils/matching.hpp:43-55,99-108
    motcpp::motion::KalmanFilterXYAH / KalmanFilterXYSR, motcpp::KalmanFilterXYWH

Everything computes on the GPU; numpy is only the host container (the reference's Eigen::MatrixXf).
std::invalid_argument maps to ValueError, std::runtime_error to RuntimeError.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import KF_XYAH, KF_XYSR, KF_XYWH, EngineConfig, MotError, check, load


def _raise(e: MotError):
    if e.code == _lib.MOT_ERR_INVALID_ARGUMENT:
        raise ValueError(str(e)) from None
    raise e


# ------------------------------------------------------------------ device / pinned memory helpers
class DeviceArray:
    """A typed device allocation with numpy-flavoured upload/download (no torch needed)."""

    def __init__(self, shape, dtype=np.float32):
        self.shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = C.c_void_p()
        check(load().mot_device_alloc(C.byref(p), max(self.nbytes, 1)))
        self.ptr = p.value

    @classmethod
    def from_host(cls, arr, dtype=None):
        arr = np.ascontiguousarray(arr, dtype=dtype)
        d = cls(arr.shape, arr.dtype)
        d.upload(arr)
        return d

    def upload(self, arr):
        arr = np.ascontiguousarray(arr, dtype=self.dtype)
        assert arr.nbytes == self.nbytes, (arr.shape, self.shape)
        if self.nbytes:
            check(load().mot_copy_h2d(self.ptr, arr.ctypes.data, self.nbytes, None))
            check(load().mot_stream_sync(None))

    def download(self) -> np.ndarray:
        out = np.empty(self.shape, self.dtype)
        if self.nbytes:
            check(load().mot_copy_d2h(out.ctypes.data, self.ptr, self.nbytes, None))
            check(load().mot_stream_sync(None))
        return out

    def free(self):
        if getattr(self, "ptr", None):
            load().mot_device_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """numpy array over page-locked host memory (so engine copies are asynchronous DMA)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape, dtype=np.int64))
    p = C.c_void_p()
    check(load().mot_host_alloc(C.byref(p), max(n * dtype.itemsize, 1)))
    buf = (C.c_char * max(n * dtype.itemsize, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
    return arr


# ------------------------------------------------------------------ utils:: free functions
def iou_batch(bboxes1, bboxes2, _mode=0, _conf=None) -> np.ndarray:
    """utils::iou_batch (iou.hpp:63-100): (N,4),(M,4) xyxy -> (N,M)."""
    a = np.ascontiguousarray(bboxes1, np.float32).reshape(-1, 4)
    b = np.ascontiguousarray(bboxes2, np.float32).reshape(-1, 4)
    n, m = a.shape[0], b.shape[0]
    if n == 0 or m == 0:
        return np.zeros((n, m), np.float32) if _mode == 0 else np.ones((n, m), np.float32)
    _lib.require_gpu()
    ld = (m + 3) // 4 * 4
    da, db, do = DeviceArray.from_host(a), DeviceArray.from_host(b), DeviceArray((n, ld))
    dc = DeviceArray.from_host(np.ascontiguousarray(_conf, np.float32)) if _conf is not None else None
    try:
        check(load().mot_cost_iou(da.ptr, n, db.ptr, m, dc.ptr if dc else None, do.ptr, ld, _mode, None))
    except MotError as e:
        _raise(e)
    return np.ascontiguousarray(do.download()[:, :m])


def iou_distance(atracks, btracks) -> np.ndarray:
    """utils::iou_distance (matching.cpp:62-65): 1 - iou_batch."""
    return iou_batch(atracks, btracks, _mode=1)


def iou_distance_fused(atracks, btracks, det_confs) -> np.ndarray:
    """fuse_score(iou_distance(a, b), det_confs) in one kernel (matching.cpp:130-143)."""
    return iou_batch(atracks, btracks, _mode=2, _conf=det_confs)


_ASSO_KINDS = {"hmiou": 3, "giou": 4, "diou": 5, "centroid": 6, "ciou": 7}


def asso_batch(asso_func: str, bboxes1, bboxes2, frame_width: int = 0, frame_height: int = 0) -> np.ndarray:
    """AssociationFunction (iou.hpp:371-411) for "iou", "hmiou", "giou", "diou", "ciou", "centroid": (N,4),(M,4) xyxy -> (N,M),
    evaluated pair-wise (the reference's hmiou / giou / diou expressions only line up for M == 1)."""
    if asso_func == "iou":
        return iou_batch(bboxes1, bboxes2)
    if asso_func not in _ASSO_KINDS:
        raise ValueError("Invalid association mode: " + asso_func)                 # iou.hpp:407
    a = np.ascontiguousarray(bboxes1, np.float32).reshape(-1, 4)
    b = np.ascontiguousarray(bboxes2, np.float32).reshape(-1, 4)
    n, m = a.shape[0], b.shape[0]
    if n == 0 or m == 0:
        return np.zeros((n, m), np.float32)
    _lib.require_gpu()
    ld = (m + 3) // 4 * 4
    da, db, do = DeviceArray.from_host(a), DeviceArray.from_host(b), DeviceArray((n, ld))
    try:
        check(load().mot_cost_iou_variant(da.ptr, n, db.ptr, m, _ASSO_KINDS[asso_func], int(frame_width), int(frame_height),
                                          do.ptr, ld, None))
    except MotError as e:
        _raise(e)
    return np.ascontiguousarray(do.download()[:, :m])


def ocm_cost(detections, trackers, velocities, previous_obs, vdc_weight: float):
    """OC-SORT association cost (ocsort_assoc::associate, ocsort.cpp:617-700): detections (N,5) [xyxy,score],
    trackers (M,4) predicted boxes, velocities (M,2) (dy,dx), previous_obs (M,5) -> (cost, iou), both (N,M),
    cost = -(iou + valid * angle * vdc_weight * score)."""
    d = np.ascontiguousarray(detections, np.float32).reshape(-1, 5)
    t = np.ascontiguousarray(trackers, np.float32).reshape(-1, 4)
    v = np.ascontiguousarray(velocities, np.float32).reshape(-1, 2)
    p = np.ascontiguousarray(previous_obs, np.float32).reshape(-1, 5)
    n, m = d.shape[0], t.shape[0]
    if v.shape[0] != m or p.shape[0] != m:
        raise ValueError("velocities / previous_obs must have one row per tracker")
    if n == 0 or m == 0:
        return np.zeros((n, m), np.float32), np.zeros((n, m), np.float32)
    _lib.require_gpu()
    ld = (m + 3) // 4 * 4
    dd, dt, dv, dp = (DeviceArray.from_host(x) for x in (d, t, v, p))
    dc, di = DeviceArray((n, ld)), DeviceArray((n, ld))
    try:
        check(load().mot_cost_ocm(dd.ptr, n, dt.ptr, dv.ptr, dp.ptr, m, float(vdc_weight), dc.ptr, di.ptr, ld, None))
    except MotError as e:
        _raise(e)
    return np.ascontiguousarray(dc.download()[:, :m]), np.ascontiguousarray(di.download()[:, :m])


def embedding_distance(track_features, det_features, metric: str = "cosine") -> np.ndarray:
    """utils::embedding_distance (matching.cpp:67-107), cosine metric on tensor cores."""
    if metric != "cosine":
        raise ValueError("Unknown metric: " + metric if metric != "euclidean" else
                         "euclidean embedding distance is not on the accelerated path")
    t = np.ascontiguousarray(track_features, np.float32)
    d = np.ascontiguousarray(det_features, np.float32)
    n, m = t.shape[0], d.shape[0]
    if n == 0 or m == 0:
        return np.zeros((n, m), np.float32)
    _lib.require_gpu()
    ld = (m + 3) // 4 * 4
    dt, dd, do = DeviceArray.from_host(t), DeviceArray.from_host(d), DeviceArray((n, ld))
    try:
        check(load().mot_cost_cosine(dt.ptr, n, dd.ptr, m, t.shape[1], do.ptr, ld, None))
    except MotError as e:
        _raise(e)
    return np.ascontiguousarray(do.download()[:, :m])


# ------------------------------------------------------------------ StrongSORT cost builders (strongsort.cpp)
INFTY_COST = 1e5          # linear_assignment::INFTY_COST (include/motcpp/trackers/strongsort.hpp)


def nn_cosine_distance(samples, targets_of_samples, n_targets: int, features) -> np.ndarray:
    """NearestNeighborDistanceMetric::distance, metric "cosine" (strongsort.cpp:240-334): samples (S,D) are the gallery
    rows of all targets, targets_of_samples[s] the output row of sample s, features (M,D) the raw detection features
    -> (n_targets, M) = min over a target's samples of 1 - cos; targets without samples get 1e5.  Tensor cores."""
    smp = np.ascontiguousarray(samples, np.float32)
    seg = np.ascontiguousarray(targets_of_samples, np.int32).reshape(-1)
    f = np.ascontiguousarray(features, np.float32)
    if smp.ndim != 2 or f.ndim != 2 or (smp.shape[0] and smp.shape[1] != f.shape[1]) or seg.shape[0] != smp.shape[0]:
        raise ValueError("samples (S,D), targets_of_samples (S,), features (M,D) expected")
    if seg.size and (seg.min() < 0 or seg.max() >= n_targets):
        raise ValueError("targets_of_samples out of range")
    m = f.shape[0]
    if n_targets == 0 or m == 0:
        return np.zeros((n_targets, m), np.float32)
    _lib.require_gpu()
    ld = (m + 3) // 4 * 4
    ds, dg, df, do = DeviceArray.from_host(smp), DeviceArray.from_host(seg), DeviceArray.from_host(f), DeviceArray((n_targets, ld))
    try:
        check(load().mot_cost_nn_cosine(ds.ptr, dg.ptr, smp.shape[0], n_targets, df.ptr, m, f.shape[1], do.ptr, ld, None))
    except MotError as e:
        _raise(e)
    return np.ascontiguousarray(do.download()[:, :m])


def gate_cost_matrix(cost_matrix, means, covariances, measurements_xyah, mc_lambda: float, gated_cost: float = INFTY_COST,
                     only_position: bool = False) -> np.ndarray:
    """linear_assignment::gate_cost_matrix (strongsort.cpp:451-492): cost (N,M), the N tracks' XYAH means (N,8) and
    covariances (N,8,8), the M detections as xyah rows -> gated + motion-blended cost (N,M)."""
    c = np.ascontiguousarray(cost_matrix, np.float32)
    mu = np.ascontiguousarray(means, np.float32).reshape(-1, 8)
    cv = np.ascontiguousarray(covariances, np.float32).reshape(-1, 64)
    z = np.ascontiguousarray(measurements_xyah, np.float32).reshape(-1, 4)
    n, m = mu.shape[0], z.shape[0]
    if c.shape != (n, m) or cv.shape[0] != n:
        raise ValueError("cost_matrix must be (tracks x measurements)")
    if n == 0 or m == 0:
        return c.copy()
    _lib.require_gpu()
    dc, dr, dz = DeviceArray.from_host(c), DeviceArray.from_host(np.concatenate([mu, cv], axis=1)), DeviceArray.from_host(z)
    try:
        check(load().mot_cost_gate(dc.ptr, m, dr.ptr, n, dz.ptr, m, float(mc_lambda), float(gated_cost), int(only_position), None))
    except MotError as e:
        _raise(e)
    return dc.download()


def iou_cost_tlwh(track_tlwh, det_tlwh, time_since_update=None) -> np.ndarray:
    """iou_matching::iou_cost (strongsort.cpp:502-585): 1 - IoU of tlwh boxes; rows with time_since_update > 1 = 1e5."""
    a = np.ascontiguousarray(track_tlwh, np.float32).reshape(-1, 4)
    b = np.ascontiguousarray(det_tlwh, np.float32).reshape(-1, 4)
    n, m = a.shape[0], b.shape[0]
    if n == 0 or m == 0:
        return np.zeros((n, m), np.float32)
    _lib.require_gpu()
    ld = (m + 3) // 4 * 4
    da, db, do = DeviceArray.from_host(a), DeviceArray.from_host(b), DeviceArray((n, ld))
    dt = DeviceArray.from_host(np.ascontiguousarray(time_since_update, np.int32)) if time_since_update is not None else None
    try:
        check(load().mot_cost_iou_tlwh(da.ptr, dt.ptr if dt else None, n, db.ptr, m, do.ptr, ld, None))
    except MotError as e:
        _raise(e)
    return np.ascontiguousarray(do.download()[:, :m])


def aw_max_metric(emb_cost, w_association_emb: float = 0.5, bottom: float = 0.5) -> np.ndarray:
    """deepocsort_assoc::compute_aw_max_metric (deepocsort.cpp:294-345): DeepOC-SORT's adaptive embedding weights."""
    e = np.ascontiguousarray(emb_cost, np.float32)
    if e.ndim != 2:
        raise ValueError("emb_cost must be 2-D")
    n, m = e.shape
    if n == 0 or m == 0:
        return e.copy()
    _lib.require_gpu()
    de, do = DeviceArray.from_host(e), DeviceArray((n, m))
    try:
        check(load().mot_cost_aw_max_metric(de.ptr, n, m, m, float(w_association_emb), float(bottom), do.ptr, m, None))
        check(load().mot_stream_sync(None))
    except MotError as ex:
        _raise(ex)
    return do.download()


@dataclass
class LinearAssignmentResult:
    """utils::LinearAssignmentResult (matching.hpp:32-36)."""
    matches: List[Tuple[int, int]] = field(default_factory=list)
    unmatched_a: List[int] = field(default_factory=list)
    unmatched_b: List[int] = field(default_factory=list)


def linear_assignment(cost_matrix, thresh: float) -> LinearAssignmentResult:
    """utils::linear_assignment (matching.cpp:14-60): matches in ascending row order."""
    cost = np.ascontiguousarray(cost_matrix, np.float32)
    if cost.ndim != 2:
        raise ValueError("cost_matrix must be 2-D")
    n, m = cost.shape
    res = LinearAssignmentResult()
    if n == 0 or m == 0:
        res.unmatched_a = list(range(n))
        res.unmatched_b = list(range(m))
        return res
    _lib.require_gpu()
    r2c = np.empty(n, np.int32)
    c2r = np.empty(m, np.int32)
    try:
        check(load().mot_lap_host(cost.ctypes.data, n, m, m, float(thresh), r2c.ctypes.data, c2r.ctypes.data))
    except MotError as e:
        _raise(e)
    res.matches = [(int(i), int(j)) for i, j in enumerate(r2c) if j >= 0]
    res.unmatched_a = [int(i) for i in np.nonzero(r2c < 0)[0]]
    res.unmatched_b = [int(j) for j in np.nonzero(c2r < 0)[0]]
    return res


def linear_assignment_arrays(cost_matrix, thresh: float):
    """Same solve, raw row2col / col2row arrays (-1 = unmatched)."""
    cost = np.ascontiguousarray(cost_matrix, np.float32)
    n, m = cost.shape
    r2c = np.full(n, -1, np.int32)
    c2r = np.full(m, -1, np.int32)
    if n and m:
        _lib.require_gpu()
        check(load().mot_lap_host(cost.ctypes.data, n, m, m, float(thresh), r2c.ctypes.data, c2r.ctypes.data))
    return r2c, c2r


def linear_assignment_reference_order(cost_matrices, thresh: float):
    """The reference's dense LAPJV itself on the GPU (mot_lap_jv_batch_device): cost (P, n, m) or (n, m) ->
    (row2col, col2row), ties resolved exactly as utils::linear_assignment resolves them (one warp per problem while
    n + m <= 384, one CTA per problem above that)."""
    cost = np.ascontiguousarray(cost_matrices, np.float32)
    single = cost.ndim == 2
    if single:
        cost = cost[None]
    P, n, m = cost.shape
    r2c = np.full((P, n), -1, np.int32)
    c2r = np.full((P, m), -1, np.int32)
    if P and n and m:
        _lib.require_gpu()
        dc, dr, dq = DeviceArray.from_host(cost), DeviceArray((P, n), np.int32), DeviceArray((P, m), np.int32)
        try:
            check(load().mot_lap_jv_batch_device(dc.ptr, n * m, P, n, m, m, float(thresh), dr.ptr, dq.ptr, None))
        except MotError as e:
            _raise(e)
        r2c, c2r = dr.download(), dq.download()
    return (r2c[0], c2r[0]) if single else (r2c, c2r)


# ------------------------------------------------------------------ Kalman filters (batched)
class _BatchedKF:
    KIND = KF_XYAH
    NX = 8

    @property
    def rec(self) -> int:
        return self.NX + self.NX * self.NX

    def _pack(self, mean, cov) -> np.ndarray:
        mean = np.ascontiguousarray(mean, np.float32).reshape(-1, self.NX)
        cov = np.ascontiguousarray(cov, np.float32).reshape(-1, self.NX * self.NX)
        return np.concatenate([mean, cov], axis=1)

    def _unpack(self, recs, single):
        mean = recs[:, :self.NX]
        cov = recs[:, self.NX:].reshape(-1, self.NX, self.NX)
        return (mean[0], cov[0]) if single else (mean, cov)

    def initiate(self, measurement):
        z = np.ascontiguousarray(measurement, np.float32)
        single = z.ndim == 1
        z = z.reshape(-1, 4)
        _lib.require_gpu()
        dz, dr = DeviceArray.from_host(z), DeviceArray((z.shape[0], self.rec))
        check(load().mot_kf_initiate(self.KIND, dr.ptr, dz.ptr, z.shape[0], None))
        return self._unpack(dr.download(), single)

    def predict(self, mean, covariance, zero_vh=None, q_xy_scaling=1.0, q_s_scaling=1.0):
        single = np.ndim(mean) == 1
        recs = self._pack(mean, covariance)
        _lib.require_gpu()
        dr = DeviceArray.from_host(recs)
        df = DeviceArray.from_host(np.ascontiguousarray(zero_vh, np.uint8)) if zero_vh is not None else None
        check(load().mot_kf_predict(self.KIND, dr.ptr, df.ptr if df else None, recs.shape[0], float(q_xy_scaling),
                                    float(q_s_scaling), None))
        return self._unpack(dr.download(), single)

    def update(self, mean, covariance, measurement, confidence=None, return_fail=False):
        single = np.ndim(mean) == 1
        recs = self._pack(mean, covariance)
        z = np.ascontiguousarray(measurement, np.float32).reshape(-1, 4)
        n = recs.shape[0]
        _lib.require_gpu()
        dr, dz = DeviceArray.from_host(recs), DeviceArray.from_host(z)
        dc = None
        if confidence is not None:
            dc = DeviceArray.from_host(np.broadcast_to(np.asarray(confidence, np.float32), (n,)).copy())
        dfail = DeviceArray((n,), np.uint8)
        check(load().mot_kf_update(self.KIND, dr.ptr, dz.ptr, dc.ptr if dc else None, n, dfail.ptr, None))
        out = self._unpack(dr.download(), single)
        return (*out, dfail.download()) if return_fail else out


class KalmanFilterXYAH(_BatchedKF):
    """motion::KalmanFilterXYAH (kalman_filter.cpp, xyah_kf.cpp), batched over tracks."""
    KIND = KF_XYAH

    def gating_distance(self, mean, covariance, measurements, only_position=False, metric="maha"):
        if metric not in ("maha", "gaussian"):
            raise ValueError("Invalid metric: " + metric)          # kalman_filter.cpp:174
        recs = self._pack(mean, covariance)
        meas = np.ascontiguousarray(measurements, np.float32).reshape(-1, 4)
        if recs.shape[0] == 0 or meas.shape[0] == 0:
            return np.zeros((recs.shape[0], meas.shape[0]), np.float32)
        _lib.require_gpu()
        dr, dm, do = DeviceArray.from_host(recs), DeviceArray.from_host(meas), DeviceArray((recs.shape[0], meas.shape[0]))
        check(load().mot_kf_gating(self.KIND, dr.ptr, recs.shape[0], dm.ptr, meas.shape[0], int(only_position),
                                   0 if metric == "maha" else 1, do.ptr, None))
        out = do.download()
        return out[0] if np.ndim(mean) == 1 else out


class KalmanFilterXYWH(KalmanFilterXYAH):
    """motcpp::KalmanFilterXYWH (xywh_kf.hpp), batched over tracks."""
    KIND = KF_XYWH

    def gating_distance(self, mean, covariance, measurements, only_position=False, metric="maha"):
        return super().gating_distance(mean, covariance, measurements, only_position, "maha")


class KalmanFilterXYSR(_BatchedKF):
    """motion::KalmanFilterXYSR (xysr_kf.cpp), batched over tracks; state 7, covariance 7x7."""
    KIND = KF_XYSR
    NX = 7

    def apply_affine_correction(self, mean, covariance, m, t):
        """KalmanFilterXYSR::apply_affine_correction (xysr_kf.cpp:114-141) for a batch of states: m (2,2), t (2,)."""
        single = np.ndim(mean) == 1
        recs = self._pack(mean, covariance)
        m2 = np.ascontiguousarray(m, np.float32).reshape(4)
        t2 = np.ascontiguousarray(t, np.float32).reshape(2)
        _lib.require_gpu()
        dr = DeviceArray.from_host(recs)
        check(load().mot_kf_xysr_affine(dr.ptr, recs.shape[0], m2.ctypes.data, t2.ctypes.data, None))
        return self._unpack(dr.download(), single)


# ------------------------------------------------------------------ tracker engine
_ERR_BITS = {1: "track capacity exceeded", 2: "too many detections", 4: "output rows truncated",
             8: "Kalman update left the Cholesky path", 16: "StrongSORT candidate table full"}


class Engine:
    """S independent trackers resident on one GPU (mot_engine_*)."""

    def __init__(self, kind: int = _lib.TRACKER_BYTETRACK, n_streams: int = 1, track_capacity: int = 0,
                 max_dets: int = 0, device: int = 0, n_chunks: int = 0, **params):
        _lib.require_gpu()
        cfg = EngineConfig()
        check(load().mot_engine_default_config(kind, C.byref(cfg)))
        cfg.n_streams, cfg.track_capacity, cfg.max_dets = n_streams, track_capacity, max_dets
        cfg.device, cfg.n_chunks = device, n_chunks
        valid = {f[0] for f in EngineConfig._fields_}
        for k, v in params.items():
            if k not in valid:
                raise ValueError(f"unknown tracker parameter {k!r}")
            setattr(cfg, k, v)
        self.cfg = cfg
        h = C.c_void_p()
        try:
            check(load().mot_engine_create(C.byref(cfg), C.byref(h)))
        except MotError as e:
            _raise(e)
        self._h = h.value
        self.n_streams = n_streams

    def close(self):
        if getattr(self, "_h", None):
            load().mot_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        check(load().mot_engine_reset(self._h))

    def update(self, dets: np.ndarray, n_dets: np.ndarray, out: Optional[np.ndarray] = None,
               n_out: Optional[np.ndarray] = None, ld_out: int = 0, embs: Optional[np.ndarray] = None):
        """dets (T,S,ld,6) or (S,ld,6) float32, n_dets (T,S) or (S,) int32 -> out (T,S,ld_out,8), n_out (T,S).
        BoT-SORT engines also take embs (T,S,ld,emb_dim): the ReID feature of every detection row."""
        dets = np.asarray(dets)
        squeeze = dets.ndim == 3
        if squeeze:
            dets = dets[None]
            if embs is not None:
                embs = np.asarray(embs)[None]
        if dets.dtype != np.float32 or not dets.flags.c_contiguous:
            dets = np.ascontiguousarray(dets, np.float32)
        T, S, ld, six = dets.shape
        if six != 6:
            raise ValueError("Detections must have 6 (AABB) or 7 (OBB) columns")     # src/tracker.cpp:110
        if S != self.n_streams:
            raise ValueError(f"expected {self.n_streams} streams, got {S}")
        n_dets = np.ascontiguousarray(n_dets, np.int32).reshape(T, S)
        if n_dets.size and (int(n_dets.max()) > ld or int(n_dets.min()) < 0):
            raise ValueError("n_dets must lie in [0, dets.shape[2]]")
        if ld_out <= 0:
            ld_out = self.cfg.track_capacity or 1536
        if out is None:
            out = np.empty((T, S, ld_out, 8), np.float32)
        if n_out is None:
            n_out = np.empty((T, S), np.int32)
        ep = None
        if embs is not None:
            embs = np.ascontiguousarray(embs, np.float32)
            if embs.shape[:3] != (T, S, ld) or embs.shape[3] != self.cfg.emb_dim:
                raise ValueError("Detections and embeddings must have same number of rows")   # src/tracker.cpp:118-121
            ep = embs.ctypes.data
        try:
            check(load().mot_engine_update_host_embs(self._h, T, dets.ctypes.data, n_dets.ctypes.data, ld, ep,
                                                     out.ctypes.data, n_out.ctypes.data, out.shape[2]))
        except MotError as e:
            _raise(e)
        return (out[0], n_out[0]) if squeeze else (out, n_out)

    def update_packed(self, dets: np.ndarray, n_dets: np.ndarray, max_rows: int = 0, out_rows: Optional[np.ndarray] = None,
                      pinned: bool = False):
        """mot_engine_update_host_packed: dets (T,S,ld,6), n_dets (T,S) -> (rows (R,8), offsets (T*S+1,), n_out (T,S)); the rows
        of frame t, stream s are rows[offsets[t*S+s] : offsets[t*S+s+1]] - exactly what update() would have returned."""
        dets = np.ascontiguousarray(dets, np.float32)
        T, S, ld, six = dets.shape
        if six != 6:
            raise ValueError("Detections must have 6 (AABB) or 7 (OBB) columns")
        if S != self.n_streams:
            raise ValueError(f"expected {self.n_streams} streams, got {S}")
        n_dets = np.ascontiguousarray(n_dets, np.int32).reshape(T, S)
        if n_dets.size and (int(n_dets.max()) > ld or int(n_dets.min()) < 0):
            raise ValueError("n_dets must lie in [0, dets.shape[2]]")
        max_rows = max_rows or (self.cfg.track_capacity or 1536)
        if out_rows is None:
            out_rows = (pinned_empty if pinned else np.empty)((T * S * max_rows, 8), np.float32)
        offsets = (pinned_empty if pinned else np.empty)((T * S + 1,), np.int64)
        n_out = np.empty((T, S), np.int32)
        try:
            check(load().mot_engine_update_host_packed(self._h, T, dets.ctypes.data, n_dets.ctypes.data, ld, max_rows,
                                                       out_rows.ctypes.data, out_rows.shape[0], offsets.ctypes.data,
                                                       n_out.ctypes.data))
        except MotError as e:
            _raise(e)
        return out_rows[:offsets[-1]], offsets, n_out

    def update_device(self, n_frames, d_dets, d_n_dets, ld_dets, d_out, d_n_out, ld_out, stream=None):
        """Raw device-pointer path (ints / c_void_p), asynchronous on `stream`."""
        check(load().mot_engine_update_device(self._h, n_frames, d_dets, d_n_dets, ld_dets, d_out, d_n_out, ld_out,
                                              stream))

    def check(self) -> np.ndarray:
        flags = np.zeros(self.n_streams, np.int32)
        rc = load().mot_engine_check(self._h, flags.ctypes.data)
        if rc != _lib.MOT_OK:
            bits = int(np.bitwise_or.reduce(flags))
            what = ", ".join(v for k, v in _ERR_BITS.items() if bits & k)
            raise RuntimeError(f"engine error flags 0x{bits:x}: {what}")
        return flags

    def header(self, stream: int = 0) -> np.ndarray:
        h = np.zeros(16, np.int32)
        check(load().mot_engine_stream_header(self._h, stream, h.ctypes.data))
        return h

    def dump(self, stream: int, which: int = 0) -> np.ndarray:
        """ByteTrack: rows of 78 floats for list `which` (0 active, 1 lost).  OC-SORT: rows of
        [id, age, hits, hit_streak, time_since_update, conf, cls, det_ind, last_obs 5, velocity 2, x 7, P 49, pad]."""
        hdr = self.header(stream)
        n = int(hdr[0] if which == 0 else hdr[1])
        if self.cfg.kind in (_lib.TRACKER_OCSORT, _lib.TRACKER_DEEPOCSORT) and which != 0:
            n = 0
        buf = np.zeros((max(n, 1), 78), np.float32)
        k = C.c_int()
        check(load().mot_engine_dump_list(self._h, stream, which, buf.ctypes.data, max(n, 1), C.byref(k)))
        return buf[:k.value]

    def dump_boost(self, stream: int) -> np.ndarray:
        """BoostTrack engines: rows of [id, age, hit_streak, time_since_update, conf, cls, det_ind, 0, x 8, P 8x8]"""
        n = max(int(self.header(stream)[0]), 1)
        buf = np.zeros((n, 80), np.float32)
        k = C.c_int()
        check(load().mot_engine_dump_boost(self._h, stream, buf.ctypes.data, n, C.byref(k)))
        return buf[:k.value]

    def dump_deep_embs(self, stream: int) -> np.ndarray:
        """DeepOC-SORT engines: the tracks' unit-length embeddings, in the row order of dump(stream)."""
        n = max(int(self.header(stream)[0]), 1)
        dim = int(self.cfg.emb_dim)
        embs = np.zeros((n, max(dim, 1)), np.float32)
        k = C.c_int()
        check(load().mot_engine_dump_deep_embs(self._h, stream, embs.ctypes.data, n, C.byref(k)))
        return embs[:k.value, :dim]

    def dump_bot(self, stream: int, which: int, with_feats: bool = False):
        """BoT-SORT engines: rows of [id, state, is_activated, frame_id, start_frame, tracklet_len, conf, cls, det_ind,
        has_feat, mean 8, cov 64] for list `which` (0 active, 1 lost) and optionally the smoothed features."""
        hdr = self.header(stream)
        n = max(int(hdr[0] if which == 0 else hdr[1]), 1)
        buf = np.zeros((n, 82), np.float32)
        dim = int(self.cfg.emb_dim)
        feats = np.zeros((n, max(dim, 1)), np.float32)
        k = C.c_int()
        check(load().mot_engine_dump_bot(self._h, stream, which, buf.ctypes.data,
                                         feats.ctypes.data if (with_feats and dim) else None, n, C.byref(k)))
        return (buf[:k.value], feats[:k.value]) if with_feats else buf[:k.value]

    def dump_strong(self, stream: int, with_feats: bool = False):
        """StrongSORT engines: the track list (reference vector order) as rows of [id, state, hits, 0, time_since_update,
        conf, cls, det_ind, has_feat, n_gallery_samples, mean 8, cov 64] and optionally the smoothed features."""
        n = max(int(self.header(stream)[0]), 1)
        buf = np.zeros((n, 82), np.float32)
        dim = int(self.cfg.emb_dim)
        feats = np.zeros((n, max(dim, 1)), np.float32)
        k = C.c_int()
        check(load().mot_engine_dump_strong(self._h, stream, buf.ctypes.data,
                                            feats.ctypes.data if (with_feats and dim) else None, n, C.byref(k)))
        return (buf[:k.value], feats[:k.value]) if with_feats else buf[:k.value]

    def info(self):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(load().mot_engine_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"threads_per_cta": a.value, "smem_bytes": b.value, "ctas": c.value, "state_bytes_per_stream": d.value}


class _Image:
    """What the hot path needs from cv::Mat: empty(), rows, cols (src/tracker.cpp:114,166-171)."""

    def __init__(self, img):
        if img is None:
            self.rows = self.cols = 0
        elif hasattr(img, "shape"):
            self.rows, self.cols = (int(img.shape[0]), int(img.shape[1])) if len(img.shape) >= 2 else (0, 0)
        else:
            self.rows, self.cols = int(img[0]), int(img[1])

    def empty(self) -> bool:
        return self.rows == 0 or self.cols == 0


class ByteTrack:
    """motcpp::trackers::ByteTrack with the reference's positional constructor
    (include/motcpp/trackers/bytetrack.hpp:97-110).  One stream; for many streams use Engine."""

    def __init__(self, det_thresh=0.3, max_age=30, max_obs=50, min_hits=3, iou_threshold=0.3, per_class=False,
                 nr_classes=80, asso_func="iou", is_obb=False, min_conf=0.1, track_thresh=0.45, match_thresh=0.8,
                 track_buffer=25, frame_rate=30, track_capacity=0, max_dets=0, device=0):
        if asso_func != "iou":
            raise ValueError("Invalid association mode: " + str(asso_func) + " (only \"iou\" is accelerated)")
        if per_class or is_obb:
            raise ValueError("per_class / OBB tracking are outside the accelerated hot path")
        self._engine = Engine(_lib.TRACKER_BYTETRACK, 1, track_capacity, max_dets, device, det_thresh=det_thresh,
                              max_age=max_age, max_obs=max_obs, min_hits=min_hits, iou_threshold=iou_threshold,
                              min_conf=min_conf, track_thresh=track_thresh, match_thresh=match_thresh,
                              track_buffer=track_buffer, frame_rate=frame_rate)
        self._max_dets = max_dets or 512
        self._cap = track_capacity or 1536
        self._dets = np.zeros((1, self._max_dets, 6), np.float32)
        self._out = np.empty((1, 1, self._cap, 8), np.float32)
        self._n_out = np.empty((1, 1), np.int32)

    def reset(self):
        self._engine.reset()

    def update(self, dets, img, embs=None) -> np.ndarray:
        dets = np.asarray(dets, np.float32)
        if dets.ndim != 2:
            dets = dets.reshape(0, 6) if dets.size == 0 else dets
        # BaseTracker::check_inputs (src/tracker.cpp:108-125)
        if dets.shape[0] > 0 and dets.shape[1] not in (6, 7):
            raise ValueError("Detections must have 6 (AABB) or 7 (OBB) columns")
        if _Image(img).empty():
            raise ValueError("Image cannot be empty")
        if embs is not None and np.shape(embs)[0] > 0 and np.shape(embs)[0] != dets.shape[0]:
            raise ValueError("Detections and embeddings must have same number of rows")
        if dets.shape[0] > 0 and dets.shape[1] == 7:
            raise ValueError("OBB detections are outside the accelerated hot path")
        n = dets.shape[0]
        if n > self._max_dets:
            raise ValueError(f"{n} detections exceed max_dets={self._max_dets}")
        self._dets[0, :n] = dets[:, :6] if n else 0
        self._engine.update(self._dets[None], np.array([[n]], np.int32), self._out, self._n_out)
        self._engine.check()
        return self._out[0, 0, :int(self._n_out[0, 0])].copy()


class Sort:
    """motcpp::trackers::Sort with the reference's positional constructor
    (include/motcpp/trackers/sort.hpp:69-77).  Like the reference it performs no input validation
    (Sort::update never calls check_inputs, src/trackers/sort.cpp:102-108) and ignores the image."""

    def __init__(self, det_thresh=0.3, max_age=1, max_obs=50, min_hits=3, iou_threshold=0.3, per_class=False,
                 nr_classes=80, asso_func="iou", is_obb=False, track_capacity=256, max_dets=64, device=0):
        if asso_func != "iou" or per_class or is_obb:
            raise ValueError("only asso_func=\"iou\", per_class=False, is_obb=False are on the accelerated path")
        self._engine = Engine(_lib.TRACKER_SORT, 1, track_capacity, max_dets, device, det_thresh=det_thresh,
                              max_age=max_age, max_obs=max_obs, min_hits=min_hits, iou_threshold=iou_threshold)
        self._max_dets = max(max_dets, 1)
        self._dets = np.zeros((1, 1, self._max_dets, 6), np.float32)
        self._out = np.empty((1, 1, max(track_capacity, 256), 8), np.float32)
        self._n_out = np.empty((1, 1), np.int32)

    def reset(self):
        self._engine.reset()

    def update(self, dets, img=None, embs=None) -> np.ndarray:
        dets = np.asarray(dets, np.float32).reshape(-1, 6) if np.size(dets) else np.zeros((0, 6), np.float32)
        n = dets.shape[0]
        if n > self._max_dets:
            raise ValueError(f"{n} detections exceed max_dets={self._max_dets}")
        self._dets[0, 0, :n] = dets
        self._engine.update(self._dets, np.array([[n]], np.int32), self._out, self._n_out)
        self._engine.check()
        return self._out[0, 0, :int(self._n_out[0, 0])].copy()


class OCSort:
    """motcpp::trackers::OCSort with the reference's positional constructor
    (include/motcpp/trackers/ocsort.hpp:88-102).  One stream; for many streams use Engine."""

    def __init__(self, det_thresh=0.2, max_age=30, max_obs=50, min_hits=3, iou_threshold=0.3, per_class=False,
                 nr_classes=80, asso_func="iou", is_obb=False, min_conf=0.1, delta_t=3, inertia=0.2, use_byte=False,
                 Q_xy_scaling=0.01, Q_s_scaling=0.0001, track_capacity=1536, max_dets=512, device=0):
        # AssociationFunction (iou.hpp:371-411): "iou" and "centroid" are wired in; the reference's hmiou / giou / diou / ciou
        # expressions are only defined when the second box set has one row (trap 11) and are refused
        if asso_func not in ("iou", "centroid"):
            raise ValueError("Invalid association mode: " + str(asso_func) + " (\"iou\" and \"centroid\" are accelerated)")
        if per_class or is_obb:
            raise ValueError("per_class / OBB tracking are outside the accelerated hot path")
        self._make = lambda w, h: Engine(_lib.TRACKER_OCSORT, 1, track_capacity, max_dets, device, det_thresh=det_thresh,
                                         max_age=max_age, max_obs=max_obs, min_hits=min_hits, iou_threshold=iou_threshold,
                                         min_conf=min_conf, delta_t=delta_t, inertia=inertia, use_byte=int(bool(use_byte)),
                                         q_xy_scaling=Q_xy_scaling, q_s_scaling=Q_s_scaling,
                                         asso_func=0 if asso_func == "iou" else 6, frame_width=w, frame_height=h)
        self._centroid = asso_func == "centroid"
        self._frame = None
        self._engine = None
        if not self._centroid:
            self._build(0, 0)

    def _build(self, w, h):
        self._engine = self._make(w, h)
        self._max_dets = self._engine.cfg.max_dets
        self._cap = self._engine.cfg.track_capacity
        self._dets = np.zeros((1, 1, self._max_dets, 6), np.float32)
        self._out = np.empty((1, 1, self._cap, 8), np.float32)
        self._n_out = np.empty((1, 1), np.int32)

    def reset(self):
        if self._engine is not None:
            self._engine.reset()

    def update(self, dets, img, embs=None) -> np.ndarray:
        dets = np.asarray(dets, np.float32)
        if dets.ndim != 2:
            dets = dets.reshape(0, 6) if dets.size == 0 else dets
        # BaseTracker::check_inputs (src/tracker.cpp:108-125)
        if dets.shape[0] > 0 and dets.shape[1] not in (6, 7):
            raise ValueError("Detections must have 6 (AABB) or 7 (OBB) columns")
        im = _Image(img)
        if im.empty():
            raise ValueError("Image cannot be empty")
        if embs is not None and np.shape(embs)[0] > 0 and np.shape(embs)[0] != dets.shape[0]:
            raise ValueError("Detections and embeddings must have same number of rows")
        if dets.shape[0] > 0 and dets.shape[1] == 7:
            raise ValueError("OBB detections are outside the accelerated hot path")
        if self._centroid:
            # the reference normalises centre distances by the diagonal of the frame it is handed (ocsort.cpp:413); the engine
            # fixes it with the first frame
            if self._engine is None:
                self._frame = (im.cols, im.rows)
                self._build(im.cols, im.rows)
            elif (im.cols, im.rows) != self._frame:
                raise ValueError(f"frame size changed from {self._frame} to {(im.cols, im.rows)}: asso_func=\"centroid\" fixes it at the first update")
        n = dets.shape[0]
        if n > self._max_dets:
            raise ValueError(f"{n} detections exceed max_dets={self._max_dets}")
        self._dets[0, 0, :n] = dets[:, :6] if n else 0
        self._engine.update(self._dets, np.array([[n]], np.int32), self._out, self._n_out)
        self._engine.check()
        return self._out[0, 0, :int(self._n_out[0, 0])].copy()


class DeepOCSort:
    """motcpp::trackers::DeepOCSort with the reference's positional constructor
    (include/motcpp/trackers/deepocsort.hpp:93-117).  The ReID network and camera-motion compensation are image
    processing outside the accelerated path: embeddings are passed to update() (the reference's `embs` argument,
    src/trackers/deepocsort.cpp:629-633) and cmc_off must be True.  With embedding_off=True no embeddings are needed
    (the reference substitutes a column of ones, :624-626, which never reaches the cost).  One stream; for many streams
    use Engine(TRACKER_DEEPOCSORT, ...).  The reference lists unmatched detections and tracks twice (:476-481, :485-501):
    size track_capacity for 2 x the tracks that can be unmatched in one frame."""

    def __init__(self, reid_weights="", use_half=False, use_gpu=False, det_thresh=0.3, max_age=30, max_obs=50,
                 min_hits=3, iou_threshold=0.3, per_class=False, nr_classes=80, asso_func="iou", is_obb=False,
                 delta_t=3, inertia=0.2, w_association_emb=0.5, alpha_fixed_emb=0.95, aw_param=0.5, embedding_off=False,
                 cmc_off=True, aw_off=False, Q_xy_scaling=0.01, Q_s_scaling=0.0001, emb_dim=0, track_capacity=1536,
                 max_dets=512, device=0):
        if not cmc_off:
            raise ValueError("camera-motion compensation is outside the accelerated hot path: pass cmc_off=True")
        if reid_weights and not embedding_off:
            raise ValueError("ReID inference is outside the accelerated hot path: pass embeddings to update()")
        if asso_func != "iou" or per_class or is_obb:
            raise ValueError("only asso_func=\"iou\", per_class=False, is_obb=False are on the accelerated path")
        self._off = bool(embedding_off)
        self._make = lambda dim: Engine(_lib.TRACKER_DEEPOCSORT, 1, track_capacity, max_dets, device, det_thresh=det_thresh,
                                        max_age=max_age, max_obs=max_obs, min_hits=min_hits, iou_threshold=iou_threshold,
                                        delta_t=delta_t, inertia=inertia, w_association_emb=w_association_emb,
                                        alpha_fixed_emb=alpha_fixed_emb, aw_param=aw_param, embedding_off=int(self._off),
                                        aw_off=int(bool(aw_off)), q_xy_scaling=Q_xy_scaling, q_s_scaling=Q_s_scaling,
                                        emb_dim=dim)
        self._engine = None
        self._dim = 0
        self._pending_empty = 0          # empty frames seen before the embedding dimension is known
        if self._off or emb_dim > 0:
            self._build(0 if self._off else emb_dim)

    def _build(self, emb_dim):
        self._engine = self._make(emb_dim)
        self._dim = emb_dim
        self._max_dets = self._engine.cfg.max_dets
        self._cap = self._engine.cfg.track_capacity
        self._dets = np.zeros((1, 1, self._max_dets, 6), np.float32)
        self._embs = np.zeros((1, 1, self._max_dets, emb_dim), np.float32) if emb_dim else None
        self._out = np.empty((1, 1, self._cap, 8), np.float32)
        self._n_out = np.empty((1, 1), np.int32)

    def reset(self):
        self._pending_empty = 0
        if self._engine is not None:
            self._engine.reset()

    def _step(self, n):
        self._engine.update(self._dets, np.array([[n]], np.int32), self._out, self._n_out, embs=self._embs)
        self._engine.check()
        return self._out[0, 0, :int(self._n_out[0, 0])].copy()

    def update(self, dets, img, embs=None) -> np.ndarray:
        dets = np.asarray(dets, np.float32)
        if dets.ndim != 2:
            dets = dets.reshape(0, 6) if dets.size == 0 else dets
        # BaseTracker::check_inputs(dets, img, embs) (src/tracker.cpp:108-125)
        if dets.shape[0] > 0 and dets.shape[1] not in (6, 7):
            raise ValueError("Detections must have 6 (AABB) or 7 (OBB) columns")
        if _Image(img).empty():
            raise ValueError("Image cannot be empty")
        n = dets.shape[0]
        have = embs is not None and np.size(embs) > 0
        if have and np.shape(embs)[0] != n:
            raise ValueError("Detections and embeddings must have same number of rows")
        if n > 0 and dets.shape[1] == 7:
            raise ValueError("OBB detections are outside the accelerated hot path")
        if not self._off:
            if n > 0 and not have:
                raise ValueError("ReID inference is outside the accelerated hot path: pass embeddings to update()")
            if have:
                embs = np.asarray(embs, np.float32)
                if embs.ndim != 2:
                    raise ValueError("Detections and embeddings must have same number of rows")
                if self._engine is None:
                    self._build(int(embs.shape[1]))
                    for _ in range(self._pending_empty):       # the empty frames that came first only advanced the clock
                        self._step(0)
                    self._pending_empty = 0
                elif embs.shape[1] != self._dim:
                    raise ValueError(f"embeddings have {embs.shape[1]} columns but the tracker was built for emb_dim={self._dim}")
            elif self._engine is None:
                self._pending_empty += 1                       # no tracks yet: frame_count_++ and an empty result (:655-668)
                return np.zeros((0, 8), np.float32)
        if n > self._max_dets:
            raise ValueError(f"{n} detections exceed max_dets={self._max_dets}")
        self._dets[0, 0, :n] = dets[:, :6] if n else 0
        if not self._off and n:
            self._embs[0, 0, :n] = embs
        return self._step(n)


class BoostTrack:
    """motcpp::trackers::BoostTrackTracker with the reference's positional constructor
    (include/motcpp/trackers/boosttrack.hpp:95-124).  Camera-motion compensation and ReID are image processing outside
    the accelerated path: use_ecc must be False (or cmc_method not "ecc"), with_reid False; use_sb (a powf) is not built.
    One stream; for many streams use Engine(TRACKER_BOOSTTRACK, ...)."""

    def __init__(self, reid_weights="", use_half=False, use_gpu=False, det_thresh=0.6, max_age=60, max_obs=50, min_hits=3,
                 iou_threshold=0.3, per_class=False, nr_classes=80, asso_func="iou", is_obb=False, use_ecc=False,
                 min_box_area=10, aspect_ratio_thresh=1.6, cmc_method="ecc", lambda_iou=0.5, lambda_mhd=0.25, lambda_shape=0.25,
                 use_dlo_boost=True, use_duo_boost=True, dlo_boost_coef=0.65, s_sim_corr=False, use_rich_s=False, use_sb=False,
                 use_vt=False, with_reid=False, track_capacity=1536, max_dets=512, device=0):
        if use_ecc and cmc_method == "ecc":
            raise ValueError("camera-motion compensation is outside the accelerated hot path: pass use_ecc=False")
        if with_reid:
            raise ValueError("BoostTrack with ReID multiplies out every detection x track pair; only with_reid=False is built")
        if use_sb:
            raise ValueError("use_sb (std::pow in the confidence boost) is not built")
        if per_class or is_obb:
            raise ValueError("per_class / OBB tracking are outside the accelerated hot path")
        self._engine = Engine(_lib.TRACKER_BOOSTTRACK, 1, track_capacity, max_dets, device, det_thresh=det_thresh, max_age=max_age,
                              max_obs=max_obs, min_hits=min_hits, iou_threshold=iou_threshold, min_box_area=int(min_box_area),
                              aspect_ratio_thresh=aspect_ratio_thresh, lambda_iou=lambda_iou, lambda_mhd=lambda_mhd,
                              lambda_shape=lambda_shape, use_dlo_boost=int(bool(use_dlo_boost)), dlo_boost_coef=dlo_boost_coef,
                              use_vt=int(bool(use_vt)))
        self._max_dets = self._engine.cfg.max_dets
        self._cap = self._engine.cfg.track_capacity
        self._dets = np.zeros((1, 1, self._max_dets, 6), np.float32)
        self._out = np.empty((1, 1, self._cap, 8), np.float32)
        self._n_out = np.empty((1, 1), np.int32)

    def reset(self):
        self._engine.reset()

    def update(self, dets, img, embs=None) -> np.ndarray:
        dets = np.asarray(dets, np.float32)
        if dets.ndim != 2:
            dets = dets.reshape(0, 6) if dets.size == 0 else dets
        # BaseTracker::check_inputs(dets, img) (src/tracker.cpp:108-125)
        if dets.shape[0] > 0 and dets.shape[1] not in (6, 7):
            raise ValueError("Detections must have 6 (AABB) or 7 (OBB) columns")
        if _Image(img).empty():
            raise ValueError("Image cannot be empty")
        if dets.shape[0] > 0 and dets.shape[1] == 7:
            raise ValueError("OBB detections are outside the accelerated hot path")
        n = dets.shape[0]
        if n > self._max_dets:
            raise ValueError(f"{n} detections exceed max_dets={self._max_dets}")
        self._dets[0, 0, :n] = dets[:, :6] if n else 0
        self._engine.update(self._dets, np.array([[n]], np.int32), self._out, self._n_out)
        self._engine.check()
        return self._out[0, 0, :int(self._n_out[0, 0])].copy()


class BotSort:
    """motcpp::trackers::BotSort with the reference's positional constructor
    (include/motcpp/trackers/botsort.hpp:108-134).  Camera-motion compensation and the ReID network are image
    processing outside the accelerated path: cmc_method must be "none" and embeddings are passed to update()
    (the reference's `embs` argument).  One stream; for many streams use Engine."""

    def __init__(self, reid_weights="", use_half=False, use_gpu=False, det_thresh=0.3, max_age=30, max_obs=50,
                 min_hits=3, iou_threshold=0.3, per_class=False, nr_classes=80, asso_func="iou", is_obb=False,
                 track_high_thresh=0.5, track_low_thresh=0.1, new_track_thresh=0.6, track_buffer=30, match_thresh=0.8,
                 proximity_thresh=0.5, appearance_thresh=0.25, cmc_method="none", frame_rate=30,
                 fuse_first_associate=False, with_reid=True, emb_dim=0, track_capacity=1536, max_dets=512, device=0):
        if cmc_method not in ("none", "", None):
            raise ValueError("camera-motion compensation (cmc_method=%r) is outside the accelerated hot path" % cmc_method)
        if reid_weights:
            raise ValueError("ReID inference is outside the accelerated hot path: pass embeddings to update()")
        if asso_func != "iou" or per_class or is_obb:
            raise ValueError("only asso_func=\"iou\", per_class=False, is_obb=False are on the accelerated path")
        self._make = lambda dim: Engine(_lib.TRACKER_BOTSORT, 1, track_capacity, max_dets, device, det_thresh=det_thresh,
                                        max_age=max_age, max_obs=max_obs, min_hits=min_hits, iou_threshold=iou_threshold,
                                        track_high_thresh=track_high_thresh, track_low_thresh=track_low_thresh,
                                        new_track_thresh=new_track_thresh, track_buffer=track_buffer, match_thresh=match_thresh,
                                        proximity_thresh=proximity_thresh, appearance_thresh=appearance_thresh,
                                        frame_rate=frame_rate, fuse_first_associate=int(bool(fuse_first_associate)),
                                        with_reid=int(bool(with_reid)), emb_dim=dim)
        self._frames = 0
        self._build(emb_dim)

    def _build(self, emb_dim):
        self._engine = self._make(emb_dim)
        self._dim = emb_dim
        self._max_dets = self._engine.cfg.max_dets
        self._cap = self._engine.cfg.track_capacity
        self._dets = np.zeros((1, 1, self._max_dets, 6), np.float32)
        self._embs = np.zeros((1, 1, self._max_dets, emb_dim), np.float32) if emb_dim else None
        self._out = np.empty((1, 1, self._cap, 8), np.float32)
        self._n_out = np.empty((1, 1), np.int32)

    def _adopt_embedding_dim(self, embs, n):
        """The reference takes whatever `embs` it is handed with the same positional constructor.  A tracker built without
        emb_dim therefore sizes itself from the first embeddings it sees; once frames have been processed the dimension
        is fixed and a mismatch is an error - never a silent fall-back to IoU-only association."""
        embs = np.asarray(embs, np.float32)
        if embs.ndim != 2 or embs.shape[0] != n:
            raise ValueError("Detections and embeddings must have same number of rows")
        if embs.shape[1] != self._dim:
            if self._dim == 0 and self._frames == 0:
                self._engine.close()
                self._build(int(embs.shape[1]))
            else:
                raise ValueError(f"embeddings have {embs.shape[1]} columns but the tracker was built for emb_dim={self._dim}")
        return embs

    def reset(self):
        self._engine.reset()

    def update(self, dets, img, embs=None) -> np.ndarray:
        dets = np.asarray(dets, np.float32)
        if dets.ndim != 2:
            dets = dets.reshape(0, 6) if dets.size == 0 else dets
        # BaseTracker::check_inputs(dets, img) (src/tracker.cpp:108-125; BotSort passes no embs to it)
        if dets.shape[0] > 0 and dets.shape[1] not in (6, 7):
            raise ValueError("Detections must have 6 (AABB) or 7 (OBB) columns")
        if _Image(img).empty():
            raise ValueError("Image cannot be empty")
        if dets.shape[0] > 0 and dets.shape[1] == 7:
            raise ValueError("OBB detections are outside the accelerated hot path")
        n = dets.shape[0]
        if n == 0:
            return np.zeros((0, 8), np.float32)                 # botsort.cpp:267-269
        if n > self._max_dets:
            raise ValueError(f"{n} detections exceed max_dets={self._max_dets}")
        self._dets[0, 0, :n] = dets[:, :6]
        e = None
        if embs is not None and np.size(embs):
            embs = self._adopt_embedding_dim(embs, n)
            self._dets[0, 0, :n] = dets[:, :6]
            self._embs[0, 0, :n] = embs
            e = self._embs
        self._engine.update(self._dets, np.array([[n]], np.int32), self._out, self._n_out, embs=e)
        self._engine.check()
        self._frames += 1
        return self._out[0, 0, :int(self._n_out[0, 0])].copy()


class StrongSort:
    """motcpp::trackers::StrongSORT with the reference's positional constructor
    (include/motcpp/trackers/strongsort.hpp:287-305).  The ReID network and ECC camera-motion estimation are image
    processing outside the accelerated path: embeddings are passed to update() (the reference's `embs` argument) and the
    camera warp is the identity (what motion::ECC yields on a static / featureless image).  One stream; for many
    streams use Engine(TRACKER_STRONGSORT, ...)."""

    def __init__(self, reid_weights="", use_half=False, use_gpu=False, det_thresh=0.3, max_age=30, max_obs=50,
                 min_hits=3, iou_threshold=0.3, per_class=False, nr_classes=80, asso_func="iou", is_obb=False,
                 min_conf=0.1, max_cos_dist=0.2, max_iou_dist=0.7, n_init=3, nn_budget=100, mc_lambda=0.98,
                 ema_alpha=0.9, emb_dim=0, track_capacity=1536, max_dets=512, device=0):
        if reid_weights:
            raise ValueError("ReID inference is outside the accelerated hot path: pass embeddings to update()")
        if per_class or is_obb:
            raise ValueError("only per_class=False, is_obb=False are on the accelerated path")
        self._make = lambda dim: Engine(_lib.TRACKER_STRONGSORT, 1, track_capacity, max_dets, device, det_thresh=det_thresh,
                                        max_age=max_age, max_obs=max_obs, min_hits=min_hits, iou_threshold=iou_threshold,
                                        min_conf=min_conf, max_cos_dist=max_cos_dist, max_iou_dist=max_iou_dist, n_init=n_init,
                                        nn_budget=nn_budget, mc_lambda=mc_lambda, ema_alpha=ema_alpha, emb_dim=dim)
        self._frames = 0
        self._build(emb_dim)

    _adopt_embedding_dim = BotSort._adopt_embedding_dim

    def _build(self, emb_dim):
        self._engine = self._make(emb_dim)
        self._dim = emb_dim
        self._max_dets = self._engine.cfg.max_dets
        self._cap = self._engine.cfg.track_capacity
        self._dets = np.zeros((1, 1, self._max_dets, 6), np.float32)
        self._embs = np.zeros((1, 1, self._max_dets, emb_dim), np.float32) if emb_dim else None
        self._out = np.empty((1, 1, self._cap, 8), np.float32)
        self._n_out = np.empty((1, 1), np.int32)

    def reset(self):
        self._engine.reset()

    def update(self, dets, img, embs=None) -> np.ndarray:
        dets = np.asarray(dets, np.float32)
        if dets.ndim != 2:
            dets = dets.reshape(0, 6) if dets.size == 0 else dets
        # BaseTracker::check_inputs(dets, img, embs) (src/tracker.cpp:108-125)
        if dets.shape[0] > 0 and dets.shape[1] not in (6, 7):
            raise ValueError("Detections must have 6 (AABB) or 7 (OBB) columns")
        if _Image(img).empty():
            raise ValueError("Image cannot be empty")
        if dets.shape[0] > 0 and dets.shape[1] == 7:
            raise ValueError("OBB detections are outside the accelerated hot path")
        n = dets.shape[0]
        if embs is not None and np.size(embs) and np.shape(embs)[0] != n:
            raise ValueError("Detections and embeddings must have same number of rows")
        if n > self._max_dets:
            raise ValueError(f"{n} detections exceed max_dets={self._max_dets}")
        self._dets[0, 0, :n] = dets[:, :6] if n else 0
        e = None
        if n and embs is not None and np.size(embs):
            embs = self._adopt_embedding_dim(embs, n)
            self._dets[0, 0, :n] = dets[:, :6]
            self._embs[0, 0, :n] = embs
            e = self._embs
        # an empty frame still advances the tracker: predict + every track missed (strongsort.cpp:833-837)
        self._engine.update(self._dets, np.array([[n]], np.int32), self._out, self._n_out, embs=e)
        self._engine.check()
        self._frames += 1
        return self._out[0, 0, :int(self._n_out[0, 0])].copy()
