"""Pins the oracle (oracle/liboracle.so) to the REFERENCE'S OWN CODE.

oracle/_ref/libref_core*.so is built by oracle/Makefile from the sources where they lie under /root/reference
(src/motion/*.cpp, src/utils/matching.cpp, include/motcpp/utils/{iou,ops,matching}.hpp, lap_solver.hpp,
src/tracker.cpp and src/trackers/{sort,bytetrack,ocsort,botsort,strongsort,deepocsort}.cpp), against the stand-in
Eigen / OpenCV headers of oracle/ref_shim/.  Nothing under /root/reference is read at test time: the binaries
travel (git-ignored, not gpurun-ignored); a checkout without them skips this module.

Two builds of the stand-in Eigen bound what cannot be pinned without real Eigen - the evaluation order inside
its kernels:
  * "textbook" order (sequential reductions)            -> the oracle must agree BIT FOR BIT, everywhere;
  * "eigen" order (Eigen 3.4 / SSE2 order, as recalled) -> identical track ids / assignment indices / integer
    columns, floating-point state within 1e-4 of the state's scale (north_star's tolerance), measured worst
    case asserted at 2e-5.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
import ref_lib as R
from motcpp_b200 import synth

pytestmark = pytest.mark.skipif(not (R.available("textbook") and R.available("eigen")),
                                reason="oracle/_ref/libref_core*.so not built (needs /root/reference at build time)")

N_RANDOM = 1000          # random inputs per function (>= 1000)
EIG_TOL = 2e-5           # norm-wise tolerance for the "eigen" evaluation order


def _close(a, b, tol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(1e-30, float(np.max(np.abs(b)))) if b.size else 1.0
    return float(np.max(np.abs(a - b))) / scale <= tol if a.size else True


class _Cmp:
    """Collects per-function agreement; textbook order must be bit-exact, eigen order within EIG_TOL (norm-wise)."""

    def __init__(self, order):
        self.order, self.bad = order, []

    def __call__(self, name, a, b):
        if self.order == "textbook":
            ok = np.array_equal(a, b)
        else:
            ok = _close(a, b, EIG_TOL)
        if not ok:
            self.bad.append(name)


def _rand_spd(rng, n, scale):
    a = rng.normal(0, 1, (n, n))
    s = a @ a.T + n * np.eye(n)
    d = np.sqrt(scale / np.diag(s))
    return (s * d[:, None] * d[None, :]).astype(np.float32)


def _boxes(rng, k, lo=0, hi=400):
    c = rng.uniform(lo, hi, (k, 2))
    w = rng.uniform(20, 120, (k, 2))
    return np.c_[c - w / 2, c + w / 2].astype(np.float32)


@pytest.mark.parametrize("order", ["textbook", "eigen"])
def test_box_conversions_and_kalman_filters_equal_reference(order):
    """ops.hpp:15-211, kalman_filter.cpp:29-176 + xyah_kf.cpp, xysr_kf.cpp:10-141, xywh_kf.hpp:41-177 on DENSE random
    SPD covariances (harder than anything a tracker produces: its covariances are block-diagonal per coordinate)."""
    L, OL, rng, cmp = R.lib(order), O.lib(), np.random.default_rng(11), _Cmp(order)
    for _ in range(N_RANDOM):
        box = np.array([rng.uniform(0, 1900), rng.uniform(0, 1000), 0, 0], np.float32)
        box[2], box[3] = box[0] + rng.uniform(5, 300), box[1] + rng.uniform(5, 300)
        for name in ("xyxy2xywh", "xywh2xyxy", "xywh2tlwh", "tlwh2xyah", "xyah2xywh", "xyxy2xysr", "xysr2xyxy"):
            a, b = np.zeros(4, np.float32), np.zeros(4, np.float32)
            getattr(OL, "orc_" + name)(box, a)
            getattr(L, "ref_" + name)(box, b)
            assert np.array_equal(a, b), name
        z = np.array([rng.uniform(0, 1900), rng.uniform(0, 1000), rng.uniform(0.2, 2), rng.uniform(20, 300)], np.float32)
        # ---- XYAH
        m1, c1, m2, c2 = np.zeros(8, np.float32), np.zeros(64, np.float32), np.zeros(8, np.float32), np.zeros(64, np.float32)
        OL.orc_kf_xyah_initiate(z, m1, c1)
        L.ref_kf_xyah_initiate(z, m2, c2)
        cmp("xyah_initiate", np.r_[m1, c1], np.r_[m2, c2])
        mean = np.r_[z, rng.normal(0, 3, 4)].astype(np.float32)
        cov = _rand_spd(rng, 8, (z[3] / 10) ** 2).reshape(-1).copy()
        m1, c1, m2, c2 = mean.copy(), cov.copy(), mean.copy(), cov.copy()
        OL.orc_kf_xyah_predict(m1, c1)
        L.ref_kf_xyah_predict(m2, c2)
        cmp("xyah_predict", np.r_[m1, c1], np.r_[m2, c2])
        conf = float(rng.uniform(0, 0.9))
        pm1, pc1, pm2, pc2 = np.zeros(4, np.float32), np.zeros(16, np.float32), np.zeros(4, np.float32), np.zeros(16, np.float32)
        OL.orc_kf_xyah_project(mean, cov, conf, pm1, pc1)
        L.ref_kf_xyah_project(mean, cov, conf, pm2, pc2)
        cmp("xyah_project", np.r_[pm1, pc1], np.r_[pm2, pc2])
        zz = (z + rng.normal(0, 2, 4) * np.array([1, 1, 0.01, 1])).astype(np.float32)
        m1, c1, m2, c2 = mean.copy(), cov.copy(), mean.copy(), cov.copy()
        assert OL.orc_kf_xyah_update(m1, c1, zz, conf) == 0 and L.ref_kf_xyah_update(m2, c2, zz, conf) == 0
        cmp("xyah_update mean", m1, m2)
        cmp("xyah_update cov", c1, c2)
        meas = (z[None, :] + rng.normal(0, 5, (7, 4)) * np.array([1, 1, 0.01, 1])).astype(np.float32)
        for only_pos in (0, 1):
            for metric in (0, 1):          # "maha" (the reference's d^T S^-2 d) and "gaussian"
                g1, g2 = np.zeros(7, np.float32), np.zeros(7, np.float32)
                OL.orc_kf_xyah_gating(mean, cov, meas, 7, only_pos, metric, g1)
                assert L.ref_kf_xyah_gating(mean, cov, meas, 7, only_pos, metric, g2) == 0
                cmp("xyah_gating", g1, g2)
        # ---- XYWH
        zw = np.array([z[0], z[1], z[2] * z[3], z[3]], np.float32)
        m1, c1, m2, c2 = np.zeros(8, np.float32), np.zeros(64, np.float32), np.zeros(8, np.float32), np.zeros(64, np.float32)
        OL.orc_kf_xywh_initiate(zw, m1, c1)
        L.ref_kf_xywh_initiate(zw, m2, c2)
        cmp("xywh_initiate", np.r_[m1, c1], np.r_[m2, c2])
        mean = np.r_[zw, rng.normal(0, 3, 4)].astype(np.float32)
        m1, c1, m2, c2 = mean.copy(), cov.copy(), mean.copy(), cov.copy()
        OL.orc_kf_xywh_predict(m1, c1)
        L.ref_kf_xywh_predict(m2, c2)
        cmp("xywh_predict", np.r_[m1, c1], np.r_[m2, c2])
        zz = (zw + rng.normal(0, 2, 4)).astype(np.float32)
        m1, c1, m2, c2 = mean.copy(), cov.copy(), mean.copy(), cov.copy()
        OL.orc_kf_xywh_update(m1, c1, zz)
        L.ref_kf_xywh_update(m2, c2, zz)
        cmp("xywh_update mean", m1, m2)
        cmp("xywh_update cov", c1, c2)
        measw = (zw[None, :] + rng.normal(0, 5, (7, 4))).astype(np.float32)
        for only_pos in (0, 1):            # only_position = top-left 2x2 OF THE 4x4 INVERSE (xywh_kf.hpp:168-171)
            g1, g2 = np.zeros(7, np.float32), np.zeros(7, np.float32)
            OL.orc_kf_xywh_gating(mean, cov, measw, 7, only_pos, g1)
            L.ref_kf_xywh_gating(mean, cov, measw, 7, only_pos, g2)
            cmp("xywh_gating", g1, g2)
        # ---- XYSR
        zs = np.array([z[0], z[1], z[2] * z[3] * z[3], z[2]], np.float32)
        x1, p1, x2, p2 = np.zeros(7, np.float32), np.zeros(49, np.float32), np.zeros(7, np.float32), np.zeros(49, np.float32)
        OL.orc_kf_xysr_init(zs, x1, p1)
        L.ref_kf_xysr_init(zs, x2, p2)
        cmp("xysr_init", np.r_[x1, p1], np.r_[x2, p2])
        x = np.r_[zs, rng.normal(0, 3, 2), rng.normal(0, 30, 1)].astype(np.float32)
        P = _rand_spd(rng, 7, 50.0).reshape(-1).copy()
        for q in ((1.0, 1.0), (0.01, 0.0001)):      # SORT / OC-SORT's Q scaling (ocsort.cpp:77-79)
            x1, p1, x2, p2 = x.copy(), P.copy(), x.copy(), P.copy()
            OL.orc_kf_xysr_predict(x1, p1, *q)
            L.ref_kf_xysr_predict(x2, p2, *q)
            cmp("xysr_predict", np.r_[x1, p1], np.r_[x2, p2])
        zz = (zs + rng.normal(0, 2, 4) * np.array([1, 1, 50, 0.01])).astype(np.float32)
        x1, p1, x2, p2 = x.copy(), P.copy(), x.copy(), P.copy()
        assert OL.orc_kf_xysr_update(x1, p1, zz) == 0 and L.ref_kf_xysr_update(x2, p2, zz) == 0
        cmp("xysr_update x", x1, x2)
        cmp("xysr_update P", p1, p2)
        th = rng.uniform(-0.1, 0.1)
        m2x2 = (np.array([np.cos(th), -np.sin(th), np.sin(th), np.cos(th)]) * rng.uniform(0.95, 1.05)).astype(np.float32)
        t2 = rng.normal(0, 5, 2).astype(np.float32)
        x1, p1, x2, p2 = x.copy(), P.copy(), x.copy(), P.copy()
        OL.orc_kf_xysr_affine(x1, p1, m2x2, t2)
        L.ref_kf_xysr_affine(x2, p2, m2x2, t2)
        cmp("xysr_affine", np.r_[x1, p1], np.r_[x2, p2])
    assert not cmp.bad, sorted(set(cmp.bad))


@pytest.mark.parametrize("order", ["textbook", "eigen"])
def test_cost_builders_and_assignment_equal_reference(order):
    """iou.hpp:63-100, matching.cpp:14-143, lap_solver.hpp (through utils::linear_assignment), iou.hpp:119-330 variants
    on their well-defined domain (1 x 1, SURVEY trap 11), deepocsort.cpp:294-345."""
    L, OL, rng, cmp = R.lib(order), O.lib(), np.random.default_rng(12), _Cmp(order)
    for it in range(N_RANDOM):
        n, m = int(rng.integers(1, 14)), int(rng.integers(1, 14))
        A, B = _boxes(rng, n), _boxes(rng, m)
        if it % 10 == 0:
            B[0] = A[0]                               # identical boxes -> IoU exactly 1
        o1, o2 = np.zeros((n, m), np.float32), np.zeros((n, m), np.float32)
        OL.orc_iou_batch(A, n, B, m, o1)
        L.ref_iou_batch(A, n, B, m, o2)
        assert np.array_equal(o1, o2)
        OL.orc_iou_distance(A, n, B, m, o1)
        L.ref_iou_distance(A, n, B, m, o2)
        assert np.array_equal(o1, o2)
        conf = rng.uniform(0.1, 1, m).astype(np.float32)
        f1, f2 = o1.copy(), o1.copy()
        OL.orc_fuse_score(f1, n, m, conf)
        L.ref_fuse_score(f2, n, m, conf)
        assert np.array_equal(f1, f2)
        # assignment through the reference's own utils::linear_assignment; every other problem quantised => heavy ties
        cost = np.ascontiguousarray(f1 if it % 2 else np.round(f1 * 4) / 4, np.float32)
        r1, c1, r2, c2 = np.zeros(n, np.int32), np.zeros(m, np.int32), np.zeros(n, np.int32), np.zeros(m, np.int32)
        thresh = float(rng.choice([0.5, 0.7, 0.8, 0.9]))
        k1 = OL.orc_linear_assignment(cost, n, m, m, thresh, r1, c1)
        k2 = L.ref_linear_assignment(cost, n, m, m, thresh, r2, c2)
        assert k1 == k2 and np.array_equal(r1, r2) and np.array_equal(c1, c2)
        dim = int(rng.choice([3, 8, 32, 128, 512, 513]))
        Tf, Df = rng.normal(0, 1, (n, dim)).astype(np.float32), rng.normal(0, 1, (m, dim)).astype(np.float32)
        for metric in (0, 1):
            e1, e2 = np.zeros((n, m), np.float32), np.zeros((n, m), np.float32)
            OL.orc_embedding_distance(Tf, n, Df, m, dim, metric, e1)
            assert L.ref_embedding_distance(Tf, n, Df, m, dim, metric, e2) == 0
            cmp("embedding_distance", e1, e2)
        emb = rng.uniform(0, 1, (n, m)).astype(np.float32)
        a1, a2 = np.zeros_like(emb), np.zeros_like(emb)
        OL.orc_aw_max_metric(emb, n, m, m, 0.5, 0.5, a1, m)
        assert L.ref_aw_max_metric(emb, n, m, m, 0.5, 0.5, a2, m) == 0
        assert np.array_equal(a1, a2)
        for kind, name in ((3, b"hmiou"), (4, b"giou"), (5, b"diou"), (6, b"centroid")):
            v1, v2 = np.zeros((1, 1), np.float32), np.zeros((1, 1), np.float32)
            OL.orc_iou_variant(A[:1], 1, B[:1], 1, kind, 1920, 1080, v1)
            assert L.ref_asso_func(name, A[:1], 1, B[:1], 1, 1920, 1080, v2) == 0, L.ref_last_error()
            assert np.array_equal(v1, v2), name
        # ciou: the stand-in Eigen's .atan() is the box's libm atanf, within 1 ulp of the oracle's correctly rounded atan
        v1, v2 = np.zeros((1, 1), np.float32), np.zeros((1, 1), np.float32)
        OL.orc_iou_variant(A[:1], 1, B[:1], 1, 7, 1920, 1080, v1)
        assert L.ref_asso_func(b"ciou", A[:1], 1, B[:1], 1, 1920, 1080, v2) == 0, L.ref_last_error()
        assert abs(float(v1[0, 0]) - float(v2[0, 0])) <= 2e-7, (v1, v2)
    assert not cmp.bad, sorted(set(cmp.bad))


def test_reference_variants_are_shape_inconsistent_beyond_one_row():
    """SURVEY trap 11: hmiou/giou/ciou/diou mix an (N,M) with an (N*M,1) replicate - only 1 x 1 is defined.  The
    stand-in Eigen refuses what a release Eigen build would read out of bounds."""
    L, rng = R.lib("textbook"), np.random.default_rng(5)
    A, B = _boxes(rng, 3), _boxes(rng, 2)
    out = np.zeros((3, 2), np.float32)
    assert L.ref_asso_func(b"iou", A, 3, B, 2, 1920, 1080, out) == 0
    for name in (b"hmiou", b"giou", b"diou", b"ciou"):
        assert L.ref_asso_func(name, A, 3, B, 2, 1920, 1080, out) == -1000
        assert b"different sizes" in L.ref_last_error()


# ---------------------------------------------------------------------------------------------- state machines
BT = [0.3, 30, 50, 3, 0.3, 0.1, 0.45, 0.8, 30, 30]
OC = dict(det_thresh=0.2, max_age=30, max_obs=50, min_hits=3, iou_threshold=0.3, min_conf=0.1, delta_t=3, inertia=0.2,
          use_byte=False, q_xy_scaling=0.01, q_s_scaling=0.0001)
BOT = dict(track_high_thresh=0.5, track_low_thresh=0.1, new_track_thresh=0.6, track_buffer=30, match_thresh=0.8,
           proximity_thresh=0.5, appearance_thresh=0.25, frame_rate=30, fuse_first_associate=False, with_reid=True)
SS = dict(max_age=30, min_conf=0.1, max_cos_dist=0.2, max_iou_dist=0.7, n_init=3, nn_budget=100, mc_lambda=0.98,
          ema_alpha=0.9)


def _run(order, orc, ref, frames):
    """frames: iterable of (dets, embs|None).  Rows must agree frame by frame: integer columns (id, cls, det_ind) and
    conf exactly in both orders; boxes bit-exact (textbook) / within EIG_TOL of the frame's scale (eigen)."""
    worst, n_rows = 0.0, 0
    for t, (d, e) in enumerate(frames):
        a = orc.update(d) if e is None else orc.update(d, e)
        b = ref.update(d, e)
        assert a.shape == b.shape, f"frame {t}: {a.shape} vs reference {b.shape}"
        assert np.array_equal(a[:, 4:], b[:, 4:]), f"frame {t}: id / conf / cls / det_ind differ"
        if order == "textbook":
            assert np.array_equal(a, b), f"frame {t}: boxes differ"
        elif a.size:
            err = float(np.max(np.abs(a[:, :4].astype(np.float64) - b[:, :4]))) / max(1.0, float(np.max(np.abs(b[:, :4]))))
            worst = max(worst, err)
            assert err <= EIG_TOL, f"frame {t}: boxes differ by {err}"
        n_rows += a.shape[0]
    assert n_rows > 0
    return worst


def _stress(sid, T):
    dets, counts = synth.stress_stream(sid, n_frames=T)
    return [(dets[t][:counts[t]], None) for t in range(T)]


def _stress_reid(sid, T, dim=32):
    dets, counts, embs = synth.stress_stream_reid(sid, n_frames=T, dim=dim)
    return [(dets[t][:counts[t]], embs[t][:counts[t]]) for t in range(T)]


@pytest.mark.parametrize("order", ["textbook", "eigen"])
@pytest.mark.parametrize("sid", [0, 1, 2])
def test_bytetrack_300_frame_stress_streams_equal_reference(order, sid):
    _run(order, O.ByteTrack(*BT), R.Tracker("bytetrack", BT, order), _stress(sid, 300))


@pytest.mark.parametrize("order", ["textbook", "eigen"])
def test_bytetrack_c2_shape_equals_reference(order):
    """BASELINE configs[1] itself: 256 objects / 512 detections per frame, 40 frames (covers the 30-frame lost window)."""
    dets = synth.bytetrack_stream(3, n_frames=40)
    _run(order, O.ByteTrack(*BT), R.Tracker("bytetrack", BT, order), [(dets[t], None) for t in range(40)])


@pytest.mark.parametrize("order", ["textbook", "eigen"])
@pytest.mark.parametrize("sid,args", [(0, (0.3, 1, 50, 3, 0.3)), (1, (0.3, 3, 50, 1, 0.3)), (2, (0.5, 30, 50, 3, 0.2))])
def test_sort_stress_streams_equal_reference(order, sid, args):
    _run(order, O.Sort(*args), R.Tracker("sort", list(args), order), _stress(20 + sid, 300))


@pytest.mark.parametrize("order", ["textbook", "eigen"])
@pytest.mark.parametrize("sid,over", [(0, {}), (1, {"use_byte": True}), (2, {"delta_t": 1, "inertia": 0.4, "min_hits": 1}),
                                      (3, {"max_age": 5, "iou_threshold": 0.2})])
def test_ocsort_stress_streams_equal_reference(order, sid, over):
    """Includes the duplicate-spawn trap (ocsort.cpp:702-735): twin tracks tie exactly, so the reference's LAPJV scan
    order decides - the oracle runs tie_mode 0 (its LAPJV restatement) and must reproduce every frame."""
    args = {**OC, **over}
    ref = R.Tracker("ocsort", [float(v) for v in args.values()], order)
    _run(order, O.OCSort(**args, tie_mode=0), ref, _stress(30 + sid, 300))


@pytest.mark.parametrize("order", ["textbook", "eigen"])
def test_ocsort_c4_shape_equals_reference(order):
    dets = synth.ocsort_stream(1, n_frames=3)
    ref = R.Tracker("ocsort", [float(v) for v in OC.values()], order)
    _run(order, O.OCSort(**OC, tie_mode=0), ref, [(dets[t], None) for t in range(3)])


@pytest.mark.parametrize("order", ["textbook", "eigen"])
@pytest.mark.parametrize("sid,over,use_embs", [(0, {}, True), (1, {"fuse_first_associate": True}, True),
                                               (2, {"with_reid": False}, False), (3, {"appearance_thresh": 0.6, "proximity_thresh": 0.8}, True)])
def test_botsort_stress_streams_equal_reference(order, sid, over, use_embs):
    args = {**BOT, **over}
    frames = _stress_reid(40 + sid, 200)
    if not use_embs:
        frames = [(d, None) for d, _ in frames]
    _run(order, O.BotSort(**args), R.Tracker("botsort", [float(v) for v in args.values()], order), frames)


@pytest.mark.parametrize("order", ["textbook", "eigen"])
def test_botsort_c3_shape_equals_reference(order):
    dets, embs = synth.embeddings_stream(2, n_frames=3)
    ref = R.Tracker("botsort", [float(v) for v in BOT.values()], order)
    _run(order, O.BotSort(**BOT), ref, [(dets[t], embs[t]) for t in range(3)])


@pytest.mark.parametrize("order", ["textbook", "eigen"])
@pytest.mark.parametrize("sid,over", [(0, {}), (1, {"max_cos_dist": 0.4, "nn_budget": 5, "n_init": 2}),
                                      (2, {"max_age": 5, "mc_lambda": 0.9, "ema_alpha": 0.8})])
def test_strongsort_stress_streams_equal_reference(order, sid, over):
    """The reference's duplicated-row quirk (strongsort.cpp:433-436, :728-733) makes exact ties systematic; oracle in
    tie_mode 0 (= LAPJV) must follow the reference's own solver through them."""
    args = {**SS, **over}
    ref = R.Tracker("strongsort", [float(v) for v in args.values()], order)
    _run(order, O.StrongSort(**args, tie_mode=0), ref, _stress_reid(50 + sid, 200))


@pytest.mark.parametrize("order", ["textbook", "eigen"])
def test_strongsort_dense_workload_equals_reference(order):
    dets, embs = synth.strongsort_stream(0, n_frames=12, n_obj=48, n_clutter=16, dim=64)
    ref = R.Tracker("strongsort", [float(v) for v in SS.values()], order)
    _run(order, O.StrongSort(**SS, tie_mode=0), ref, [(dets[t], embs[t]) for t in range(12)])


DOC = dict(det_thresh=0.3, max_age=30, max_obs=50, min_hits=3, iou_threshold=0.3, delta_t=3, inertia=0.2, w_association_emb=0.5,
           alpha_fixed_emb=0.95, aw_param=0.5, embedding_off=False, aw_off=False, q_xy_scaling=0.01, q_s_scaling=0.0001)


@pytest.mark.parametrize("order", ["textbook", "eigen"])
@pytest.mark.parametrize("sid,over,use_embs", [(0, {}, True), (1, {"aw_off": True}, True), (2, {"embedding_off": True}, False),
                                               (3, {"inertia": 0.9, "min_hits": 1, "max_age": 5, "delta_t": 1}, True)])
def test_deepocsort_stress_streams_equal_reference(order, sid, over, use_embs):
    """SURVEY 8f-2.  The reference lists every detection its assignment leaves unmatched twice (deepocsort.cpp:476-478 and
    :491-495), so two bit-identical tracks are spawned per new object whenever the assignment branch runs: exact ties in
    nearly every frame, resolved by the reference's LAPJV - which the oracle must follow step for step."""
    args = {**DOC, **over}
    frames = _stress_reid(70 + sid, 150)
    if not use_embs:
        frames = [(d, None) for d, _ in frames]
    _run(order, O.DeepOCSort(**args), R.Tracker("deepocsort", [float(v) for v in args.values()], order), frames)


def test_reset_and_empty_frames_equal_reference():
    """reset() and empty inputs: BoT-SORT returns early WITHOUT advancing (botsort.cpp:267-269), ByteTrack advances."""
    for kind, orc, params in (("bytetrack", O.ByteTrack(*BT), BT), ("botsort", O.BotSort(**BOT), [float(v) for v in BOT.values()]),
                              ("ocsort", O.OCSort(**OC), [float(v) for v in OC.values()]),
                              ("sort", O.Sort(0.3, 1, 50, 3, 0.3), [0.3, 1, 50, 3, 0.3])):
        ref = R.Tracker(kind, params, "textbook")
        frames = _stress(60, 40)
        empty = np.zeros((0, 6), np.float32)
        seq = frames[:10] + [(empty, None)] * 3 + frames[10:25]
        for d, _ in seq:
            a, b = orc.update(d), ref.update(d)
            assert a.shape == b.shape and np.array_equal(a[:, :4], b[:, :4]) and np.array_equal(a[:, 5:], b[:, 5:]), kind
        orc.reset()
        ref.reset()
        ids_o, ids_r = [], []
        for d, _ in frames[25:]:
            a, b = orc.update(d), ref.update(d)
            assert np.array_equal(a[:, :4], b[:, :4]) and np.array_equal(a[:, 5:], b[:, 5:]), kind
            ids_o.append(a[:, 4])
            ids_r.append(b[:, 4])
        # after reset the reference keeps counting ids where it was (process-global statics; BoT-SORT restarts at 1);
        # the id sequences must be equal up to that constant offset
        io, ir = np.concatenate(ids_o), np.concatenate(ids_r)
        assert io.size and np.all(ir - io == ir[0] - io[0]), kind


@pytest.mark.parametrize("order", ["textbook", "eigen"])
@pytest.mark.parametrize("sid,over", [(0, {}), (1, {"use_byte": True, "iou_threshold": 0.9}), (2, {"iou_threshold": 0.95, "inertia": 0.5})])
def test_ocsort_centroid_association_equals_reference(order, sid, over):
    """asso_func = "centroid" (iou.hpp:298-330, the one variant whose expression is defined for every N x M): the similarity is
    1 - centre distance / frame diagonal for EVERY pair, so thresholds near 1 are the meaningful ones and nothing is sparse."""
    args = {**OC, **over}
    ref = R.Tracker("ocsort", [float(v) for v in args.values()], order, asso_func="centroid")
    _run(order, O.OCSort(**args, tie_mode=0, asso_func="centroid", frame=(1920, 1080)), ref, _stress(90 + sid, 120))


BOOST = dict(det_thresh=0.6, max_age=60, max_obs=50, min_hits=3, iou_threshold=0.3, min_box_area=10, aspect_ratio_thresh=1.6,
             lambda_iou=0.5, lambda_mhd=0.25, lambda_shape=0.25, use_dlo_boost=True, dlo_boost_coef=0.65, use_vt=False)


@pytest.mark.parametrize("order", ["textbook", "eigen"])
@pytest.mark.parametrize("sid,over", [(0, {}), (1, {"det_thresh": 0.4, "max_age": 8, "min_hits": 1, "use_vt": True}),
                                      (2, {"use_dlo_boost": False, "lambda_mhd": 0.6, "iou_threshold": 0.5, "aspect_ratio_thresh": 0.5}),
                                      (3, {"det_thresh": 0.3, "dlo_boost_coef": 0.9, "max_age": 3, "min_box_area": 4000})])
def test_boosttrack_stress_streams_equal_reference(order, sid, over):
    """SURVEY 8f-1, second half: BoostTrack (ECC / ReID off) - Kalman filter through Eigen's dynamic 4 x 4 inverse, IoU +
    diagonal-Mahalanobis cost blend, detection-confidence boost, output filter."""
    args = {**BOOST, **over}
    ref = R.Tracker("boosttrack", [float(v) for v in args.values()], order)
    _run(order, O.BoostTrack(**args), ref, _stress(110 + sid, 200))
