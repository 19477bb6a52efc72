"""Parity tests proper: the sm_100a kernels, called through the C ABI, against the CPU oracle on
the same seeded inputs.  Bar: bit-exact for assignment indices, track IDs and (because the kernels
reproduce the oracle's operation order) every fp32 value; Kalman state additionally checked at the
1e-4 relative tolerance north_star states."""
import ctypes as C

import numpy as np
import pytest

from motcpp_b200 import _lib, api, synth

pytestmark = pytest.mark.gpu

KF_RTOL = 1e-4      # north_star: "within 1e-4 rel on Kalman state"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    from motcpp_b200 import build
    build.build()
    _lib.require_gpu()          # raises (does not skip): -m gpu on a box without a GPU is an error


# ------------------------------------------------------------------ linear assignment
def _rand_cost(rng, kind):
    if kind == 0:
        n, m = rng.integers(1, 12, 2)
        return rng.random((n, m)).astype(np.float32), 0.5
    if kind == 1:
        n, m = rng.integers(5, 60, 2)
        return np.where(rng.random((n, m)) < 0.9, 1.0, rng.random((n, m))).astype(np.float32), 0.8
    if kind == 2:
        n, m = rng.integers(20, 70, 2)
        return rng.random((n, m)).astype(np.float32), 0.7
    if kind == 3:
        n, m = rng.integers(100, 600, 2)
        return np.where(rng.random((n, m)) < 0.995, 1.0, rng.random((n, m))).astype(np.float32), 0.8
    n, m = rng.integers(1, 40, 2)
    return -(rng.random((n, m)) * 1.3).astype(np.float32), -0.3


def test_lap_matches_oracle(oracle):
    rng = np.random.default_rng(11)
    for trial in range(300):
        c, th = _rand_cost(rng, trial % 5)
        got = api.linear_assignment_arrays(c, th)
        ref = oracle.linear_assignment(c, th)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), (trial, c.shape)


def test_lap_reference_kats():
    # reference tests/test_matching.cpp:14-110
    la = api.linear_assignment
    r = la(np.zeros((0, 0), np.float32), 0.5)
    assert (r.matches, r.unmatched_a, r.unmatched_b) == ([], [], [])
    assert la([[0.1]], 0.5).matches == [(0, 0)]
    r = la([[0.9]], 0.5)
    assert r.matches == [] and r.unmatched_a == [0] and r.unmatched_b == [0]
    assert la([[.1, .9, .9], [.9, .1, .9], [.9, .9, .1]], 0.5).matches == [(0, 0), (1, 1), (2, 2)]
    r = la([[.1, .9], [.9, .1], [.9, .9]], 0.5)
    assert r.matches == [(0, 0), (1, 1)] and r.unmatched_a == [2] and r.unmatched_b == []
    r = la([[.1, .9, .9], [.9, .1, .9]], 0.5)
    assert r.matches == [(0, 0), (1, 1)] and r.unmatched_b == [2]
    assert la([[.1, .2], [.3, .1]], 0.5).matches == [(0, 0), (1, 1)]


def test_lap_headline_shape_and_adversarial(oracle):
    """256 x 512 from a real C2 frame, and the all-overlapping adversarial case (one component)."""
    dets = synth.bytetrack_stream(0, n_frames=2)
    a = dets[0, :256, :4]
    b = dets[1, :, :4]
    cost = oracle.fuse_score(oracle.iou_distance(a, b), dets[1, :, 4])
    got = api.linear_assignment_arrays(cost, 0.8)
    ref = oracle.linear_assignment(cost, 0.8)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    rng = np.random.default_rng(5)
    dense = rng.random((96, 160)).astype(np.float32) * 0.7          # every pair is a candidate
    got = api.linear_assignment_arrays(dense, 0.8)
    ref = oracle.linear_assignment(dense, 0.8)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])


def test_lap_batch_device(oracle):
    rng = np.random.default_rng(2)
    P, n, m = 37, 48, 80
    costs = np.where(rng.random((P, n, m)) < 0.93, 1.0, rng.random((P, n, m))).astype(np.float32)
    nr = rng.integers(0, n + 1, P).astype(np.int32)
    nc = rng.integers(0, m + 1, P).astype(np.int32)
    dc, dnr, dnc = api.DeviceArray.from_host(costs), api.DeviceArray.from_host(nr), api.DeviceArray.from_host(nc)
    dr, dq = api.DeviceArray((P, n), np.int32), api.DeviceArray((P, m), np.int32)
    _lib.check(_lib.load().mot_lap_batch_device(dc.ptr, n * m, P, dnr.ptr, dnc.ptr, n, m, m, 0.8, dr.ptr, dq.ptr, None))
    r2c, c2r = dr.download(), dq.download()
    for p in range(P):
        ref = oracle.linear_assignment(costs[p, :nr[p], :nc[p]], 0.8)
        assert np.array_equal(r2c[p, :nr[p]], ref[0]) and np.array_equal(c2r[p, :nc[p]], ref[1]), p


def _tie_cost(rng, kind, n, m):
    """the generators of tests/test_oracle_kats.py::test_lap_oracle_equals_reference_solver: kinds 1 and 3 are heavy ties"""
    if kind == 0:
        c = rng.random((n, m))
    elif kind == 1:
        c = rng.integers(0, 4, (n, m)) / 4
    elif kind == 2:
        c = np.where(rng.random((n, m)) < 0.7, 1.0, rng.random((n, m)))
    elif kind == 3:
        c = rng.integers(0, 10, (n, m)) / 10
    else:
        c = -(rng.random((n, m)) * 1.2)
    th = -0.3 if kind == 4 else [0.5, 0.8, 0.7, 0.3][rng.integers(0, 4)]
    return c.astype(np.float32), float(th)


def _reference_lap(oracle, c, th):
    """the reference's REAL solver when its binary travelled (oracle/_ref/libref_lap.so), else the oracle's restatement"""
    return oracle.linear_assignment(c, th, use_ref=oracle.ref_lap() is not None)


def test_device_lapjv_reproduces_reference_ties_one_warp(oracle):
    """mot_lap_jv_batch_device, rows + columns <= 384 (jv_device.cuh): heavy-tie matrices, batched, against the
    reference's own lap_solver.hpp."""
    rng = np.random.default_rng(21)
    for trial in range(60):
        n, m = (int(v) for v in rng.integers(1, 40, 2)) if trial % 3 else (int(rng.integers(40, 190)), int(rng.integers(40, 190)))
        kind = trial % 5
        P = 5
        cs, th = [], None
        for _ in range(P):
            c, th0 = _tie_cost(rng, kind, n, m)
            th = th if th is not None else th0
            cs.append(c)
        costs = np.stack(cs)
        r2c, c2r = api.linear_assignment_reference_order(costs, th)
        for p in range(P):
            ref = _reference_lap(oracle, costs[p], th)
            assert np.array_equal(r2c[p], ref[0]) and np.array_equal(c2r[p], ref[1]), (trial, n, m, kind, p)


@pytest.mark.parametrize("n,m,kind", [(200, 300, 1), (300, 200, 3), (448, 256, 1), (256, 448, 2), (500, 700, 3), (640, 384, 4),
                                      (385, 1, 0), (1, 400, 1)])
def test_device_lapjv_reproduces_reference_ties_cta_wide(oracle, n, m, kind):
    """rows + columns > 384: the CTA-wide dense LAPJV (jv_block_device.cuh), at C2-like sizes, ties included."""
    rng = np.random.default_rng(1000 * n + m)
    cs = []
    c, th = _tie_cost(rng, kind, n, m)
    cs.append(c)
    cs.append(_tie_cost(rng, kind, n, m)[0])
    costs = np.stack(cs)
    r2c, c2r = api.linear_assignment_reference_order(costs, th)
    for p in range(2):
        ref = _reference_lap(oracle, costs[p], th)
        assert np.array_equal(r2c[p], ref[0]) and np.array_equal(c2r[p], ref[1]), (n, m, kind, p)


def test_device_lapjv_with_work_arrays_in_global_scratch(oracle, monkeypatch):
    """The CTA-wide LAPJV keeps its work arrays in shared memory when they fit (always, at these sizes, for the stand-alone
    entry point); the engines fall back to per-stream global scratch for big problems - force that path here."""
    monkeypatch.setenv("MOT_LAPJV_GLOBAL_WORK", "1")
    rng = np.random.default_rng(77)
    for n, m, kind in ((300, 200, 3), (256, 448, 1)):
        c, th = _tie_cost(rng, kind, n, m)
        r2c, c2r = api.linear_assignment_reference_order(c, th)
        ref = _reference_lap(oracle, c, th)
        assert np.array_equal(r2c, ref[0]) and np.array_equal(c2r, ref[1]), (n, m, kind)


def test_device_lapjv_on_a_c2_frame_with_duplicated_detections(oracle):
    """A real C2 cost matrix (256 tracks x 448 detections, 1 - IoU * conf) in which 40 detections appear twice: exactly
    tied optima at the headline size, resolved as the reference resolves them."""
    dets = synth.bytetrack_stream(5, n_frames=2)
    a = dets[0, :256, :4]
    b = dets[1, :448].copy()
    b[400:440] = b[100:140]                                   # bit-identical duplicates
    cost = oracle.fuse_score(oracle.iou_distance(a, b[:, :4]), b[:, 4])
    r2c, c2r = api.linear_assignment_reference_order(cost, 0.8)
    ref = _reference_lap(oracle, cost, 0.8)
    assert np.array_equal(r2c, ref[0]) and np.array_equal(c2r, ref[1])
    sparse = api.linear_assignment_arrays(cost, 0.8)          # the sparse solver: same matched SET sizes, ties may differ
    assert (sparse[0] >= 0).sum() == (ref[0] >= 0).sum()


# ------------------------------------------------------------------ cost matrices
@pytest.mark.parametrize("n,m", [(1, 1), (7, 5), (256, 512), (300, 1031), (2048, 2048)])
def test_iou_costs_bit_exact(oracle, n, m):
    rng = np.random.default_rng(n * 7 + m)
    def boxes(k):
        c = rng.uniform(0, 2000, (k, 2)); w = rng.uniform(20, 200, (k, 2))
        return np.concatenate([c - w / 2, c + w / 2], 1).astype(np.float32)
    a, b = boxes(n), boxes(m)
    conf = rng.uniform(0.1, 1.0, m).astype(np.float32)
    assert np.array_equal(api.iou_batch(a, b), oracle.iou_batch(a, b))
    d = oracle.iou_distance(a, b)
    assert np.array_equal(api.iou_distance(a, b), d)
    assert np.array_equal(api.iou_distance_fused(a, b, conf), oracle.fuse_score(d, conf))


def test_iou_reference_kats():
    # reference tests/test_iou.cpp:27-73
    assert api.iou_batch([[0, 0, 100, 100]], [[0, 0, 100, 100]])[0, 0] == 1.0
    assert api.iou_batch([[0, 0, 100, 100]], [[200, 200, 300, 300]])[0, 0] == 0.0
    assert api.iou_batch([[0, 0, 100, 100]], [[50, 50, 150, 150]])[0, 0] == np.float32(2500.0) / np.float32(17500.0)
    assert api.iou_batch(np.zeros((0, 4)), [[0, 0, 1, 1]]).shape == (0, 1)
    # degenerate (zero-area) boxes: union > 0 guard
    assert api.iou_batch([[5, 5, 5, 5]], [[5, 5, 5, 5]])[0, 0] == 0.0


# ------------------------------------------------------------------ Kalman filters
def _chain_states(oracle, kind, n, rng, steps=3):
    """Realistic (mean, cov) pairs: oracle initiate then a few predict/update rounds."""
    means, covs = [], []
    for _ in range(n):
        h = rng.uniform(40, 260)
        if kind == "xyah":
            z = np.array([rng.uniform(0, 1920), rng.uniform(0, 1080), rng.uniform(0.3, 0.6), h], np.float32)
            m, P = oracle.KFXYAH.initiate(z)
            for _ in range(int(rng.integers(0, steps + 1))):
                m, P = oracle.KFXYAH.predict(m, P)
                m, P, _ = oracle.KFXYAH.update(m, P, z + rng.normal(0, [2, 2, 0.01, 2]).astype(np.float32))
        elif kind == "xywh":
            z = np.array([rng.uniform(0, 1920), rng.uniform(0, 1080), h * 0.45, h], np.float32)
            m, P = oracle.KFXYWH.initiate(z)
            for _ in range(int(rng.integers(0, steps + 1))):
                m, P = oracle.KFXYWH.predict(m, P)
                m, P = oracle.KFXYWH.update(m, P, z + rng.normal(0, 2, 4).astype(np.float32))
        else:
            z = np.array([rng.uniform(0, 1920), rng.uniform(0, 1080), h * h * 0.45, 0.45], np.float32)
            m, P = oracle.KFXYSR.init(z)
            for _ in range(int(rng.integers(0, steps + 1))):
                m, P = oracle.KFXYSR.predict(m, P, 0.01, 0.0001)
                m, P, _ = oracle.KFXYSR.update(m, P, z + rng.normal(0, [2, 2, 50, 0.01]).astype(np.float32))
        means.append(m); covs.append(P)
    return np.stack(means), np.stack(covs)


def _assert_kf(got, want, what):
    assert np.array_equal(got, want) or np.allclose(got, want, rtol=KF_RTOL, atol=1e-6), what
    assert np.array_equal(got, want), f"{what}: within tolerance but not bit-identical to the oracle"


def test_kf_xyah(oracle):
    rng = np.random.default_rng(21)
    n = 203
    mean, cov = _chain_states(oracle, "xyah", n, rng)
    kf = api.KalmanFilterXYAH()
    z = (mean[:, :4] + rng.normal(0, [3, 3, 0.01, 3], (n, 4))).astype(np.float32)
    gm, gc = kf.initiate(z)
    for k in range(n):
        m, P = oracle.KFXYAH.initiate(z[k])
        _assert_kf(gm[k], m, "initiate mean"); _assert_kf(gc[k], P, "initiate cov")
    flags = (rng.random(n) < 0.4).astype(np.uint8)
    pm, pc = kf.predict(mean, cov, zero_vh=flags)
    for k in range(n):
        mk = mean[k].copy()
        if flags[k]:
            mk[7] = 0
        m, P = oracle.KFXYAH.predict(mk, cov[k])
        _assert_kf(pm[k], m, "predict mean"); _assert_kf(pc[k], P, "predict cov")
    conf = rng.uniform(0, 0.9, n).astype(np.float32)
    um, uc, fail = kf.update(pm, pc, z, confidence=conf, return_fail=True)
    assert not fail.any()
    for k in range(n):
        m, P, rc = oracle.KFXYAH.update(pm[k], pc[k], z[k], conf[k])
        assert rc == 0
        _assert_kf(um[k], m, "update mean"); _assert_kf(uc[k], P, "update cov")
    meas = (mean[:17, :4] + rng.normal(0, 4, (17, 4))).astype(np.float32)
    for only_pos in (False, True):
        for metric in ("maha", "gaussian"):
            g = kf.gating_distance(pm[:50], pc[:50], meas, only_pos, metric)
            for k in range(50):
                assert np.array_equal(g[k], oracle.KFXYAH.gating(pm[k], pc[k], meas, only_pos, metric))
    with pytest.raises(ValueError):
        kf.gating_distance(pm[:1], pc[:1], meas, False, "euclid")


def test_kf_xyah_non_spd_is_flagged_not_updated(oracle):
    kf = api.KalmanFilterXYAH()
    mean = np.array([[10, 10, 0.5, 50, 0, 0, 0, 0]], np.float32)
    cov = -np.eye(8, dtype=np.float32)[None] * 1e6              # S not positive definite
    um, uc, fail = kf.update(mean, cov, np.array([[11, 11, 0.5, 50]], np.float32), return_fail=True)
    assert fail[0] == 1 and np.array_equal(um, mean) and np.array_equal(uc, cov)
    assert oracle.KFXYAH.update(mean[0], cov[0], [11, 11, 0.5, 50])[2] == 1


def test_kf_xysr(oracle):
    rng = np.random.default_rng(22)
    n = 150
    mean, cov = _chain_states(oracle, "xysr", n, rng)
    kf = api.KalmanFilterXYSR()
    z = (mean[:, :4] + rng.normal(0, [3, 3, 80, 0.01], (n, 4))).astype(np.float32)
    gm, gc = kf.initiate(z)
    for k in range(n):
        m, P = oracle.KFXYSR.init(z[k])
        _assert_kf(gm[k], m, "init x"); _assert_kf(gc[k], P, "init P")
    for qxy, qs in ((1.0, 1.0), (0.01, 0.0001)):
        pm, pc = kf.predict(mean, cov, q_xy_scaling=qxy, q_s_scaling=qs)
        for k in range(n):
            m, P = oracle.KFXYSR.predict(mean[k], cov[k], qxy, qs)
            _assert_kf(pm[k], m, "predict x"); _assert_kf(pc[k], P, "predict P")
    um, uc = kf.update(pm, pc, z)
    for k in range(n):
        m, P, rc = oracle.KFXYSR.update(pm[k], pc[k], z[k])
        assert rc == 0
        _assert_kf(um[k], m, "update x"); _assert_kf(uc[k], P, "update P")
    # reference KATs (tests/test_kalman_filter.cpp:34-57)
    x0 = np.array([100, 100, 1000, .5, 10, 10, 0], np.float32)
    _, P0 = kf.initiate(np.zeros(4, np.float32))
    x1, P1 = kf.predict(x0, P0)
    assert np.array_equal(x1, [110, 110, 1000, .5, 10, 10, 0]) and P1[0, 0] == 1011.0
    x2, _ = kf.update(np.array([100, 100, 1000, .5, 0, 0, 0], np.float32), P0, [110, 110, 1100, .5])
    assert 100 < x2[0] < 110 and abs(x2[0] - (100 + 100 / 11)) < 1e-4


def test_kf_xywh(oracle):
    rng = np.random.default_rng(23)
    n = 150
    mean, cov = _chain_states(oracle, "xywh", n, rng)
    kf = api.KalmanFilterXYWH()
    z = (mean[:, :4] + rng.normal(0, 3, (n, 4))).astype(np.float32)
    gm, gc = kf.initiate(z)
    pm, pc = kf.predict(mean, cov)
    um, uc = kf.update(pm, pc, z)
    for k in range(n):
        m, P = oracle.KFXYWH.initiate(z[k])
        _assert_kf(gm[k], m, "initiate mean"); _assert_kf(gc[k], P, "initiate cov")
        m, P = oracle.KFXYWH.predict(mean[k], cov[k])
        _assert_kf(pm[k], m, "predict mean"); _assert_kf(pc[k], P, "predict cov")
        m, P = oracle.KFXYWH.update(pm[k], pc[k], z[k])
        _assert_kf(um[k], m, "update mean"); _assert_kf(uc[k], P, "update cov")
    meas = (mean[:9, :4] + rng.normal(0, 4, (9, 4))).astype(np.float32)
    for only_pos in (False, True):
        g = kf.gating_distance(pm[:40], pc[:40], meas, only_pos)
        for k in range(40):
            assert np.array_equal(g[k], oracle.KFXYWH.gating(pm[k], pc[k], meas, only_pos))


# ------------------------------------------------------------------ ByteTrack engine
BT_ARGS = dict(det_thresh=0.3, max_age=30, max_obs=50, min_hits=3, iou_threshold=0.3, min_conf=0.1,
               track_thresh=0.45, match_thresh=0.8, track_buffer=30, frame_rate=30)   # tools/motcpp_eval.cpp:133-148


def _oracle_bt(oracle):
    return oracle.ByteTrack(0.3, 30, 50, 3, 0.3, 0.1, 0.45, 0.8, 30, 30)


def _check_stream(oracle, eng, s, ref, dets, counts, out, n_out, t0, check_state_every=10):
    T = dets.shape[0]
    for t in range(T):
        want = ref.update(dets[t, :counts[t]])
        got = out[t, s, :n_out[t, s]]
        assert got.shape == want.shape, f"stream {s} frame {t0 + t}: {got.shape} vs {want.shape}"
        if not np.array_equal(got, want):
            bad = np.nonzero(~(got == want).all(1))[0]
            raise AssertionError(f"stream {s} frame {t0 + t}: rows {bad[:5]} differ\n{got[bad[:3]]}\n{want[bad[:3]]}")


@pytest.mark.parametrize("sid", [0, 1, 2, 3])
def test_bytetrack_stress_stream_frame_by_frame(oracle, sid):
    dets, counts = synth.stress_stream(sid, n_frames=300)
    eng = api.Engine(_lib.TRACKER_BYTETRACK, 1, 256, dets.shape[1], **BT_ARGS)
    ref = _oracle_bt(oracle)
    for t in range(dets.shape[0]):
        out, n_out = eng.update(dets[t][None], counts[t:t + 1], ld_out=256)
        want = ref.update(dets[t, :counts[t]])
        got = out[0, :n_out[0]]
        assert got.shape == want.shape and np.array_equal(got, want), f"frame {t}"
        if t % 25 == 0 or t == dets.shape[0] - 1:
            for which in (0, 1):
                g, w = eng.dump(0, which), ref.dump(which)
                assert g.shape == w.shape, (t, which)
                assert np.allclose(g[:, 6:], w[:, 6:], rtol=KF_RTOL, atol=1e-6), "Kalman state beyond 1e-4 rel"
                assert np.array_equal(g, w), f"frame {t} list {which}: state not bit-identical"
    eng.check()
    eng.close()


def test_bytetrack_headline_config_sequence(oracle):
    """C2 (256 objects, 512 detections/frame): 4 streams x 80 frames in ONE call vs 4 oracles."""
    S, T = 4, 80
    dets = np.stack([synth.bytetrack_stream(s, n_frames=T) for s in range(S)], 1)     # (T,S,512,6)
    counts = np.full((T, S), 512, np.int32)
    eng = api.Engine(_lib.TRACKER_BYTETRACK, S, 1536, 512, **BT_ARGS)
    out, n_out = eng.update(dets, counts, ld_out=1024)
    eng.check()
    for s in range(S):
        ref = _oracle_bt(oracle)
        _check_stream(oracle, eng, s, ref, dets[:, s], counts[:, s], out, n_out, 0)
        for which in (0, 1):
            assert np.array_equal(eng.dump(s, which), ref.dump(which))
        hdr = eng.header(s)
        assert list(hdr[6:12]) == list(ref.last_sizes()[:6])        # the three LAP sub-problem sizes
    eng.close()


def test_bytetrack_chunked_host_path_equals_single_launch(oracle):
    """n_chunks pipelining and frame-by-frame calls are pure scheduling: identical outputs."""
    S, T = 9, 30
    streams = [synth.stress_stream(40 + s, n_frames=T) for s in range(S)]
    dets = np.stack([st[0] for st in streams], 1)
    counts = np.stack([st[1] for st in streams], 1).astype(np.int32)
    ld = dets.shape[2]
    e1 = api.Engine(_lib.TRACKER_BYTETRACK, S, 256, ld, n_chunks=1, **BT_ARGS)
    e4 = api.Engine(_lib.TRACKER_BYTETRACK, S, 256, ld, n_chunks=4, **BT_ARGS)
    o1, n1 = e1.update(dets, counts, ld_out=128)
    outs, ns = [], []
    for t in range(T):
        o, n = e4.update(dets[t], counts[t], ld_out=128)
        outs.append(o.copy()); ns.append(n.copy())
    o4, n4 = np.stack(outs), np.stack(ns)
    assert np.array_equal(n1, n4)
    for t in range(T):
        for s in range(S):
            assert np.array_equal(o1[t, s, :n1[t, s]], o4[t, s, :n4[t, s]])
    ref = _oracle_bt(oracle)
    _check_stream(oracle, e1, 5, ref, dets[:, 5], counts[:, 5], o1, n1, 0)
    e1.close(); e4.close()


def test_packed_host_path_equals_padded_path(oracle):
    """mot_engine_update_host_packed: the same rows as the padded call, back to back, plus exclusive offsets - for many
    frames per call (frame-chunk pipeline), a single frame, and an engine of every box-only tracker kind."""
    S, T = 7, 41
    streams = [synth.stress_stream(60 + s, n_frames=T) for s in range(S)]
    dets = np.stack([st[0] for st in streams], 1)
    counts = np.stack([st[1] for st in streams], 1).astype(np.int32)
    ld = dets.shape[2]
    for kind, kw in ((_lib.TRACKER_BYTETRACK, BT_ARGS), (_lib.TRACKER_SORT, dict(det_thresh=0.3, max_age=3, min_hits=1, iou_threshold=0.3)),
                     (_lib.TRACKER_OCSORT, dict(det_thresh=0.2, max_age=30, min_hits=3, iou_threshold=0.3, min_conf=0.1, delta_t=3,
                                                inertia=0.2, use_byte=0, q_xy_scaling=0.01, q_s_scaling=0.0001))):
        ea = api.Engine(kind, S, 256, ld, **kw)
        eb = api.Engine(kind, S, 256, ld, **kw)
        o, n = ea.update(dets, counts, ld_out=256)
        ec = api.Engine(kind, S, 256, ld, **kw)
        rows_p, off_p, n_p = ec.update_packed(dets[:30], counts[:30], max_rows=256, pinned=True)     # pinned result buffers
        rows, off, n2 = eb.update_packed(dets[:30], counts[:30], max_rows=256)                         # pageable result buffers
        assert np.array_equal(rows_p, rows) and np.array_equal(off_p, off) and np.array_equal(n_p, n2)
        assert np.array_equal(n2, n[:30]) and off[0] == 0 and off[-1] == n[:30].sum() == len(rows)
        assert np.array_equal(np.diff(off), n[:30].reshape(-1))
        for t in range(30):
            for s_ in range(S):
                f = t * S + s_
                assert np.array_equal(rows[off[f]:off[f + 1]], o[t, s_, :n[t, s_]]), (kind, t, s_)
        for t in range(30, T):                                    # one frame per call
            rows, off, n2 = eb.update_packed(dets[t:t + 1], counts[t:t + 1], max_rows=256)
            for s_ in range(S):
                assert np.array_equal(rows[off[s_]:off[s_ + 1]], o[t, s_, :n[t, s_]]), (kind, t, s_)
        ea.check(); eb.check(); ec.check()
        with pytest.raises(ValueError, match="out_rows holds"):
            eb.update_packed(dets[:4], counts[:4], max_rows=256, out_rows=np.empty((3, 8), np.float32))
        with pytest.raises(ValueError, match="out_rows holds"):
            ec.update_packed(dets[:4], counts[:4], max_rows=256, out_rows=api.pinned_empty((3, 8), np.float32), pinned=True)
        ea.close(); eb.close(); ec.close()


def test_engine_rejects_counts_beyond_the_leading_dimension_and_reports_flags_once():
    """n_dets > ld_dets used to read the next stream's rows; the Python mirror refuses it, the kernels clamp and
    flag it.  Error bits are read-and-clear: reported by the check that follows, not by every later one."""
    dets, counts = synth.stress_stream(4, n_frames=4)
    eng = api.Engine(_lib.TRACKER_BYTETRACK, 2, 256, 64, **BT_ARGS)
    d2 = np.stack([dets[:, :32], dets[:, :32]], 1)                 # ld_dets = 32 < the 64 the shape allows
    with pytest.raises(ValueError, match="n_dets"):
        eng.update(d2, np.full((4, 2), 40, np.int32), ld_out=64)
    dd, nn = api.DeviceArray.from_host(d2), api.DeviceArray.from_host(np.full((4, 2), 40, np.int32))
    do, dn = api.DeviceArray((4, 2, 64, 8), np.float32), api.DeviceArray((4, 2), np.int32)
    eng.update_device(4, dd.ptr, nn.ptr, 32, do.ptr, dn.ptr, 64)
    with pytest.raises(RuntimeError, match="too many detections|detections"):
        eng.check()
    eng.check()                                                   # cleared by the failing check
    eng.update(d2, np.full((4, 2), 32, np.int32), ld_out=64)
    eng.check()
    eng.close()


def test_reid_wrappers_size_themselves_from_the_first_embeddings(oracle):
    """a default-constructed BotSort / StrongSort must not drop `embs` silently - it adopts their dimension on
    the first frame (the reference uses whatever it is handed) and matches the oracle that uses them."""
    d, c, e = synth.stress_stream_reid(3, n_frames=25, dim=32)
    trk, ref = api.BotSort(track_capacity=256, max_dets=64), oracle.BotSort()
    for t in range(25):
        assert np.array_equal(trk.update(d[t, :c[t]], (540, 960), e[t, :c[t]]), ref.update(d[t, :c[t]], e[t, :c[t]])), t
    with pytest.raises(ValueError, match="emb_dim"):
        trk.update(d[0, :c[0]], (540, 960), e[0, :c[0], :16])
    trk, ref = api.StrongSort(track_capacity=256, max_dets=64), oracle.StrongSort(tie_mode=0)
    for t in range(25):
        assert np.array_equal(trk.update(d[t, :c[t]], (540, 960), e[t, :c[t]]), ref.update(d[t, :c[t]], e[t, :c[t]])), t


def test_bytetrack_reset_keeps_id_counter(oracle):
    # reference: ByteTrack::reset clears lists but STrack::clear_count is empty (bytetrack.hpp:38-40)
    dets, counts = synth.stress_stream(4, n_frames=20)
    eng = api.Engine(_lib.TRACKER_BYTETRACK, 1, 256, dets.shape[1], **BT_ARGS)
    ref = _oracle_bt(oracle)
    for t in range(10):
        eng.update(dets[t][None], counts[t:t + 1], ld_out=256)
        ref.update(dets[t, :counts[t]])
    eng.reset(); ref.reset()
    assert eng.header(0)[0] == 0 and eng.header(0)[4] == 0 and eng.header(0)[3] > 0
    for t in range(10, 20):
        out, n_out = eng.update(dets[t][None], counts[t:t + 1], ld_out=256)
        want = ref.update(dets[t, :counts[t]])
        assert np.array_equal(out[0, :n_out[0]], want)
    eng.close()


def test_bytetrack_capacity_flag_is_loud():
    # every frame: 32 fresh boxes + last frame's 32 again (which confirms them).  With no low-score
    # detections confirmed tracks are never marked lost (bytetrack.cpp:387), so they pile up: the 256
    # slots of the smallest shape are gone after ~8 frames and the engine must say so.
    rng = np.random.default_rng(0)
    eng = api.Engine(_lib.TRACKER_BYTETRACK, 1, 256, 64, **BT_ARGS)

    def boxes(t):
        c = np.random.default_rng(100 + t).uniform(0, 4000, (32, 2)) + 10000 * t
        return np.concatenate([c, c + [50, 110], np.full((32, 1), 0.9), np.zeros((32, 1))], 1).astype(np.float32)

    for t in range(12):
        dets = np.concatenate([boxes(t), boxes(t - 1)], 0)
        eng.update(dets[None], np.array([64], np.int32), ld_out=256)
    with pytest.raises(RuntimeError, match="capacity"):
        eng.check()
    eng.close()


def test_bytetrack_facade_mirrors_reference_api(oracle):
    """motcpp::trackers::ByteTrack(...).update(dets, img) incl. check_inputs' exceptions."""
    trk = api.ByteTrack(0.3, 30, 50, 3, 0.3, False, 80, "iou", False, 0.1, 0.45, 0.8, 30, 30, track_capacity=256, max_dets=64)
    ref = _oracle_bt(oracle)
    img = np.zeros((480, 640, 3), np.uint8)
    dets, counts = synth.stress_stream(5, n_frames=40)
    for t in range(40):
        got = trk.update(dets[t, :counts[t]], img)
        assert np.array_equal(got, ref.update(dets[t, :counts[t]]))
        assert got.shape[1] == 8
    with pytest.raises(ValueError, match="6 .AABB. or 7"):
        trk.update(np.zeros((2, 5), np.float32), img)                 # src/tracker.cpp:110-112
    with pytest.raises(ValueError, match="Image cannot be empty"):
        trk.update(dets[0, :3], np.zeros((0, 0, 3), np.uint8))        # src/tracker.cpp:114-116
    with pytest.raises(ValueError, match="same number of rows"):
        trk.update(dets[0, :3], img, embs=np.zeros((2, 8), np.float32))
    assert trk.update(np.zeros((0, 6), np.float32), img).shape[1] == 8
    # reference tests/test_trackers.cpp:52-80: ids persist over identical frames
    t2 = api.ByteTrack(track_capacity=64, max_dets=16)
    two = np.array([[100, 100, 200, 200, 0.9, 0], [300, 300, 400, 420, 0.8, 0]], np.float32)
    ids = [sorted(t2.update(two, img)[:, 4].astype(int)) for _ in range(3)]
    assert ids[0] == ids[1] == ids[2] and len(ids[0]) == 2


def test_bytetrack_full_size_properties():
    """BASELINE configs[1] at bench scale through size-independent properties: batching idempotence
    (one 40-frame call == 40 one-frame calls), unique IDs per frame, det_ind points at a detection
    whose confidence is the reported one, IDs never reused."""
    S, T = 64, 40
    base = [synth.bytetrack_stream(s, n_frames=T) for s in range(4)]
    dets = np.stack([base[s % 4][:, np.random.default_rng(s).permutation(512)] for s in range(S)], 1)
    counts = np.full((T, S), 512, np.int32)
    a = api.Engine(_lib.TRACKER_BYTETRACK, S, 1536, 512, **BT_ARGS)
    b = api.Engine(_lib.TRACKER_BYTETRACK, S, 1536, 512, **BT_ARGS)
    oa, na = a.update(dets, counts, ld_out=768)
    seen_max = np.zeros(S, np.int64)
    for t in range(T):
        ob, nb = b.update(dets[t], counts[t], ld_out=768)
        assert np.array_equal(nb, na[t])
        for s in range(S):
            rows = oa[t, s, :na[t, s]]
            assert np.array_equal(rows, ob[s, :nb[s]])
            ids = rows[:, 4].astype(np.int64)
            assert len(np.unique(ids)) == len(ids)
            di = rows[:, 7].astype(np.int64)
            assert di.min() >= 0 and di.max() < 512
            fresh = rows[:, 5] == dets[t, s, di, 4]                  # tracks updated this frame carry their det's score
            assert fresh.mean() > 0.5
            seen_max[s] = max(seen_max[s], ids.max())
    a.check(); b.check()
    hdr = a.header(0)
    assert hdr[3] >= seen_max[0] and hdr[4] == T
    a.close(); b.close()


def test_bytetrack_both_cta_widths_agree_with_the_oracle(oracle):
    """The C2 shape has two kernels: 512 threads x 2 CTAs per SM when the streams fill the machine (the bench), one
    1024-thread CTA per SM when there are fewer streams than SMs (BASELINE configs[4]).  160 streams select the first,
    8 the second; both must equal the oracle (4 distinct streams, the others are row permutations checked against each other
    through their first copy)."""
    T = 45
    base = [synth.bytetrack_stream(20 + s, n_frames=T) for s in range(4)]
    refs = []
    for s in range(4):
        r = _oracle_bt(oracle)
        refs.append([r.update(base[s][t]) for t in range(T)])
    for S in (160, 8):
        dets = np.stack([base[s % 4] for s in range(S)], 1)
        counts = np.full((T, S), 512, np.int32)
        eng = api.Engine(_lib.TRACKER_BYTETRACK, S, 1536, 512, **BT_ARGS)
        assert eng.info()["threads_per_cta"] == (512 if S > 148 else 1024)
        out, n_out = eng.update(dets, counts, ld_out=640)
        eng.check()
        for s in range(S):
            for t in range(T):
                assert np.array_equal(out[t, s, :n_out[t, s]], refs[s % 4][t]), (S, s, t)
        eng.close()


# ------------------------------------------------------------------ cosine embedding cost (tcgen05)
COSINE_ATOL = 2e-5     # 3-term bf16 split + fp32 tensor-core accumulation vs the oracle's sequential fp32 sum


@pytest.mark.parametrize("n,m,dim", [(1, 1, 8), (7, 5, 33), (128, 64, 64), (130, 70, 128), (300, 517, 512), (1024, 1024, 512)])
def test_cosine_cost_matches_oracle(oracle, n, m, dim):
    rng = np.random.default_rng(n + 3 * m + dim)
    t = rng.normal(0, 1, (n, dim)).astype(np.float32)
    d = rng.normal(0, 1, (m, dim)).astype(np.float32)
    if n > 4:                                   # some near-duplicates (cost ~ 0) and an exact copy (cost clamps at 0)
        d[0] = t[1] + 0.05 * rng.normal(0, 1, dim).astype(np.float32)
        d[m - 1] = 3.0 * t[2]
    got = api.embedding_distance(t, d)
    want = oracle.embedding_distance(t, d)
    assert got.shape == want.shape
    assert np.all(got >= 0.0)
    np.testing.assert_allclose(got, want, atol=COSINE_ATOL, rtol=0)
    if n > 4:
        assert got[2, m - 1] <= COSINE_ATOL


@pytest.mark.parametrize("n,m,dim", [(2048, 2100, 64), (4096, 4800, 32), (5000, 9500, 16)])
def test_cosine_cost_large_tiles(n, m, dim):
    """Sizes that select the 128 x 128 and 128 x 256 persistent-tile variants (several tiles per CTA, both TMEM
    accumulator stages in use, ragged right / bottom edges).  Checked against a float64 numpy evaluation of the
    same formula (the C oracle is the same arithmetic in fp32; it would take minutes at these sizes)."""
    rng = np.random.default_rng(n + m + dim)
    t = rng.normal(0, 1, (n, dim)).astype(np.float32)
    d = rng.normal(0, 1, (m, dim)).astype(np.float32)
    got = api.embedding_distance(t, d)
    t64, d64 = t.astype(np.float64), d.astype(np.float64)
    sim = (t64 @ d64.T) / (np.linalg.norm(t64, axis=1)[:, None] * np.linalg.norm(d64, axis=1)[None] + 1e-10)
    np.testing.assert_allclose(got, np.maximum(0.0, 1.0 - sim), atol=COSINE_ATOL, rtol=0)


def test_cosine_botsort_embeddings(oracle):
    """C3-style inputs: unit-norm identity embeddings, detections = identity + noise."""
    dets, embs = synth.embeddings_stream(0, n_frames=2, n_obj=256, dim=512)
    got = api.embedding_distance(embs[0], embs[1])
    want = oracle.embedding_distance(embs[0], embs[1])
    np.testing.assert_allclose(got, want, atol=COSINE_ATOL, rtol=0)
    # the matching a downstream LAP would make is unchanged by the tolerance
    assert np.array_equal(np.argmin(got, 1), np.argmin(want, 1))


# ------------------------------------------------------------------ SORT engine
@pytest.mark.parametrize("sid,args", [(0, (0.3, 1, 3, 0.3)), (1, (0.3, 3, 1, 0.3)), (2, (0.5, 30, 3, 0.2))])
def test_sort_stress_streams(oracle, sid, args):
    dets, counts = synth.stress_stream(20 + sid, n_frames=250)
    det_thresh, max_age, min_hits, iou_thr = args
    eng = api.Engine(_lib.TRACKER_SORT, 1, 256, dets.shape[1], det_thresh=det_thresh, max_age=max_age,
                     min_hits=min_hits, iou_threshold=iou_thr)
    ref = oracle.Sort(det_thresh, max_age, 50, min_hits, iou_thr)
    out, n_out = eng.update(dets[:, None], counts[:, None], ld_out=256)        # all frames in one launch
    eng.check()
    for t in range(dets.shape[0]):
        want = ref.update(dets[t, :counts[t]])
        got = out[t, 0, :n_out[t, 0]]
        assert got.shape == want.shape and np.array_equal(got, want), f"frame {t}"
    eng.close()


def test_sort_reference_kats():
    # reference tests/test_sort.cpp:50-68, :70-85, :126-148
    det = np.array([[100, 100, 200, 200, 0.9, 0]], np.float32)
    s = api.Sort(0.3, 3, 50, 1)
    s.update(det); s.update(det)
    out = s.update(np.array([[110, 110, 210, 210, 0.9, 0]], np.float32))
    assert out.shape == (1, 8) and int(out[0, 4]) == 1 and out[0, 2] > out[0, 0] and out[0, 3] > out[0, 1]
    s = api.Sort(0.3, 2, 50, 1)
    s.update(det); s.update(np.zeros((0, 6), np.float32))
    assert s.update(np.zeros((0, 6), np.float32)).shape[0] == 0
    s = api.Sort(0.3, 3, 50, 1)
    s.update(det); s.update(np.zeros((0, 6), np.float32))
    out = s.update(det)
    assert out.shape[0] == 1 and int(out[0, 4]) == 1


@pytest.mark.parametrize("seq", ["MOT17_02_FRCNN", "MOT17_04_FRCNN"])
def test_mot17_mini_sort_and_bytetrack(oracle, seq):
    """BASELINE configs[0]: SORT on assets/MOT17-mini (committed fixture), every frame in order, plus
    ByteTrack on the same detections; the whole sequence runs as one launch and must equal the oracle."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "mot17_mini_dets.npz"))
    frames, dets = z[seq + "_frames"], z[seq + "_dets"]
    f0, f1 = int(frames.min()), int(frames.max())
    T = f1 - f0 + 1
    batch = np.zeros((T, 1, 64, 6), np.float32)
    counts = np.zeros((T, 1), np.int32)
    for t in range(T):
        d = dets[frames == f0 + t]
        batch[t, 0, :len(d)] = d
        counts[t, 0] = len(d)
    for kind, ref, kw in ((_lib.TRACKER_SORT, oracle.Sort(0.3, 1, 50, 3, 0.3),
                           dict(det_thresh=0.3, max_age=1, max_obs=50, min_hits=3, iou_threshold=0.3)),
                          (_lib.TRACKER_BYTETRACK, _oracle_bt(oracle), BT_ARGS)):
        eng = api.Engine(kind, 1, 256, 64, **kw)
        out, n_out = eng.update(batch, counts, ld_out=256)
        eng.check()
        total = 0
        for t in range(T):
            want = ref.update(batch[t, 0, :counts[t, 0]])
            got = out[t, 0, :n_out[t, 0]]
            assert got.shape == want.shape and np.array_equal(got, want), f"{seq} kind {kind} frame {f0 + t}"
            total += len(want)
        assert total == int(z[f"{seq}_{'sort' if kind == _lib.TRACKER_SORT else 'bytetrack'}_digest"][1])
        eng.close()
