"""Pins the CPU oracle: known-answer tests taken from the reference's own test fixtures
(tests/test_iou.cpp, tests/test_matching.cpp, tests/test_kalman_filter.cpp, tests/test_sort.cpp,
tests/test_trackers.cpp of motcpp) plus the hand-derived KATs of SURVEY.md section 8c, and the
reference's REAL LAP solver (oracle/_ref/libref_lap.so, built in place from /root/reference)."""
import numpy as np
import pytest


# ---------------------------------------------------------------- IoU (reference tests/test_iou.cpp:27-73)
def test_iou_kats(oracle):
    b1 = [[0, 0, 100, 100]]
    b2 = [[50, 50, 150, 150]]
    b3 = [[200, 200, 300, 300]]
    assert oracle.iou_batch(b1, b1)[0, 0] == np.float32(1.0)
    assert oracle.iou_batch(b1, b3)[0, 0] == np.float32(0.0)
    assert oracle.iou_batch(b1, b2)[0, 0] == np.float32(np.float32(2500.0) / np.float32(17500.0))
    m = oracle.iou_batch([[0, 0, 100, 100], [50, 50, 150, 150]], [[0, 0, 100, 100], [200, 200, 300, 300]])
    assert m.shape == (2, 2) and m[0, 0] == 1.0 and m[0, 1] == 0.0
    assert oracle.iou_batch(np.zeros((0, 4)), b1).shape == (0, 1)


def test_iou_distance_and_fuse(oracle):
    a = np.array([[0, 0, 100, 100], [10, 10, 60, 90]], np.float32)
    b = np.array([[50, 50, 150, 150], [0, 0, 100, 100], [500, 500, 600, 600]], np.float32)
    iou = oracle.iou_batch(a, b)
    d = oracle.iou_distance(a, b)
    assert np.array_equal(d, np.float32(1.0) - iou)
    conf = np.array([0.9, 0.5, 0.7], np.float32)
    f = oracle.fuse_score(d, conf)
    expect = np.float32(1.0) - (np.float32(1.0) - d) * conf[None, :]
    assert np.array_equal(f, expect.astype(np.float32))
    assert f[0, 2] == 1.0          # disjoint pair stays at exactly 1


# ---------------------------------------------------------------- LAP (reference tests/test_matching.cpp:14-110)
def _sets(r2c, c2r):
    matches = {(i, int(j)) for i, j in enumerate(r2c) if j >= 0}
    ua = [i for i, j in enumerate(r2c) if j < 0]
    ub = [j for j, i in enumerate(c2r) if i < 0]
    return matches, ua, ub


@pytest.mark.parametrize("use_ref", [False, True])
def test_lap_kats(oracle, use_ref):
    if use_ref and oracle.ref_lap() is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    la = lambda c, t: _sets(*oracle.linear_assignment(np.array(c, np.float32), t, use_ref=use_ref))
    assert la(np.zeros((0, 0)), 0.5) == (set(), [], [])
    assert la([[0.1]], 0.5) == ({(0, 0)}, [], [])
    assert la([[0.9]], 0.5) == (set(), [0], [0])
    assert la([[.1, .9, .9], [.9, .1, .9], [.9, .9, .1]], 0.5) == ({(0, 0), (1, 1), (2, 2)}, [], [])
    assert la([[.1, .9], [.9, .1], [.9, .9]], 0.5) == ({(0, 0), (1, 1)}, [2], [])
    assert la([[.1, .9, .9], [.9, .1, .9]], 0.5) == ({(0, 0), (1, 1)}, [], [2])
    assert la([[.1, .2], [.3, .1]], 0.5) == ({(0, 0), (1, 1)}, [], [])


def test_lap_oracle_equals_reference_solver(oracle):
    """The restated LAPJV must reproduce the reference's real lap_solver.hpp bit for bit, ties included."""
    if oracle.ref_lap() is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(0)
    for trial in range(3000):
        n, m = rng.integers(1, 14, 2)
        kind = trial % 5
        if kind == 0:
            c = rng.random((n, m))
        elif kind == 1:
            c = rng.integers(0, 4, (n, m)) / 4          # heavy ties
        elif kind == 2:
            c = np.where(rng.random((n, m)) < 0.7, 1.0, rng.random((n, m)))
        elif kind == 3:
            c = rng.integers(0, 10, (n, m)) / 10
        else:
            c = -(rng.random((n, m)) * 1.2)             # OC-SORT style negative costs
        th = -0.3 if kind == 4 else [0.5, 0.8, 0.7, 0.3][rng.integers(0, 4)]
        c = c.astype(np.float32)
        a = oracle.linear_assignment(c, th)
        b = oracle.linear_assignment(c, th, use_ref=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (trial, c, th)
    for trial in range(30):
        n, m = rng.integers(40, 160, 2)
        c = np.where(rng.random((n, m)) < 0.93, 1.0, rng.random((n, m))).astype(np.float32)
        a = oracle.linear_assignment(c, 0.8)
        b = oracle.linear_assignment(c, 0.8, use_ref=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_lap_matches_scipy_objective(oracle):
    """Independent cross-check of the 'minimise sum(c - thresh) over partial matchings' reading."""
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(3)
    for _ in range(200):
        n, m = rng.integers(1, 20, 2)
        c = rng.random((n, m)).astype(np.float32)
        th = np.float32(0.6)
        r2c, _ = oracle.linear_assignment(c, th)
        got = sum(float(c[i, j]) - float(th) for i, j in enumerate(r2c) if j >= 0)
        ext = np.zeros((n, m + n))
        ext[:, :m] = c.astype(np.float64) - float(th)
        ext[:, m:] = 1e9
        ext[np.arange(n), m + np.arange(n)] = 0.0
        ri, ci = linear_sum_assignment(ext)
        best = ext[ri, ci].sum()
        assert abs(got - best) < 1e-9


# ---------------------------------------------------------------- Kalman filters
def test_xysr_kats(oracle):
    # reference tests/test_kalman_filter.cpp:34-57 and SURVEY.md 8c
    x = np.array([100, 100, 1000, .5, 10, 10, 0], np.float32)
    _, P0 = oracle.KFXYSR.init([0, 0, 0, 0])
    assert np.array_equal(np.diag(P0), np.array([10, 10, 10, 10, 1000, 1000, 1000], np.float32))
    x1, P1 = oracle.KFXYSR.predict(x, P0)
    assert np.array_equal(x1, np.array([110, 110, 1000, .5, 10, 10, 0], np.float32))
    assert P1[0, 0] == 1011.0
    x = np.array([100, 100, 1000, .5, 0, 0, 0], np.float32)
    x2, P2, rc = oracle.KFXYSR.update(x, P0, [110, 110, 1100, .5])
    assert rc == 0
    assert abs(x2[0] - (100 + 10 * 10 / 11)) < 1e-4 and abs(x2[1] - x2[0]) == 0
    assert abs(x2[2] - 1050) < 1e-3 and x2[3] == 0.5
    assert 100 < x2[0] < 110


def test_xyah_kats(oracle):
    # initiate z=[cx,cy,a,100]
    m, P = oracle.KFXYAH.initiate([50, 60, 0.5, 100])
    assert np.array_equal(m, np.array([50, 60, 0.5, 100, 0, 0, 0, 0], np.float32))
    d = np.diag(P)
    np.testing.assert_allclose(d, [100, 100, 1e-4, 100, 39.0625, 39.0625, 1e-10, 39.0625], rtol=1e-6)
    m1, P1 = oracle.KFXYAH.predict(m, P)
    np.testing.assert_allclose(P1[0, 0], 100 + 39.0625 + 25, rtol=1e-6)
    np.testing.assert_allclose(P1[0, 4], 39.0625, rtol=1e-6)
    np.testing.assert_allclose(P1[4, 4], 39.0625 + 0.390625, rtol=1e-6)
    # update against a float64 numpy re-derivation (parity unpinned in the reference: no golden values)
    z = np.array([52, 61, 0.52, 101], np.float32)
    m2, P2, rc = oracle.KFXYAH.update(m1, P1, z)
    assert rc == 0
    H = np.eye(4, 8)
    h = float(m1[3])
    R = np.diag([(h / 20) ** 2, (h / 20) ** 2, 1e-2, (h / 20) ** 2])
    Pd = P1.astype(np.float64)
    S = H @ Pd @ H.T + R
    K = Pd @ H.T @ np.linalg.inv(S)
    np.testing.assert_allclose(m2, m1 + K @ (z - H @ m1), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(P2, Pd - K @ S @ K.T, rtol=1e-4, atol=1e-4)


def test_xyah_gating_is_s_inverse_squared(oracle):
    """The reference's "maha" gating returns d^T S^-2 d (kalman_filter.cpp:166-172)."""
    m, P = oracle.KFXYAH.initiate([50, 60, 0.5, 100])
    m, P = oracle.KFXYAH.predict(m, P)
    meas = np.array([[55, 58, 0.5, 104], [50, 60, 0.5, 100]], np.float32)
    g = oracle.KFXYAH.gating(m, P, meas)
    h = float(m[3])
    S = P[:4, :4].astype(np.float64) + np.diag([(h / 20) ** 2, (h / 20) ** 2, 1e-2, (h / 20) ** 2])
    for k in range(2):
        d = meas[k].astype(np.float64) - m[:4]
        zz = np.linalg.solve(S, d)
        np.testing.assert_allclose(g[k], zz @ zz, rtol=1e-4, atol=1e-9)
    g2 = oracle.KFXYAH.gating(m, P, meas, metric="gaussian")
    np.testing.assert_allclose(g2[0], np.sum((meas[0] - m[:4]) ** 2), rtol=1e-6)


def test_xywh_against_numpy(oracle):
    m, P = oracle.KFXYWH.initiate([50, 60, 40, 100])
    np.testing.assert_allclose(np.diag(P), [100] * 4 + [39.0625] * 4, rtol=1e-6)
    m1, P1 = oracle.KFXYWH.predict(m, P)
    z = np.array([52, 61, 41, 101], np.float32)
    m2, P2 = oracle.KFXYWH.update(m1, P1, z)
    H = np.eye(4, 8)
    h = float(m1[3])
    Pd = P1.astype(np.float64)
    S = H @ Pd @ H.T + np.eye(4) * (h / 20) ** 2
    K = Pd @ H.T @ np.linalg.inv(S)
    np.testing.assert_allclose(m2, m1 + K @ (z - H @ m1), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(P2, Pd - K @ S @ K.T, rtol=1e-4, atol=1e-4)
    g = oracle.KFXYWH.gating(m1, P1, z[None])
    d = z - m1[:4]
    np.testing.assert_allclose(g[0], d @ np.linalg.inv(S) @ d, rtol=1e-4)
    gp = oracle.KFXYWH.gating(m1, P1, z[None], only_position=True)
    np.testing.assert_allclose(gp[0], d[:2] @ np.linalg.inv(S)[:2, :2] @ d[:2], rtol=1e-4)


# ---------------------------------------------------------------- trackers
def test_sort_kats(oracle):
    # reference tests/test_sort.cpp:50-68 (id 1 persists), :70-85 (deleted after max_age)
    det = np.array([[100, 100, 200, 200, 0.9, 0]], np.float32)
    s = oracle.Sort(0.3, 3, 50, 1)
    s.update(det)
    s.update(det)
    out = s.update(np.array([[110, 110, 210, 210, 0.9, 0]], np.float32))
    assert out.shape == (1, 8) and int(out[0, 4]) == 1
    assert out[0, 2] > out[0, 0] and out[0, 3] > out[0, 1]
    s = oracle.Sort(0.3, 2, 50, 1)
    s.update(det)
    s.update(np.zeros((0, 6), np.float32))
    assert s.update(np.zeros((0, 6), np.float32)).shape[0] == 0
    # :126-148 one missed frame keeps the id
    s = oracle.Sort(0.3, 3, 50, 1)
    s.update(det)
    s.update(np.zeros((0, 6), np.float32))
    out = s.update(det)
    assert out.shape[0] == 1 and int(out[0, 4]) == 1


def test_bytetrack_kats(oracle):
    # reference tests/test_trackers.cpp:52-80 (ID persistence over identical frames), :98-105 (empty -> 0 rows)
    bt = oracle.ByteTrack()
    dets = np.array([[100, 100, 200, 200, 0.9, 0], [300, 300, 400, 420, 0.8, 0]], np.float32)
    ids = []
    for _ in range(3):
        out = bt.update(dets)
        assert out.shape == (2, 8)
        ids.append(sorted(out[:, 4].astype(int)))
    assert ids[0] == ids[1] == ids[2] == [1, 2]
    assert oracle.ByteTrack().update(np.zeros((0, 6), np.float32)).shape == (0, 8)
    # a detection AT track_thresh belongs to neither set (strict comparisons, bytetrack.cpp:190-193)
    bt = oracle.ByteTrack(track_thresh=0.5)
    assert bt.update(np.array([[0, 0, 10, 20, 0.5, 0]], np.float32)).shape[0] == 0


# ---------------------------------------------------------------- golden fixture (BASELINE configs[0])
def test_mot17_mini_fixture_digests(oracle):
    """tests/golden/mot17_mini_dets.npz: the reference's MOT17-mini detections + digests of the oracle's
    SORT / ByteTrack outputs recorded when the fixture was made (make_mot17_mini_fixture.py)."""
    import hashlib
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "mot17_mini_dets.npz"))
    for seq in ("MOT17_02_FRCNN", "MOT17_04_FRCNN"):
        frames, dets = z[seq + "_frames"], z[seq + "_dets"]
        for name, trk in (("sort", oracle.Sort(0.3, 1, 50, 3, 0.3)),
                          ("bytetrack", oracle.ByteTrack(0.3, 30, 50, 3, 0.3, 0.1, 0.45, 0.8, 30, 30))):
            h = hashlib.sha256()
            rows = 0
            for f in range(int(frames.min()), int(frames.max()) + 1):
                out = trk.update(dets[frames == f])
                h.update(np.int32(f).tobytes())
                h.update(out.tobytes())
                rows += len(out)
            want_digest, want_rows = z[f"{seq}_{name}_digest"]
            assert rows == int(want_rows)
            assert h.hexdigest() == str(want_digest), f"{seq} {name}: oracle output changed"
