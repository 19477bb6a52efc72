"""StrongSORT cost builders and the XYSR affine correction (8a9): oracle KATs on the CPU, and the sm_100a
kernels through the C ABI against the oracle on the GPU.  Reference: src/trackers/strongsort.cpp:240-334 (nearest-
neighbour cosine), :451-492 (gate_cost_matrix), :502-585 (tlwh IoU cost), src/motion/kalman_filters/xysr_kf.cpp:114-141."""
import numpy as np
import pytest

from motcpp_b200 import _lib, api

NN_COSINE_ATOL = 2e-5      # 3-term bf16 split + fp32 tensor-core accumulation vs the oracle's sequential fp32 sums


def _tracks(rng, n):
    """n plausible XYAH states after a few predict/update rounds (numpy, float32)."""
    means = np.zeros((n, 8), np.float32)
    covs = np.zeros((n, 8, 8), np.float32)
    for i in range(n):
        h = rng.uniform(60, 260)
        means[i] = [rng.uniform(0, 1920), rng.uniform(0, 1080), rng.uniform(0.3, 0.6), h, *rng.normal(0, 2, 4)]
        a = rng.normal(0, 1, (8, 8)) * np.array([h / 20] * 2 + [1e-2] + [h / 20] + [h / 160] * 2 + [1e-5] + [h / 160])[:, None]
        covs[i] = (a @ a.T + np.diag(np.array([h / 10] * 2 + [1e-2] + [h / 10] + [h / 16] * 2 + [1e-5] + [h / 16]) ** 2)).astype(np.float32)
    return means, covs


# ------------------------------------------------------------------ oracle KATs (CPU)
def test_oracle_iou_cost_tlwh_kats(oracle):
    c = oracle.iou_cost_tlwh([[0, 0, 100, 100], [0, 0, 100, 100], [0, 0, 0, 0]],
                             [[50, 50, 100, 100], [0, 0, 100, 100], [500, 500, 10, 10]], tsu=[1, 2, 1])
    assert c[0, 0] == np.float32(1.0) - np.float32(2500.0) / np.float32(17500.0)      # strongsort.cpp:502-536
    assert c[0, 1] == 0.0 and c[0, 2] == 1.0
    assert np.all(c[1] == np.float32(1e5))                                             # time_since_update > 1 (:567-570)
    assert c[2, 2] == 1.0                                                              # zero-area track, disjoint


def test_oracle_gate_cost_matrix_kats(oracle):
    # mean (0, 0, 0.5, 100), P = initiate()'s diag(100, 100, 1e-4, 100, ...): S = diag(125, 125, 1e-4 + 1e-2, 125)
    mean, cov = oracle.KFXYAH.initiate(np.array([0, 0, 0.5, 100], np.float32))
    meas = np.array([[125, 0, 0.5, 100], [387.5, 0, 0.5, 100], [0, 0, 0.5, 100]], np.float32)
    gd = oracle.KFXYAH.gating(mean, cov, meas)
    np.testing.assert_allclose(gd, [1.0, 9.61, 0.0], rtol=1e-6)                         # (dx / 125)^2: the S^-2 quirk
    out = oracle.gate_cost_matrix(np.full((1, 3), 0.25, np.float32), mean[None], cov[None], meas, 0.98)
    lam = np.float32(0.98)
    want = [lam * np.float32(0.25) + (np.float32(1) - lam) * gd[0], lam * np.float32(1e5) + (np.float32(1) - lam) * gd[1],
            lam * np.float32(0.25) + (np.float32(1) - lam) * gd[2]]
    assert np.array_equal(out[0], np.array(want, np.float32))                          # strongsort.cpp:477-487


def test_oracle_nn_cosine_vs_numpy(oracle):
    rng = np.random.default_rng(5)
    smp = rng.normal(0, 1, (37, 48)).astype(np.float32)
    seg = rng.integers(0, 6, 37).astype(np.int32)
    seg[seg == 4] = 3                                                                  # target 4 has no samples
    f = rng.normal(0, 1, (11, 48)).astype(np.float32)
    smp[5] = 0.0                                                                       # zero row stays unnormalised: cost 1
    got = oracle.nn_cosine_distance(smp, seg, 6, f)
    s64 = smp.astype(np.float64); f64 = f.astype(np.float64)
    sn = np.linalg.norm(s64, axis=1, keepdims=True); sn[sn <= 1e-10] = 1.0
    d = 1.0 - (s64 / sn) @ (f64 / np.linalg.norm(f64, axis=1, keepdims=True)).T
    want = np.full((6, 11), 1e5)
    for t in range(6):
        if np.any(seg == t):
            want[t] = d[seg == t].min(axis=0)
    np.testing.assert_allclose(got, want, atol=2e-6, rtol=0)
    assert np.all(got[4] == np.float32(1e5))


def test_oracle_xysr_affine_vs_numpy(oracle):
    rng = np.random.default_rng(9)
    x = rng.normal(0, 50, 7).astype(np.float32)
    a = rng.normal(0, 1, (7, 7)); P = (a @ a.T).astype(np.float32)
    m = np.array([[1.01, -0.02], [0.03, 0.99]], np.float32); t = np.array([3.5, -1.25], np.float32)
    gx, gP = oracle.kf_xysr_affine(x, P, m, t)
    wx = x.astype(np.float64).copy(); wP = P.astype(np.float64).copy(); m64 = m.astype(np.float64)
    wx[0:2] = m64 @ x[0:2] + t; wx[4:6] = m64 @ x[4:6]
    wP[0:2, 0:2] = m64 @ P[0:2, 0:2] @ m64.T; wP[4:6, 4:6] = m64 @ P[4:6, 4:6] @ m64.T
    wP[0:2, 4:6] = m64 @ P[0:2, 4:6] @ m64.T; wP[4:6, 0:2] = wP[0:2, 4:6].T
    np.testing.assert_allclose(gx, wx, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(gP, wP, rtol=1e-5, atol=1e-5)
    # identity warp: nothing changes, bit for bit (xysr_kf.cpp:114-141 with m = I, t = 0)
    ix, iP = oracle.kf_xysr_affine(x, P, np.eye(2, dtype=np.float32), np.zeros(2, np.float32))
    assert np.array_equal(ix, x) and np.array_equal(iP[0:2, 0:2], P[0:2, 0:2])


def test_oracle_aw_max_metric_kats(oracle):
    # deepocsort.cpp:294-345 by hand: a distinctive row keeps its weight, an ambiguous one loses it
    e = np.array([[0.9, 0.1, 0.0], [0.8, 0.8, 0.2]], np.float32)
    out = oracle.aw_max_metric(e, 0.5, 0.5)
    f = np.float32
    def wt(mx, se):
        return f(1) - max(f(f(se) / f(mx)) - f(0.5), f(0)) / f(0.5)
    rw = [wt(0.9, 0.1), wt(0.8, 0.8)]                       # 1.0 and 0.0
    cw = [wt(0.9, 0.8), wt(0.8, 0.1), wt(0.2, 0.0)]
    assert rw[0] == 1.0 and rw[1] == 0.0
    want = np.array([[f(f(f(0.5) * rw[i]) * cw[j]) * e[i, j] for j in range(3)] for i in range(2)], np.float32)
    assert np.array_equal(out, want)
    z = oracle.aw_max_metric(np.array([[0.0, -1.0], [0.5, 0.25]], np.float32))
    assert np.all(z[0] == 0.0)                             # row maximum 0: weights zeroed (:315-316)
    one = oracle.aw_max_metric(np.array([[0.3, 0.7]], np.float32), 0.5, 0.5)          # a single row: columns are not weighted
    assert np.array_equal(one, np.array([[f(0.5) * wt(0.7, 0.3) * f(0.3), f(0.5) * wt(0.7, 0.3) * f(0.7)]], np.float32))


def test_oracle_iou_variant_kats(oracle):
    # reference tests/test_iou.cpp:74-115 (box1 = [0,0,100,100], box2 = [50,50,150,150], box3 = [200,200,300,300]) + by hand
    b1, b2, b3 = [[0, 0, 100, 100]], [[50, 50, 150, 150]], [[200, 200, 300, 300]]
    f = np.float32
    iou = f(2500) / f(17500)
    g = oracle.iou_variant(4, b1, b2)[0, 0]
    inter = iou * f(20000) / (iou + f(1e-10))
    want_g = (iou - (f(22500) - (f(20000) - inter)) / (f(22500) + f(1e-10)) + f(1)) / f(2)
    assert 0.0 <= g <= 1.0 and g == want_g                                        # enclosing box 150 x 150
    d = oracle.iou_variant(5, b1, b2)[0, 0]
    assert 0.0 <= d <= 1.0 and d == (iou - f(5000) / (f(45000) + f(1e-10)) + f(1)) / f(2)   # centres 50 apart per axis
    c = oracle.iou_variant(6, b1, b3, 640, 480)[0, 0]
    assert 0.0 < c < 1.0 and c == f(1) - f(np.sqrt(f(80000))) / f(800.0)           # sqrt(640^2 + 480^2) = 800
    h = oracle.iou_variant(3, b1, b2)[0, 0]
    assert h == iou * (f(50) / f(150))
    # identical boxes: diou is 1; giou is 0.5, because the reference recovers the intersection as iou (a1 + a2) / (iou + 1e-10)
    # (iou.hpp:181) - twice the true value at iou = 1 - and its union collapses to 0
    assert oracle.iou_variant(5, b1, b1)[0, 0] == 1.0 and oracle.iou_variant(4, b1, b1)[0, 0] == 0.5


def test_atanf_is_correctly_rounded_and_ciou_follows_the_formula(oracle):
    """ciou (iou.hpp:197-253) needs an arc tangent; the contract is the correctly rounded fp32 value (see oracle/cost.cpp)."""
    import ctypes as C
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-4, 4, 20000), 10.0 ** rng.uniform(-12, 25, 4000), -10.0 ** rng.uniform(-12, 25, 1000),
                         [0, 1, -1, 0.4375, 0.6875, 1.1875, 2.4375, 1e30, -1e30, np.inf, -np.inf]]).astype(np.float32)
    mine = np.array([oracle.lib().orc_atanf(float(x)) for x in xs], np.float32)
    assert np.array_equal(mine, np.arctan(xs.astype(np.float64)).astype(np.float32))
    libm = C.CDLL("libm.so.6")
    libm.atanf.argtypes, libm.atanf.restype = [C.c_float], C.c_float
    lm = np.array([libm.atanf(float(x)) for x in xs[:5000]], np.float32)
    assert np.abs(mine[:5000].view(np.int32) - lm.view(np.int32)).max() <= 1     # the box's libm: within 1 ulp of it
    # ciou against a float64 re-derivation, and its fixed points
    def boxes(k):
        xy = rng.uniform(0, 600, (k, 2)); wh = rng.uniform(5, 200, (k, 2))
        return np.concatenate([xy, xy + wh], 1).astype(np.float32)
    a, b = boxes(40), boxes(30)
    got = oracle.iou_variant(7, a, b)
    A, B = a.astype(np.float64)[:, None], b.astype(np.float64)[None]
    iou = oracle.iou_batch(a, b).astype(np.float64)
    inner = ((A[..., 0] + A[..., 2]) / 2 - (B[..., 0] + B[..., 2]) / 2) ** 2 + ((A[..., 1] + A[..., 3]) / 2 - (B[..., 1] + B[..., 3]) / 2) ** 2
    outer = (np.maximum(A[..., 2], B[..., 2]) - np.minimum(A[..., 0], B[..., 0])) ** 2 + \
            (np.maximum(A[..., 3], B[..., 3]) - np.minimum(A[..., 1], B[..., 1])) ** 2 + 1e-7
    ad = np.arctan((B[..., 2] - B[..., 0]) / (B[..., 3] - B[..., 1] + 1e-7)) - np.arctan((A[..., 2] - A[..., 0]) / (A[..., 3] - A[..., 1] + 1e-7))
    v = 4 / np.pi ** 2 * ad ** 2
    want = (iou - inner / outer + v / (1 - iou + v + 1e-7) * v + 1) / 2
    assert np.allclose(got, want, rtol=0, atol=3e-6)
    assert oracle.iou_variant(7, [[0, 0, 100, 100]], [[0, 0, 100, 100]])[0, 0] == 1.0          # identical boxes: v = 0, ciou = iou = 1
    assert oracle.iou_variant(7, [[0, 0, 100, 100]], [[50, 50, 150, 150]])[0, 0] == oracle.iou_variant(5, [[0, 0, 100, 100]], [[50, 50, 150, 150]])[0, 0]


# ------------------------------------------------------------------ GPU parity (through the C ABI)
@pytest.fixture()
def _gpu():
    from motcpp_b200 import build
    build.build()
    _lib.require_gpu()


@pytest.mark.gpu
@pytest.mark.parametrize("n,m", [(1, 1), (9, 70), (256, 512), (517, 1031)])
def test_gate_cost_matrix_bit_exact(oracle, _gpu, n, m):
    rng = np.random.default_rng(100 + n + m)
    means, covs = _tracks(rng, n)
    meas = np.stack([rng.uniform(0, 1920, m), rng.uniform(0, 1080, m), rng.uniform(0.3, 0.6, m), rng.uniform(60, 260, m)], 1).astype(np.float32)
    for k in range(min(n, m)):                       # make a band of pairs pass the gate
        meas[k] = means[k, :4] + rng.normal(0, 1, 4).astype(np.float32) * np.array([3, 3, 0.01, 3], np.float32)
    cost = rng.random((n, m)).astype(np.float32)
    for only_position in (False, True):
        got = api.gate_cost_matrix(cost, means, covs, meas, 0.98, only_position=only_position)
        want = oracle.gate_cost_matrix(cost, means, covs, meas, 0.98, only_position=only_position)
        assert np.array_equal(got, want)
        if not only_position and min(n, m) > 4:
            assert np.any(want < 1.0) and np.any(want > 1e4)          # both sides of the gate are exercised


@pytest.mark.gpu
@pytest.mark.parametrize("n,m", [(1, 1), (7, 5), (256, 512), (300, 1031)])
def test_iou_cost_tlwh_bit_exact(oracle, _gpu, n, m):
    rng = np.random.default_rng(n * 7 + m)
    def boxes(k):
        return np.stack([rng.uniform(0, 900, k), rng.uniform(0, 500, k), rng.uniform(0, 200, k), rng.uniform(0, 300, k)], 1).astype(np.float32)
    a, b = boxes(n), boxes(m)
    b[0] = a[0]
    a[n - 1, 2:] = 0.0
    tsu = rng.integers(0, 3, n).astype(np.int32)
    assert np.array_equal(api.iou_cost_tlwh(a, b, tsu), oracle.iou_cost_tlwh(a, b, tsu))
    assert np.array_equal(api.iou_cost_tlwh(a, b), oracle.iou_cost_tlwh(a, b))


@pytest.mark.gpu
@pytest.mark.parametrize("targets,budget,m,dim", [(1, 1, 1, 8), (5, 3, 7, 33), (40, 100, 70, 128), (256, 100, 512, 512), (300, 17, 200, 64)])
def test_nn_cosine_distance_matches_oracle(oracle, _gpu, targets, budget, m, dim):
    rng = np.random.default_rng(targets + budget + m + dim)
    sizes = rng.integers(0, budget + 1, targets)
    sizes[0] = budget
    if targets > 2:
        sizes[1] = 0                                                  # a target without samples -> 1e5
    seg = np.repeat(np.arange(targets), sizes).astype(np.int32)
    big = targets * budget * m * dim > 2e9                            # the C oracle is a scalar triple loop
    ident = rng.normal(0, 1, (targets, dim)).astype(np.float32)
    smp = (ident[seg] + 0.3 * rng.normal(0, 1, (seg.size, dim))).astype(np.float32)
    f = rng.normal(0, 1, (m, dim)).astype(np.float32)
    k = min(targets, m)
    f[:k] = (ident[:k] * rng.uniform(0.5, 2.0, (k, 1)) + 0.3 * rng.normal(0, 1, (k, dim))).astype(np.float32)
    got = api.nn_cosine_distance(smp, seg, targets, f)
    if big:
        s64, f64 = smp.astype(np.float64), f.astype(np.float64)
        d = 1.0 - (s64 / np.linalg.norm(s64, axis=1, keepdims=True)) @ (f64 / np.linalg.norm(f64, axis=1, keepdims=True)).T
        want = np.full((targets, m), 1e5)
        start = np.concatenate([[0], np.cumsum(sizes)])
        for t in range(targets):
            if sizes[t]:
                want[t] = d[start[t]:start[t + 1]].min(axis=0)
    else:
        want = oracle.nn_cosine_distance(smp, seg, targets, f)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, atol=NN_COSINE_ATOL, rtol=0)
    has = sizes > 0
    assert np.all(got[~has] == np.float32(1e5))
    # the decisions downstream are unchanged by the tolerance: same nearest detection per target, same side of 0.2
    assert np.array_equal(np.argmin(got[has], 1), np.argmin(want[has], 1))
    far = np.abs(want - 0.2) > 1e-4
    assert np.array_equal((got <= 0.2)[far], (want <= 0.2)[far])
    # unordered sample rows (seg not sorted) give the same answer
    perm = rng.permutation(seg.size)
    got2 = api.nn_cosine_distance(smp[perm], seg[perm], targets, f)
    np.testing.assert_allclose(got2, got, atol=NN_COSINE_ATOL, rtol=0)


@pytest.mark.gpu
def test_xysr_affine_bit_exact(oracle, _gpu):
    rng = np.random.default_rng(3)
    n = 1000
    xs = rng.normal(0, 100, (n, 7)).astype(np.float32)
    a = rng.normal(0, 1, (n, 7, 7)); Ps = (a @ a.transpose(0, 2, 1)).astype(np.float32)
    m = np.array([[1.002, -0.013], [0.011, 0.997]], np.float32); t = np.array([4.25, -2.5], np.float32)
    gx, gP = api.KalmanFilterXYSR().apply_affine_correction(xs, Ps, m, t)
    for k in range(n):
        wx, wP = oracle.kf_xysr_affine(xs[k], Ps[k], m, t)
        assert np.array_equal(gx[k], wx) and np.array_equal(gP[k], wP), k


@pytest.mark.gpu
@pytest.mark.parametrize("n,m", [(1, 1), (1, 9), (7, 1), (33, 70), (512, 300), (2048, 2048)])
def test_aw_max_metric_bit_exact(oracle, _gpu, n, m):
    rng = np.random.default_rng(n * 31 + m)
    e = rng.random((n, m)).astype(np.float32)
    if n > 4 and m > 4:
        e[1] = 0.0                      # an all-zero row and column: weights zeroed
        e[:, 2] = 0.0
        e[3, 4] = e[3].max()            # a duplicated maximum: second == max
        e[0, 0] = -0.5
    for w, bottom in ((0.5, 0.5), (0.75, 0.2)):
        assert np.array_equal(api.aw_max_metric(e, w, bottom), oracle.aw_max_metric(e, w, bottom)), (n, m, w)


@pytest.mark.gpu
@pytest.mark.parametrize("n,m", [(1, 1), (40, 1), (7, 5), (300, 517)])
def test_iou_variants_bit_exact(oracle, _gpu, n, m):
    rng = np.random.default_rng(n * 13 + m)
    def boxes(k):
        xy = rng.uniform(0, 600, (k, 2)); wh = rng.uniform(5, 200, (k, 2))
        return np.concatenate([xy, xy + wh], 1).astype(np.float32)
    a, b = boxes(n), boxes(m)
    b[0] = a[0]
    for name, kind in (("hmiou", 3), ("giou", 4), ("diou", 5), ("centroid", 6), ("ciou", 7)):
        got = api.asso_batch(name, a, b, 1920, 1080)
        assert np.array_equal(got, oracle.iou_variant(kind, a, b, 1920, 1080)), (name, n, m)
    assert np.array_equal(api.asso_batch("iou", a, b), oracle.iou_batch(a, b))
    with pytest.raises(ValueError):
        api.asso_batch("ct_dist", a, b)
