"""OC-SORT: oracle known-answer tests, kernel logic under the SIMT emulator (CPU) and parity of the
sm_100a kernels through the C ABI (GPU) - reference src/trackers/ocsort.cpp.

Ties.  The reference spawns bit-identical "twin" tracks (parity trap 8), so exactly tied
assignment optima are systematic in OC-SORT.  Oracle modes: tie_mode=0 resolves them with the reference's LAPJV
scan order (pinned to the real lap_solver.hpp and to the reference's compiled ocsort.cpp); tie_mode=1 with the "prefer
the higher column" infinitesimal of the sparse CUDA solver.  The CUDA kernel's policy is tie_mode=0 at ANY size: every
frame that can tie is re-solved with the reference's own dense LAPJV on the device (one warp, csrc/jv_device.cuh, while
rows + columns <= 384; the whole CTA, csrc/jv_block_device.cuh, above).  Kernel parity is asserted bit for bit against
tie_mode=0; test_tie_modes_differ_only_on_twin_ties measures how modes 0 and 1 relate.
"""
import ctypes as C

import numpy as np
import pytest

import sim_lib
from motcpp_b200 import _lib, api, synth

OC_ARGS = dict(det_thresh=0.2, max_age=30, max_obs=50, min_hits=3, iou_threshold=0.3, min_conf=0.1, delta_t=3,
               inertia=0.2, use_byte=False, q_xy_scaling=0.01, q_s_scaling=0.0001)


# ------------------------------------------------------------------ oracle KATs (CPU)
def test_acosf_is_correctly_rounded(oracle):
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-1, 1, 20000), [1, -1, 0, 0.5, -0.5, 0.25, 0.99999994, -0.99999994]]).astype(np.float32)
    mine = np.array([oracle.lib().orc_acosf(float(x)) for x in xs], np.float32)
    want = np.arccos(xs.astype(np.float64)).astype(np.float32)
    assert np.array_equal(mine, want)
    # the box's libm (what the reference would link here) is within 1 ulp of it
    libm = C.CDLL("libm.so.6")
    libm.acosf.argtypes, libm.acosf.restype = [C.c_float], C.c_float
    lm = np.array([libm.acosf(float(x)) for x in xs[:5000]], np.float32)
    assert np.abs(mine[:5000].view(np.int32) - lm.view(np.int32)).max() <= 1


def test_ocm_cost_against_float64_numpy(oracle):
    rng = np.random.default_rng(1)
    n, m = 23, 31
    def boxes(k):
        c = rng.uniform(0, 600, (k, 2)); wh = rng.uniform(30, 120, (k, 2))
        return np.concatenate([c - wh / 2, c + wh / 2], 1)
    dets = np.concatenate([boxes(n), rng.uniform(0.3, 1.0, (n, 1))], 1).astype(np.float32)
    trks = boxes(m).astype(np.float32)
    vel = rng.normal(0, 1, (m, 2)); vel /= np.linalg.norm(vel, axis=1, keepdims=True)
    vel[::5] = 0.0
    prev = np.concatenate([boxes(m), rng.uniform(0.3, 1, (m, 1))], 1)
    prev[::7] = -1.0
    cost, iou = oracle.ocm_cost(dets, trks, vel.astype(np.float32), prev.astype(np.float32), 0.2)
    assert np.array_equal(iou, oracle.iou_batch(dets[:, :4], trks))
    d, p, v = dets.astype(np.float64), prev.astype(np.float32).astype(np.float64), vel.astype(np.float32).astype(np.float64)
    cx1, cy1 = (d[:, 0] + d[:, 2]) / 2, (d[:, 1] + d[:, 3]) / 2
    cx2, cy2 = (p[:, 0] + p[:, 2]) / 2, (p[:, 1] + p[:, 3]) / 2
    dx, dy = cx1[:, None] - cx2[None], cy1[:, None] - cy2[None]
    norm = np.sqrt(dx * dx + dy * dy) + 1e-6
    cosang = np.clip(v[None, :, 1] * dx / norm + v[None, :, 0] * dy / norm, -1, 1)
    ang = (np.pi / 2 - np.abs(np.arccos(cosang))) / np.pi
    want = -(iou.astype(np.float64) + (p[None, :, 4] >= 0) * ang * 0.2 * d[:, 4:5])
    assert np.allclose(cost, want, rtol=0, atol=2e-6)


def test_ocsort_reference_style_kats(oracle):
    # tests/test_trackers.cpp:113-119 (valid 8-column output) + the SORT-style persistence scenario
    t = oracle.OCSort()
    dets = np.array([[100, 100, 200, 200, .9, 0], [300, 300, 400, 400, .8, 0], [500, 100, 600, 200, .7, 1]], np.float32)
    assert t.update(dets).shape == (0, 8)                   # first frame: tracks are created, nothing is emitted
    out = t.update(dets)
    assert out.shape == (3, 8)                              # frame_count <= min_hits: emitted, REVERSE track order
    assert list(out[:, 4]) == [4.0, 3.0, 2.0]               # ids are id()+1 (ocsort.cpp:576)
    assert np.array_equal(out[:, :4], dets[::-1, :4]) and list(out[:, 7]) == [2.0, 1.0, 0.0]
    assert t.update(np.zeros((0, 6), np.float32)).shape == (0, 8)
    out = t.update(dets)                                    # hit_streak was reset by the missed frame, frame 4 > min_hits
    assert out.shape == (0, 8)


def test_ocsort_duplicate_spawn_trap(oracle):
    """Parity trap 8: an assignment pair rejected by the IoU filter ends up twice in the unmatched lists, so the
    detection spawns TWO tracks with consecutive ids."""
    found = False
    for shift in range(45, 80):
        t = oracle.OCSort(**{**OC_ARGS, "min_hits": 1})
        for k in range(6):                                  # a track moving right at 10 px / frame
            t.update(np.array([[100 + 10 * k, 100, 200 + 10 * k, 300, .9, 0]], np.float32))
        n0 = len(t.dump())
        # a jump in the direction of motion: IoU with the prediction falls below 0.3 but the momentum term lifts
        # iou + angle cost over it => the assignment pairs them, the IoU filter rejects the pair (ocsort.cpp:702-712)
        t.update(np.array([[150 + shift, 100, 250 + shift, 300, .9, 0]], np.float32))
        ls, d = t.last_sizes(), t.dump()
        if ls[2] == 1 and ls[3] == 0 and ls[7] == 2:
            assert len(d) == n0 + 2 and d[-1, 0] == d[-2, 0] + 1
            assert np.array_equal(d[-1, 1:], d[-2, 1:])     # bit-identical twins
            found = True
            break
    assert found


# ------------------------------------------------------------------ kernel logic under the emulator (CPU)
def _sim_vs_oracle(oracle, seed, T, args, n_obj=40, canvas=(960, 540), threads=128, asso_func="iou"):
    d, c = synth.stress_stream(seed, n_frames=T, n_obj=n_obj, canvas=canvas)
    ref = oracle.OCSort(**args, tie_mode=0, asso_func=asso_func, frame=canvas)
    sim = sim_lib.SimOCSort(1, args["det_thresh"], args["max_age"], args["min_hits"], args["iou_threshold"],
                            args["min_conf"], args["delta_t"], args["inertia"], args["use_byte"], args["q_xy_scaling"],
                            args["q_s_scaling"], asso_func=asso_func, frame=canvas)
    stats = np.zeros(8, np.int64)
    for t in range(T):
        n = int(c[t])
        want = ref.update(d[t, :n])
        out, n_out = sim.update(d[t][None, None], np.array([[n]]), threads)
        got = out[0, 0, :n_out[0, 0]]
        h = sim.header()
        assert h[5] == 0
        assert np.array_equal(h[6:14], ref.last_sizes()), (seed, t)
        assert got.shape == want.shape and np.array_equal(got, want), (seed, t)
        stats += ref.last_sizes()
        if t % 5 == 0 or t == T - 1:
            dm, sd = ref.dump(), sim.dump()
            assert np.array_equal(sd[:, :15], dm[:, :15]) and np.array_equal(sd[:, 15:], dm[:, 20:]), (seed, t)
    return stats


def test_ocsort_kernel_logic_under_emulator(oracle):
    st = _sim_vs_oracle(oracle, 0, 110, OC_ARGS)          # frame 105 of this stream has a twin tie the two rules disagree on
    assert st[2] > 50 and st[6] > 0 and st[7] > 20          # assignments, re-matches and spawns all happened
    _sim_vs_oracle(oracle, 7, 70, {**OC_ARGS, "use_byte": True}, threads=64)
    _sim_vs_oracle(oracle, 8, 60, {**OC_ARGS, "use_byte": True, "inertia": 0.9})        # dense (unpruned) path
    _sim_vs_oracle(oracle, 9, 50, {**OC_ARGS, "iou_threshold": 0.1, "inertia": 0.5}, n_obj=48, canvas=(480, 270))


def test_ocsort_centroid_association_under_emulator(oracle):
    """asso_func = "centroid" (iou.hpp:298-330): similarity 1 - centre distance / frame diagonal for EVERY pair - no
    pruning, dense candidate sets, thresholds near 1."""
    st = _sim_vs_oracle(oracle, 11, 80, {**OC_ARGS, "iou_threshold": 0.95}, asso_func="centroid")
    assert st[2] > 30 and st[3] > 500                       # assignments ran and matched
    _sim_vs_oracle(oracle, 12, 60, {**OC_ARGS, "iou_threshold": 0.9, "use_byte": True, "inertia": 0.5}, asso_func="centroid", threads=64)
    _sim_vs_oracle(oracle, 13, 40, {**OC_ARGS, "iou_threshold": 0.3}, asso_func="centroid", n_obj=20)   # nearly everything is a candidate
    with sim_lib.variant("jvblock"):
        _sim_vs_oracle(oracle, 11, 80, {**OC_ARGS, "iou_threshold": 0.95}, asso_func="centroid")


def test_ocsort_cta_wide_lapjv_under_emulator(oracle):
    """The same streams with the one-warp LAPJV limited to 24 rows + columns: every larger twin-tie frame goes through
    the CTA-wide dense LAPJV (csrc/jv_block_device.cuh), which must reproduce the reference's LAPJV (oracle tie_mode 0)."""
    with sim_lib.variant("jvblock"):
        _sim_vs_oracle(oracle, 0, 110, OC_ARGS)
        _sim_vs_oracle(oracle, 7, 70, {**OC_ARGS, "use_byte": True}, threads=64)
        _sim_vs_oracle(oracle, 9, 50, {**OC_ARGS, "iou_threshold": 0.1, "inertia": 0.5}, n_obj=48, canvas=(480, 270))


def test_tie_modes_differ_only_on_twin_ties(oracle):
    """tie_mode 0 (reference LAPJV order) and 1 (kernel rule) may only part ways at a frame whose first
    association had two bit-identical candidate columns (twin tracks); on such frames the kernel rule agrees
    with the reference most of the time."""
    diverged = tie_frames = agree = 0
    for seed in range(6):
        d, c = synth.stress_stream(seed, n_frames=150, n_obj=40)
        a, b = oracle.OCSort(**OC_ARGS, tie_mode=0), oracle.OCSort(**OC_ARGS, tie_mode=1)
        a.capture(True)
        for t in range(150):
            oa, ob = a.update(d[t, :c[t]]), b.update(d[t, :c[t]])
            cost = a.last_cost()
            twins = False
            if a.last_sizes()[2] == 1 and cost.size:
                cols = [cost[:, j].tobytes() for j in range(cost.shape[1]) if (cost[:, j] <= -0.3).any()]
                twins = len(set(cols)) < len(cols)
            same = oa.shape == ob.shape and np.array_equal(oa, ob) and np.array_equal(a.dump(), b.dump())
            tie_frames += twins
            agree += twins and same
            if not same:
                assert twins, (seed, t)
                diverged += 1
                break                                        # states differ from here on
    assert tie_frames > 20 and agree >= 0.9 * (tie_frames - diverged)


# ------------------------------------------------------------------ GPU parity through the C ABI
@pytest.fixture
def gpu():
    from motcpp_b200 import build
    build.build()
    _lib.require_gpu()


@pytest.mark.gpu
def test_gpu_acos_and_ocm_cost_match_oracle(oracle, gpu):
    rng = np.random.default_rng(5)
    for n, m in ((1, 1), (17, 33), (300, 700), (1030, 517)):
        def boxes(k):
            c = rng.uniform(0, 3000, (k, 2)); wh = rng.uniform(30, 160, (k, 2))
            return np.concatenate([c - wh / 2, c + wh / 2], 1)
        dets = np.concatenate([boxes(n), rng.uniform(0.2, 1.0, (n, 1))], 1).astype(np.float32)
        trks = boxes(m).astype(np.float32)
        trks[: min(n, m)] = dets[: min(n, m), :4] + rng.normal(0, 6, (min(n, m), 4)).astype(np.float32)
        vel = rng.normal(0, 1, (m, 2)); vel /= np.linalg.norm(vel, axis=1, keepdims=True)
        vel[::5] = 0.0
        prev = np.concatenate([boxes(m), rng.uniform(0.3, 1, (m, 1))], 1).astype(np.float32)
        prev[::7] = -1.0
        got_c, got_i = api.ocm_cost(dets, trks, vel, prev, 0.2)
        want_c, want_i = oracle.ocm_cost(dets, trks, vel.astype(np.float32), prev, 0.2)
        assert np.array_equal(got_i, want_i) and np.array_equal(got_c.view(np.int32), want_c.view(np.int32)), (n, m)


def _engine_vs_oracle(oracle, streams, args, cap, d_max, T_chunk=None, check_state_every=10, centroid_frame=None):
    S = len(streams)
    T = streams[0][0].shape[0]
    dets = np.stack([s[0] for s in streams], 1)
    counts = np.stack([s[1] for s in streams], 1).astype(np.int32)
    if centroid_frame:                                   # asso_func = "centroid", frames of this (width, height)
        eng = api.Engine(_lib.TRACKER_OCSORT, S, cap, d_max, **args, asso_func=6, frame_width=centroid_frame[0], frame_height=centroid_frame[1])
        refs = [oracle.OCSort(**args, tie_mode=0, asso_func="centroid", frame=centroid_frame) for _ in range(S)]
    else:
        eng = api.Engine(_lib.TRACKER_OCSORT, S, cap, d_max, **args)
        refs = [oracle.OCSort(**args, tie_mode=0) for _ in range(S)]
    T_chunk = T_chunk or T
    for t0 in range(0, T, T_chunk):
        t1 = min(T, t0 + T_chunk)
        out, n_out = eng.update(dets[t0:t1], counts[t0:t1], ld_out=cap)
        eng.check()
        for s in range(S):
            for t in range(t0, t1):
                want = refs[s].update(dets[t, s, :counts[t, s]])
                got = out[t - t0, s, :n_out[t - t0, s]]
                assert got.shape == want.shape and np.array_equal(got, want), (s, t)
            dm, gd = refs[s].dump(), eng.dump(s, 0)
            assert np.array_equal(gd[:, :15], dm[:, :15]) and np.array_equal(gd[:, 15:71], dm[:, 20:]), (s, t1)
            assert np.array_equal(eng.header(s)[6:14], refs[s].last_sizes())
    eng.close()


@pytest.mark.gpu
def test_gpu_ocsort_engine_matches_oracle_stress(oracle, gpu):
    streams = [synth.stress_stream(200 + s, n_frames=160) for s in range(4)]
    for d, c in streams:
        c[20::23] = 0                                   # empty frames; stream 0 also loses every detection for a while
    streams[0][1][60:95] = 0                            # longer than max_age: every track ages out, ids keep counting
    _engine_vs_oracle(oracle, streams, OC_ARGS, 256, 64, T_chunk=40)
    streams = [synth.stress_stream(300 + s, n_frames=100) for s in range(3)]
    _engine_vs_oracle(oracle, streams, {**OC_ARGS, "use_byte": True}, 256, 64, T_chunk=1)
    _engine_vs_oracle(oracle, streams, {**OC_ARGS, "use_byte": True, "inertia": 0.9, "delta_t": 1}, 256, 64)


@pytest.mark.gpu
def test_gpu_ocsort_c2_shape_and_api_mirror(oracle, gpu):
    d = synth.bytetrack_stream(3, n_frames=45, n_clutter=24, n_low=40, config=4)   # 320 detections / frame
    streams = [(d, np.full(d.shape[0], d.shape[1], np.int32))]
    _engine_vs_oracle(oracle, streams, OC_ARGS, 1536, 512, T_chunk=15)
    trk, ref = api.OCSort(), oracle.OCSort(tie_mode=0)
    dd, cc = synth.stress_stream(77, n_frames=50)
    for t in range(50):
        assert np.array_equal(trk.update(dd[t, :cc[t]], (540, 960)), ref.update(dd[t, :cc[t]]))
    with pytest.raises(ValueError):
        trk.update(np.zeros((2, 5), np.float32), (540, 960))
    with pytest.raises(ValueError):
        trk.update(np.zeros((0, 6), np.float32), None)


@pytest.mark.gpu
def test_gpu_ocsort_centroid_association(oracle, gpu):
    """asso_func = "centroid" wired into the engine (reference iou.hpp:298-330 through ocsort.cpp:413, :438, :494): every pair
    is a candidate above the threshold, nothing is pruned; engine = oracle = the reference's compiled ocsort.cpp (test_ref_pin)."""
    streams = [synth.stress_stream(700 + s, n_frames=100) for s in range(3)]
    _engine_vs_oracle(oracle, streams, {**OC_ARGS, "iou_threshold": 0.95}, 256, 64, T_chunk=25, centroid_frame=(960, 540))
    _engine_vs_oracle(oracle, streams, {**OC_ARGS, "iou_threshold": 0.9, "use_byte": True, "inertia": 0.5}, 256, 64, centroid_frame=(960, 540))
    d = synth.bytetrack_stream(4, n_frames=12, n_clutter=24, n_low=40, config=4)       # 320 detections per frame, C2 canvas
    _engine_vs_oracle(oracle, [(d, np.full(d.shape[0], d.shape[1], np.int32))], {**OC_ARGS, "iou_threshold": 0.97}, 1536, 512,
                      centroid_frame=(3840, 2160))
    trk, ref = api.OCSort(iou_threshold=0.95, asso_func="centroid", track_capacity=256, max_dets=64), \
        oracle.OCSort(**{**OC_ARGS, "iou_threshold": 0.95}, tie_mode=0, asso_func="centroid", frame=(960, 540))
    dd, cc = synth.stress_stream(78, n_frames=40)
    for t in range(40):
        assert np.array_equal(trk.update(dd[t, :cc[t]], (540, 960)), ref.update(dd[t, :cc[t]])), t
    with pytest.raises(ValueError):
        trk.update(dd[0, :cc[0]], (720, 1280))                                        # the frame size is fixed by the first update
    with pytest.raises(ValueError):
        api.OCSort(asso_func="giou")                                                  # undefined in the reference beyond one row
    with pytest.raises(ValueError):
        api.Engine(_lib.TRACKER_OCSORT, 1, 256, 64, **OC_ARGS, asso_func=6)           # centroid without a frame size
    with pytest.raises(_lib.MotError):
        api.Engine(_lib.TRACKER_BYTETRACK, 1, 256, 64, asso_func=6, frame_width=640, frame_height=480)


@pytest.mark.gpu
def test_gpu_ocsort_twin_ties_above_the_one_warp_limit(oracle, gpu):
    """Crowded scenes (rows + columns well above 384) in which the duplicate-spawn trap keeps producing twin tracks: the
    kernel must follow the reference's LAPJV (oracle tie_mode 0, pinned to the reference's own solver) through every tie
    - the CTA-wide dense LAPJV of csrc/jv_block_device.cuh - and the test checks that it actually ran."""
    streams = [synth.stress_stream(400 + s, n_frames=70, n_obj=260, canvas=(1600, 900)) for s in range(2)]
    S = len(streams)
    dets = np.stack([s[0] for s in streams], 1)
    counts = np.stack([s[1] for s in streams], 1).astype(np.int32)
    for args in (OC_ARGS, {**OC_ARGS, "use_byte": True, "iou_threshold": 0.2}):
        eng = api.Engine(_lib.TRACKER_OCSORT, S, 1536, 512, **args)
        refs = [oracle.OCSort(**args, tie_mode=0) for _ in range(S)]
        out, n_out = eng.update(dets, counts, ld_out=1536)
        eng.check()
        exact = 0
        for s in range(S):
            for t in range(dets.shape[0]):
                want = refs[s].update(dets[t, s, :counts[t, s]])
                got = out[t, s, :n_out[t, s]]
                assert got.shape == want.shape and np.array_equal(got, want), (s, t)
            dm, gd = refs[s].dump(), eng.dump(s, 0)
            assert np.array_equal(gd[:, :15], dm[:, :15]) and np.array_equal(gd[:, 15:71], dm[:, 20:]), s
            exact += int(eng.header(s)[14])
        eng.close()
        assert exact > 0, "no twin-tie frame in these streams: the CTA-wide LAPJV never ran"


@pytest.mark.gpu
def test_gpu_ocsort_c4_full_size(oracle, gpu):
    """BASELINE configs[3]: 2048 tracks x 2048 detections (one stream, a few frames: the dense oracle needs ~1 s / frame)."""
    d = synth.ocsort_stream(0, n_frames=8)
    streams = [(d, np.full(d.shape[0], d.shape[1], np.int32))]
    _engine_vs_oracle(oracle, streams, OC_ARGS, 3072, 2048)


@pytest.mark.gpu
def test_gpu_ocsort_capacity_and_argument_errors(oracle, gpu):
    d = synth.bytetrack_stream(1, n_frames=12, n_clutter=192, n_low=0, config=4)     # 448 fresh detections per frame
    eng = api.Engine(_lib.TRACKER_OCSORT, 1, 256, 512, **OC_ARGS)                    # rounded up to 1536 tracks
    eng.update(d[:, None], np.full((12, 1), d.shape[1], np.int32), ld_out=64)        # far more rows than ld_out
    with pytest.raises(RuntimeError, match="output rows truncated"):
        eng.check()
    eng.close()
    with pytest.raises(_lib.MotError):
        api.Engine(_lib.TRACKER_OCSORT, 1, 256, 64, **{**OC_ARGS, "delta_t": 8})     # observation ring holds 8 ages: the current one and the 7 before it
    with pytest.raises(ValueError):
        api.Engine(_lib.TRACKER_OCSORT, 1, 4096, 4096, **OC_ARGS)                    # beyond the largest built shape


def test_small_shape_kernel_equals_reference_tie_breaking(oracle):
    """rows + columns <= 320 for the 256-track / 64-detection shape, so the kernel's policy IS the reference's LAPJV:
    two independent oracle instances in tie_mode 0 stay identical frame after frame on streams full of twin ties."""
    for seed in (0, 3):
        d, c = synth.stress_stream(seed, n_frames=150, n_obj=40)
        a, b = oracle.OCSort(**OC_ARGS, tie_mode=0), oracle.OCSort(**OC_ARGS, tie_mode=0)
        for t in range(150):
            assert np.array_equal(a.update(d[t, :c[t]]), b.update(d[t, :c[t]]))
        assert np.array_equal(a.dump(), b.dump())
