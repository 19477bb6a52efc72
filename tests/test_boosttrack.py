"""BoostTrack (second half; default options, ECC / ReID off): kernel logic under the SIMT emulator (CPU) and
parity of the sm_100a kernel through the C ABI (GPU) against oracle/boosttrack.cpp - which tests/test_ref_pin.py pins bit for
bit to the reference's own src/trackers/boosttrack.cpp compiled in place."""
import numpy as np
import pytest

import sim_lib
from motcpp_b200 import _lib, api, synth

BOOST = dict(det_thresh=0.6, max_age=60, max_obs=50, min_hits=3, iou_threshold=0.3, min_box_area=10, aspect_ratio_thresh=1.6,
             lambda_iou=0.5, lambda_mhd=0.25, lambda_shape=0.25, use_dlo_boost=True, dlo_boost_coef=0.65, use_vt=False)


def test_boost_cost_pieces_against_float64_numpy(oracle):
    rng = np.random.default_rng(2)
    n, m = 19, 27
    def boxes(k):
        c = rng.uniform(0, 600, (k, 2)); wh = rng.uniform(30, 120, (k, 2))
        return np.concatenate([c - wh / 2, c + wh / 2], 1)
    d, t = boxes(n).astype(np.float32), boxes(m).astype(np.float32)
    t[:10] = d[:10] + rng.normal(0, 4, (10, 4)).astype(np.float32)
    w, h = t[:, 2] - t[:, 0], t[:, 3] - t[:, 1]
    mean = np.stack([t[:, 0] + w / 2, t[:, 1] + h / 2, h, w / h], 1).astype(np.float32)
    var = rng.uniform(5, 200, (m, 4)).astype(np.float32)
    cost, iou, mh = oracle.boost_cost(d, t, mean, var)
    assert np.allclose(iou, 1 - oracle.iou_batch(d, t), atol=1e-6)
    dd = d.astype(np.float64)
    z = np.stack([dd[:, 0] + (dd[:, 2] - dd[:, 0]) / 2, dd[:, 1] + (dd[:, 3] - dd[:, 1]) / 2, dd[:, 3] - dd[:, 1],
                  (dd[:, 2] - dd[:, 0]) / (dd[:, 3] - dd[:, 1])], 1)
    want_mh = (((z[:, None] - mean[None].astype(np.float64)) ** 2) / var[None].astype(np.float64)).sum(-1)
    assert np.allclose(mh, want_mh, rtol=2e-6, atol=1e-5)
    sim = (13.2767 - np.minimum(want_mh, 13.2767)) / 13.2767
    assert np.allclose(cost, iou - 0.25 * sim, atol=2e-6)


def _sim_vs_oracle(oracle, seed, T, over, threads=128, **scene):
    args = {**BOOST, **over}
    d, c = synth.stress_stream(seed, n_frames=T, **scene)
    ref = oracle.BoostTrack(**args)
    sim = sim_lib.SimBoostTrack(1, **args)
    stats = np.zeros(4, np.int64)
    for t in range(T):
        n = int(c[t])
        want = ref.update(d[t, :n])
        out, n_out = sim.update(d[t][None, None], np.array([[n]]), threads)
        got = out[0, 0, :n_out[0, 0]]
        h = sim.header()
        assert h[5] == 0, (seed, t, h[5])
        assert np.array_equal(h[6:10], ref.last_sizes()), (seed, t, h[6:10], ref.last_sizes())
        assert got.shape == want.shape and np.array_equal(got, want), (seed, t)
        stats += ref.last_sizes()
        if t % 5 == 0 or t == T - 1:
            assert np.array_equal(sim.dump(), ref.dump()), (seed, t)
    return stats


def test_boosttrack_kernel_logic_under_emulator(oracle):
    st = _sim_vs_oracle(oracle, 120, 120, {"max_age": 20})
    assert st[2] > 1000 and st[3] > 30                                       # matches and births happened
    _sim_vs_oracle(oracle, 121, 80, {"det_thresh": 0.4, "max_age": 8, "min_hits": 1, "use_vt": True}, threads=64)
    _sim_vs_oracle(oracle, 122, 80, {"use_dlo_boost": False, "lambda_mhd": 0.6, "iou_threshold": 0.5, "aspect_ratio_thresh": 0.5, "max_age": 10})
    _sim_vs_oracle(oracle, 123, 60, {"det_thresh": 0.3, "dlo_boost_coef": 0.9, "max_age": 3, "min_box_area": 4000, "lambda_mhd": 0.9},
                   n_obj=48, canvas=(480, 270))                              # 1 - lambda_mhd < threshold: nothing is pruned


# ------------------------------------------------------------------ the sm_100a kernel through the C ABI (GPU)
@pytest.fixture
def gpu():
    from motcpp_b200 import build
    build.build()
    _lib.require_gpu()


def _engine_vs_oracle(oracle, streams, over, cap, d_max, T_chunk=None):
    args = {**BOOST, **over}
    S, T = len(streams), streams[0][0].shape[0]
    dets = np.stack([s[0] for s in streams], 1)
    counts = np.stack([s[1] for s in streams], 1).astype(np.int32)
    eng = api.Engine(_lib.TRACKER_BOOSTTRACK, S, cap, d_max, **args)
    refs = [oracle.BoostTrack(**args) for _ in range(S)]
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
            assert np.array_equal(eng.dump_boost(s), refs[s].dump()), (s, t1)
            assert np.array_equal(eng.header(s)[6:10], refs[s].last_sizes()), (s, t1)
    eng.close()


@pytest.mark.gpu
def test_gpu_boosttrack_engine_matches_oracle(oracle, gpu):
    streams = [synth.stress_stream(800 + s, n_frames=150) for s in range(4)]
    for d, c in streams:
        c[20::23] = 0
    streams[0][1][60:95] = 0
    _engine_vs_oracle(oracle, streams, {"max_age": 20}, 256, 64, T_chunk=30)
    streams = [synth.stress_stream(820 + s, n_frames=80) for s in range(3)]
    _engine_vs_oracle(oracle, streams, {"det_thresh": 0.4, "max_age": 8, "min_hits": 1, "use_vt": True}, 256, 64, T_chunk=1)
    _engine_vs_oracle(oracle, streams, {"use_dlo_boost": False, "lambda_mhd": 0.9, "iou_threshold": 0.5, "max_age": 10}, 256, 64)
    d = synth.bytetrack_stream(6, n_frames=30)                               # the C2 scene: 512 detections per frame
    _engine_vs_oracle(oracle, [(d, np.full(d.shape[0], d.shape[1], np.int32))], {"max_age": 4}, 1536, 512, T_chunk=10)
    # the microbenchmark's scene: clutter tracks coast for 30 frames, some with runaway height / ratio velocities - the
    # track boxes the grid keeps in its overflow list or drops as outside the detections' hull (grid_device.cuh)
    streams = []
    for s in range(2):
        d = synth.bytetrack_stream(s, n_frames=70, n_clutter=32, n_low=32, config=1)
        streams.append((d, np.full(d.shape[0], d.shape[1], np.int32)))
    _engine_vs_oracle(oracle, streams, {"max_age": 30}, 1536, 512, T_chunk=35)


@pytest.mark.gpu
def test_gpu_boosttrack_api_mirror_and_reset(oracle, gpu):
    d, c = synth.stress_stream(79, n_frames=60)
    trk, ref = api.BoostTrack(max_age=15, track_capacity=256, max_dets=64), oracle.BoostTrack(**{**BOOST, "max_age": 15})
    for t in range(30):
        assert np.array_equal(trk.update(d[t, :c[t]], (540, 960)), ref.update(d[t, :c[t]])), t
    trk.reset(); ref.reset()                                                 # ids restart at 1 (boosttrack.cpp:272-277)
    for t in range(30, 60):
        assert np.array_equal(trk.update(d[t, :c[t]], (540, 960)), ref.update(d[t, :c[t]])), t
    with pytest.raises(ValueError):
        api.BoostTrack(use_ecc=True)
    with pytest.raises(ValueError):
        api.BoostTrack(with_reid=True)
    with pytest.raises(_lib.MotError):
        api.Engine(_lib.TRACKER_BOOSTTRACK, 1, 256, 64, use_sb=1)
