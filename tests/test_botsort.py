"""BoT-SORT: oracle known-answer tests, kernel logic under the SIMT emulator (CPU) and parity of the sm_100a
kernel through the C ABI (GPU) - reference src/trackers/botsort.cpp with cmc_method = "none" and embeddings passed
in.  Costs, Kalman state, smoothed features, ids and rows are compared bit for bit."""
import numpy as np
import pytest

import sim_lib
from motcpp_b200 import _lib, api, synth

# BotSort-specific ctor arguments of tools/motcpp_eval.cpp:222-246 (the CLI's BoT-SORT set)
CLI = dict(track_high_thresh=0.6, track_low_thresh=0.1, new_track_thresh=0.7, track_buffer=30, match_thresh=0.8,
           proximity_thresh=0.5, appearance_thresh=0.25, frame_rate=30, fuse_first_associate=False, with_reid=True)


# ------------------------------------------------------------------ oracle KATs (CPU)
def test_botsort_oracle_kats(oracle):
    dets = np.array([[100, 100, 200, 200, .9, 0], [300, 300, 400, 400, .8, 0], [500, 100, 600, 200, .65, 1],
                     [700, 100, 800, 200, .3, 1]], np.float32)
    t = oracle.BotSort(**CLI)
    out = t.update(dets)
    # frame 1: tracks from detections with conf >= new_track_thresh are activated at once (botsort.cpp:103-105)
    assert out.shape == (2, 8) and list(out[:, 4]) == [1.0, 2.0] and list(out[:, 7]) == [0.0, 1.0]
    assert np.array_equal(out[:, :4], dets[:2, :4])
    # an empty frame returns nothing and does not advance the frame counter (:267-269)
    assert t.update(np.zeros((0, 6), np.float32)).shape == (0, 8) and t.counts() == (2, 0)
    out = t.update(dets)
    assert list(out[:, 4]) == [1.0, 2.0]
    # ids restart at 0 on reset (:257)
    t.reset()
    assert list(t.update(dets)[:, 4]) == [1.0, 2.0]


def test_botsort_refound_lost_track_vanishes(oracle):
    """prepare_output (botsort.cpp:712-744) drops a re-found lost track from the lost list without ever copying
    it to the active list."""
    a = np.array([[100, 100, 200, 300, .9, 0]], np.float32)
    low = np.array([[600, 100, 700, 300, .3, 0]], np.float32)      # a low-confidence detection elsewhere
    t = oracle.BotSort(**CLI)
    for _ in range(3):
        t.update(a)
    assert t.counts() == (1, 0)
    t.update(low)                       # second association runs, the track is unmatched -> Lost
    assert t.counts() == (0, 1)
    out = t.update(a)                   # re-found: re_activate()d inside the lost list ... and gone
    assert out.shape == (0, 8) and t.counts() == (0, 0)
    out = t.update(a)                   # the object comes back under a NEW id one frame later
    assert t.counts() == (1, 0) and t.dump(0)[0, 0] == 2.0


def test_botsort_embedding_gate(oracle):
    """With with_reid the cost is min(iou_dist, emb/2) where emb survives only if it is <= appearance_thresh and
    iou_dist <= proximity_thresh: a matching appearance rescues a pair whose IoU alone would fail match_thresh."""
    rng = np.random.default_rng(0)
    f = rng.normal(size=16).astype(np.float32)
    g = rng.normal(size=16).astype(np.float32)
    a = np.array([[100, 100, 200, 300, .9, 0]], np.float32)
    b = np.array([[135, 100, 235, 300, .9, 0]], np.float32)          # IoU with the track ~0.48: iou_dist ~0.52 > 0.5
    c = np.array([[130, 100, 230, 300, .9, 0]], np.float32)          # IoU ~0.54: iou_dist ~0.46 <= proximity_thresh
    for det, feat in ((c, f), (c, g), (b, f)):
        t = oracle.BotSort(**{**CLI, "match_thresh": 0.4})
        t.update(a, f[None]); t.update(a, f[None])
        t.update(det, feat[None])
        matched = t.counts()[0] == 1        # an unmatched detection (conf 0.9) would have started a second track
        # (c, f): emb ~ 0 rescues it; (c, g): appearance differs, iou_dist 0.46 > 0.4 fails; (b, f): masked by proximity
        assert matched == (det is c and feat is f)


# ------------------------------------------------------------------ kernel logic under the emulator (CPU)
def _sim_vs_oracle(oracle, seed, T, dim, args, use_embs=True, threads=128, n_obj=40, canvas=(960, 540)):
    d, c, e = synth.stress_stream_reid(seed, n_frames=T, dim=max(dim, 4), n_obj=n_obj, canvas=canvas)
    ref = oracle.BotSort(**args)
    sim = sim_lib.SimBotSort(1, dim if use_embs else 0, *args.values())
    stats = np.zeros(8, np.int64)
    for t in range(T):
        n = 0 if t % 17 == 13 else int(c[t])                 # some empty frames
        want = ref.update(d[t, :n], e[t, :n] if use_embs else None)
        out, n_out = sim.update(d[t][None, None], np.array([[n]]), e[t][None, None] if use_embs else None, threads)
        got = out[0, 0, :n_out[0, 0]]
        h = sim.header()
        assert h[5] == 0
        assert got.shape == want.shape and np.array_equal(got, want), (seed, t)
        if n:
            assert np.array_equal(h[6:14], ref.last_sizes()), (seed, t)
            stats += ref.last_sizes()
        if t % 4 == 0 or t == T - 1:
            for which in (0, 1):
                if use_embs:
                    rb, rf = ref.dump(which, dim)
                    sb, sf = sim.dump(0, which)
                    assert np.array_equal(rb, sb) and np.array_equal(rf, sf), (seed, t, which)
                else:
                    assert np.array_equal(ref.dump(which), sim.dump(0, which)[0]), (seed, t, which)
    return stats


def test_botsort_kernel_logic_under_emulator(oracle):
    st = _sim_vs_oracle(oracle, 0, 90, 32, CLI)
    assert st[2] > 100 and st[4] > 100 and st[6] > 100 and st[7] > 100    # every stage ran, tracks were lost
    _sim_vs_oracle(oracle, 5, 60, 0, CLI, use_embs=False)
    _sim_vs_oracle(oracle, 6, 60, 32, {**CLI, "fuse_first_associate": True}, threads=64)
    _sim_vs_oracle(oracle, 7, 60, 32, {**CLI, "track_high_thresh": 0.5, "new_track_thresh": 0.6, "track_buffer": 10,
                                      "appearance_thresh": 0.6, "with_reid": False})
    _sim_vs_oracle(oracle, 8, 50, 64, CLI, n_obj=48, canvas=(480, 270))
    # proximity_thresh = 1: no pair is gated out, nothing can be pruned or tabulated -> dense scan, per-thread dot fallback
    _sim_vs_oracle(oracle, 9, 30, 32, {**CLI, "proximity_thresh": 1.0})
    _sim_vs_oracle(oracle, 10, 12, 516, CLI)         # more than 512 floats: the streaming (not register-resident) feature passes


# ------------------------------------------------------------------ GPU parity through the C ABI
@pytest.fixture
def gpu():
    from motcpp_b200 import build
    build.build()
    _lib.require_gpu()


def _engine_vs_oracle(oracle, streams, args, cap, d_max, dim, T_chunk=None):
    S = len(streams)
    T = streams[0][0].shape[0]
    dets = np.stack([s[0] for s in streams], 1)
    counts = np.stack([s[1] for s in streams], 1).astype(np.int32)
    embs = np.stack([s[2] for s in streams], 1) if dim else None
    eng = api.Engine(_lib.TRACKER_BOTSORT, S, cap, d_max, emb_dim=dim, **args)
    refs = [oracle.BotSort(**args) for _ in range(S)]
    T_chunk = T_chunk or T
    for t0 in range(0, T, T_chunk):
        t1 = min(T, t0 + T_chunk)
        out, n_out = eng.update(dets[t0:t1], counts[t0:t1], ld_out=cap, embs=embs[t0:t1] if dim else None)
        eng.check()
        for s in range(S):
            for t in range(t0, t1):
                n = counts[t, s]
                want = refs[s].update(dets[t, s, :n], embs[t, s, :n] if dim else None)
                got = out[t - t0, s, :n_out[t - t0, s]]
                assert got.shape == want.shape and np.array_equal(got, want), (s, t)
            for which in (0, 1):
                if dim:
                    rb, rf = refs[s].dump(which, dim)
                    gb, gf = eng.dump_bot(s, which, with_feats=True)
                    assert np.array_equal(rb, gb) and np.array_equal(rf, gf), (s, t1, which)
                else:
                    assert np.array_equal(refs[s].dump(which), eng.dump_bot(s, which)), (s, t1, which)
    eng.close()


@pytest.mark.gpu
def test_gpu_botsort_engine_matches_oracle_stress(oracle, gpu):
    streams = []
    for s in range(4):
        d, c, e = synth.stress_stream_reid(400 + s, n_frames=150, dim=32)
        c = c.copy(); c[13::17] = 0
        streams.append((d, c, e))
    _engine_vs_oracle(oracle, streams, CLI, 256, 64, 32, T_chunk=50)
    _engine_vs_oracle(oracle, streams[:2], {**CLI, "fuse_first_associate": True}, 256, 64, 32, T_chunk=1)
    _engine_vs_oracle(oracle, streams[:2], {**CLI, "with_reid": False}, 256, 64, 0)


@pytest.mark.gpu
def test_gpu_botsort_api_mirror(oracle, gpu):
    d, c, e = synth.stress_stream_reid(77, n_frames=60, dim=64)
    trk = api.BotSort(**CLI, emb_dim=64, track_capacity=256, max_dets=64)
    ref = oracle.BotSort(**CLI)
    for t in range(60):
        n = 0 if t % 11 == 5 else c[t]
        assert np.array_equal(trk.update(d[t, :n], (540, 960), e[t, :n]), ref.update(d[t, :n], e[t, :n]))
    with pytest.raises(ValueError):
        trk.update(d[0, :3], (540, 960), e[0, :2])
    with pytest.raises(ValueError):
        api.BotSort(cmc_method="ecc")


@pytest.mark.gpu
def test_gpu_botsort_c3_full_size(oracle, gpu):
    """BASELINE configs[2]: BoT-SORT, 1024 objects x 1024 detections per frame, 512-d embeddings."""
    d, e = synth.embeddings_stream(0, n_frames=6)
    streams = [(d, np.full(d.shape[0], d.shape[1], np.int32), e)]
    _engine_vs_oracle(oracle, streams, CLI, 2048, 1024, 512)
