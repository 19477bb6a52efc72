"""DeepOC-SORT: kernel logic under the SIMT emulator (CPU) and parity of the sm_100a kernel through the
C ABI (GPU) against oracle/deepocsort.cpp - which tests/test_ref_pin.py pins bit for bit to the reference's own
src/trackers/deepocsort.cpp compiled in place.

The reference lists everything its assignment leaves unmatched TWICE (deepocsort.cpp:476-481 and :485-501), so a new
object spawns two bit-identical tracks whenever the assignment branch ran and the re-match sees every row and column
twice: exactly tied optima in nearly every frame.  The kernel resolves them as the reference does - with the
reference's own dense LAPJV (csrc/jv_device.cuh / jv_block_device.cuh) - so parity here is bit for bit, ids included.
"""
import numpy as np
import pytest

import sim_lib
from motcpp_b200 import _lib, api, synth

DOC = dict(det_thresh=0.3, max_age=30, max_obs=50, min_hits=3, iou_threshold=0.3, delta_t=3, inertia=0.2,
           w_association_emb=0.5, alpha_fixed_emb=0.95, aw_param=0.5, embedding_off=False, aw_off=False,
           q_xy_scaling=0.01, q_s_scaling=0.0001)


def _stream(seed, T, dim, n_obj=40, canvas=(960, 540)):
    return synth.stress_stream_reid(seed, n_frames=T, dim=dim, n_obj=n_obj, canvas=canvas)


def _compare_state(ref, got_rows, got_embs, dim, where):
    if dim:
        want_rows, want_embs = ref.dump(dim)
    else:
        want_rows, want_embs = ref.dump(), None
    assert got_rows.shape[0] == want_rows.shape[0], where
    # [id, age, hits, streak, tsu, conf, cls, det_ind, last_obs 5, velocity 2] then x 7, P 49
    assert np.array_equal(got_rows[:, :15], want_rows[:, :15]), where
    assert np.array_equal(got_rows[:, 15:71], want_rows[:, 15:71]), where
    if dim:
        assert np.array_equal(got_embs, want_embs), where


def _sim_vs_oracle(oracle, seed, T, over, dim=16, threads=128, **scene):
    args = {**DOC, **over}
    use_embs = not args["embedding_off"]
    d, c, e = _stream(seed, T, dim, **scene)
    ref = oracle.DeepOCSort(**args)
    sim = sim_lib.SimDeepOCSort(1, dim, **{k: v for k, v in args.items() if k != "max_obs"})
    stats = np.zeros(8, np.int64)
    for t in range(T):
        n = int(c[t])
        want = ref.update(d[t, :n], e[t, :n] if use_embs else None)
        out, n_out = sim.update(d[t][None, None], np.array([[n]]), e[t][None, None] if use_embs else None, threads)
        got = out[0, 0, :n_out[0, 0]]
        h = sim.header()
        assert h[5] == 0, (seed, t, h[5])
        assert np.array_equal(h[6:14], ref.last_sizes()), (seed, t, h[6:14], ref.last_sizes())
        assert got.shape == want.shape and np.array_equal(got, want), (seed, t)
        stats += ref.last_sizes()
        if t % 5 == 0 or t == T - 1:
            rows, embs = sim.dump()
            _compare_state(ref, rows, embs, dim if use_embs else 0, (seed, t))
    return stats, sim.header()[14]


def test_deepocsort_kernel_logic_under_emulator(oracle):
    st, exact = _sim_vs_oracle(oracle, 70, 90, {})
    assert st[2] > 40 and st[6] > 0 and st[7] > 20 and exact > 20      # assignments, re-matches, spawns, exact LAPJV solves
    # (the emulator builds the 256-track / 64-detection shape; the twice-listed leftovers need 2 x unmatched <= 256)
    _sim_vs_oracle(oracle, 71, 60, {"aw_off": True, "max_age": 10}, threads=64)
    _sim_vs_oracle(oracle, 72, 60, {"embedding_off": True, "max_age": 12})
    _sim_vs_oracle(oracle, 73, 60, {"inertia": 0.9, "min_hits": 1, "max_age": 5, "delta_t": 1}, dim=5)   # odd dim: scalar dot path
    _sim_vs_oracle(oracle, 74, 50, {"iou_threshold": 0.1, "w_association_emb": 1.5, "aw_param": 0.8}, n_obj=48, canvas=(480, 270))


def test_deepocsort_cta_wide_lapjv_under_emulator(oracle):
    """The one-warp LAPJV limited to 24 rows + columns: the duplicated lists of larger frames go through the CTA-wide LAPJV."""
    with sim_lib.variant("jvblock"):
        _sim_vs_oracle(oracle, 70, 90, {})
        _sim_vs_oracle(oracle, 74, 50, {"iou_threshold": 0.1, "w_association_emb": 1.5, "aw_param": 0.8}, n_obj=48, canvas=(480, 270))


# ------------------------------------------------------------------ the sm_100a kernel through the C ABI (GPU)
@pytest.fixture
def gpu():
    from motcpp_b200 import build
    build.build()
    _lib.require_gpu()


def _engine_vs_oracle(oracle, streams, over, cap, d_max, dim, T_chunk=None):
    """streams: [(dets (T, ld, 6), counts (T,), embs (T, ld, dim))]; returns the number of exact LAPJV re-solves."""
    args = {**DOC, **over}
    use_embs = not args["embedding_off"]
    S, T = len(streams), streams[0][0].shape[0]
    dets = np.stack([s[0] for s in streams], 1)
    counts = np.stack([s[1] for s in streams], 1).astype(np.int32)
    embs = np.stack([s[2] for s in streams], 1) if use_embs else None
    eng = api.Engine(_lib.TRACKER_DEEPOCSORT, S, cap, d_max, emb_dim=dim if use_embs else 0,
                     **{k: v for k, v in args.items() if k != "max_obs"}, max_obs=args["max_obs"])
    refs = [oracle.DeepOCSort(**args) for _ in range(S)]
    T_chunk = T_chunk or T
    for t0 in range(0, T, T_chunk):
        t1 = min(T, t0 + T_chunk)
        out, n_out = eng.update(dets[t0:t1], counts[t0:t1], ld_out=cap, embs=embs[t0:t1] if use_embs else None)
        eng.check()
        for s in range(S):
            for t in range(t0, t1):
                n = counts[t, s]
                want = refs[s].update(dets[t, s, :n], embs[t, s, :n] if use_embs else None)
                got = out[t - t0, s, :n_out[t - t0, s]]
                assert got.shape == want.shape and np.array_equal(got, want), (s, t)
            _compare_state(refs[s], eng.dump(s, 0), eng.dump_deep_embs(s) if use_embs else None, dim if use_embs else 0, (s, t1))
            assert np.array_equal(eng.header(s)[6:14], refs[s].last_sizes()), (s, t1)
    exact = sum(int(eng.header(s)[14]) for s in range(S))
    eng.close()
    return exact


@pytest.mark.gpu
def test_gpu_deepocsort_engine_matches_oracle_stress(oracle, gpu):
    streams = [_stream(500 + s, 120, 32) for s in range(4)]
    for d, c, e in streams:
        c[20::23] = 0                                   # empty frames
    streams[0][1][50:85] = 0                            # longer than max_age: every track ages out, ids keep counting
    assert _engine_vs_oracle(oracle, streams, {"max_age": 12}, 256, 64, 32, T_chunk=30) > 50
    streams = [_stream(520 + s, 80, 8) for s in range(3)]
    _engine_vs_oracle(oracle, streams, {"aw_off": True, "max_age": 10}, 256, 64, 8, T_chunk=1)
    _engine_vs_oracle(oracle, streams, {"embedding_off": True, "max_age": 10}, 256, 64, 8)
    _engine_vs_oracle(oracle, streams, {"inertia": 0.9, "delta_t": 1, "min_hits": 1, "max_age": 6, "w_association_emb": 2.0,
                                        "aw_param": 0.8, "alpha_fixed_emb": 0.5}, 256, 64, 8)
    streams = [_stream(540 + s, 60, 5) for s in range(2)]                               # odd width: the scalar dot path
    _engine_vs_oracle(oracle, streams, {"max_age": 8}, 256, 64, 5)


@pytest.mark.gpu
def test_gpu_deepocsort_crowded_scene_cta_wide_lapjv(oracle, gpu):
    """260 objects per frame on the 1536-track / 512-detection shape: the first association and the twice-listed
    re-match are far above the one-warp LAPJV's 384 rows + columns, so the CTA-wide LAPJV resolves the ties."""
    streams = [_stream(600 + s, 40, 64, n_obj=260, canvas=(1600, 900)) for s in range(2)]
    assert _engine_vs_oracle(oracle, streams, {"max_age": 4}, 1536, 512, 64) > 0
    # C3-like shape: 1024 detections with 512-float embeddings on the largest shape, a few frames
    dets, embs = synth.embeddings_stream(5, n_frames=4, n_obj=1024, dim=512)
    streams = [(dets, np.full(4, 1024, np.int32), embs)]
    _engine_vs_oracle(oracle, streams, {}, 3072, 2048, 512)


@pytest.mark.gpu
def test_gpu_deepocsort_api_mirror(oracle, gpu):
    d, c, e = _stream(77, 60, 16)
    trk, ref = api.DeepOCSort(max_age=10), oracle.DeepOCSort(**{**DOC, "max_age": 10})
    assert trk.update(np.zeros((0, 6), np.float32), (540, 960)).shape == (0, 8)        # before the width is known
    assert ref.update(np.zeros((0, 6), np.float32)).shape == (0, 8)
    for t in range(60):
        assert np.array_equal(trk.update(d[t, :c[t]], (540, 960), e[t, :c[t]]), ref.update(d[t, :c[t]], e[t, :c[t]])), t
    with pytest.raises(ValueError):
        trk.update(d[0, :3], (540, 960))                                               # the reference would run its ReID net
    with pytest.raises(ValueError):
        trk.update(d[0, :3], (540, 960), np.zeros((3, 8), np.float32))                 # another width
    with pytest.raises(ValueError):
        trk.update(d[0, :3], (540, 960), e[0, :2])
    with pytest.raises(ValueError):
        api.DeepOCSort(cmc_off=False)
    off, ref = api.DeepOCSort(embedding_off=True, max_age=10), oracle.DeepOCSort(**{**DOC, "embedding_off": True, "max_age": 10})
    for t in range(40):
        assert np.array_equal(off.update(d[t, :c[t]], (540, 960)), ref.update(d[t, :c[t]])), t
    trk.reset(); ref = oracle.DeepOCSort(**{**DOC, "max_age": 10})
    with pytest.raises((_lib.MotError, ValueError)):
        api.Engine(_lib.TRACKER_DEEPOCSORT, 1, 256, 64)                                # embeddings on but no width


@pytest.mark.gpu
def test_gpu_deepocsort_reset_capacity_flag_and_packed_path(oracle, gpu):
    """reset() clears the tracks and keeps the id counter (deepocsort.cpp:575-579); the reference's twice-listed leftovers
    overflow a too-small engine loudly (flag bit 1), never silently; the packed host path carries no embeddings."""
    d, c, e = _stream(91, 50, 8)
    args = {**DOC, "max_age": 8}
    eng = api.Engine(_lib.TRACKER_DEEPOCSORT, 1, 256, 64, emb_dim=8, **{k: v for k, v in args.items()})
    ref = oracle.DeepOCSort(**args)
    def run(t0, t1):
        out, n_out = eng.update(d[t0:t1, None], c[t0:t1, None].astype(np.int32), ld_out=256, embs=e[t0:t1, None])
        eng.check()
        for t in range(t0, t1):
            want = ref.update(d[t, :c[t]], e[t, :c[t]])
            assert np.array_equal(out[t - t0, 0, :n_out[t - t0, 0]], want), t
    run(0, 20)
    eng.reset(); ref.reset()
    run(20, 50)                                            # fresh tracks, ids continue where they stopped
    assert eng.dump(0, 0)[:, 0].min() > 20
    eng.close()
    # capacity: 40 objects + clutter with max_age 30 need far more than 2 x unmatched <= 256 after a while
    eng = api.Engine(_lib.TRACKER_DEEPOCSORT, 1, 256, 64, emb_dim=16, **{**DOC, "aw_off": True})
    d2, c2, e2 = _stream(71, 60, 16)
    eng.update(d2[:, None], c2[:, None].astype(np.int32), ld_out=256, embs=e2[:, None])
    with pytest.raises(RuntimeError, match="capacity"):
        eng.check()
    with pytest.raises(_lib.MotError):
        eng.update_packed(d2[:4, None], c2[:4, None].astype(np.int32), max_rows=64)
    eng.close()
    off = api.Engine(_lib.TRACKER_DEEPOCSORT, 1, 256, 64, embedding_off=1, max_age=6)
    rows, offsets, n_out = off.update_packed(d[:10, None], c[:10, None].astype(np.int32), max_rows=256)
    ref = oracle.DeepOCSort(**{**DOC, "embedding_off": True, "max_age": 6})
    for t in range(10):
        assert np.array_equal(rows[offsets[t]:offsets[t + 1]], ref.update(d[t, :c[t]])), t
    off.close()
