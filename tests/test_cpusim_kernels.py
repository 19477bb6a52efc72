"""Kernel LOGIC under the SIMT emulator (tests/cpusim): the same .cuh sources the product builds
with nvcc, run fiber-per-thread on the CPU and compared bit-for-bit with the oracle.  Small sizes
only; the real parity tests are the -m gpu ones."""
import numpy as np
import pytest

import sim_lib
from motcpp_b200 import synth


def _rand_cost(rng, kind):
    if kind == 0:
        n, m = rng.integers(1, 12, 2)
        return rng.random((n, m)).astype(np.float32), 0.5
    if kind == 1:
        n, m = rng.integers(5, 60, 2)
        return np.where(rng.random((n, m)) < 0.9, 1.0, rng.random((n, m))).astype(np.float32), 0.8
    if kind == 2:                       # dense: one big component -> global-scratch solver
        n, m = rng.integers(20, 50, 2)
        return rng.random((n, m)).astype(np.float32), 0.7
    if kind == 3:
        n, m = rng.integers(30, 200, 2)
        return np.where(rng.random((n, m)) < 0.985, 1.0, rng.random((n, m))).astype(np.float32), 0.8
    n, m = rng.integers(1, 40, 2)
    return -(rng.random((n, m)) * 1.3).astype(np.float32), -0.3


def test_block_lap_matches_oracle(oracle):
    rng = np.random.default_rng(1)
    for trial in range(150):
        c, th = _rand_cost(rng, trial % 5)
        e_cap = 4096 if trial % 7 else 16          # tiny edge buffer -> overflow (single component) path
        got = sim_lib.sim_lap(c, th, e_cap, [32, 64, 128, 256][trial % 4])
        ref = oracle.linear_assignment(c, th)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), (trial, c.shape)


def test_block_lap_empty_sides():
    r, c = sim_lib.sim_lap(np.zeros((0, 5), np.float32), 0.5)
    assert r.size == 0 and np.all(c == -1)
    r, c = sim_lib.sim_lap(np.zeros((4, 0), np.float32), 0.5)
    assert c.size == 0 and np.all(r == -1)


def _run_stream(oracle, dets, counts, cap, d_max, threads, e_cap=4096):
    T = dets.shape[0]
    sim = sim_lib.SimByteTrack(1, cap, d_max, e_cap, 0.1, 0.45, 0.8, 30, 30)
    ref = oracle.ByteTrack(0.3, 30, 50, 3, 0.3, 0.1, 0.45, 0.8, 30, 30)
    for t in range(T):
        n = int(counts[t])
        want = ref.update(dets[t, :n])
        out, n_out = sim.update(dets[t][None, None], np.array([[n]]), threads)
        got = out[0, 0, :n_out[0, 0]]
        assert got.shape == want.shape and np.array_equal(got, want), f"frame {t}"
        assert sim.header()[5] == 0
        if t % 10 == 0 or t == T - 1:               # full state, covariances included
            for which in (0, 1):
                assert np.array_equal(sim.dump(0, which), ref.dump(which)), f"frame {t} list {which}"
    return sim.header()


@pytest.mark.parametrize("sid,threads", [(0, 128), (1, 64), (2, 256)])
def test_bytetrack_kernel_stress_streams(oracle, sid, threads):
    dets, counts = synth.stress_stream(sid, n_frames=120)
    hdr = _run_stream(oracle, dets, counts, 256, 64, threads)
    assert hdr[4] == 120


def test_bytetrack_kernel_detection_gaps(oracle):
    dets, counts = synth.stress_stream(7, n_frames=130)
    counts = counts.copy()
    counts[40:75] = 0                              # > max_time_lost empty frames: everything expires
    counts[100] = 0
    _run_stream(oracle, dets, counts, 256, 64, 128)


def test_bytetrack_kernel_edge_overflow_path(oracle):
    dets, counts = synth.stress_stream(3, n_frames=40)
    _run_stream(oracle, dets, counts, 256, 64, 128, e_cap=8)    # forces the dense single-component solve


def test_bytetrack_kernel_headline_shape(oracle):
    dets = synth.bytetrack_stream(0, n_frames=12)
    _run_stream(oracle, dets, np.full(12, 512), 1536, 512, 256)


def test_bytetrack_kernel_crowded_grid_overflow(oracle):
    """512 detections on a 960x540 canvas: the spatial index overflows and the kernel must fall back
    to visiting all pairs; components get large (exercises the >32-node solver too)."""
    dets = synth.bytetrack_stream(1, n_frames=6, canvas=(960, 540))
    _run_stream(oracle, dets, np.full(6, 512), 2048, 512, 256)


def test_bytetrack_kernel_multi_stream_sequence(oracle):
    """T frames x S streams in one launch == S oracles stepped frame by frame."""
    S, T = 3, 25
    streams = [synth.stress_stream(10 + s, n_frames=T) for s in range(S)]
    ld = streams[0][0].shape[1]
    dets = np.stack([st[0] for st in streams], 1)            # (T,S,ld,6)
    counts = np.stack([st[1] for st in streams], 1)          # (T,S)
    sim = sim_lib.SimByteTrack(S, 256, ld, 4096)
    out, n_out = sim.update(dets, counts, 128, os_threads=3)
    for s in range(S):
        ref = oracle.ByteTrack(0.3, 30, 50, 3, 0.3, 0.1, 0.45, 0.8, 30, 30)
        for t in range(T):
            want = ref.update(dets[t, s, :counts[t, s]])
            assert np.array_equal(out[t, s, :n_out[t, s]], want), (s, t)


# ------------------------------------------------------------------ SORT kernel
@pytest.mark.parametrize("sid,args,threads", [(0, (0.3, 1, 3, 0.3), 128), (1, (0.3, 3, 1, 0.3), 64), (2, (0.5, 30, 3, 0.2), 256)])
def test_sort_kernel_stress_streams(oracle, sid, args, threads):
    dets, counts = synth.stress_stream(20 + sid, n_frames=150)
    det_thresh, max_age, min_hits, iou_thr = args
    sim = sim_lib.SimSort(1, det_thresh, max_age, min_hits, iou_thr)
    ref = oracle.Sort(det_thresh, max_age, 50, min_hits, iou_thr)
    seen = 0
    for t in range(dets.shape[0]):
        n = int(counts[t])
        want = ref.update(dets[t, :n])
        out, n_out = sim.update(dets[t][None, None], np.array([[n]]), threads)
        got = out[0, 0, :n_out[0, 0]]
        assert got.shape == want.shape and np.array_equal(got, want), f"frame {t}"
        assert sim.header()[5] == 0
        seen += len(want)
    assert seen > 100


def test_sort_kernel_reference_kats():
    # reference tests/test_sort.cpp:50-68 and :70-85
    det = np.zeros((1, 1, 64, 6), np.float32)
    det[0, 0, 0] = [100, 100, 200, 200, 0.9, 0]
    one, zero = np.array([[1]]), np.array([[0]])
    s = sim_lib.SimSort(1, 0.3, 3, 1, 0.3)
    s.update(det, one)
    s.update(det, one)
    det[0, 0, 0, :4] = [110, 110, 210, 210]
    out, n = s.update(det, one)
    assert n[0, 0] == 1 and int(out[0, 0, 0, 4]) == 1
    s = sim_lib.SimSort(1, 0.3, 2, 1, 0.3)
    s.update(det, one)
    s.update(det, zero)
    out, n = s.update(det, zero)
    assert n[0, 0] == 0


def test_reference_order_lapjv_kernels_under_emulator(oracle):
    """csrc/jv_device.cuh (one warp) and csrc/jv_block_device.cuh (whole CTA; work arrays in global scratch or all in shared
    memory) must reproduce the reference's LAPJV - ties included - on the extended matrix (oracle pinned to the real
    lap_solver.hpp)."""
    rng = np.random.default_rng(5)
    for trial in range(40):
        n, m = int(rng.integers(1, 28)), int(rng.integers(1, 28))
        kind = trial % 4
        if kind == 0:
            c = rng.random((n, m))
        elif kind == 1:
            c = rng.integers(0, 4, (n, m)) / 4                       # heavy ties
        elif kind == 2:
            c = np.where(rng.random((n, m)) < 0.7, 1.0, rng.random((n, m)))
        else:
            c = -(rng.integers(0, 6, (n, m)) / 5.0)                  # negative costs with ties (the OC-SORT family's -IoU)
        th = -0.3 if kind == 3 else float([0.5, 0.8, 0.3][trial % 3])
        c = c.astype(np.float32)
        want = oracle.linear_assignment(c, th)
        for block, threads in ((0, 32), (1, 64), (2, 64), (2, 128)):
            got = sim_lib.sim_lap_jv(c, th, block, threads)
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), (trial, n, m, kind, block)
        with sim_lib.variant("jvblock"):            # level-opening records applied as parallel permutations (serial limit 2)
            for block, threads in ((1, 64), (2, 96)):
                got = sim_lib.sim_lap_jv(c, th, block, threads)
                assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), (trial, n, m, kind, block, "parallel records")
    # larger problems: more than 40 records per level with the product's own limit
    for trial, (n, m, kind) in enumerate(((60, 90, 1), (20, 140, 3), (100, 70, 2), (12, 150, 3))):
        if kind == 1:
            c = rng.integers(0, 4, (n, m)) / 4
        elif kind == 2:
            c = np.where(rng.random((n, m)) < 0.7, 1.0, rng.random((n, m)))
        else:
            c = -(rng.integers(0, 6, (n, m)) / 5.0)
        th = -0.3 if kind == 3 else 0.8
        c = c.astype(np.float32)
        want = oracle.linear_assignment(c, th)
        for block, threads in ((1, 128), (2, 128)):
            got = sim_lib.sim_lap_jv(c, th, block, threads)
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), ("large", trial, block)


def _iou(a, b):
    x1, y1 = np.maximum(a[:, None, 0], b[None, :, 0]), np.maximum(a[:, None, 1], b[None, :, 1])
    x2, y2 = np.minimum(a[:, None, 2], b[None, :, 2]), np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    ar = lambda q: (q[:, 2] - q[:, 0]) * (q[:, 3] - q[:, 1])
    return inter / (ar(a)[:, None] + ar(b)[None] - inter)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_corner_grid_with_overflow_list_visits_exactly_the_pairs_that_matter(seed):
    """grid_device.cuh: runaway column boxes (inflated, or far off the canvas) go to the overflow list / are dropped; the
    overlap query still sees every overlapping pair exactly once, the IoU-floor query every pair above the floor at most once."""
    rng = np.random.default_rng(seed)
    def boxes(k, lo, hi):
        c = rng.uniform(0, [1920, 1080], (k, 2)); w = rng.uniform(lo, hi, (k, 1)); wh = np.concatenate([w, 2.2 * w], 1)
        return np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    rows = boxes(120, 40, 120)
    cols = boxes(300, 40, 120)
    cols[:90] = rows[:90] + rng.normal(0, 6, (90, 4)).astype(np.float32)          # near-duplicates of rows: high IoU pairs
    cols[100:108, 3] += rng.uniform(2000, 9000, 8).astype(np.float32)             # inflated heights
    cols[108:112] = np.array([-500, -400, 2500, 1600], np.float32)                # covers the whole canvas
    cols[112:120] += np.float32(30000)                                            # coasted far off the canvas
    cols[120] = [np.nan, 0, 10, 10]                                               # non-finite: never visited
    cols[121] = [500, 500, 400, 400]                                              # inverted: empty interior
    big_w, big_h = 2 * (rows[:, 2] - rows[:, 0]).max(), 2 * (rows[:, 3] - rows[:, 1]).max()
    overlap = (np.minimum(rows[:, None, 2], cols[None, :, 2]) > np.maximum(rows[:, None, 0], cols[None, :, 0])) & \
              (np.minimum(rows[:, None, 3], cols[None, :, 3]) > np.maximum(rows[:, None, 1], cols[None, :, 1]))
    for use_roi in (False, True):
        v, n_big = sim_lib.sim_grid_pairs(rows, cols, big_w, big_h, use_roi)
        assert n_big == 12 and np.array_equal(v, overlap.astype(np.int32))
    v, n_big = sim_lib.sim_grid_pairs(rows, cols)                                 # thresholds off: everything in the cells, same visits
    assert n_big == 0 and np.array_equal(v, overlap.astype(np.int32))
    with np.errstate(invalid="ignore", divide="ignore"):
        iou = np.nan_to_num(_iou(rows.astype(np.float64), cols.astype(np.float64)), nan=0.0)
    for t in (0.2, 0.45, 0.79):
        v, _ = sim_lib.sim_grid_pairs(rows, cols, big_w, big_h, True, t)
        assert v.max() <= 1 and not (v.astype(bool) & ~overlap).any()
        assert v[iou > t].all(), t                                                # nothing above the floor is missed
        assert v.sum() < overlap.sum()                                            # and the window is tighter
