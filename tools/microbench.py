#!/usr/bin/env python
"""Kernel-level microbenchmarks with roofline fractions: batched Kalman kernels and
the IoU cost kernel against measured HBM bandwidth, the cosine GEMM against measured bf16 tensor
throughput, the assignment kernel as microseconds per solve.  CUDA events on torch's current
stream, >= 3 warm-up iterations, working sets larger than the 126 MB L2 where HBM is the bound.

    python tools/microbench.py [--quick] > profiles/r1_microbench.jsonl
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from motcpp_b200 import _lib, api, build, synth  # noqa: E402

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = float(PEAKS.get("hbm_gbs", 6650.0))
TENSOR = float(PEAKS.get("bf16_tflops", 1590.0))


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    return float(np.median(ms)), float(np.min(ms))


def emit(name, ms, **kw):
    rec = {"kernel": name, "ms_median": ms[0], "ms_min": ms[1]}
    rec.update(kw)
    print(json.dumps(rec), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    build.build()
    _lib.require_gpu()
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    want = lambda k: (not args.only) or (args.only in k)

    # ---------------- Kalman kernels: n tracks, 288 B in + 288 B out (224 + 224 for XYSR)
    n = 500_000 if args.quick else 2_000_000
    for kind, name, recf in ((0, "xyah", 72), (1, "xysr", 56), (2, "xywh", 72)):
        if not want("kf"):
            break
        z = torch.empty((n, 4), device=dev)
        z[:, 0].uniform_(0, 1920); z[:, 1].uniform_(0, 1080)
        if kind == 1:
            z[:, 2].uniform_(2000, 30000); z[:, 3].fill_(0.45)
        elif kind == 0:
            z[:, 2].fill_(0.45); z[:, 3].uniform_(40, 260)
        else:
            z[:, 2].uniform_(20, 120); z[:, 3].uniform_(40, 260)
        recs = torch.empty((n, recf), device=dev)
        api.check(lib.mot_kf_initiate(kind, recs.data_ptr(), z.data_ptr(), n, st))
        api.check(lib.mot_kf_predict(kind, recs.data_ptr(), None, n, 1.0, 1.0, st))
        base = recs.clone()
        ms = timeit(lambda: api.check(lib.mot_kf_predict(kind, recs.data_ptr(), None, n, 1.0, 1.0, st)))
        by = n * recf * 4 * 2
        emit(f"kf_predict_{name}", ms, tracks=n, algorithmic_bytes=by, achieved_gbs=by / ms[0] / 1e6,
             peak_gbs=HBM, frac=by / ms[0] / 1e6 / HBM, bound="hbm")
        recs.copy_(base)

        def upd():
            api.check(lib.mot_kf_update(kind, recs.data_ptr(), z.data_ptr(), None, n, None, st))
        ms = timeit(upd, iters=5, warm=1)          # repeated updates keep P positive definite
        by = n * (recf * 4 * 2 + 16)
        emit(f"kf_update_{name}", ms, tracks=n, algorithmic_bytes=by, achieved_gbs=by / ms[0] / 1e6,
             peak_gbs=HBM, frac=by / ms[0] / 1e6 / HBM, bound="hbm")
        del recs, base, z

    # ---------------- IoU cost matrix (output stream dominates: 4 B per pair)
    for (N, M) in ((256, 512), (2048, 2048), (8192, 8192)):
        if not want("iou") or (args.quick and N > 2048):
            continue
        c = torch.rand((N + M, 2), device=dev) * 8000
        w = torch.rand((N + M, 2), device=dev) * 100 + 40
        boxes = torch.cat([c, c + w], 1).contiguous()
        a, b = boxes[:N].contiguous(), boxes[N:].contiguous()
        conf = torch.rand(M, device=dev)
        out = torch.empty((N, M), device=dev)
        ms = timeit(lambda: api.check(lib.mot_cost_iou(a.data_ptr(), N, b.data_ptr(), M, conf.data_ptr(), out.data_ptr(), M, 2, st)))
        by = 4 * N * M + 16 * (N + M) + 4 * M
        emit(f"iou_cost_fused_{N}x{M}", ms, algorithmic_bytes=by, achieved_gbs=by / ms[0] / 1e6, peak_gbs=HBM,
             frac=by / ms[0] / 1e6 / HBM, bound="hbm", note="fits in L2 below ~5000^2: L2-resident, not an HBM number" if N < 5000 else "")

    # ---------------- exact assignment: microseconds per solve
    if want("lap"):
        import oracle_lib as O
        dets = synth.bytetrack_stream(0, n_frames=2)
        cost = O.fuse_score(O.iou_distance(dets[0, :256, :4], dets[1, :, :4]), dets[1, :, 4])
        P = 296
        costs = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(cost, (P,) + cost.shape))).to(dev).contiguous()
        r2c = torch.empty((P, 256), dtype=torch.int32, device=dev)
        c2r = torch.empty((P, 512), dtype=torch.int32, device=dev)
        ms = timeit(lambda: api.check(lib.mot_lap_batch_device(costs.data_ptr(), 256 * 512, P, None, None, 256, 512, 512, 0.8,
                                                               r2c.data_ptr(), c2r.data_ptr(), st)))
        ref = O.linear_assignment(cost, 0.8)
        ok = bool(np.array_equal(r2c[0].cpu().numpy(), ref[0]))
        emit("lap_256x512_c2_frame", ms, problems=P, us_per_solve=1e3 * ms[0] / P, matches_oracle=ok,
             bound="latency (no HBM/tensor roofline)", matched=int((ref[0] >= 0).sum()))
        one = costs[:1].contiguous()
        ms = timeit(lambda: api.check(lib.mot_lap_device(one.data_ptr(), 256, 512, 512, 0.8, r2c.data_ptr(), c2r.data_ptr(), st)))
        emit("lap_256x512_single", ms, us_per_solve=1e3 * ms[0], bound="latency")
        rng = np.random.default_rng(0)
        dense = (rng.random((256, 512)) * 0.7).astype(np.float32)          # adversarial: every pair is a candidate
        dd = torch.from_numpy(dense).to(dev)
        ms = timeit(lambda: api.check(lib.mot_lap_device(dd.data_ptr(), 256, 512, 512, 0.8, r2c.data_ptr(), c2r.data_ptr(), st)), iters=3, warm=1)
        refd = O.linear_assignment(dense, 0.8)
        emit("lap_256x512_dense_adversarial", ms, us_per_solve=1e3 * ms[0],
             matches_oracle=bool(np.array_equal(r2c[0].cpu().numpy(), refd[0])), bound="latency")

    # ---------------- cosine embedding cost on tensor cores
    for (N, M, D) in ((1024, 1024, 512), (4096, 4096, 512)):
        if not want("cos") or (args.quick and N > 1024):
            continue
        t = torch.randn((N, D), device=dev)
        d = torch.randn((M, D), device=dev)
        out = torch.empty((N, M), device=dev)
        ms = timeit(lambda: api.check(lib.mot_cost_cosine(t.data_ptr(), N, d.data_ptr(), M, D, out.data_ptr(), M, st)))
        useful = 2.0 * N * M * D
        emit(f"cosine_{N}x{M}x{D}", ms, useful_flop=useful, executed_bf16_flop=6 * useful,
             useful_tflops=useful / ms[0] / 1e9, executed_tflops=6 * useful / ms[0] / 1e9, peak_tflops=TENSOR,
             frac_executed=6 * useful / ms[0] / 1e9 / TENSOR, bound="tensor",
             note="time includes the fp32->3xbf16 split pre-pass and stream-ordered allocation")


if __name__ == "__main__":
    main()
