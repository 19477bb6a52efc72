#!/usr/bin/env python
"""Longer randomized parity soak of the StrongSORT engine against the oracle (not part of the test suite):
many seeds x parameter sets, outputs + full state + features compared bit for bit."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from motcpp_b200 import synth  # noqa: E402
from test_strongsort import ARGS, LOOSE, _engine_vs_oracle  # noqa: E402

t0 = time.time()
rows = 0
sets = [LOOSE, {**ARGS, "nn_budget": 12}, {**LOOSE, "mc_lambda": 0.5, "max_cos_dist": 0.6, "nn_budget": 3},
        {**ARGS, "n_init": 2, "max_age": 4, "ema_alpha": 0.5, "nn_budget": 40, "max_cos_dist": 0.35}]
for k, args in enumerate(sets):
    for dim, noise in ((16, 0.2), (64, 0.35), (128, 0.15)):
        streams = [synth.stress_stream_reid(100 + 10 * k + s, n_frames=150, n_obj=28, dim=dim, noise=noise) for s in range(6)]
        rows += _engine_vs_oracle(O, streams, args, 256, 64, dim, T_chunk=25, check_state_every=1)
        print(f"set {k} dim {dim}: ok ({rows} rows so far, {time.time() - t0:.0f} s)", flush=True)
d, e = synth.strongsort_stream(3, 30, n_obj=60, n_clutter=20, dim=64)
c = np.full(30, d.shape[1], np.int32)
rows += _engine_vs_oracle(O, [(d, c, e)], {**ARGS, "nn_budget": 25}, 1536, 512, 64, T_chunk=10, check_state_every=1)
print(f"strongsort_stream workload (60 objects seen twice): ok; total rows {rows}, {time.time() - t0:.0f} s")
