#!/usr/bin/env python
"""Per-phase cycle shares of bytetrack_step_kernel on the C2 workload (mot_engine_profile).  Diagnostics, not a benchmark:
the counters cost one clock read + one atomic per phase and frame."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motcpp_b200 import _lib, api, synth

NAMES = ["A dets+split", "B pool lists", "C pred boxes", "D1 candidates", "D2 components", "D3 grouping", "D4 solves", "harvest",
         "E kalman", "F second assoc", "G unconfirmed", "H new tracks", "I+J lists", "K duplicates", "L output"]

def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 296
    T, warm = 20, 150
    B = 8
    base = np.stack([synth.bytetrack_stream(b, n_frames=warm + T) for b in range(B)], 1)
    dets = base[:, np.arange(S) % B].copy()
    counts = np.full((warm + T, S), 512, np.int32)
    args = dict(det_thresh=0.3, max_age=30, max_obs=50, min_hits=3, iou_threshold=0.3, min_conf=0.1, track_thresh=0.45,
                match_thresh=0.8, track_buffer=30, frame_rate=30)
    eng = api.Engine(_lib.TRACKER_BYTETRACK, S, 1536, 512, **args)
    for t0 in range(0, warm, 50):
        eng.update(dets[t0:t0 + 50], counts[t0:t0 + 50], ld_out=512)
    lib = _lib.load()
    api.check(lib.mot_engine_profile(eng._h, 1, None))
    eng.update(dets[warm:], counts[warm:], ld_out=512)
    cyc = np.zeros(32, np.uint64)
    api.check(lib.mot_engine_profile(eng._h, 0, cyc.ctypes.data))
    raw = cyc.copy()
    extra = raw[16:].astype(np.float64) / (S * T)
    sub = raw[20:28].astype(np.float64) / (S * T)      # F: 20 solves, 21 candidates, 22 components, 23 grouping; G: 24..27 likewise
    cyc = raw[:16].copy()
    cyc[9] += raw[20:24].sum()                         # the association's own sub-phases belong to F / G
    cyc[10] += raw[24:28].sum()
    tot = float(cyc.sum())
    per_frame = cyc.astype(np.float64) / (S * T)
    print("streams %d, frames %d: %.0f cycles per frame per CTA (%.1f us at 1.965 GHz)" % (S, T, tot / (S * T), tot / (S * T) / 1965.0))
    for k, n in enumerate(NAMES):
        print("  %-16s %8.0f cycles  %5.1f %%" % (n, per_frame[k], 100.0 * cyc[k] / tot))
    print("  F inside block_lap: candidates %.0f, components %.0f, grouping %.0f, solves %.0f cycles" % (sub[1], sub[2], sub[3], sub[0]))
    print("  G inside block_lap: candidates %.0f, components %.0f, grouping %.0f, solves %.0f cycles" % (sub[5], sub[6], sub[7], sub[4]))
    print("  D4 split (with extra barriers): class lists %.0f, trivial %.0f, teams %.0f, warps = D4 solves above" % tuple(raw[28:31].astype(np.float64) / (S * T)))
    print("  D1 split: init %.0f, grid build %.0f, pair collection %.0f, pair evaluation %.0f cycles" % (extra[3], extra[0], extra[1], extra[2]))
    print(json.dumps({"streams": S, "frames": T, "cycles_per_frame": tot / (S * T), "share": {n: float(cyc[k] / tot) for k, n in enumerate(NAMES)}}))

if __name__ == "__main__":
    main()
