#!/usr/bin/env python
"""Replay MOT sequences through the B200 engine and write MOT-format result files - the role of the reference's
tools/motcpp_eval.cpp (:19-468) for the front-ends on the accelerated path, with the argument sets it passes
(:118-246).  Every sequence becomes one stream; ALL frames of ALL sequences go through the engine in one batched call.

    python tools/motb200_eval.py <mot_root> <output_dir> [sort|bytetrack|ocsort|botsort|strongsort] [det_emb_root model reid]

<mot_root>/<seq>/det/det.txt is read unless det_emb_root is given, in which case detections come from
<det_emb_root>/dets/<model>/<seq>.txt and embeddings from <det_emb_root>/embs/<model>/<reid>/<seq>.txt (the layout
motcpp_eval expects).  Unlike motcpp_eval, every frame that has detections is processed (no "ablation offset")."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motcpp_b200 import _lib, api, mot_io  # noqa: E402

PRESETS = {   # tools/motcpp_eval.cpp:118-246
    "sort": (_lib.TRACKER_SORT, dict(det_thresh=0.3, max_age=1, max_obs=50, min_hits=3, iou_threshold=0.3)),
    "bytetrack": (_lib.TRACKER_BYTETRACK, dict(det_thresh=0.3, max_age=30, max_obs=50, min_hits=3, iou_threshold=0.3, min_conf=0.1,
                                               track_thresh=0.45, match_thresh=0.8, track_buffer=30, frame_rate=30)),
    "ocsort": (_lib.TRACKER_OCSORT, dict(det_thresh=0.2, max_age=30, max_obs=50, min_hits=3, iou_threshold=0.3, min_conf=0.1, delta_t=3,
                                         inertia=0.2, use_byte=0, q_xy_scaling=0.01, q_s_scaling=0.0001)),
    "botsort": (_lib.TRACKER_BOTSORT, dict(det_thresh=0.3, max_age=30, max_obs=50, min_hits=3, iou_threshold=0.3, track_high_thresh=0.6,
                                           track_low_thresh=0.1, new_track_thresh=0.7, track_buffer=30, match_thresh=0.8,
                                           proximity_thresh=0.5, appearance_thresh=0.25, frame_rate=30, fuse_first_associate=0,
                                           with_reid=1)),
    "strongsort": (_lib.TRACKER_STRONGSORT, dict(det_thresh=0.3, max_age=30, max_obs=50, min_hits=3, iou_threshold=0.3, min_conf=0.1,
                                                 max_cos_dist=0.2, max_iou_dist=0.7, n_init=3, nn_budget=100, mc_lambda=0.98,
                                                 ema_alpha=0.9)),
}


def replay(sequences, method, out_dir=None):
    """sequences: {name: (dets {frame: (n,6)}, embs {frame: (n,D)} or None)} -> {name: [(frame, tracks (M,8))]}; writes
    <out_dir>/<name>.txt when out_dir is given."""
    kind, params = PRESETS[method]
    names = sorted(sequences)
    S = len(names)
    T = max(max(sequences[n][0]) for n in names)
    d_need = max(max(len(v) for v in sequences[n][0].values()) for n in names)
    dim = 0
    if kind in (_lib.TRACKER_BOTSORT, _lib.TRACKER_STRONGSORT):
        dims = {v.shape[1] for n in names if sequences[n][1] for v in sequences[n][1].values()}
        dim = dims.pop() if len(dims) == 1 else 0
    d_max = 64 if d_need <= 64 else (512 if d_need <= 512 else 1024)
    cap = 256 if d_max == 64 else 1536
    dets = np.zeros((T, S, d_max, 6), np.float32)
    counts = np.zeros((T, S), np.int32)
    embs = np.zeros((T, S, d_max, dim), np.float32) if dim else None
    for s, n in enumerate(names):
        for f, v in sequences[n][0].items():
            dets[f - 1, s, :len(v)] = v
            counts[f - 1, s] = len(v)
            if dim and sequences[n][1] and f in sequences[n][1] and len(sequences[n][1][f]) == len(v):
                embs[f - 1, s, :len(v)] = sequences[n][1][f]
    eng = api.Engine(kind, S, cap, d_max, emb_dim=dim, **params)
    out, n_out = eng.update(dets, counts, ld_out=cap, embs=embs)
    eng.check()
    eng.close()
    results = {}
    for s, n in enumerate(names):
        rows = []
        path = os.path.join(out_dir, n + ".txt") if out_dir else None
        if path and os.path.exists(path):
            os.remove(path)
        for f in sorted(sequences[n][0]):
            tr = out[f - 1, s, :n_out[f - 1, s]]
            rows.append((f, tr.copy()))
            if path:
                mot_io.write_mot_results(path, mot_io.convert_to_mot_format(tr, f))
        results[n] = rows
    return results


def main(argv):
    if len(argv) < 3:
        print(__doc__)
        return 1
    mot_root, out_dir = argv[1], argv[2]
    method = argv[3] if len(argv) > 3 else "bytetrack"
    det_emb_root, model, reid = (argv[4:7] + [None] * 3)[:3] if len(argv) > 4 else (None, None, None)
    seqs = {}
    for name in sorted(os.listdir(mot_root)):
        if not os.path.isdir(os.path.join(mot_root, name)):
            continue
        det_path = os.path.join(det_emb_root, "dets", model, name + ".txt") if det_emb_root else os.path.join(mot_root, name, "det", "det.txt")
        d = mot_io.load_detections(det_path)
        if not d:
            continue
        e = mot_io.load_embeddings(os.path.join(det_emb_root, "embs", model, reid, name + ".txt"), d) if det_emb_root and reid else None
        seqs[name] = (d, e)
    if not seqs:
        print("no sequences with detections under", mot_root)
        return 1
    res = replay(seqs, method, out_dir)
    for n, rows in res.items():
        print(f"{n}: {len(rows)} frames, {sum(len(r[1]) for r in rows)} track rows -> {os.path.join(out_dir, n + '.txt')}")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
