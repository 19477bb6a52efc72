#!/usr/bin/env python
"""Builds tests/golden/mot17_mini_dets.npz from the reference's own data asset
(/root/reference/assets/MOT17-mini/train/*/det/det.txt, public MOT17 FRCNN detections) the way
MOT17Dataset::load_detections parses it (reference src/data/mot17_dataset.cpp:176-209:
`frame,-1,x,y,w,h,conf` -> [x1, y1, x1+w, y1+h, conf, cls=0]), and records digests of what the
CPU oracle's SORT (BASELINE configs[0]: Sort(0.3, 1, 50, 3, 0.3)) and ByteTrack produce on it, so
that a later change of the oracle is noticed.  Run in the build container only (needs
/root/reference); the GPU box reads the committed .npz.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
REF = "/root/reference/assets/MOT17-mini/train"


def load(path):
    rows = []
    for line in open(path):
        v = [np.float32(x) for x in line.strip().split(",") if x]
        if len(v) < 7:
            continue
        x1, y1, w, h, conf = v[2], v[3], v[4], v[5], v[6]
        rows.append((int(v[0]), x1, y1, np.float32(x1 + w), np.float32(y1 + h), conf, np.float32(0)))
    rows.sort(key=lambda r: r[0])           # stable: keeps file order inside a frame
    frames = np.array([r[0] for r in rows], np.int32)
    dets = np.array([r[1:] for r in rows], np.float32)
    return frames, dets


def digest(tracker, frames, dets):
    h = hashlib.sha256()
    n_rows = 0
    for f in range(int(frames.min()), int(frames.max()) + 1):
        out = tracker.update(dets[frames == f])
        h.update(np.int32(f).tobytes())
        h.update(out.tobytes())
        n_rows += len(out)
    return h.hexdigest(), n_rows


def main():
    import oracle_lib as O
    out = {}
    for seq in sorted(os.listdir(REF)):
        frames, dets = load(os.path.join(REF, seq, "det", "det.txt"))
        key = seq.replace("-", "_")
        out[key + "_frames"] = frames
        out[key + "_dets"] = dets
        d, n = digest(O.Sort(0.3, 1, 50, 3, 0.3), frames, dets)
        out[key + "_sort_digest"] = np.array([d, str(n)])
        d, n = digest(O.ByteTrack(0.3, 30, 50, 3, 0.3, 0.1, 0.45, 0.8, 30, 30), frames, dets)
        out[key + "_bytetrack_digest"] = np.array([d, str(n)])
        print(seq, "frames", frames.min(), "-", frames.max(), "dets", len(dets), "max/frame",
              np.bincount(frames).max(), "sort rows", out[key + "_sort_digest"][1], "bytetrack rows", n)
    np.savez_compressed(os.path.join(HERE, "mot17_mini_dets.npz"), **out)


if __name__ == "__main__":
    main()
