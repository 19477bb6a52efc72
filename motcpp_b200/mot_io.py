"""MOT-format I/O either side of update(), mirroring the reference's readers / writer:
    MOT17Dataset::load_detections   src/data/mot17_dataset.cpp:149-238
    MOT17Dataset::load_embeddings   src/data/mot17_dataset.cpp:240-289
    utils::convert_to_mot_format    include/motcpp/utils/mot_format.hpp:20-50
    utils::write_mot_results        include/motcpp/utils/mot_format.hpp:57-74
Host-side text handling only (the reference's is C++ iostreams); it exists so that real sequences can be replayed
through the engine and scored with TrackEval, and it is what tools/motb200_eval.py is made of."""
from __future__ import annotations

import os
from typing import Dict

import numpy as np


def _floats(tokens):
    vals = []
    for t in tokens:
        try:
            vals.append(np.float32(t))          # std::stof
        except ValueError:
            break                                # mot17_dataset.cpp:181-185: stop at the first bad token
    return vals


def load_detections(det_path: str) -> Dict[int, np.ndarray]:
    """frame_id -> (n, 6) float32 [x1, y1, x2, y2, conf, cls].  Comma-separated files are MOT17 `frame,-1,x,y,w,h,conf[,cls]`
    (tlwh -> xyxy as x + w, y + h in fp32); space-separated files are `frame x1 y1 x2 y2 conf cls`.  The format is decided by
    the first line; empty and `#` lines and rows with fewer than 7 values are skipped; a missing file gives {}."""
    frames: Dict[int, list] = {}
    if not os.path.exists(det_path):
        return {}
    with open(det_path) as f:
        lines = f.read().split("\n")
    comma = bool(lines) and "," in lines[0]
    for line in lines:
        if not line or line[0] == "#":
            continue
        if comma:
            v = _floats(line.split(","))
            if len(v) < 7:
                continue
            x1, y1, w, h, conf = v[2], v[3], v[4], v[5], v[6]
            cls = v[7] if len(v) > 7 else np.float32(0)
            row = (x1, y1, np.float32(x1 + w), np.float32(y1 + h), conf, cls)
        else:
            v = _floats(line.split())
            if len(v) < 7:
                continue
            row = tuple(v[1:7])
        frames.setdefault(int(v[0]), []).append(row)
    return {k: np.array(r, np.float32).reshape(-1, 6) for k, r in frames.items()}


def load_embeddings(emb_path: str, detections: Dict[int, np.ndarray]) -> Dict[int, np.ndarray]:
    """One whitespace-separated vector per line, line k belonging to the k-th detection.  The reference walks its
    std::unordered_map of detections in hash order to number the detections (:251-256); here frames are numbered in
    ASCENDING frame order, which is the order the pre-generated files are written in."""
    out: Dict[int, list] = {}
    if not os.path.exists(emb_path):
        return {}
    order = [(f, i) for f in sorted(detections) for i in range(len(detections[f]))]
    k = 0
    with open(emb_path) as fh:
        for line in fh:
            line = line.strip()
            if not line or line[0] == "#":
                continue
            if k >= len(order):
                break
            v = _floats(line.split())
            if not v:
                continue
            out.setdefault(order[k][0], []).append(v)
            k += 1
    return {f: np.array(r, np.float32) for f, r in out.items()}


def convert_to_mot_format(tracks: np.ndarray, frame_id: int) -> np.ndarray:
    """(N, 8) [x1,y1,x2,y2,id,conf,cls,det_ind] -> (N, 10) [frame,id,x1,y1,w,h,conf,-1,-1,-1] (mot_format.hpp:20-50)."""
    t = np.asarray(tracks, np.float32).reshape(-1, 8)
    out = np.full((t.shape[0], 10), -1.0, np.float32)
    out[:, 0] = np.float32(frame_id)
    out[:, 1] = t[:, 4]
    out[:, 2] = t[:, 0]
    out[:, 3] = t[:, 1]
    out[:, 4] = t[:, 2] - t[:, 0]
    out[:, 5] = t[:, 3] - t[:, 1]
    out[:, 6] = t[:, 5]
    return out


def format_mot_rows(mot_results: np.ndarray) -> str:
    """The text write_mot_results appends (mot_format.hpp:57-74): ints by truncation towards zero, conf with 6 decimals."""
    m = np.asarray(mot_results, np.float32).reshape(-1, 10)
    lines = []
    for r in m:
        i = [int(np.trunc(x)) for x in r]
        lines.append(f"{i[0]},{i[1]},{i[2]},{i[3]},{i[4]},{i[5]},{float(r[6]):.6f},{i[7]},{i[8]},{i[9]}\n")
    return "".join(lines)


def write_mot_results(output_path: str, mot_results: np.ndarray) -> None:
    """Appends to output_path, creating its directory (mot_format.hpp:57-74)."""
    d = os.path.dirname(output_path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(output_path, "a") as f:
        f.write(format_mot_rows(mot_results))
