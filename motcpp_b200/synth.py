"""Synthetic detection streams for the BASELINE.json configs.

All randomness is numpy PCG64 seeded with ``1000 * config + stream_id``.  Values are continuous
fp32, so exact cost ties have probability ~0.  Nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

import numpy as np

# canvas presets named in SURVEY.md 8d (headline first)
CANVAS = {"headline": (3840, 2160), "crowded": (1920, 1080), "sparse": (7680, 4320)}


def _rng(config: int, stream_id: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(1000 * config + stream_id))


def _boxes(cx, cy, w, h):
    return np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], axis=-1)


def bytetrack_stream(stream_id: int = 0, n_frames: int = 1050, n_obj: int = 256, n_clutter: int = 192,
                     n_low: int = 64, canvas=(3840, 2160), config: int = 2) -> np.ndarray:
    """C2: (n_frames, n_obj + n_clutter + n_low, 6) float32 rows [x1,y1,x2,y2,conf,cls].

    256 objects, w~U(40,120), h=2.2w, centre~U(canvas), velocity~N(0,3^2) px/frame, reflecting
    borders.  Each frame: true boxes with N(0,2^2) jitter and conf~U(.55,.99); high-confidence
    clutter (uniform position, conf~U(.46,.99)); low-confidence boxes near random objects
    (jitter N(0,6^2), conf~U(.11,.44)); rows in a seeded random permutation; cls = 0.
    """
    rng = _rng(config, stream_id)
    W, H = canvas
    w = rng.uniform(40, 120, n_obj)
    h = 2.2 * w
    cx = rng.uniform(0, W, n_obj)
    cy = rng.uniform(0, H, n_obj)
    vx = rng.normal(0, 3, n_obj)
    vy = rng.normal(0, 3, n_obj)
    n_det = n_obj + n_clutter + n_low
    out = np.empty((n_frames, n_det, 6), np.float32)
    for t in range(n_frames):
        cx += vx
        cy += vy
        flip = (cx < 0) | (cx > W)
        vx[flip] = -vx[flip]
        cx = np.clip(cx, 0, W)
        flip = (cy < 0) | (cy > H)
        vy[flip] = -vy[flip]
        cy = np.clip(cy, 0, H)
        true = _boxes(cx, cy, w, h) + rng.normal(0, 2, (n_obj, 4))
        true_c = rng.uniform(0.55, 0.99, n_obj)
        cw = rng.uniform(40, 120, n_clutter)
        clutter = _boxes(rng.uniform(0, W, n_clutter), rng.uniform(0, H, n_clutter), cw, 2.2 * cw)
        clutter_c = rng.uniform(0.46, 0.99, n_clutter)
        pick = rng.integers(0, n_obj, n_low)
        low = _boxes(cx[pick], cy[pick], w[pick], h[pick]) + rng.normal(0, 6, (n_low, 4))
        low_c = rng.uniform(0.11, 0.44, n_low)
        boxes = np.concatenate([true, clutter, low], 0)
        conf = np.concatenate([true_c, clutter_c, low_c], 0)
        perm = rng.permutation(n_det)
        out[t, :, :4] = boxes[perm]
        out[t, :, 4] = conf[perm]
        out[t, :, 5] = 0.0
    return out


def stress_stream(stream_id: int = 0, n_frames: int = 300, n_obj: int = 48, canvas=(960, 540),
                  config: int = 9) -> np.ndarray:
    """Parity-stress generator (not timed): missed detections (10 %), confidence drops into the low
    band (5 %), births/deaths and crossing trajectories on a small canvas, so that the lost /
    re-found / duplicate-removal paths of the state machines are exercised.  Returns
    (n_frames, max_dets, 6) padded with conf = 0 rows plus a per-frame count array.
    """
    rng = _rng(config, stream_id)
    W, H = canvas
    w = rng.uniform(30, 90, n_obj)
    h = 2.2 * w
    cx = rng.uniform(0, W, n_obj)
    cy = rng.uniform(0, H, n_obj)
    vx = rng.normal(0, 4, n_obj)
    vy = rng.normal(0, 2, n_obj)
    alive = rng.random(n_obj) < 0.7
    max_dets = n_obj + 16
    out = np.zeros((n_frames, max_dets, 6), np.float32)
    counts = np.zeros(n_frames, np.int32)
    for t in range(n_frames):
        cx += vx
        cy += vy
        flip = (cx < 0) | (cx > W)
        vx[flip] = -vx[flip]
        flip = (cy < 0) | (cy > H)
        vy[flip] = -vy[flip]
        toggle = rng.random(n_obj) < 0.01          # births / deaths
        alive ^= toggle
        rows = []
        for k in np.nonzero(alive)[0]:
            if rng.random() < 0.10:
                continue                           # missed detection
            conf = rng.uniform(0.5, 0.99)
            if rng.random() < 0.05:
                conf = rng.uniform(0.12, 0.44)     # drop into the low band
            b = _boxes(cx[k], cy[k], w[k], h[k]) + rng.normal(0, 1.5, 4)
            rows.append([*b, conf, 0.0])
        for _ in range(int(rng.integers(0, 6))):   # clutter, any confidence
            cw = rng.uniform(30, 90)
            b = _boxes(rng.uniform(0, W), rng.uniform(0, H), cw, 2.2 * cw)
            rows.append([*b, rng.uniform(0.05, 0.99), 0.0])
        rows = np.asarray(rows, np.float32).reshape(-1, 6)
        rows = rows[rng.permutation(len(rows))][:max_dets]
        out[t, :len(rows)] = rows
        counts[t] = len(rows)
    return out, counts


def embeddings_stream(stream_id: int, n_frames: int, n_obj: int = 1024, dim: int = 512,
                      canvas=(7680, 4320), config: int = 3):
    """C3: BoT-SORT stream.  Returns dets (T, n_obj, 6) all conf > 0.6 and embs (T, n_obj, dim):
    identity vectors e_k ~ N(0, I) normalised, detection emb = normalise(e_k + 0.35 N(0, I))."""
    rng = _rng(config, stream_id)
    W, H = canvas
    w = rng.uniform(40, 120, n_obj)
    h = 2.2 * w
    cx = rng.uniform(0, W, n_obj)
    cy = rng.uniform(0, H, n_obj)
    vx = rng.normal(0, 3, n_obj)
    vy = rng.normal(0, 3, n_obj)
    ident = rng.normal(0, 1, (n_obj, dim))
    ident /= np.linalg.norm(ident, axis=1, keepdims=True)
    dets = np.empty((n_frames, n_obj, 6), np.float32)
    embs = np.empty((n_frames, n_obj, dim), np.float32)
    for t in range(n_frames):
        cx += vx
        cy += vy
        flip = (cx < 0) | (cx > W)
        vx[flip] = -vx[flip]
        flip = (cy < 0) | (cy > H)
        vy[flip] = -vy[flip]
        b = _boxes(cx, cy, w, h) + rng.normal(0, 2, (n_obj, 4))
        e = ident + 0.35 * rng.normal(0, 1, (n_obj, dim))
        e /= np.linalg.norm(e, axis=1, keepdims=True)
        perm = rng.permutation(n_obj)
        dets[t, :, :4] = b[perm]
        dets[t, :, 4] = rng.uniform(0.61, 0.99, n_obj)
        dets[t, :, 5] = 0.0
        embs[t] = e[perm]
    return dets, embs


def ocsort_stream(stream_id: int = 0, n_frames: int = 250, n_obj: int = 2048, canvas=(15360, 8640),
                  config: int = 4) -> np.ndarray:
    """C4: (n_frames, n_obj, 6); conf ~ U(0.25, 0.99)."""
    rng = _rng(config, stream_id)
    W, H = canvas
    w = rng.uniform(40, 120, n_obj)
    h = 2.2 * w
    cx = rng.uniform(0, W, n_obj)
    cy = rng.uniform(0, H, n_obj)
    vx = rng.normal(0, 3, n_obj)
    vy = rng.normal(0, 3, n_obj)
    out = np.empty((n_frames, n_obj, 6), np.float32)
    for t in range(n_frames):
        cx += vx
        cy += vy
        flip = (cx < 0) | (cx > W)
        vx[flip] = -vx[flip]
        flip = (cy < 0) | (cy > H)
        vy[flip] = -vy[flip]
        b = _boxes(cx, cy, w, h) + rng.normal(0, 2, (n_obj, 4))
        perm = rng.permutation(n_obj)
        out[t, :, :4] = b[perm]
        out[t, :, 4] = rng.uniform(0.25, 0.99, n_obj)
        out[t, :, 5] = 0.0
    return out


def stress_stream_reid(stream_id: int = 0, n_frames: int = 200, n_obj: int = 40, dim: int = 32, canvas=(960, 540),
                       noise: float = 0.5, config: int = 8):
    """Parity-stress generator with ReID embeddings for BoT-SORT (not timed): the stress_stream scenario
    (misses, confidence drops, births/deaths, crossings, clutter) plus one feature vector per detection:
    normalise(identity_k + noise * N(0, I)) for true objects (deliberately NOT unit length: a random scale is
    applied so the tracker's own normalisation is exercised), random vectors for clutter.
    Returns dets (T, max_dets, 6), counts (T,), embs (T, max_dets, dim)."""
    rng = _rng(config, stream_id)
    W, H = canvas
    w = rng.uniform(30, 90, n_obj)
    h = 2.2 * w
    cx = rng.uniform(0, W, n_obj)
    cy = rng.uniform(0, H, n_obj)
    vx = rng.normal(0, 4, n_obj)
    vy = rng.normal(0, 2, n_obj)
    ident = rng.normal(0, 1, (n_obj, dim))
    ident /= np.linalg.norm(ident, axis=1, keepdims=True)
    alive = rng.random(n_obj) < 0.7
    max_dets = n_obj + 16
    out = np.zeros((n_frames, max_dets, 6), np.float32)
    embs = np.zeros((n_frames, max_dets, dim), np.float32)
    counts = np.zeros(n_frames, np.int32)
    for t in range(n_frames):
        cx += vx
        cy += vy
        flip = (cx < 0) | (cx > W)
        vx[flip] = -vx[flip]
        flip = (cy < 0) | (cy > H)
        vy[flip] = -vy[flip]
        alive ^= rng.random(n_obj) < 0.01
        rows, feats = [], []
        for k in np.nonzero(alive)[0]:
            if rng.random() < 0.10:
                continue
            conf = rng.uniform(0.55, 0.99)
            if rng.random() < 0.08:
                conf = rng.uniform(0.12, 0.5)
            b = _boxes(cx[k], cy[k], w[k], h[k]) + rng.normal(0, 1.5, 4)
            rows.append([*b, conf, 0.0])
            e = ident[k] + noise * rng.normal(0, 1, dim) / np.sqrt(dim)
            feats.append(e / np.linalg.norm(e) * rng.uniform(0.5, 2.0))
        for _ in range(int(rng.integers(0, 6))):
            cw = rng.uniform(30, 90)
            b = _boxes(rng.uniform(0, W), rng.uniform(0, H), cw, 2.2 * cw)
            rows.append([*b, rng.uniform(0.05, 0.99), 0.0])
            feats.append(rng.normal(0, 1, dim))
        rows = np.asarray(rows, np.float32).reshape(-1, 6)
        feats = np.asarray(feats, np.float32).reshape(-1, dim)
        perm = rng.permutation(len(rows))[:max_dets]
        out[t, :len(perm)] = rows[perm]
        embs[t, :len(perm)] = feats[perm]
        counts[t] = len(perm)
    return out, counts, embs


def strongsort_stream(stream_id: int, n_frames: int, n_obj: int = 192, n_clutter: int = 64, dim: int = 128,
                      canvas=(3840, 2160), noise: float = 0.3, config: int = 6):
    """StrongSORT workload: n_obj moving objects, each seen TWICE per frame - a confident box and an
    overlapping low-confidence near-duplicate - plus n_clutter random boxes; 2 n_obj + n_clutter detections per frame.
    The duplicate matters: with no confirmed track the reference feeds every tentative track to the IoU stage twice and
    deletes it unless both copies find a detection (oracle/strongsort.cpp q1), so a scene with one clean box per object
    never confirms anything.  Embeddings: unit identity vectors + noise of total norm `noise`, rescaled by U(0.5, 2) so the
    tracker's own normalisation is exercised; clutter gets random vectors.  Returns dets (T, D, 6), embs (T, D, dim)."""
    rng = _rng(config, stream_id)
    W, H = canvas
    w = rng.uniform(40, 120, n_obj)
    h = 2.2 * w
    cx = rng.uniform(0, W, n_obj)
    cy = rng.uniform(0, H, n_obj)
    vx = rng.normal(0, 3, n_obj)
    vy = rng.normal(0, 3, n_obj)
    ident = rng.normal(0, 1, (n_obj, dim))
    ident /= np.linalg.norm(ident, axis=1, keepdims=True)
    D = 2 * n_obj + n_clutter
    dets = np.zeros((n_frames, D, 6), np.float32)
    embs = np.empty((n_frames, D, dim), np.float32)
    sigma = noise / np.sqrt(dim)
    for t in range(n_frames):
        cx += vx
        cy += vy
        flip = (cx < 0) | (cx > W)
        vx[flip] = -vx[flip]
        flip = (cy < 0) | (cy > H)
        vy[flip] = -vy[flip]
        b = np.empty((D, 4))
        b[:n_obj] = _boxes(cx, cy, w, h) + rng.normal(0, 2, (n_obj, 4))
        b[n_obj:2 * n_obj] = _boxes(cx, cy, w, h) + rng.normal(0, 4, (n_obj, 4))
        cw = rng.uniform(40, 120, n_clutter)
        b[2 * n_obj:] = _boxes(rng.uniform(0, W, n_clutter), rng.uniform(0, H, n_clutter), cw, 2.2 * cw)
        e = np.empty((D, dim))
        e[:n_obj] = ident + sigma * rng.normal(0, 1, (n_obj, dim))
        e[n_obj:2 * n_obj] = ident + sigma * rng.normal(0, 1, (n_obj, dim))
        e[2 * n_obj:] = rng.normal(0, 1, (n_clutter, dim))
        e *= rng.uniform(0.5, 2.0, (D, 1))
        conf = np.concatenate([rng.uniform(0.6, 0.99, n_obj), rng.uniform(0.15, 0.45, n_obj), rng.uniform(0.3, 0.9, n_clutter)])
        perm = rng.permutation(D)
        dets[t, :, :4] = b[perm]
        dets[t, :, 4] = conf[perm]
        embs[t] = e[perm]
    return dets, embs
