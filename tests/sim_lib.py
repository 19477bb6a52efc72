"""ctypes loader for tests/cpusim/libmotb200_cpusim.so: the product's kernel sources compiled
against the SIMT emulator.  TEST INFRASTRUCTURE (kernel-logic unit tests without a GPU)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SIM_DIR = os.path.join(HERE, "cpusim")
f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_LIBS = {}
# "" = the product's constants; "jvblock" = the same sources built with MOT_OC_JVMAX = MOT_SS_JVMAX = 24, so that the
# CTA-wide dense LAPJV (csrc/jv_block_device.cuh, the path for rows + columns > 384 in the product) runs at emulator sizes
VARIANT = ""


class variant:
    """with sim_lib.variant("jvblock"): ... - objects must be created AND used inside the block"""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        global VARIANT
        self.prev, VARIANT = VARIANT, self.name

    def __exit__(self, *exc):
        global VARIANT
        VARIANT = self.prev


def lib():
    _LIB = _LIBS.get(VARIANT)
    if _LIB is None:
        subprocess.check_call(["make", "-C", SIM_DIR, "-s", "-j2"], stdout=subprocess.DEVNULL)
        S = C.CDLL(os.path.join(SIM_DIR, "libmotb200_cpusim%s.so" % ("_" + VARIANT if VARIANT else "")))
        S.sim_lap.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_float, i32p, i32p, C.c_int, C.c_int]
        S.sim_lap_jv.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_float, i32p, i32p, C.c_int, C.c_int]
        S.sim_grid_pairs.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float, i32p, i32p, C.c_int]
        S.sim_bt_create.argtypes = [C.c_int] * 4 + [C.c_float] * 3 + [C.c_int] * 2
        S.sim_bt_create.restype = C.c_void_p
        S.sim_bt_update.argtypes = [C.c_void_p, f32p, i32p, C.c_int, C.c_int, f32p, i32p, C.c_int, C.c_int, C.c_int]
        S.sim_bt_header.argtypes = [C.c_void_p, C.c_int, i32p]
        S.sim_bt_dump.argtypes = [C.c_void_p, C.c_int, C.c_int, f32p, C.c_int]
        S.sim_bt_dump.restype = C.c_int
        S.sim_bt_destroy.argtypes = [C.c_void_p]
        S.sim_sort_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_float]
        S.sim_sort_create.restype = C.c_void_p
        S.sim_sort_update.argtypes = [C.c_void_p, f32p, i32p, C.c_int, C.c_int, f32p, i32p, C.c_int, C.c_int]
        S.sim_sort_header.argtypes = [C.c_void_p, C.c_int, i32p]
        S.sim_sort_destroy.argtypes = [C.c_void_p]
        S.sim_acosf.argtypes = [C.c_float]
        S.sim_acosf.restype = C.c_float
        S.sim_oc_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float,
                                    C.c_int, C.c_float, C.c_float]
        S.sim_oc_create.restype = C.c_void_p
        S.sim_oc_set_asso.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        S.sim_oc_update.argtypes = [C.c_void_p, f32p, i32p, C.c_int, C.c_int, f32p, i32p, C.c_int, C.c_int]
        S.sim_oc_header.argtypes = [C.c_void_p, C.c_int, i32p]
        S.sim_oc_dump.argtypes = [C.c_void_p, C.c_int, f32p, C.c_int]
        S.sim_oc_dump.restype = C.c_int
        S.sim_oc_destroy.argtypes = [C.c_void_p]
        S.sim_deepoc_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float,
                                        C.c_float, C.c_float, C.c_int, C.c_int, C.c_float, C.c_float]
        S.sim_deepoc_create.restype = C.c_void_p
        S.sim_deepoc_update.argtypes = [C.c_void_p, f32p, i32p, C.c_void_p, C.c_int, C.c_int, f32p, i32p, C.c_int, C.c_int]
        S.sim_deepoc_header.argtypes = [C.c_void_p, C.c_int, i32p]
        S.sim_deepoc_dump.argtypes = [C.c_void_p, C.c_int, f32p, C.c_void_p, C.c_int]
        S.sim_deepoc_dump.restype = C.c_int
        S.sim_deepoc_destroy.argtypes = [C.c_void_p]
        S.sim_boost_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, C.c_int,
                                       C.c_float, C.c_int]
        S.sim_boost_create.restype = C.c_void_p
        S.sim_boost_update.argtypes = [C.c_void_p, f32p, i32p, C.c_int, C.c_int, f32p, i32p, C.c_int, C.c_int]
        S.sim_boost_header.argtypes = [C.c_void_p, C.c_int, i32p]
        S.sim_boost_dump.argtypes = [C.c_void_p, C.c_int, f32p, C.c_int]
        S.sim_boost_dump.restype = C.c_int
        S.sim_boost_destroy.argtypes = [C.c_void_p]
        S.sim_bot_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float,
                                     C.c_float, C.c_int, C.c_int, C.c_int]
        S.sim_bot_create.restype = C.c_void_p
        S.sim_bot_update.argtypes = [C.c_void_p, f32p, i32p, C.c_void_p, C.c_int, C.c_int, f32p, i32p, C.c_int, C.c_int]
        S.sim_bot_header.argtypes = [C.c_void_p, C.c_int, i32p]
        S.sim_bot_dump.argtypes = [C.c_void_p, C.c_int, C.c_int, f32p, C.c_void_p, C.c_int]
        S.sim_bot_dump.restype = C.c_int
        S.sim_bot_destroy.argtypes = [C.c_void_p]
        S.sim_ss_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_float, C.c_float]
        S.sim_ss_create.restype = C.c_void_p
        S.sim_ss_update.argtypes = [C.c_void_p, f32p, i32p, C.c_void_p, C.c_int, C.c_int, f32p, i32p, C.c_int, C.c_int]
        S.sim_ss_header.argtypes = [C.c_void_p, C.c_int, i32p]
        S.sim_ss_dump.argtypes = [C.c_void_p, C.c_int, f32p, C.c_void_p, C.c_int]
        S.sim_ss_dump.restype = C.c_int
        S.sim_ss_destroy.argtypes = [C.c_void_p]
        _LIBS[VARIANT] = _LIB = S
    return _LIB


def sim_lap_jv(cost, thresh, block=0, threads=128):
    """the reference-order dense LAPJV kernels under the emulator: block 0 = one warp, 1 = whole CTA with the work arrays in
    global scratch, 2 = whole CTA with everything in shared memory"""
    cost = np.ascontiguousarray(cost, np.float32)
    n, m = cost.shape
    r = np.full(max(n, 1), -7, np.int32)
    q = np.full(max(m, 1), -7, np.int32)
    assert lib().sim_lap_jv(cost, n, m, max(m, 1), float(thresh), r, q, block, threads) == 0
    return r[:n], q[:m]


def sim_lap(cost, thresh, e_cap=4096, threads=128):
    cost = np.ascontiguousarray(cost, np.float32)
    n, m = cost.shape
    r = np.full(max(n, 1), -7, np.int32)
    q = np.full(max(m, 1), -7, np.int32)
    lib().sim_lap(cost, n, m, max(m, 1), float(thresh), r, q, e_cap, threads)
    return r[:n], q[:m]


def sim_grid_pairs(rows, cols, big_w=3.0e38, big_h=3.0e38, use_roi=False, t=0.0, threads=64):
    """visits[i, j] = how often row i's grid query saw column j; also the length of the grid's overflow list"""
    rows, cols = np.ascontiguousarray(rows, np.float32), np.ascontiguousarray(cols, np.float32)
    n, m = rows.shape[0], cols.shape[0]
    visits = np.zeros((n, max(m, 1)), np.int32)
    n_big = np.zeros(1, np.int32)
    lib().sim_grid_pairs(rows, n, cols, m, float(big_w), float(big_h), int(use_roi), float(t), visits, n_big, threads)
    return visits[:, :m], int(n_big[0])


class SimByteTrack:
    def __init__(self, n_streams=1, cap=256, d_max=64, e_cap=4096, min_conf=0.1, track_thresh=0.45,
                 match_thresh=0.8, track_buffer=30, frame_rate=30):
        self.S, self.cap, self.d_max = n_streams, cap, d_max
        self.h = lib().sim_bt_create(n_streams, cap, d_max, e_cap, min_conf, track_thresh, match_thresh,
                                     track_buffer, frame_rate)

    def __del__(self):
        if getattr(self, "h", None):
            lib().sim_bt_destroy(self.h)
            self.h = None

    def update(self, dets, n_dets, threads=128, os_threads=1):
        """dets (T,S,ld,6), n_dets (T,S) -> out (T,S,cap,8), n_out (T,S)"""
        dets = np.ascontiguousarray(dets, np.float32)
        T, S, ld, _ = dets.shape
        n_dets = np.ascontiguousarray(n_dets, np.int32).reshape(T, S)
        out = np.zeros((T, S, self.cap, 8), np.float32)
        n_out = np.zeros((T, S), np.int32)
        lib().sim_bt_update(self.h, dets, n_dets, T, ld, out, n_out, self.cap, threads, os_threads)
        return out, n_out

    def header(self, s=0):
        h = np.zeros(16, np.int32)
        lib().sim_bt_header(self.h, s, h)
        return h

    def dump(self, s, which):
        buf = np.zeros((max(self.cap, 1), 78), np.float32)
        k = lib().sim_bt_dump(self.h, s, which, buf, self.cap)
        return buf[:k]


class SimSort:
    def __init__(self, n_streams=1, det_thresh=0.3, max_age=1, min_hits=3, iou_threshold=0.3):
        self.S, self.cap = n_streams, 256
        self.h = lib().sim_sort_create(n_streams, det_thresh, max_age, min_hits, iou_threshold)

    def __del__(self):
        if getattr(self, "h", None):
            lib().sim_sort_destroy(self.h)
            self.h = None

    def update(self, dets, n_dets, threads=128):
        dets = np.ascontiguousarray(dets, np.float32)
        T, S, ld, _ = dets.shape
        n_dets = np.ascontiguousarray(n_dets, np.int32).reshape(T, S)
        out = np.zeros((T, S, self.cap, 8), np.float32)
        n_out = np.zeros((T, S), np.int32)
        lib().sim_sort_update(self.h, dets, n_dets, T, ld, out, n_out, self.cap, threads)
        return out, n_out

    def header(self, s=0):
        h = np.zeros(16, np.int32)
        lib().sim_sort_header(self.h, s, h)
        return h


class SimOCSort:
    def __init__(self, n_streams=1, det_thresh=0.2, max_age=30, min_hits=3, iou_threshold=0.3, min_conf=0.1, delta_t=3,
                 inertia=0.2, use_byte=False, q_xy_scaling=0.01, q_s_scaling=0.0001, asso_func="iou", frame=(1920, 1080)):
        self.S, self.cap = n_streams, 256
        self.h = lib().sim_oc_create(n_streams, det_thresh, max_age, min_hits, iou_threshold, min_conf, delta_t, inertia,
                                     int(use_byte), q_xy_scaling, q_s_scaling)
        if asso_func != "iou":
            lib().sim_oc_set_asso(self.h, {"centroid": 6}[asso_func], int(frame[0]), int(frame[1]))

    def __del__(self):
        if getattr(self, "h", None):
            lib().sim_oc_destroy(self.h)
            self.h = None

    def update(self, dets, n_dets, threads=128):
        dets = np.ascontiguousarray(dets, np.float32)
        T, S, ld, _ = dets.shape
        n_dets = np.ascontiguousarray(n_dets, np.int32).reshape(T, S)
        out = np.zeros((T, S, self.cap, 8), np.float32)
        n_out = np.zeros((T, S), np.int32)
        lib().sim_oc_update(self.h, dets, n_dets, T, ld, out, n_out, self.cap, threads)
        return out, n_out

    def header(self, s=0):
        h = np.zeros(16, np.int32)
        lib().sim_oc_header(self.h, s, h)
        return h

    def dump(self, s=0):
        buf = np.zeros((self.cap, 71), np.float32)
        k = lib().sim_oc_dump(self.h, s, buf, self.cap)
        return buf[:k]


class SimDeepOCSort:
    def __init__(self, n_streams=1, emb_dim=0, det_thresh=0.3, max_age=30, min_hits=3, iou_threshold=0.3, delta_t=3,
                 inertia=0.2, w_association_emb=0.5, alpha_fixed_emb=0.95, aw_param=0.5, embedding_off=False, aw_off=False,
                 q_xy_scaling=0.01, q_s_scaling=0.0001):
        self.S, self.cap, self.dim = n_streams, 256, 0 if embedding_off else emb_dim
        self.h = lib().sim_deepoc_create(n_streams, emb_dim, det_thresh, max_age, min_hits, iou_threshold, delta_t, inertia,
                                         w_association_emb, alpha_fixed_emb, aw_param, int(embedding_off), int(aw_off),
                                         q_xy_scaling, q_s_scaling)

    def __del__(self):
        if getattr(self, "h", None):
            lib().sim_deepoc_destroy(self.h)
            self.h = None

    def update(self, dets, n_dets, embs=None, threads=128):
        dets = np.ascontiguousarray(dets, np.float32)
        T, S, ld, _ = dets.shape
        n_dets = np.ascontiguousarray(n_dets, np.int32).reshape(T, S)
        out = np.zeros((T, S, self.cap, 8), np.float32)
        n_out = np.zeros((T, S), np.int32)
        ep = None
        if embs is not None:
            embs = np.ascontiguousarray(embs, np.float32)
            assert embs.shape == (T, S, ld, self.dim)
            ep = embs.ctypes.data_as(C.c_void_p)
        lib().sim_deepoc_update(self.h, dets, n_dets, ep, T, ld, out, n_out, self.cap, threads)
        return out, n_out

    def header(self, s=0):
        h = np.zeros(16, np.int32)
        lib().sim_deepoc_header(self.h, s, h)
        return h

    def dump(self, s=0):
        buf = np.zeros((self.cap, 71), np.float32)
        emb = np.zeros((self.cap, max(self.dim, 1)), np.float32)
        k = lib().sim_deepoc_dump(self.h, s, buf, emb.ctypes.data_as(C.c_void_p), self.cap)
        return buf[:k], emb[:k, :self.dim]


class SimBoostTrack:
    def __init__(self, n_streams=1, det_thresh=0.6, max_age=60, min_hits=3, iou_threshold=0.3, min_box_area=10,
                 aspect_ratio_thresh=1.6, lambda_mhd=0.25, use_dlo_boost=True, dlo_boost_coef=0.65, use_vt=False, **_ignored):
        self.S, self.cap = n_streams, 256
        self.h = lib().sim_boost_create(n_streams, det_thresh, max_age, min_hits, iou_threshold, int(min_box_area),
                                        aspect_ratio_thresh, lambda_mhd, int(use_dlo_boost), dlo_boost_coef, int(use_vt))

    def __del__(self):
        if getattr(self, "h", None):
            lib().sim_boost_destroy(self.h)
            self.h = None

    def update(self, dets, n_dets, threads=128):
        dets = np.ascontiguousarray(dets, np.float32)
        T, S, ld, _ = dets.shape
        n_dets = np.ascontiguousarray(n_dets, np.int32).reshape(T, S)
        out = np.zeros((T, S, self.cap, 8), np.float32)
        n_out = np.zeros((T, S), np.int32)
        lib().sim_boost_update(self.h, dets, n_dets, T, ld, out, n_out, self.cap, threads)
        return out, n_out

    def header(self, s=0):
        h = np.zeros(16, np.int32)
        lib().sim_boost_header(self.h, s, h)
        return h

    def dump(self, s=0):
        buf = np.zeros((self.cap, 80), np.float32)
        return buf[:lib().sim_boost_dump(self.h, s, buf, self.cap)]


class SimBotSort:
    def __init__(self, n_streams=1, dim=0, track_high_thresh=0.5, track_low_thresh=0.1, new_track_thresh=0.6,
                 track_buffer=30, match_thresh=0.8, proximity_thresh=0.5, appearance_thresh=0.25, frame_rate=30,
                 fuse_first_associate=False, with_reid=True):
        self.S, self.cap, self.dim = n_streams, 256, dim
        self.h = lib().sim_bot_create(n_streams, dim, track_high_thresh, track_low_thresh, new_track_thresh, track_buffer,
                                      match_thresh, proximity_thresh, appearance_thresh, frame_rate,
                                      int(fuse_first_associate), int(with_reid))

    def __del__(self):
        if getattr(self, "h", None):
            lib().sim_bot_destroy(self.h)
            self.h = None

    def update(self, dets, n_dets, embs=None, threads=128):
        """dets (T,S,ld,6), n_dets (T,S), embs (T,S,ld,dim) or None"""
        dets = np.ascontiguousarray(dets, np.float32)
        T, S, ld, _ = dets.shape
        n_dets = np.ascontiguousarray(n_dets, np.int32).reshape(T, S)
        ep = None
        if embs is not None and self.dim > 0:
            embs = np.ascontiguousarray(embs, np.float32)
            assert embs.shape == (T, S, ld, self.dim)
            ep = embs.ctypes.data_as(C.c_void_p)
        out = np.zeros((T, S, self.cap, 8), np.float32)
        n_out = np.zeros((T, S), np.int32)
        lib().sim_bot_update(self.h, dets, n_dets, ep, T, ld, out, n_out, self.cap, threads)
        return out, n_out

    def header(self, s=0):
        h = np.zeros(16, np.int32)
        lib().sim_bot_header(self.h, s, h)
        return h

    def dump(self, s, which):
        buf = np.zeros((self.cap, 82), np.float32)
        feats = np.zeros((self.cap, max(self.dim, 1)), np.float32)
        k = lib().sim_bot_dump(self.h, s, which, buf, feats.ctypes.data_as(C.c_void_p) if self.dim else None, self.cap)
        return buf[:k], feats[:k]


class SimStrongSort:
    def __init__(self, n_streams=1, dim=0, max_age=30, min_conf=0.1, max_cos_dist=0.2, max_iou_dist=0.7, n_init=3,
                 nn_budget=100, mc_lambda=0.98, ema_alpha=0.9):
        self.S, self.cap, self.dim = n_streams, 256, dim
        self.h = lib().sim_ss_create(n_streams, dim, max_age, min_conf, max_cos_dist, max_iou_dist, n_init, nn_budget,
                                     mc_lambda, ema_alpha)

    def __del__(self):
        if getattr(self, "h", None):
            lib().sim_ss_destroy(self.h)
            self.h = None

    def update(self, dets, n_dets, embs=None, threads=128):
        """dets (T,S,ld,6), n_dets (T,S), embs (T,S,ld,dim) or None"""
        dets = np.ascontiguousarray(dets, np.float32)
        T, S, ld, _ = dets.shape
        n_dets = np.ascontiguousarray(n_dets, np.int32).reshape(T, S)
        ep = None
        if embs is not None and self.dim > 0:
            embs = np.ascontiguousarray(embs, np.float32)
            assert embs.shape == (T, S, ld, self.dim)
            ep = embs.ctypes.data_as(C.c_void_p)
        out = np.zeros((T, S, self.cap, 8), np.float32)
        n_out = np.zeros((T, S), np.int32)
        lib().sim_ss_update(self.h, dets, n_dets, ep, T, ld, out, n_out, self.cap, threads)
        return out, n_out

    def header(self, s=0):
        h = np.zeros(16, np.int32)
        lib().sim_ss_header(self.h, s, h)
        return h

    def dump(self, s=0):
        buf = np.zeros((self.cap, 82), np.float32)
        feats = np.zeros((self.cap, max(self.dim, 1)), np.float32)
        k = lib().sim_ss_dump(self.h, s, buf, feats.ctypes.data_as(C.c_void_p) if self.dim else None, self.cap)
        return buf[:k], feats[:k]
