"""ctypes binding of libmotb200.so (include/motb200.h).  No fallback of any kind: if the library is
missing, or the machine has no CUDA device, every compute call raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MOT_LIB") or os.path.join(HERE, "libmotb200.so")

MOT_OK = 0
MOT_ERR_INVALID_ARGUMENT = 1
MOT_ERR_CUDA = 2
MOT_ERR_NO_DEVICE = 3
MOT_ERR_CAPACITY = 4
MOT_ERR_NUMERIC = 5
MOT_ERR_UNSUPPORTED = 6

TRACKER_SORT, TRACKER_BYTETRACK, TRACKER_OCSORT, TRACKER_BOTSORT, TRACKER_STRONGSORT, TRACKER_DEEPOCSORT = 0, 1, 2, 3, 4, 5
TRACKER_BOOSTTRACK = 6
KF_XYAH, KF_XYSR, KF_XYWH = 0, 1, 2


class MotError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libmotb200 error {code}: {msg}")
        self.code = code


class EngineConfig(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("n_streams", C.c_int), ("track_capacity", C.c_int), ("max_dets", C.c_int),
        ("device", C.c_int), ("n_chunks", C.c_int),
        ("det_thresh", C.c_float), ("max_age", C.c_int), ("max_obs", C.c_int), ("min_hits", C.c_int),
        ("iou_threshold", C.c_float),
        ("min_conf", C.c_float), ("track_thresh", C.c_float), ("match_thresh", C.c_float),
        ("track_buffer", C.c_int), ("frame_rate", C.c_int),
        ("delta_t", C.c_int), ("inertia", C.c_float), ("use_byte", C.c_int),
        ("q_xy_scaling", C.c_float), ("q_s_scaling", C.c_float),
        ("track_high_thresh", C.c_float), ("track_low_thresh", C.c_float), ("new_track_thresh", C.c_float),
        ("proximity_thresh", C.c_float), ("appearance_thresh", C.c_float),
        ("fuse_first_associate", C.c_int), ("with_reid", C.c_int), ("emb_dim", C.c_int),
        ("max_cos_dist", C.c_float), ("max_iou_dist", C.c_float), ("n_init", C.c_int), ("nn_budget", C.c_int),
        ("mc_lambda", C.c_float), ("ema_alpha", C.c_float),
        ("w_association_emb", C.c_float), ("alpha_fixed_emb", C.c_float), ("aw_param", C.c_float),
        ("embedding_off", C.c_int), ("aw_off", C.c_int),
        ("asso_func", C.c_int), ("frame_width", C.c_int), ("frame_height", C.c_int),
        ("min_box_area", C.c_int), ("aspect_ratio_thresh", C.c_float), ("lambda_iou", C.c_float), ("lambda_mhd", C.c_float),
        ("lambda_shape", C.c_float), ("use_dlo_boost", C.c_int), ("dlo_boost_coef", C.c_float), ("use_sb", C.c_int), ("use_vt", C.c_int),
    ]


# every symbol include/motb200.h declares: name -> (restype, argtypes)
_VP, _I, _LL, _F, _SZ = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t
SYMBOLS = {
    "mot_last_error": (C.c_char_p, []),
    "mot_version": (_I, []),
    "mot_device_count": (_I, []),
    "mot_device_alloc": (_I, [C.POINTER(_VP), _SZ]),
    "mot_device_free": (_I, [_VP]),
    "mot_host_alloc": (_I, [C.POINTER(_VP), _SZ]),
    "mot_host_free": (_I, [_VP]),
    "mot_copy_h2d": (_I, [_VP, _VP, _SZ, _VP]),
    "mot_copy_d2h": (_I, [_VP, _VP, _SZ, _VP]),
    "mot_memset_device": (_I, [_VP, _I, _SZ, _VP]),
    "mot_stream_sync": (_I, [_VP]),
    "mot_engine_default_config": (_I, [_I, C.POINTER(EngineConfig)]),
    "mot_engine_create": (_I, [C.POINTER(EngineConfig), C.POINTER(_VP)]),
    "mot_engine_destroy": (_I, [_VP]),
    "mot_engine_reset": (_I, [_VP]),
    "mot_engine_update_host": (_I, [_VP, _I, _VP, _VP, _I, _VP, _VP, _I]),
    "mot_engine_update_host_packed": (_I, [_VP, _I, _VP, _VP, _I, _I, _VP, C.c_longlong, _VP, _VP]),
    "mot_engine_profile": (_I, [_VP, _I, _VP]),
    "mot_engine_update_device": (_I, [_VP, _I, _VP, _VP, _I, _VP, _VP, _I, _VP]),
    "mot_engine_update_host_embs": (_I, [_VP, _I, _VP, _VP, _I, _VP, _VP, _VP, _I]),
    "mot_engine_update_device_embs": (_I, [_VP, _I, _VP, _VP, _I, _VP, _VP, _VP, _I, _VP]),
    "mot_engine_dump_bot": (_I, [_VP, _I, _I, _VP, _VP, _I, C.POINTER(_I)]),
    "mot_engine_dump_strong": (_I, [_VP, _I, _VP, _VP, _I, C.POINTER(_I)]),
    "mot_engine_check": (_I, [_VP, _VP]),
    "mot_engine_stream_header": (_I, [_VP, _I, _VP]),
    "mot_engine_dump_list": (_I, [_VP, _I, _I, _VP, _I, C.POINTER(_I)]),
    "mot_engine_dump_deep_embs": (_I, [_VP, _I, _VP, _I, C.POINTER(_I)]),
    "mot_engine_dump_boost": (_I, [_VP, _I, _VP, _I, C.POINTER(_I)]),
    "mot_engine_info": (_I, [_VP, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "mot_kf_initiate": (_I, [_I, _VP, _VP, _LL, _VP]),
    "mot_kf_predict": (_I, [_I, _VP, _VP, _LL, _F, _F, _VP]),
    "mot_kf_update": (_I, [_I, _VP, _VP, _VP, _LL, _VP, _VP]),
    "mot_kf_gating": (_I, [_I, _VP, _I, _VP, _I, _I, _I, _VP, _VP]),
    "mot_cost_iou": (_I, [_VP, _I, _VP, _I, _VP, _VP, _I, _I, _VP]),
    "mot_cost_iou_variant": (_I, [_VP, _I, _VP, _I, _I, _I, _I, _VP, _I, _VP]),
    "mot_cost_ocm": (_I, [_VP, _I, _VP, _VP, _VP, _I, _F, _VP, _VP, _I, _VP]),
    "mot_cost_cosine": (_I, [_VP, _I, _VP, _I, _I, _VP, _I, _VP]),
    "mot_cost_nn_cosine": (_I, [_VP, _VP, _I, _I, _VP, _I, _I, _VP, _I, _VP]),
    "mot_cost_gate": (_I, [_VP, _I, _VP, _I, _VP, _I, _F, _F, _I, _VP]),
    "mot_cost_iou_tlwh": (_I, [_VP, _VP, _I, _VP, _I, _VP, _I, _VP]),
    "mot_kf_xysr_affine": (_I, [_VP, _LL, _VP, _VP, _VP]),
    "mot_cost_aw_max_metric": (_I, [_VP, _I, _I, _I, _F, _F, _VP, _I, _VP]),
    "mot_lap_device": (_I, [_VP, _I, _I, _I, _F, _VP, _VP, _VP]),
    "mot_lap_batch_device": (_I, [_VP, _LL, _I, _VP, _VP, _I, _I, _I, _F, _VP, _VP, _VP]),
    "mot_lap_jv_batch_device": (_I, [_VP, _LL, _I, _I, _I, _I, _F, _VP, _VP, _VP]),
    "mot_lap_host": (_I, [_VP, _I, _I, _I, _F, _VP, _VP]),
}

_lib = None


def load() -> C.CDLL:
    """Load libmotb200.so (built in-tree by motcpp_b200.build).  Raises if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MotError(MOT_ERR_UNSUPPORTED,
                           f"{LIB_PATH} is missing - run `python -m motcpp_b200.build` (nvcc, sm_100a). "
                           "There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)          # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != MOT_OK:
        raise MotError(rc, load().mot_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    return int(load().mot_device_count())


def require_gpu() -> None:
    if device_count() <= 0:
        raise MotError(MOT_ERR_NO_DEVICE, "no CUDA device: motcpp_b200 has no CPU path")
