"""ctypes bindings for oracle/_ref/libref_core*.so - the reference's OWN sources (Kalman filters, IoU /
matching utilities, LAP solver, tracker state machines) compiled in place from /root/reference against the
stand-in Eigen / OpenCV headers of oracle/ref_shim/ (oracle/Makefile target `ref`).

TEST INFRASTRUCTURE: imported only by tests/.  Two builds exist (see oracle/ref_shim/Eigen/Dense):
  order="eigen"     libref_core.so     reductions / triangular solves in Eigen 3.4's evaluation order (as recalled)
  order="textbook"  libref_core_tb.so  every reduction sequential - the order oracle/smallmat.hpp documents
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
_NAMES = {"eigen": "libref_core.so", "textbook": "libref_core_tb.so"}
_LIBS: dict = {}
_TMP = None

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

OPS = ("xyxy2xywh", "xywh2xyxy", "xywh2tlwh", "tlwh2xywh", "tlwh2xyxy", "xyxy2tlwh", "tlwh2xyah", "xyah2tlwh",
       "xywh2xyah", "xyah2xywh", "xyxy2xysr", "xysr2xyxy")


def available(order: str = "eigen") -> bool:
    return os.path.exists(os.path.join(REF_DIR, _NAMES[order]))


def _bind(L):
    for name in OPS:
        getattr(L, "ref_" + name).argtypes = [f32p, f32p]
        getattr(L, "ref_" + name).restype = None
    L.ref_last_error.restype = C.c_char_p
    L.ref_kf_xyah_initiate.argtypes = [f32p, f32p, f32p]
    L.ref_kf_xyah_predict.argtypes = [f32p, f32p]
    L.ref_kf_xyah_project.argtypes = [f32p, f32p, C.c_float, f32p, f32p]
    L.ref_kf_xyah_update.argtypes = [f32p, f32p, f32p, C.c_float]
    L.ref_kf_xyah_gating.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, f32p]
    L.ref_kf_xysr_init.argtypes = [f32p, f32p, f32p]
    L.ref_kf_xysr_predict.argtypes = [f32p, f32p, C.c_float, C.c_float]
    L.ref_kf_xysr_update.argtypes = [f32p, f32p, f32p]
    L.ref_kf_xysr_affine.argtypes = [f32p, f32p, f32p, f32p]
    L.ref_kf_xywh_initiate.argtypes = [f32p, f32p, f32p]
    L.ref_kf_xywh_predict.argtypes = [f32p, f32p]
    L.ref_kf_xywh_update.argtypes = [f32p, f32p, f32p]
    L.ref_kf_xywh_gating.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, f32p]
    L.ref_iou_batch.argtypes = [f32p, C.c_int, f32p, C.c_int, f32p]
    L.ref_iou_distance.argtypes = [f32p, C.c_int, f32p, C.c_int, f32p]
    L.ref_fuse_score.argtypes = [f32p, C.c_int, C.c_int, f32p]
    L.ref_embedding_distance.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_int, C.c_int, f32p]
    L.ref_asso_func.argtypes = [C.c_char_p, f32p, C.c_int, f32p, C.c_int, C.c_int, C.c_int, f32p]
    L.ref_linear_assignment.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_float, i32p, i32p]
    L.ref_aw_max_metric.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, f32p, C.c_int]
    L.ref_set_asso_func.argtypes = [C.c_char_p]
    L.ref_tracker_create.argtypes = [C.c_char_p, f32p, C.c_int]
    L.ref_tracker_create.restype = C.c_void_p
    L.ref_tracker_destroy.argtypes = [C.c_void_p]
    L.ref_tracker_reset.argtypes = [C.c_void_p]
    L.ref_tracker_update.argtypes = [C.c_void_p, f32p, C.c_int, C.c_void_p, C.c_int, f32p, C.c_int]
    return L


def lib(order: str = "eigen"):
    """The shared instance (None when it was never built, e.g. in a checkout without /root/reference)."""
    if order not in _LIBS:
        so = os.path.join(REF_DIR, _NAMES[order])
        _LIBS[order] = _bind(C.CDLL(so)) if os.path.exists(so) else None
    return _LIBS[order]


def private_lib(order: str = "eigen"):
    """A PRIVATE copy of the library (dlopen of a fresh file): the reference keeps its track-ID counters in
    process-global statics (bytetrack.hpp:33-40, sort.cpp:16-19, ocsort.hpp:32-39), so a tracker whose IDs are
    to start at 1 - like a freshly started reference process - needs its own image of those statics."""
    global _TMP
    so = os.path.join(REF_DIR, _NAMES[order])
    if not os.path.exists(so):
        return None
    if _TMP is None:
        _TMP = tempfile.mkdtemp(prefix="refcore_")
    fd, path = tempfile.mkstemp(suffix=".so", dir=_TMP)
    os.close(fd)
    shutil.copyfile(so, path)
    L = _bind(C.CDLL(path))
    os.unlink(path)            # the mapping stays valid
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Tracker:
    """One reference tracker (kind = sort | bytetrack | ocsort | botsort | strongsort | deepocsort) in a private
    image of the library.  `params` = the numeric constructor arguments in the order of ref_tracker_create."""

    def __init__(self, kind: str, params, order: str = "eigen", asso_func: str = "iou"):
        self.L = private_lib(order)
        p = _f32(params)
        self.L.ref_set_asso_func(asso_func.encode())          # OC-SORT's asso_func ctor argument; frames are 1920 x 1080
        self.h = self.L.ref_tracker_create(kind.encode(), p, p.size)
        self.L.ref_set_asso_func(b"iou")
        if not self.h:
            raise RuntimeError(self.L.ref_last_error().decode())
        self._out = np.zeros((8192, 8), np.float32)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_tracker_destroy(self.h)
            self.h = None

    def reset(self):
        self.L.ref_tracker_reset(self.h)

    def update(self, dets, embs=None):
        dets = _f32(dets).reshape(-1, 6)
        if embs is not None and np.size(embs):
            embs = _f32(embs).reshape(dets.shape[0], -1)
            ep, dim = embs.ctypes.data_as(C.c_void_p), embs.shape[1]
        else:
            ep, dim = None, 0
        n = self.L.ref_tracker_update(self.h, dets, dets.shape[0], ep, dim, self._out, self._out.shape[0])
        if n == -1000:
            raise RuntimeError(self.L.ref_last_error().decode())
        assert n >= 0
        return self._out[:n].copy()
