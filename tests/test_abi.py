"""The C-ABI library loads and exports every symbol include/motb200.h declares; without a GPU the
compute entry points fail loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from motcpp_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def _declared():
    text = open(os.path.join(ROOT, "include", "motb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mot_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/motb200.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)


def test_config_struct_matches_header(lib):
    text = open(os.path.join(ROOT, "include", "motb200.h")).read()
    body = re.search(r"typedef struct mot_engine_config \{(.*?)\} mot_engine_config;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        typ, rest = decl.split(None, 1)
        fields += [(n.strip(), typ) for n in rest.split(",")]
    want = [(n, {"c_int": "int", "c_float": "float"}[t.__name__]) for n, t in _lib.EngineConfig._fields_]
    assert fields == want


def test_default_config_is_the_reference_ctor(lib):
    cfg = _lib.EngineConfig()
    assert lib.mot_engine_default_config(_lib.TRACKER_BYTETRACK, C.byref(cfg)) == 0
    # include/motcpp/trackers/bytetrack.hpp:97-110
    assert (cfg.max_age, cfg.max_obs, cfg.min_hits, cfg.track_buffer, cfg.frame_rate) == (30, 50, 3, 25, 30)
    assert abs(cfg.track_thresh - 0.45) < 1e-7 and abs(cfg.match_thresh - 0.8) < 1e-7 and abs(cfg.min_conf - 0.1) < 1e-7
    assert lib.mot_engine_default_config(_lib.TRACKER_SORT, C.byref(cfg)) == 0 and cfg.max_age == 1
    assert lib.mot_engine_default_config(99, C.byref(cfg)) == _lib.MOT_ERR_INVALID_ARGUMENT
    assert b"unknown tracker kind" in lib.mot_last_error()


def test_no_device_means_loud_failure(lib):
    if lib.mot_device_count() > 0:
        pytest.skip("a GPU is present")
    cfg = _lib.EngineConfig()
    lib.mot_engine_default_config(_lib.TRACKER_BYTETRACK, C.byref(cfg))
    h = C.c_void_p()
    assert lib.mot_engine_create(C.byref(cfg), C.byref(h)) == _lib.MOT_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.mot_last_error()
    cost = np.zeros((2, 2), np.float32)
    r = np.zeros(2, np.int32)
    assert lib.mot_lap_host(cost.ctypes.data, 2, 2, 2, 0.5, r.ctypes.data, r.ctypes.data) == _lib.MOT_ERR_NO_DEVICE
    from motcpp_b200 import api
    with pytest.raises(_lib.MotError):
        api.ByteTrack()
    with pytest.raises(_lib.MotError):
        api.linear_assignment(np.array([[0.1]], np.float32), 0.5)


def test_host_side_argument_checks():
    from motcpp_b200 import api
    # empty problems never reach the device (reference src/utils/matching.cpp:20-28)
    res = api.linear_assignment(np.zeros((0, 3), np.float32), 0.5)
    assert res.matches == [] and res.unmatched_a == [] and res.unmatched_b == [0, 1, 2]
    assert api.iou_batch(np.zeros((0, 4)), np.zeros((2, 4))).shape == (0, 2)
    with pytest.raises(ValueError):
        api.embedding_distance(np.zeros((1, 4)), np.zeros((1, 4)), metric="manhattan")


def _build_cpp_example(tmp_path):
    import subprocess
    exe = str(tmp_path / "simple_tracking")
    lib_dir = os.path.join(ROOT, "motcpp_b200")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "simple_tracking.cpp"), "-L" + lib_dir, "-lmotb200",
                           "-Wl,-rpath," + lib_dir, "-o", exe])
    return exe


def test_cpp_binding_compiles_and_fails_loudly_without_gpu(lib, tmp_path):
    """include/motcpp_b200/trackers.hpp (Sort / ByteTrack / OCSort / BotSort with the reference's positional
    constructors) builds with plain g++ against the C ABI."""
    import subprocess
    exe = _build_cpp_example(tmp_path)
    if lib.mot_device_count() > 0:
        pytest.skip("a GPU is present")
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 1 and "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_cpp_binding_runs_on_gpu(lib, tmp_path):
    import subprocess
    p = subprocess.run([_build_cpp_example(tmp_path)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "frame 2 rows: sort 2 ocsort 2 botsort 2" in p.stdout and "frame 2 id 1" in p.stdout
    assert "frame 2 deepocsort rows 2" in p.stdout
