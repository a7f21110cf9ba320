"""The C-ABI library on a machine without a GPU: it builds, loads, exports every symbol the header
declares, mirrors the reference's PSOConfig, and refuses to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ndtpso_slam_b200 import build as nbuild
from ndtpso_slam_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ndtpso_b200.h")


@pytest.fixture(scope="module")
def lib():
    nbuild.build()
    return capi.load_library()


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ndtpso_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(lib):
    declared = header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ndtpso_b200.h but not exported"
    assert sorted(capi.EXPORTS) == declared


def test_abi_version(lib):
    assert lib.ndtpso_abi_version() == 1


def test_default_config_is_reference_psoconfig(lib):
    # include/ndtpso_slam/config.h:20-37
    cf = capi.PsoConfig()
    lib.ndtpso_pso_config_default(C.byref(cf))
    assert (cf.iterations, cf.population, cf.num_threads) == (50, 30, -1)
    assert (cf.w, cf.c1, cf.c2, cf.w_dumping) == (0.8, 2.0, 2.0, 1.0)
    assert C.sizeof(capi.PsoConfig) == 48  # same size as the reference's PSOConfig POD
    assert lib.ndtpso_rand_draws(C.byref(cf)) == 3 + 3 * 30 + 6 * 30 * 50
    cf2 = capi.PsoConfig.make(population=70, iterations=50)
    assert lib.ndtpso_rand_draws(C.byref(cf2)) == 21213


def test_struct_layout():
    assert C.sizeof(capi.MapView) == 104
    assert C.sizeof(capi.Problem) == 104 + 8 + 8 + 24 + 24 + 8 + 8
    assert capi.Problem.guess.offset == 120


def test_problem_packing(golden):
    c, flats = golden.problems("cfg1")
    ps = capi.ProblemSet(flats)
    assert ps.n == len(c["seeds"])
    assert ps.array[0].map.mean == ps.array[1].map.mean  # shared table -> same pointer
    assert ps.array[0].n_points == 361 and ps.array[0].map.n_sparse == -1
    assert ps.array[3].seed == c["seeds"][3]
    _, sp = golden.problems("cfg1", sparse=True)
    ps2 = capi.ProblemSet(sp)
    assert ps2.array[0].map.n_sparse == 53


@pytest.mark.skipif(capi.load_library().ndtpso_device_count() > 0, reason="a CUDA device is present")
def test_no_cpu_fallback(lib):
    h = C.c_void_p()
    assert lib.ndtpso_ctx_create(0, C.byref(h)) == capi.ERR_NODEVICE
    assert not h.value
    with pytest.raises(capi.NdtpsoError):
        capi.Context(0)


def test_null_arguments_are_errors_not_crashes(lib):
    assert lib.ndtpso_ctx_create(0, None) == capi.ERR_ARG
    assert lib.ndtpso_batch_solve(None) == capi.ERR_ARG
    assert lib.ndtpso_align_batch(None, 0, None, None, None, None) == capi.ERR_ARG
    lib.ndtpso_ctx_destroy(None)
    lib.ndtpso_batch_destroy(None)
    assert lib.ndtpso_rand_draws(None) == 0


def test_dframes_header_exports_and_layout(lib):
    """include/ndtpso_dframes.h: every declared entry point is exported; the config struct mirrors the header."""
    from ndtpso_slam_b200 import dframes
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ndtpso_dframes.h")).read(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(ndtpso_dframes_[a-z0-9_]+)\s*\(", src)))
    assert len(declared) == 19
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ndtpso_dframes.h but not exported"
    assert sorted(dframes.EXPORTS) == declared
    assert C.sizeof(dframes.DFramesConfig) == 48
    cfg = dframes.DFramesConfig()
    dframes._lib().ndtpso_dframes_config_default(C.byref(cfg))
    # NDTFrame's constructor defaults (ndtframe.h:39-40) and LASER_IGNORE_EPSILON (config.h:6)
    assert (cfg.width_m, cfg.height_m, cfg.cell_side, cfg.max_beams) == (20, 20, 1.0, 1081)
    assert abs(cfg.laser_ignore_epsilon - 0.1) < 1e-7
    # null handles are errors, not crashes
    L = dframes._lib()
    assert L.ndtpso_dframes_create(None, None, None) == capi.ERR_ARG
    assert L.ndtpso_dframes_update(None, None) == capi.ERR_ARG
    assert L.ndtpso_dframes_align(None, None, None, 0, None, None, None) == capi.ERR_ARG
    assert L.ndtpso_dframes_device_bytes(None) == 0
    L.ndtpso_dframes_destroy(None)
