"""In-tree build of libndtpso_b200.so (sm_100a only; nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libndtpso_b200.so")
SOURCES = [os.path.join(CSRC, "ndtpso_capi.cu")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("ndtpso_kernels.cuh", "ndtpso_pso_sliced.cuh", "ndtpso_dframes.cuh", "ndtpso_dframes_host.inc", "fast_exp.h")]
DEPS += [os.path.join(PKG, "..", "include", f) for f in ("ndtpso_b200.h", "ndtpso_dframes.h")]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libndtpso_b200.so cannot be built (there is no CPU fallback)")
    return exe


SHIM = os.path.join(PKG, "shim")
SHIM_LIB = os.path.join(LIB_DIR, "libndtpso_slam.so")
SHIM_SOURCES = [os.path.join(SHIM, "src", f) for f in ("ndtcell.cpp", "ndtframe.cpp", "ndtframe_device.cpp", "core.cpp", "frame_capi.cpp")]
EIGEN_STANDIN = os.path.join(PKG, "..", "third_party", "eigen_standin")


def eigen_include() -> str:
    """A system Eigen if there is one, else the fixed-size stand-in (third_party/eigen_standin)."""
    for root in ("/usr/include", "/usr/local/include"):
        if os.path.exists(os.path.join(root, "eigen3", "Eigen", "Core")):
            return root
    return EIGEN_STANDIN


def build_shim(force: bool = False) -> str:
    """libndtpso_slam.so: the drop-in NDTFrame / pso_optimization host library (g++, no FMA contraction,
    like the reference's build) linked against libndtpso_b200.so."""
    deps = SHIM_SOURCES + [os.path.join(SHIM, "include", "ndtpso_slam", f) for f in ("config.h", "core.h", "ndtcell.h", "ndtframe.h")]
    deps += [os.path.join(PKG, "..", "include", f) for f in ("ndtpso_frames.h", "ndtpso_dframes.h", "ndtpso_b200.h")] + [LIB_PATH]
    if not force and os.path.exists(SHIM_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(SHIM_LIB) for d in deps if os.path.exists(d)):
        return SHIM_LIB
    cmd = ["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-I" + os.path.join(SHIM, "include"),
           "-I" + os.path.join(PKG, "..", "include"), "-I" + eigen_include(), "-o", SHIM_LIB] + SHIM_SOURCES
    cmd += ["-L" + LIB_DIR, "-lndtpso_b200", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return SHIM_LIB


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        build_shim()
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    build_shim(force=True)
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
