"""Builds ``libb200nerf.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""

from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.environ.get("B200NERF_LIB") or os.path.join(PKG_DIR, "libb200nerf.so")
SOURCES = ["b200nerf.cu", "sampling.cu", "train.cu"]
HEADERS = ["ptx.cuh", "umma_selftest.cuh", "mlp_fast.cuh", "mlp_exact.cuh", "composite_tma.cuh", "tgemm.cuh", "tgemm_reg.cuh", "host_common.h", os.path.join("..", "..", "include", "b200nerf.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libb200nerf.so")
    return exe


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, extra_flags=(), out_path: str = None) -> str:
    """Compile the CUDA library if it is missing or older than its sources; returns its path."""
    out_path = out_path or LIB_PATH
    if not force and not is_stale() and out_path == LIB_PATH:
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out_path] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out_path


if __name__ == "__main__":
    print(build(force=True, verbose=True))
