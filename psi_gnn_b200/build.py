"""In-tree build of the sm_100a extension: ``psi_gnn_b200/lib/libpsignn_b200.so``.

Plain ``nvcc`` (cross-compiles without a GPU); the shared library exports only the C ABI
declared in ``include/psignn_b200.h`` and is loaded with ``ctypes`` (psi_gnn_b200/_native.py).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libpsignn_b200.so")
STAMP = os.path.join(LIB_DIR, "libpsignn_b200.stamp")
SOURCES = ["psignn_b200.cu", "common.cuh", "weights.cuh", "graph.cuh", "layer.cuh", "vjp.cuh", "broyden.cuh", "anderson.cuh", "qn_tma.cuh", "comm.cuh", "pgrad.cuh", "baseline_bwd.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the PSI-GNN B200 extension cannot be built")


def source_hash() -> str:
    h = hashlib.sha256()
    for name in SOURCES + [os.path.join(ROOT, "include", "psignn_b200.h")]:
        with open(name if os.path.isabs(name) else os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == source_hash()


def build_variant(path: str, defines) -> str:
    """an experimental build with extra -D macros into another file (select it with PSI_GNN_B200_LIB=<path>); used for A/B timing"""
    cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", path, os.path.join(CSRC, "psignn_b200.cu")]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    return path


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the extension if the sources changed since the last build; returns the .so path."""
    if not force and is_current():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, os.path.join(CSRC, "psignn_b200.cu")]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    with open(STAMP, "w") as f:
        f.write(source_hash())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
