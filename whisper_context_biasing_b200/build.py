"""Build libwlm.so in-tree with nvcc for sm_100a (no torch C++ ABI involved).

    python -m whisper_context_biasing_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libwlm.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def _deps():
    out = [os.path.join(ROOT, "include", "wlm.h")]
    for f in sorted(os.listdir(CSRC)):
        out.append(os.path.join(CSRC, f))
    return out


def kernel_sources_sha() -> str:
    """sha256 over the sources libwlm.so is compiled from (csrc/* + include/wlm.h), in name order.  bench.py prints ncu
    traffic figures only when the committed profile summary was taken from a build of exactly these sources."""
    import hashlib

    h = hashlib.sha256()
    for d in _deps():
        h.update(os.path.basename(d).encode() + b"\0")
        h.update(open(d, "rb").read())
    return h.hexdigest()[:16]


UBENCH_SRC = os.path.join(ROOT, "tools", "ubench_peak.cu")
UBENCH_LIB = os.path.join(ROOT, "tools", "libwlm_ubench.so")


def build_ubench(force: bool = False) -> str:
    """tools/libwlm_ubench.so: the FFMA peak microbenchmark bench.py uses for the "of measured" FP32 roofline."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not force and os.path.exists(UBENCH_LIB) and os.path.getmtime(UBENCH_LIB) >= os.path.getmtime(UBENCH_SRC):
        return UBENCH_LIB
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-shared", "-Xcompiler", "-fPIC",
           "-o", UBENCH_LIB, UBENCH_SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return UBENCH_LIB


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in _deps())


def build_library(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libwlm.so cannot be built (and there is no CPU fallback)")
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    flags = list(NVCC_FLAGS)
    cmd = [nvcc, *flags, *extra_flags, "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           "-o", LIB_PATH, *_sources()]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_ubench(force="--force" in sys.argv))
    print("kernel sources sha:", kernel_sources_sha())
