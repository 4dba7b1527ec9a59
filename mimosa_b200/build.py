"""Build libmimosa_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libmimosa_b200.so")
SOURCES = ["mb_ctx.cu", "mb_map.cu", "mb_factor.cu", "mb_scan.cu", "mb_decode.cu"]
# -fmad=false: no multiply-add contraction — the op-by-op IEEE behaviour the parity contract relies on.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build_native(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(HERE, "..", "include", "mimosa_b200.h"))
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest(deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("MB_NVCC_EXTRA", "").split(), "-o", OUT, *srcs, "-lnccl"]  # MB_NVCC_EXTRA: development
    if verbose:
        cmd += ["-Xptxas", "-v"]
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    r = subprocess.run(cmd, cwd=CSRC, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libmimosa_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    build_e2e_caller()
    return OUT


E2E_OUT = os.path.join(HERE, "lib", "libmb_e2e_caller.so")


def build_e2e_caller() -> str:
    """host/e2e_caller.cpp -> lib/libmb_e2e_caller.so: plain C++ over the C ABI (bench.py's e2e loop)."""
    src = os.path.join(HERE, "host", "e2e_caller.cpp")
    lib_dir = os.path.dirname(OUT)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", src, "-L", lib_dir, "-lmimosa_b200", "-Wl,-rpath,$ORIGIN", "-o", E2E_OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building libmb_e2e_caller.so")
    return E2E_OUT


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
