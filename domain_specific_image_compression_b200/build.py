"""Build libsic.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m domain_specific_image_compression_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the repo snapshot.  No torch involvement: the library is plain
CUDA runtime + extern "C".
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libsic.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]

# (source, extra flags).  tables.cu holds the IEEE-only CDF spec: no FMA contraction there.
SOURCES = [
    ("api.cu", []),
    ("bottleneck.cu", []),
    ("bottleneck_cdf.cu", []),
    ("gdn.cu", []),
    ("gdn_dense.cu", []),
    ("gdn_dense_ws.cu", []),
    ("gdn_dense_bwd.cu", []),
    ("gdn_dense_dgamma.cu", []),
    ("conv0_gdn.cu", []),
    ("deconv_rgb.cu", []),
    ("msssim.cu", []),
    ("hyper_tail.cu", []),
    ("train_step.cu", []),
    ("bias_act.cu", []),
    ("tables.cu", ["-fmad=false"]),
    ("rans_host.cpp", []),
    ("rans_device.cu", []),
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "sic.h"))
    headers.append(os.path.abspath(__file__))
    objs, jobs = [], []
    for src, extra in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            if s.endswith(".cpp"):
                cmd.insert(1, "-x")
                cmd.insert(2, "cu")
            jobs.append(cmd)

    def _compile(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)

    if verbose or len(jobs) <= 1:                      # verbose: keep each file's ptxas report together
        for cmd in jobs:
            _compile(cmd)
    else:                                              # translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            list(pool.map(_compile, jobs))
    if force or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
