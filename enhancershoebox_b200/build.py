"""Builds libesb.so (CUDA, sm_100a only) in-tree with nvcc.  `python -m enhancershoebox_b200.build`.

esb_kernels.cu is compiled three times (16-warp CTAs x 2 per SM for full batches, 32-warp CTAs x 1 per SM for batches
smaller than the SM count, and the same CTA in thread-block clusters of 2/4/8 per replica for still smaller ones); the
five translation units are compiled in parallel and linked into one library."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libesb.so")
SOURCES = ["esb_api.cu", "esb_kernels.cu", "esb_aggregate.cu"]
HEADERS = ["esb_device.cuh", os.path.join("..", "..", "include", "esb.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
UNITS = [   # (object, source, extra defines)
    ("esb_api.o", "esb_api.cu", []),
    ("esb_kernels_512.o", "esb_kernels.cu", ["-DESB_TAG=512", "-DESB_THREADS=512", "-DESB_MIN_CTAS=2"]),
    ("esb_kernels_1024.o", "esb_kernels.cu", ["-DESB_TAG=1024", "-DESB_THREADS=1024", "-DESB_MIN_CTAS=1"]),
    ("esb_kernels_2048.o", "esb_kernels.cu", ["-DESB_TAG=2048", "-DESB_THREADS=1024", "-DESB_MIN_CTAS=1", "-DESB_CLUSTER=1"]),
    ("esb_aggregate.o", "esb_aggregate.cu", []),
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, s)) > t for s in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("ESB_NVCC_EXTRA", "").split()           # e.g. "-DESB_PREFETCH=4" for experiments
    os.makedirs(OBJ, exist_ok=True)
    procs = []
    for obj, src, defs in UNITS:
        cmd = [nvcc, *NVCC_FLAGS, *defs, *extra, *(["-Xptxas", "-v"] if verbose else []), "-c", os.path.join(CSRC, src), "-o", os.path.join(OBJ, obj)]
        procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *[os.path.join(OBJ, u[0]) for u in UNITS]])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
