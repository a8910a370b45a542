"""Build recipes: the product library (CUDA, sm_100a only) and -- separately -- the checkers.

`build_product()` compiles emfusion_b200/csrc/*.cu into emfusion_b200/lib/libemf_b200.so with
explicit nvcc flags (no torch, no JIT cache: the .so lives in-tree and travels to the GPU box).
`build_checkers()` runs oracle/Makefile (C oracle; and the reference kernels when
/root/reference is present).  Building a checker is not using it: nothing in the product
imports oracle/.
"""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "emfusion_b200", "csrc")
LIBDIR = os.path.join(ROOT, "emfusion_b200", "lib")
LIB = os.path.join(LIBDIR, "libemf_b200.so")
SOURCES = ["integrate.cu", "raycast.cu", "bricks.cu", "assoc.cu", "fgprob.cu", "track.cu", "resize.cu", "mcubes.cu", "preprocess.cu", "xchg.cu", "host.cu", "engine.cu"]
NVCC_FLAGS = [
    "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--default-stream", "legacy",
]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build_product(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "emf_b200.h"))
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        if force or not _newer(o, [s] + hdrs):
            cmd = ["nvcc", *NVCC_FLAGS, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            subprocess.run(cmd, check=True)
        objs.append(o)
    if force or not _newer(LIB, objs):
        subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-lcudart"], check=True)
    return LIB


def build_checkers() -> None:
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)
    # the reference's launch structure over this repo's operators (a measurement: bench.py --impl unchanged-caller)
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "b200ops"], check=True)


if __name__ == "__main__":
    build_product(force="--force" in sys.argv, verbose="-v" in sys.argv)
    if "--checkers" in sys.argv:
        build_checkers()
    print(LIB)
