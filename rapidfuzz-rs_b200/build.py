#!/usr/bin/env python3
"""Builds rapidfuzz-rs_b200/lib/librfgpu.so in-tree: nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "librfgpu.so")
NVCC = os.environ.get("RF_NVCC", "/usr/local/cuda/bin/nvcc")
HOSTCXX = os.environ.get("RF_HOSTCXX", "/usr/bin/g++")
SOURCES = ["rf_kernels.cu", "rf_layout.cu", "rf_select.cu", "rf_api.cu", "rf_sharded.cu", "rf_io.cpp"]
HEADERS = ["rf_core.cuh", "rf_kernels.cuh", "rf_internal.h", os.path.join("..", "..", "include", "rfgpu.h")]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [NVCC, "-ccbin", HOSTCXX, "-std=c++17", "-O3", "-lineinfo",
           "-gencode", "arch=compute_100a,code=sm_100a",
           # Rust never contracts a*b+c; the f64 Jaro/normalisation epilogues must match it bit for bit
           "--fmad=false",
           "-Xcompiler", "-fPIC,-fopenmp,-ffp-contract=off,-Wall",
           "-shared", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lgomp", "-lnccl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    if os.environ.get("RF_DEBUG_MW"):
        cmd.insert(1, "-DRF_DEBUG_MW")
    for d in os.environ.get("RF_DEFINES", "").split():   # dev A/B builds, e.g. RF_DEFINES=RF_LEV_NO_DP4A RF_LIB_OUT=...
        cmd.insert(1, "-D" + d)
    if os.environ.get("RF_LIB_OUT"):
        cmd[cmd.index("-o") + 1] = os.environ["RF_LIB_OUT"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
