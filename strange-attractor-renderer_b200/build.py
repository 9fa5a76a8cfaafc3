"""Builds libsar_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(_HERE, "libsar_b200.so")
# the same sources with -DSAR_DIAGNOSTICS: roofline-experiment variants of the iterate kernel, for tools/ only
OUT_DIAG = os.path.join(_HERE, "libsar_b200_diag.so")
SOURCES = ["sar_kernels.cu", "sar_deflate.cu", "sar_abi.cu"]
HEADERS = ["sar_device.cuh", "sar_deflate.cuh", os.path.join("..", "..", "include", "sar.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # bit parity with the reference's f64 arithmetic: never contract a*b+c (device and host)
    "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall",
    "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, diagnostics: bool = False, defines=()) -> str:
    out = OUT_DIAG if diagnostics else OUT
    if not force and not diagnostics and not needs_build():
        return out
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    if diagnostics:
        cmd += ["-DSAR_DIAGNOSTICS"]
    cmd += [f"-D{d}" for d in defines]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
