"""Builds libckks_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python liberate-fhe_b200/csrc/build.py [--force] [--verbose]
"""
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
SOURCES = ["ckks_b200.cu"]
HEADERS = ["mont.cuh", "ntt_kernels.cuh", "ntt_fast.cuh", "csprng.cuh", "../../include/ckks_b200.h"]
LIB = HERE / "libckks_b200.so"
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-split-compile", "0"]


def stale():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any((HERE / f).stat().st_mtime > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, lab=False):
    """lab=True: libckks_b200_lab.so with -DCKKS_LAB (measurement knob 5 = skip one pass of a transform; never the product
    library -- load it with CKKS_B200_LIB=... from scripts/ntt_lab.py)"""
    out = HERE / "libckks_b200_lab.so" if lab else LIB
    if not force and not lab and not stale():
        return LIB
    tmp = out.with_suffix(f".tmp{__import__('os').getpid()}.so")     # atomic replace: other ranks may be loading the old file
    cmd = (["nvcc"] + NVCC_FLAGS + (["-DCKKS_LAB"] if lab else []) + (["-Xptxas", "-v"] if verbose else []) +
           ["-o", str(tmp)] + [str(HERE / s) for s in SOURCES])
    subprocess.check_call(cmd)
    tmp.replace(out)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, lab="--lab" in sys.argv))
