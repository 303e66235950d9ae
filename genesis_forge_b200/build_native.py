"""
Build libgfb200.so in-tree with nvcc for sm_100a (B200).

    python -m genesis_forge_b200.build_native [--force]

nvcc cross-compiles without a GPU.  The result is git-ignored but travels to the GPU box with the
repo snapshot.
"""
from __future__ import annotations

import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
SRC = PKG / "csrc" / "gfb200.cu"
OUT = PKG / "libgfb200.so"
DEPS = [SRC, *sorted((PKG / "csrc").glob("*.cuh")), *sorted((PKG / "csrc").glob("*.h")), PKG.parent / "include" / "gfb200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--fmad=false",  # every FMA in the kernels is explicit (bit-exact recipes); never contract implicitly
]


def needs_build() -> bool:
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    return any(d.stat().st_mtime > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *NVCC_FLAGS, str(SRC), "-o", str(OUT)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr)
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
