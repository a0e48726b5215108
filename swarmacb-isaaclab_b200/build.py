"""In-tree nvcc build of the C-ABI library (csrc/swarm_step.cu -> libswarmstep.so) for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "swarm_step.cu")
HEADER = os.path.join(os.path.dirname(HERE), "include", "swarm_abi.h")
LIB = os.path.join(HERE, "libswarmstep.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    # the pose path must not contract a*b+c into FMA (reference = separate float32 torch ops)
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in (SRC, HEADER, __file__))


def build_variant(out: str, defines: list[str]) -> str:
    """Tuning aid: compile csrc/swarm_step.cu with extra -D flags into ``out`` (not used by the product path)."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", out, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if it is missing or older than its sources; return its path."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libswarmstep.so")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
