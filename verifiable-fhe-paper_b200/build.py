"""Build recipe for libvpbs_commit.so (hand-written CUDA for sm_100a, in-tree)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvpbs_commit.so")
SOURCES = ["vpbs_commit.cu"]
HEADERS = ["gl64.cuh", "poseidon.cuh", "poseidon_rc.inc", "poseidon_rcd.inc", "poseidon_rcs.inc", "poseidon_rcp.inc", "ntt.cuh", "merkle.cuh",
           "permutation.cuh", "openings.cuh", "host_stage.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared"]


def nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libvpbs_commit.so cannot be built (there is no CPU fallback)")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(HERE, "..", "include", "vpbs_commit.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into verifiable-fhe-paper_b200/libvpbs_commit.so."""
    if not force and not stale():
        return LIB
    cmd = [nvcc(), *NVCC_FLAGS, "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    # the image exports CC=/opt/gcc/bin/gcc for other tools; nvcc should use the system g++
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.run(cmd, check=True, env=env)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
