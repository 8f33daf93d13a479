"""Build libscb.so in-tree with nvcc for sm_100a (no torch involved).

    python -m safe_control_b200.build [--force] [--verbose]

The .so lands next to this file (git-ignored, but it travels to the GPU box with the
gpurun snapshot).  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libscb.so")
SRCS = [os.path.join(CSRC, "scb_api.cu"), os.path.join(CSRC, "scb_params.cc")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr"]


def _deps():
    d = list(SRCS) + [os.path.join(HERE, "..", "include", "scb.h")]
    d += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return d


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(f) > t for f in _deps())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    extra = os.environ.get("SCB_EXTRA_NVCC_FLAGS", "").split()          # e.g. -DSCB_MPC_PROFILE (debug)
    cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu"] + SRCS + ["-o", OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libscb.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
