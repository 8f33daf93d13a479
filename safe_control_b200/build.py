"""Build libscb.so in-tree with nvcc for sm_100a (no torch involved).

    python -m safe_control_b200.build [--force] [--verbose] [--out PATH]

The .so lands next to this file (git-ignored, but it travels to the GPU box with the
gpurun snapshot).  nvcc cross-compiles without a GPU.  The translation units (the C ABI +
QP / closed-loop kernels, and one MPC kernel instantiation per model) compile in parallel
into csrc/_obj/ and are linked by nvcc.

Environment: NVCC, SCB_EXTRA_NVCC_FLAGS (e.g. "-DSCB_MPC_LANES=16"; with -DSCB_MPC_PROFILE or
-DSCB_SINGLE_TU everything is compiled as one translation unit).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libscb.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
          "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
# scb_model ids with an MPC kernel (include/scb.h): all but Manipulator2D; 100 + id = the general-row variants of SI / DU / DI
# for superellipsoid obstacles; 200 + id = the optimal-decay MPC variants (omega1, omega2 as stage inputs)
MPC_MODELS = [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 100, 101, 105, 201, 202, 206, 210]


def _units(single_tu):
    """-> [(object name, source, extra flags)]"""
    u = [("scb_api", os.path.join(CSRC, "scb_api.cu"), []),
         ("scb_params", os.path.join(CSRC, "scb_params.cc"), [])]
    if not single_tu:
        u += [(f"scb_mpc_inst_{m}", os.path.join(CSRC, "scb_mpc_inst.cu"), [f"-DSCB_MPC_INST={m}"]) for m in MPC_MODELS]
    return u


def _deps():
    d = [os.path.join(HERE, "..", "include", "scb.h"), os.path.abspath(__file__)]
    d += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".cu", ".cc"))]
    return d


def needs_build(out=OUT):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(f) > t for f in _deps())


def build(force=False, verbose=False, out=OUT, extra=None):
    if extra is None:
        extra = os.environ.get("SCB_EXTRA_NVCC_FLAGS", "").split()
    if not force and not extra and not needs_build(out):
        return out
    single = any(f in ("-DSCB_SINGLE_TU", "-DSCB_MPC_PROFILE") for f in extra)
    tag = hashlib.sha1(" ".join(extra).encode()).hexdigest()[:8] if extra else "default"
    objdir = os.path.join(CSRC, "_obj", tag)
    os.makedirs(objdir, exist_ok=True)
    vflags = ["-Xptxas", "-v"] if verbose else []

    def compile_one(unit):
        name, src, fl = unit
        obj = os.path.join(objdir, name + ".o")
        cmd = [NVCC] + CFLAGS + extra + fl + vflags + ["-x", "cu", "-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return obj, res

    units = _units(single)
    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, units))
    objs = []
    for obj, res in results:
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed building libscb.so")
        objs.append(obj)
    res = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", out],
                         capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libscb.so")
    return out


if __name__ == "__main__":
    o = OUT
    if "--out" in sys.argv:
        o = os.path.abspath(sys.argv[sys.argv.index("--out") + 1])
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, out=o))
