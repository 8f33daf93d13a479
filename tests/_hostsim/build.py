"""Build the CPU host-sim of the kernel bodies (test aid; see hostsim.cpp)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "build", "libscb_hostsim.so")
OUT_FMA = os.path.join(HERE, "build", "libscb_hostsim_fma.so")
SRCS = [os.path.join(HERE, "hostsim.cpp"), os.path.join(ROOT, "safe_control_b200", "csrc", "scb_params.cc")]
DEPS = SRCS + [os.path.join(ROOT, "safe_control_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "safe_control_b200", "csrc"))
               if f.endswith((".cuh", ".h"))] + [os.path.join(ROOT, "include", "scb.h")]


def build(force=False, mpc=True, fma=False):
    """fma=True contracts a*b+c into FMAs like nvcc does (-fmad=true): rounding-path coverage."""
    out = OUT_FMA if fma else OUT
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in DEPS):
        return out
    fp = ["-ffp-contract=fast", "-mfma"] if fma else ["-ffp-contract=off"]
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++"] + fp + ["-o", out]
    if mpc and os.path.exists(os.path.join(ROOT, "safe_control_b200", "csrc", "scb_mpc.cuh")):
        cmd.append("-DSCB_HOSTSIM_MPC")
    cmd += SRCS + ["-lm"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force=True))
    print(build(force=True, fma=True))
