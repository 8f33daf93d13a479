"""Build the CPU host-sim of the kernel bodies (test aid; see hostsim.cpp)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "build", "libscb_hostsim.so")
SRCS = [os.path.join(HERE, "hostsim.cpp"), os.path.join(ROOT, "safe_control_b200", "csrc", "scb_params.cc")]
DEPS = SRCS + [os.path.join(ROOT, "safe_control_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "safe_control_b200", "csrc"))
               if f.endswith((".cuh", ".h"))] + [os.path.join(ROOT, "include", "scb.h")]


def build(force=False, mpc=True):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-ffp-contract=off", "-o", OUT]
    if mpc and os.path.exists(os.path.join(ROOT, "safe_control_b200", "csrc", "scb_mpc.cuh")):
        cmd.append("-DSCB_HOSTSIM_MPC")
    cmd += SRCS + ["-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
