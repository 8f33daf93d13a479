#!/usr/bin/env python
"""Time the device-side closed loop (scb_run_all_steps): python tools/prof_loop.py [N] [M] [steps] [controller] [model]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from safe_control_b200 import BatchedTrackingController  # noqa: E402
from track_util import random_closed_loop_case  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
M = int(sys.argv[2]) if len(sys.argv) > 2 else 16
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
controller = sys.argv[4] if len(sys.argv) > 4 else "cbf_qp"
model = sys.argv[5] if len(sys.argv) > 5 else "DynamicUnicycle2D"
X0, scene, wps = random_closed_loop_case(model, N, M, seed=1234)
tc = BatchedTrackingController(X0, {"model": model, "num_constraints": M, "mpc_horizon": 8}, {"pos": controller}, obs=scene)
tc.set_waypoints(wps)
tc.run_steps(20); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record(); tc.run_steps(steps); e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
ms = e0.elapsed_time(e1)
print(f"eager  N={N} M={M} {controller} {model}: {ms / steps * 1e3:.2f} us/step  ({N * steps / ms * 1e3:.3e} agent-steps/s) wall {1e3*(t1-t0):.1f} ms")
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    tc.run_steps(steps)
g.replay(); torch.cuda.synchronize()
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"graph  N={N} M={M} {controller} {model}: {ms / steps * 1e3:.2f} us/step  ({N * steps / ms * 1e3:.3e} agent-steps/s)")
b = tc.buffers()
print("done", int(b["done"].sum()), "ret -2:", int((b["ret"] == -2).sum()), "ret -1:", int((b["ret"] == -1).sum()), "nsteps mean", float(b["nsteps"].float().mean()))
