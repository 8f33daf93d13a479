"""Driver for tools/count_flops.sh: ONE solve of each MPC workload the bench reports (cfg3; config 5's three model
groups at one GPU's share), so that ncu can count the FP64 instructions of each mpc_kernel launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safe_control_b200 import BatchedMPCCBF, scenes

t = lambda a: torch.from_numpy(a).cuda()
CASES = [("cfg3", "DynamicUnicycle2D", 4096, 16, 8), ("cfg5", "DynamicUnicycle2D", 2731, 64, 10),
         ("cfg5", "KinematicBicycle2D", 2731, 64, 10), ("cfg5", "Quad3D", 2730, 64, 10)]
for key, model, N, M, H in CASES:
    sc = scenes.make_scene(model, N, M, seed=1234)
    ctrl = BatchedMPCCBF(sc["spec"], num_obs=M, horizon=H)
    out = ctrl.solve(t(sc["X"]), t(sc["goal"]), t(sc["u_prev"]), t(sc["OBS"]), t(sc["nobs"]))
    torch.cuda.synchronize()
    print("CASE", key, model, N, M, H, "iters_mean", float(out["iters"].float().mean()), flush=True)
