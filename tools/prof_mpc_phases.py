"""Per-phase cycle breakdown of the MPC kernel (needs a -DSCB_MPC_PROFILE build of libscb.so)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from safe_control_b200 import BatchedMPCCBF, scenes
from safe_control_b200._lib import lib
t = lambda a: torch.from_numpy(a).cuda()
N = 888
sc = scenes.make_scene("DynamicUnicycle2D", N, 16, seed=1234)
ctrl = BatchedMPCCBF(sc["spec"], num_obs=16, horizon=8)
ctrl.schedule = False          # index order: group 0 of block 0 solves agent 0 only (N < one wave)
a = [t(sc[k]) for k in ("X", "goal", "u_prev", "OBS", "nobs")]
ctrl.solve(*a); torch.cuda.synchronize()
buf = (C.c_longlong * 24)()
l = lib()
l.scb_debug_mpc_profile.argtypes = [C.c_void_p, C.c_int]
l.scb_debug_mpc_profile(None, 1)
out = ctrl.solve(*a); torch.cuda.synchronize()
l.scb_debug_mpc_profile(buf, 0)
names = ["stage_derivatives(jets)", "stage_sums#1", "grad+adjoint+rate", "residuals+mu+slacks", "stage_hessians",
         "rhs stage gradients", "riccati backward", "(GN / shift fallback)", "riccati forward", "stage_sums#2", "directions", "dJ+merit0",
         "backtracking", "accept"]
v = np.array(list(buf)[:14], dtype=float)
it = out["iters"].cpu().numpy()
# block 0 / group 0 handles agents 0, 148*6, ... ; report per-iteration averages over what it did
print("total cycles", v.sum(), "agent 0 iterations", int(it[0]), "-> cycles per iteration %.0f" % (v.sum() / max(int(it[0]), 1)), "; share by phase:")
for n, c in zip(names, v):
    print(f"  {n:28s} {c:14.0f}  {100*c/v.sum():5.1f}%")
