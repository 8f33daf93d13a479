"""Driver for ncu: two launches of the MPC kernel (cfg3 shapes, one persistent wave)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from safe_control_b200 import BatchedMPCCBF, scenes
t = lambda a: torch.from_numpy(a).cuda()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 888
sc = scenes.make_scene("DynamicUnicycle2D", N, 16, seed=1234)
ctrl = BatchedMPCCBF(sc["spec"], num_obs=16, horizon=8)
a = [t(sc[k]) for k in ("X", "goal", "u_prev", "OBS", "nobs")]
for _ in range(2):
    out = ctrl.solve(*a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = ctrl.solve(*a); e1.record(); torch.cuda.synchronize()
it = out["iters"].cpu().numpy(); st = out["status"].cpu().numpy()
print("N", N, "ms", e0.elapsed_time(e1), "iters mean %.1f max %d" % (it.mean(), it.max()), "status", np.bincount(st, minlength=4))
