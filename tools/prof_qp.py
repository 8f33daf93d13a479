"""Tiny driver for ncu: 3 launches of the CBF-QP kernel at N=1024 (32 lanes/QP) then 3 at N=1M
(8 lanes/QP), then 3 optimal-decay launches at N=8192.  Inputs differ per launch (cold caches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from safe_control_b200 import BatchedCBFQP, BatchedOptimalDecayCBFQP, scenes

t = lambda a: torch.from_numpy(a).cuda()
M = 16
sc = scenes.make_scene("DynamicUnicycle2D", 1024 * 3, M, seed=1234)
ctrl = BatchedCBFQP(sc["spec"], num_obs=M)
X, Ur, OBS, nobs = t(sc["X"]), t(sc["U_ref"]), t(sc["OBS"]), t(sc["nobs"])
for k in range(3):
    s = slice(k * 1024, (k + 1) * 1024)
    ctrl.solve(X[s], Ur[s], OBS[s], nobs[s])
torch.cuda.synchronize()
NB = 1 << 20
big = scenes.make_scene("DynamicUnicycle2D", NB // 4, M, seed=77)
Xb, Ub, Ob, nb = [t(np.tile(big[k], (4,) + (1,) * (big[k].ndim - 1))) for k in ("X", "U_ref", "OBS", "nobs")]
for k in range(3):
    ctrl.solve(Xb, Ub, Ob, nb)
torch.cuda.synchronize()
sc4 = scenes.make_scene("KinematicBicycle2D_C3BF", 8192, 32, seed=1234, optimal_decay=True)
od = BatchedOptimalDecayCBFQP(sc4["spec"], num_obs=32)
args = [t(sc4[k]) for k in ("X", "U_ref", "OBS", "nobs")]
for k in range(3):
    od.solve(*args)
torch.cuda.synchronize()
print("done")
