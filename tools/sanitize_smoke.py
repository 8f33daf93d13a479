"""Small invocation of every kernel family for compute-sanitizer (tools/sanitize.sh): CBF-QP (warp-per-agent LDG kernel,
lane-group bulk-async kernel), optimal-decay (both), MPC (fast path, general rows, schedule), closed loop (fused + per-step), Backup-CBF QP (split / fused), gatekeeper / MPS."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from safe_control_b200 import BatchedCBFQP, BatchedOptimalDecayCBFQP, BatchedMPCCBF, BatchedTrackingController, scenes

t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
big = int(os.environ.get("SAN_N", "6000"))
for N in (257, big):                                    # 257: warp per QP; big (> 16 x SMs): lane groups + bulk copies
    sc = scenes.make_scene("DynamicUnicycle2D", N, 16, seed=1)
    U, st, act = BatchedCBFQP(sc["spec"], num_obs=16).solve(t(sc["X"]), t(sc["U_ref"]), t(sc["OBS"]), t(sc["nobs"]))
    sc = scenes.make_scene("KinematicBicycle2D_C3BF", N, 32, seed=2, optimal_decay=True)
    BatchedOptimalDecayCBFQP(sc["spec"], num_obs=32).solve(t(sc["X"]), t(sc["U_ref"]), t(sc["OBS"]), t(sc["nobs"]))
torch.cuda.synchronize()
for model, N, M, H in (("DynamicUnicycle2D", 1300, 8, 6), ("Quad3D", 40, 8, 5), ("KinematicBicycle2D_C3BF", 40, 6, 5)):
    sc = scenes.make_scene(model, N, M, seed=3)
    BatchedMPCCBF(sc["spec"], num_obs=M, horizon=H).solve(t(sc["X"]), t(sc["goal"]), t(sc["u_prev"]), t(sc["OBS"]), t(sc["nobs"]),
                                                          want_pred=True, want_active=True)
sc = scenes.with_superellipsoids(scenes.make_scene("DynamicUnicycle2D", 48, 8, seed=4))
BatchedMPCCBF(sc["spec"], num_obs=8, horizon=5).solve(t(sc["X"]), t(sc["goal"]), t(sc["u_prev"]), t(sc["OBS"]), t(sc["nobs"]))
torch.cuda.synchronize()
rng = np.random.default_rng(0)
scene = np.zeros((12, 7)); scene[:, :2] = rng.uniform(0, 12, (12, 2)); scene[:, 2] = 0.3
X0 = np.hstack([rng.uniform(0, 12, (64, 2)), rng.uniform(-3, 3, (64, 1)), rng.uniform(0, 1, (64, 1))])
wps = np.zeros((64, 3, 3)); wps[:, :, :2] = rng.uniform(0, 12, (64, 3, 2))
for ctrl in ("cbf_qp", "mpc_cbf"):
    tc = BatchedTrackingController(X0, {"model": "DynamicUnicycle2D", "num_constraints": 6, "mpc_horizon": 5}, {"pos": ctrl}, obs=scene)
    tc.set_waypoints(wps)
    tc.run_steps(4)
    tc.control_step()
torch.cuda.synchronize()
# Backup-CBF QP (split and fused launches, both lane-group geometries) and the gatekeeper / MPS step kernels (evade scene)
from safe_control_b200 import BatchedBackupCBF, BatchedShield, EvadeSceneParams
for n in (37, 700):
    X, Ur, MOV = scenes.make_evade_batch(n, seed=5, k_mov=2)
    for lanes, fused in ((5, 0), (8, 0), (32, 0), (8, 1), (32, 1)):
        os.environ["SCB_BK_LANES"] = str(lanes)
        os.environ.pop("SCB_BK_FUSED", None)
        if fused:
            os.environ["SCB_BK_FUSED"] = "1"
        BatchedBackupCBF(EvadeSceneParams(dt=0.1, backup_horizon=4.0)).solve(t(X), t(Ur), t(MOV), want_phi=True, want_active=True)
    os.environ.pop("SCB_BK_LANES", None); os.environ.pop("SCB_BK_FUSED", None)
    NOMX, NOMU = scenes.make_evade_plans(X, 30)
    STAT = scenes.evade_static_rects(MOV)
    for mode, lanes, two in (("gatekeeper", 32, 0), ("gatekeeper", 8, 0), ("gatekeeper", 1, 0), ("gatekeeper", 8, 1), ("gatekeeper", 32, 1), ("mps", 1, 0)):
        os.environ["SCB_SHIELD_LANES"] = str(lanes)
        os.environ["SCB_SHIELD_TWO_PHASE"] = str(two)
        sh = BatchedShield(n, mode, EvadeSceneParams(dt=0.1, backup_horizon=4.0), 0.05, None, 30, keep_states=True)
        for _ in range(3):
            sh.step(t(X), t(NOMX), t(NOMU), t(MOV), t(STAT))
    os.environ.pop("SCB_SHIELD_LANES", None); os.environ.pop("SCB_SHIELD_TWO_PHASE", None)
torch.cuda.synchronize()
print("sanitize smoke done")
