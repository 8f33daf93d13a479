"""Time the MPC kernel on the BASELINE shapes (cfg3; cfg5 per model group) with the library selected by SCB_LIB.
    SCB_LIB=safe_control_b200/libscb_l16.so python tools/mpc_variants.py [cfg3 du5 kb5 q5]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from safe_control_b200 import BatchedMPCCBF, scenes
t = lambda a: torch.from_numpy(a).cuda()
CASES = {"cfg3": ("DynamicUnicycle2D", 4096, 16, 8), "du5": ("DynamicUnicycle2D", 2731, 64, 10),
         "kb5": ("KinematicBicycle2D", 2731, 64, 10), "q5": ("Quad3D", 2730, 64, 10), "si": ("SingleIntegrator2D", 4096, 16, 10),
         "di": ("DoubleIntegrator2D", 4096, 16, 10), "q2d": ("Quad2D", 4096, 16, 10)}
for name in (sys.argv[1:] or ["cfg3", "du5", "kb5", "q5"]):
    model, N, M, H = CASES[name]
    sc = scenes.make_scene(model, N, M, seed=1234)
    ctrl = BatchedMPCCBF(sc["spec"], num_obs=M, horizon=H)
    ctrl.schedule = os.environ.get("SCB_MPC_SCHEDULE", "1") != "0"
    a = [t(sc[k]) for k in ("X", "goal", "u_prev", "OBS", "nobs")]
    for _ in range(2):
        out = ctrl.solve(*a)
    torch.cuda.synchronize()
    ms = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = ctrl.solve(*a); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    it = out["iters"].cpu().numpy(); st = out["status"].cpu().numpy(); U = out["U"].cpu().numpy()
    print(f"{os.path.basename(os.environ.get('SCB_LIB', 'default')):18s} sched {int(ctrl.schedule)} {name:5s} N {N} ms {min(ms):8.3f}  agents/s {N / min(ms) * 1e3:10.0f}  iters mean {it.mean():.2f} max {it.max()}"
          f"  status {np.bincount(st, minlength=4)}  sumU {np.nansum(U):.9f}", flush=True)
