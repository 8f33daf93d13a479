"""A/B of the bulk-async (TMA) staged QP kernels against the LDG kernels (SCB_QP_TMA=0), CUDA events, inputs > L2.
cbf_qp: DynamicUnicycle2D, M = 16, N in {8192, 65536, 1 Mi};  optimal decay: KinematicBicycle2D_C3BF, M = 32, N in {8192 (config 4), 262144}."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from safe_control_b200 import BatchedCBFQP, BatchedOptimalDecayCBFQP, scenes
t = lambda a: torch.from_numpy(a).cuda()


def run(name, ctrl, base, N, B, lanes_list):
    n0 = base["X"].shape[0]
    rep = max(1, N // n0)
    pool = max(2, int(np.ceil(2.2 * 126e6 / (B * N)))) if N * B < 300e6 else 2
    pool = min(pool, 40)
    ins = []
    for q in range(pool):
        lo = (q * 7919) % max(n0 - min(N, n0), 1)
        sl = slice(lo, lo + min(N, n0))
        ins.append([t(np.tile(base[k][sl], (rep,) + (1,) * (base[k].ndim - 1))[:N]) for k in ("X", "U_ref", "OBS", "nobs")])
    for lanes in lanes_list:
        for tma in ("1", "0"):
            os.environ["SCB_QP_TMA"] = tma
            if lanes:
                os.environ["SCB_QP_LANES"] = str(lanes)
            else:
                os.environ.pop("SCB_QP_LANES", None)
            for q in range(3):
                out = ctrl.solve(*ins[q % pool])
            torch.cuda.synchronize()
            K = 200 if N <= 65536 else 20
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for q in range(K):
                    ctrl.solve(*ins[q % pool])
            g.replay(); torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / K)
            chk = float(out[0].double().nan_to_num().sum())
            print(f"{name} N={N:8d} lanes={lanes or 'auto':>4} tma={tma}  {best*1e3:9.2f} us/launch  {N/best/1e3:9.1f} M steps/s  "
                  f"{B*N/best/1e6:8.1f} GB/s ({B*N/best/1e6/6458.1:.3f} of measured peak)  sumU {chk:.6f}", flush=True)


base = scenes.make_scene("DynamicUnicycle2D", 1 << 18, 16, seed=1234)
ctrl = BatchedCBFQP(base["spec"], num_obs=16)
ONLY_BIG = os.environ.get("SWEEP_ONLY") == "big"
for N in ((1 << 20,) if ONLY_BIG else (8192, 65536, 1 << 20)):
    run("cbfqp", ctrl, base, N, 972, [8, 4] if N > 8192 else [8])
base4 = scenes.make_scene("KinematicBicycle2D_C3BF", 1 << 17, 32, seed=1234, optimal_decay=True)
od = BatchedOptimalDecayCBFQP(base4["spec"], num_obs=32)
for N in (8192, 1 << 18):
    run("odcbf", od, base4, N, 1868, [0])
