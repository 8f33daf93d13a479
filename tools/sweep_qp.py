"""GPU tuning sweep for the CBF-QP kernel: lanes per QP x batch size (CUDA events, cold inputs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from safe_control_b200 import BatchedCBFQP, scenes
t = lambda a: torch.from_numpy(a).cuda()
M = 16
base = scenes.make_scene("DynamicUnicycle2D", 1 << 18, M, seed=1234)
ctrl = BatchedCBFQP(base["spec"], num_obs=M)
SIZES = [int(x) for x in os.environ.get('SWEEP_N', '1024,8192,65536,1048576').split(',')]
LANES = [int(x) for x in os.environ.get('SWEEP_LANES', '32,8,4').split(',')]
for N in SIZES:
    rep = max(1, N // (1 << 18))
    pool = 8 if N <= 65536 else 2
    ins = []
    for q in range(pool):
        sl = slice((q * N) % (1 << 18), (q * N) % (1 << 18) + min(N, 1 << 18))
        ins.append([t(np.tile(base[k][sl], (rep,) + (1,) * (base[k].ndim - 1))) for k in ("X", "U_ref", "OBS", "nobs")])
    for lanes in LANES:
        os.environ["SCB_QP_LANES"] = str(lanes)
        for q in range(3):
            ctrl.solve(*ins[q % pool])
        torch.cuda.synchronize()
        K = 200 if N <= 65536 else 10
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for q in range(K):
                ctrl.solve(*ins[q % pool])
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"N={N:8d} lanes={lanes:2d}  {ms*1e3:9.2f} us/launch  {N/ms/1e3:9.1f} M steps/s  {972*N/ms/1e6:8.1f} GB/s", flush=True)
