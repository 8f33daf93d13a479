"""Event-time one gatekeeper / MPS control step at N agents (steady state: every agent re-plans each step).
    python tools/time_shield.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_shield import full_batch, make_shield, Lanes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
sc, X, NOMX, NOMU, MOV, STAT = full_batch(n, 100, seed=5)
d = [torch.from_numpy(v).cuda() for v in (X, NOMX, NOMU, MOV, STAT)]
for mode, lanes, two in (("gatekeeper", 8, False), ("gatekeeper", 32, False), ("gatekeeper", 8, True), ("gatekeeper", 32, True), ("gatekeeper", None, None), ("mps", 1, None)):
    sh = make_shield(sc, mode, n, 100, keep_states=False)
    with Lanes(lanes, two):
        for _ in range(3): sh.step(*d)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): o = sh.step(*d)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"shield {mode} N {n} lanes {lanes} two-launch {two}: ms {ms:.4f} agent-steps/s {n / ms * 1e3:.0f}  committed nominal legs {float((sh.nsteps > 0).float().mean()):.3f} "
          f"mean nominal steps {float(sh.nsteps.float().mean()):.1f} using_backup {float(o['using_backup'].float().mean()):.3f}", flush=True)
