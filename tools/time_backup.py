"""Event-time the Backup-CBF launch variants at N agents; print nvidia-smi clocks alongside.
    python tools/time_backup.py [N] [iters]"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import backup_cbf as B
from test_backupcbf import c_params, random_batch
from safe_control_b200 import BatchedBackupCBF
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
sc = B.EvadeScene()
X, Ur, MOV = random_batch(sc, n, seed=11, k_mov=1)
a = [torch.from_numpy(v).cuda() for v in (X, Ur, MOV)]
def clocks():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader"],
                              capture_output=True, text=True, timeout=10).stdout.strip()
    except Exception as e:
        return str(e)
for lanes, fused in ((5, 0), (8, 0), (8, 1), (32, 0), (0, 0)):
    os.environ.pop("SCB_BK_LANES", None); os.environ.pop("SCB_BK_FUSED", None)
    if lanes: os.environ["SCB_BK_LANES"] = str(lanes)
    if fused: os.environ["SCB_BK_FUSED"] = "1"
    ctrl = BatchedBackupCBF(c_params(sc))
    for _ in range(3): ctrl.solve(*a)
    torch.cuda.synchronize()
    per = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = ctrl.solve(*a); e1.record(); torch.cuda.synchronize()
        per.append(e0.elapsed_time(e1))
    print(f"backupcbf N {n} lanes {lanes or 'auto'} fused {fused}: ms min {min(per):.3f} median {sorted(per)[len(per)//2]:.3f} max {max(per):.3f}  "
          f"agents/s {n / min(per) * 1e3:.0f}  | {clocks()}", flush=True)
