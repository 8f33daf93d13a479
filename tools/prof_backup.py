"""One Backup-CBF solve per launch variant at N agents (for ncu): split (rollout + QP) and fused, 8 lanes per agent.
    python tools/prof_backup.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import backup_cbf as B
from test_backupcbf import c_params, random_batch
from safe_control_b200 import BatchedBackupCBF
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
sc = B.EvadeScene()
X, Ur, MOV = random_batch(sc, n, seed=11, k_mov=1)
a = [torch.from_numpy(v).cuda() for v in (X, Ur, MOV)]
for fused in (0, 1):
    os.environ["SCB_BK_LANES"] = "8"
    os.environ.pop("SCB_BK_FUSED", None)
    if fused:
        os.environ["SCB_BK_FUSED"] = "1"
    ctrl = BatchedBackupCBF(c_params(sc))
    ctrl.solve(*a)
    torch.cuda.synchronize()
