#!/bin/bash
# 1 GPU: Backup-CBF QP parity tests + a first timing of the kernel at 65 536 agents (both lane-group geometries)
O=gpurun_out/r2; mkdir -p $O
(time timeout 900 python -m pytest tests/test_gpu_backup.py -x -q -m gpu) > $O/pytest_gpu_backup.log 2>&1
tail -25 $O/pytest_gpu_backup.log
timeout 300 python - <<'PY' 2>&1 | tee $O/backup_timing.txt
import os, sys, numpy as np, torch
sys.path.insert(0, "tests")
from oracle import backup_cbf as B
from test_backupcbf import c_params, random_batch
from safe_control_b200 import BatchedBackupCBF
sc = B.EvadeScene()
for n in (1, 1024, 65536):
    X, Ur, MOV = random_batch(sc, n, seed=11, k_mov=1)
    a = [torch.from_numpy(v).cuda() for v in (X, Ur, MOV)]
    for lanes, fused in ((8, 0), (32, 0), (8, 1), (32, 1)):
        os.environ["SCB_BK_LANES"] = str(lanes)
        os.environ.pop("SCB_BK_FUSED", None)
        if fused: os.environ["SCB_BK_FUSED"] = "1"
        ctrl = BatchedBackupCBF(c_params(sc))
        for _ in range(3): ctrl.solve(*a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): o = ctrl.solve(*a)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"backupcbf N {n} lanes {lanes} fused {fused} ms {ms:.4f} agents/s {n / ms * 1e3:.0f} optimal {float((o['status'] == 0).float().mean()):.3f}")
PY
