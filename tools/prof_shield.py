"""Two gatekeeper steps and two MPS steps at N agents (for ncu).   python tools/prof_shield.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_shield import full_batch, make_shield
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
sc, X, NOMX, NOMU, MOV, STAT = full_batch(n, 100, seed=5)
d = [torch.from_numpy(v).cuda() for v in (X, NOMX, NOMU, MOV, STAT)]
for mode in ("gatekeeper", "mps"):
    sh = make_shield(sc, mode, n, 100, keep_states=False)
    sh.step(*d); sh.step(*d)
    torch.cuda.synchronize()
