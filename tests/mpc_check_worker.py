"""Worker process of parity_util.check_mpc_parallel: oracle check of every (r mod procs)-th sampled agent.
    python tests/mpc_check_worker.py <dir with in.npz + meta.json> <r> <procs>   -> one JSON line of check_mpc stats"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))

if __name__ == "__main__":
    import torch
    torch.set_num_threads(1)
    from parity_util import check_mpc
    tmp, r, procs = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    z = np.load(os.path.join(tmp, "in.npz"))
    meta = json.load(open(os.path.join(tmp, "meta.json")))
    out = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    n = z["X"].shape[0]
    stats = check_mpc(meta["spec"], meta["M"], meta["H"], z["X"], z["goal"], z["u_prev"], z["OBS"], z["nobs"], out,
                      sample=range(r, n, procs), min_agree=0.0, **meta["kw"])
    print(json.dumps(stats))
