"""torchrun --nproc-per-node N tools/check_sharded_backup.py : Backup-CBF QPs of one batch held by rank 0 solved by N ranks over
NCCL (ShardedBackupCBF) == the same batch solved by rank 0 alone; prints the step time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from safe_control_b200 import BatchedBackupCBF, scenes
from safe_control_b200.backup import ShardedBackupCBF
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
N = 65536
sh = ShardedBackupCBF(N, 1, None, dev)
ins = None
if rank == 0:
    X, Ur, MOV = scenes.make_evade_batch(N, seed=1234)
    ins = {"X": torch.from_numpy(X).to(dev), "U_ref": torch.from_numpy(Ur).to(dev), "MOV": torch.from_numpy(MOV).to(dev)}
for _ in range(3):
    out = sh.solve(ins)
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = sh.solve(ins)
e1.record(); dist.barrier(); torch.cuda.synchronize()
if rank == 0:
    ref = BatchedBackupCBF().solve(ins["X"], ins["U_ref"], ins["MOV"])
    same = all(torch.equal(out[k], ref[k]) for k in ("U", "status", "intervene", "h_min"))
    ms = e0.elapsed_time(e1) / 10
    print(f"sharded backup-cbf over {world} ranks: identical to one GPU: {same}; {ms:.3f} ms per 65536-agent step incl. NCCL scatter / gather = {N / ms * 1e3:.3g} agents/s")
    assert same
dist.destroy_process_group()
