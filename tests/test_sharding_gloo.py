"""world_size-2 gloo test (CPU) of the multi-GPU host logic: block bounds, ragged scatter/gather order.
The per-rank solve is a stand-in (the CUDA kernels need a GPU); what is under test is the plumbing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from safe_control_b200.sharding import shard_bounds, solve_sharded


def test_shard_bounds():
    assert shard_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert shard_bounds(0, 2) == [(0, 0), (0, 0)]
    for n in (1, 7, 1024, 65536):
        for w in (1, 2, 4, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [h - l for l, h in b]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_agents, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        X = torch.rand((n_agents, 4), generator=g, dtype=torch.float64)
        OBS = torch.rand((n_agents, 3, 7), generator=g, dtype=torch.float64)
        nobs = torch.arange(n_agents, dtype=torch.int32) % 4

        def fake_solve(b):           # deterministic per-agent function of the block's rows only
            U = torch.stack([b["X"][:, 0] + b["OBS"][:, 1, 2], b["X"][:, 3] * b["nobs"].double()], dim=1)
            return {"U": U, "status": (b["nobs"] > 2).to(torch.int32)}

        specs = {"X": ((4,), torch.float64), "OBS": ((3, 7), torch.float64), "nobs": ((), torch.int32)}
        out = solve_sharded(fake_solve, {"X": X, "OBS": OBS, "nobs": nobs} if rank == 0 else None, specs, n_agents, "cpu")
        if rank == 0:
            ref = fake_solve({"X": X, "OBS": OBS, "nobs": nobs})
            q.put((torch.equal(out["U"], ref["U"]), torch.equal(out["status"], ref["status"]), tuple(out["U"].shape)))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_agents", [5, 1, 64])
def test_scatter_solve_gather_two_ranks(n_agents):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_agents, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    same_u, same_s, shape = q.get(timeout=10)
    assert same_u and same_s and shape == (n_agents, 2)
