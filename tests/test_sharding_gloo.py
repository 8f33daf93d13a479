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


def _plan_worker(rank, world, port, n_agents, q):
    """ShardPlan reused over several steps (buffers allocated once), even and ragged blocks, and the config-5 wrapper
    (ShardedMixedMPCCBF) with a stand-in for the per-rank CUDA solve."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from safe_control_b200.sharding import ShardPlan
        F64, I32 = torch.float64, torch.int32
        plan = ShardPlan(n_agents, {"X": ((4,), F64), "nobs": ((), I32)}, {"U": ((2,), F64), "status": ((), I32)}, "cpu")
        ok = True
        ptrs = None
        for step in range(3):
            g = torch.Generator().manual_seed(step)
            X = torch.rand((n_agents, 4), generator=g, dtype=F64)
            nobs = (torch.arange(n_agents, dtype=I32) + step) % 5
            blk = plan.scatter({"X": X, "nobs": nobs} if rank == 0 else None)
            assert blk["X"].shape[0] == plan.n_local
            out = {"U": torch.stack([blk["X"][:, 0] * 2, blk["X"][:, 1] + blk["nobs"].double()], 1), "status": (blk["nobs"] > 2).to(I32)}
            full = plan.gather(out)
            if rank == 0:
                ref_u = torch.stack([X[:, 0] * 2, X[:, 1] + nobs.double()], 1)
                ok = ok and torch.equal(full["U"], ref_u) and torch.equal(full["status"], (nobs > 2).to(I32))
                p = (full["U"].data_ptr(), full["status"].data_ptr())
                ok = ok and (ptrs is None or ptrs == p)           # same pre-sized output buffers every step
                ptrs = p
        # point-to-point form of the same plan (exact blocks, one coalesced group)
        from safe_control_b200.sharding import run_p2p
        X = torch.rand((n_agents, 4), generator=torch.Generator().manual_seed(9), dtype=F64)
        nobs = torch.arange(n_agents, dtype=I32) % 3
        ops, blk = plan.scatter_ops({"X": X, "nobs": nobs} if rank == 0 else None)
        run_p2p(ops)
        out = {"U": torch.stack([blk["X"][:, 2], blk["X"][:, 3] - blk["nobs"].double()], 1), "status": blk["nobs"].clone()}
        ops, full = plan.gather_ops(out)
        run_p2p(ops)
        if rank == 0:
            ok = ok and torch.equal(full["U"], torch.stack([X[:, 2], X[:, 3] - nobs.double()], 1)) and torch.equal(full["status"], nobs)
        # config-5 wrapper: three model groups with their own row shapes, stand-in solve
        from safe_control_b200.mixed import ShardedMixedMPCCBF
        specs = [{"model": "DynamicUnicycle2D"}, {"model": "KinematicBicycle2D"}, {"model": "Quad3D"}]
        counts = [n_agents, n_agents + 1, max(n_agents - 1, 1)]
        sh = ShardedMixedMPCCBF(specs, counts, num_obs=3, horizon=2, device="cpu", want_active=True)

        def fake(blocks, want_active=False):
            outs = []
            for grp, b in zip(sh.mixed.groups, blocks):
                n = b["X"].shape[0]
                outs.append({"U": b["X"][:, : grp.nu] + b["goal"][:, :1], "status": b["nobs"].clone(), "iters": b["nobs"] * 2,
                             "active": torch.full((n, grp.active_words), 7, dtype=torch.int64)})
            return outs
        sh.mixed.solve = fake
        ins = None
        if rank == 0:
            ins = []
            for grp, c in zip(sh.mixed.groups, counts):
                g = torch.Generator().manual_seed(c)
                ins.append({"X": torch.rand((c, grp.nx), generator=g, dtype=F64), "goal": torch.rand((c, grp.ngoal), generator=g, dtype=F64),
                            "u_prev": torch.zeros((c, grp.nu), dtype=F64), "OBS": torch.rand((c, 3, 7), generator=g, dtype=F64),
                            "nobs": (torch.arange(c, dtype=I32) % 4)})
        res = sh.solve(ins)
        if rank == 0:
            for grp, r, a in zip(sh.mixed.groups, res, ins):
                ok = ok and torch.equal(r["U"], a["X"][:, : grp.nu] + a["goal"][:, :1]) and torch.equal(r["status"], a["nobs"])
                ok = ok and torch.equal(r["iters"], a["nobs"] * 2) and tuple(r["active"].shape) == (a["X"].shape[0], grp.active_words)
                ok = ok and bool((r["active"] == 7).all())
            q.put(ok)
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_agents", [6, 7])
def test_shard_plan_reuse_and_config5_wrapper(n_agents):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_plan_worker, args=(r, 2, port, n_agents, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=10)


# ---- Backup-CBF QP sharded over two ranks: the per-rank solve is the CPU build of the REAL kernel body ----------------------
def _backup_worker(rank, world, port, n_agents, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import hostsim_util as H
        from oracle import backup_cbf as B
        from test_backupcbf import hs_solve, random_batch
        from safe_control_b200.backup import ShardedBackupCBF
        sc = B.EvadeScene(dt=0.1, backup_horizon=4.0)

        def solve_block(b):
            o = hs_solve(H.hostsim(False), sc, b["X"].numpy(), b["U_ref"].numpy(), b["MOV"].numpy())
            return {"U": torch.from_numpy(o["U"]), "status": torch.from_numpy(o["status"]), "intervene": torch.from_numpy(o["intervene"]),
                    "h_min": torch.from_numpy(o["h_min"])}

        sh = ShardedBackupCBF(n_agents, 2, None, "cpu", solve_block=solve_block)
        X, Ur, MOV = random_batch(sc, n_agents, seed=9)
        ins = {"X": torch.from_numpy(X), "U_ref": torch.from_numpy(Ur), "MOV": torch.from_numpy(MOV)}
        out = sh.solve(ins if rank == 0 else None)
        if rank == 0:
            ref = solve_block(ins)
            q.put(all(torch.equal(out[k], ref[k]) for k in ref) and tuple(out["U"].shape) == (n_agents, 2) and
                  int((ref["status"] == 0).sum()) > 0)
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_agents", [9, 32])
def test_backupcbf_sharded_two_ranks(n_agents):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_backup_worker, args=(r, 2, port, n_agents, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=10)
