"""Multi-GPU plumbing: agents are independent (examples/test_multi_robot.py:77-85 steps its agents one
after the other and they never interact), so the batch is cut into contiguous blocks, one per rank, and
the ONLY communication is moving inputs to the ranks and the controls back.
There is no collective inside the solve.

Works with any torch.distributed backend: NCCL over NVLink on CUDA tensors in production, gloo on CPU
tensors in the tests (tests/test_sharding_gloo.py).
"""
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_agents: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced blocks: the first (n % world) ranks get one extra agent."""
    if world < 1:
        raise ValueError("world must be >= 1")
    base, extra = divmod(max(n_agents, 0), world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def scatter_rows(full: Optional[torch.Tensor], row_shape: Sequence[int], dtype, device, n_agents: int,
                 src: int = 0, group=None) -> torch.Tensor:
    """Rank `src` holds `full` [N, *row_shape]; every rank gets its block [n_r, *row_shape]."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(n_agents, world)
    lo, hi = bounds[rank]
    mine = torch.empty((hi - lo, *row_shape), dtype=dtype, device=device)
    # ragged blocks: pad to the largest so one scatter suffices
    width = max(h - l for l, h in bounds)
    buf = torch.empty((width, *row_shape), dtype=dtype, device=device)
    if rank == src:
        chunks = []
        for l, h in bounds:
            c = torch.zeros((width, *row_shape), dtype=dtype, device=device)
            c[: h - l] = full[l:h].to(device)
            chunks.append(c)
        dist.scatter(buf, chunks, src=src, group=group)
    else:
        dist.scatter(buf, None, src=src, group=group)
    mine.copy_(buf[: hi - lo])
    return mine


def gather_rows(mine: torch.Tensor, n_agents: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Inverse of scatter_rows: rank `dst` gets [N, *row_shape] in agent order, the others None."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(n_agents, world)
    width = max(h - l for l, h in bounds)
    row_shape = tuple(mine.shape[1:])
    buf = torch.zeros((width, *row_shape), dtype=mine.dtype, device=mine.device)
    buf[: mine.shape[0]] = mine
    if rank == dst:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.gather(buf, parts, dst=dst, group=group)
        return torch.cat([p[: h - l] for p, (l, h) in zip(parts, bounds)], dim=0)
    dist.gather(buf, None, dst=dst, group=group)
    return None


def solve_sharded(solve_block: Callable[[Dict[str, torch.Tensor]], Dict[str, torch.Tensor]],
                  inputs: Optional[Dict[str, torch.Tensor]], specs: Dict[str, Tuple[Tuple[int, ...], torch.dtype]],
                  n_agents: int, device, src: int = 0, group=None) -> Optional[Dict[str, torch.Tensor]]:
    """scatter -> per-rank solve on its block -> gather.  `inputs` only needs to exist on rank `src`;
    `specs` maps input name -> (row shape, dtype) so the other ranks can allocate.  Shared (per-scene)
    tensors are broadcast by the caller.  Returns the gathered outputs on rank `src`."""
    rank = dist.get_rank(group)
    block = {k: scatter_rows(inputs[k] if rank == src else None, shp, dt, device, n_agents, src, group)
             for k, (shp, dt) in specs.items()}
    out = solve_block(block)
    gathered = {k: gather_rows(v, n_agents, src, group) for k, v in out.items()}
    return gathered if rank == src else None
