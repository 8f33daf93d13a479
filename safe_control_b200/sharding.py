"""Multi-GPU plumbing: agents are independent (examples/test_multi_robot.py:77-85 steps its agents one
after the other and they never interact), so the batch is cut into contiguous blocks, one per rank, and
the ONLY communication is moving inputs to the ranks and the controls back.
There is no collective inside the solve.

    plan = ShardPlan(n_agents, in_specs, out_specs, device)        # every rank, once: all buffers pre-sized
    block = plan.scatter(inputs_on_src)                            # NCCL scatter: rank r gets rows [lo_r, hi_r)
    out = solve(block)                                             # per-rank launch(es), no communication
    full = plan.gather(out)                                        # NCCL gather -> [n_agents, ...] on src

or, for several tensors / several plans at once, the point-to-point form (exact unpadded blocks, everything coalesced
into one NCCL group -- what mixed.ShardedMixedMPCCBF uses for BASELINE config 5):

    ops, block = plan.scatter_ops(inputs_on_src); run_p2p(ops)
    ops, full = plan.gather_ops(out);             run_p2p(ops)

Works with any torch.distributed backend: NCCL over NVLink on CUDA tensors in production (bench.py --gpus N runs
BASELINE config 5 through it), gloo on CPU tensors in the tests (tests/test_sharding_gloo.py).
"""
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

Spec = Dict[str, Tuple[Tuple[int, ...], torch.dtype]]


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


class ShardPlan:
    """Scatter / gather of per-agent rows between rank `src` and all ranks, with every buffer allocated ONCE.

    Blocks are padded to the widest one (`width`) so a single collective per tensor suffices.  When the blocks
    are even (n_agents % world == 0) the collectives read / write the caller's [n_agents, ...] tensors in
    place through views -- no staging copy on either side; ragged batches go through one pre-allocated
    [world, width, ...] staging buffer per tensor on `src`.
    """

    def __init__(self, n_agents: int, in_specs: Spec, out_specs: Spec, device, src: int = 0, group=None):
        self.group, self.src = group, src
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n = int(n_agents)
        self.bounds = shard_bounds(self.n, self.world)
        self.lo, self.hi = self.bounds[self.rank]
        self.width = max(h - l for l, h in self.bounds)
        self.even = all(h - l == self.width for l, h in self.bounds)
        self.device = torch.device(device)
        self.in_specs, self.out_specs = dict(in_specs), dict(out_specs)
        mk = lambda shape, dt: torch.zeros(shape, dtype=dt, device=self.device)
        # per-rank receive blocks (inputs) -- solve() sees views [:n_r] of them
        self._in_block = {k: mk((self.width, *shp), dt) for k, (shp, dt) in self.in_specs.items()}
        # per-rank send blocks (outputs), only needed when the solve's output is narrower than `width`
        self._out_block = {k: mk((self.width, *shp), dt) for k, (shp, dt) in self.out_specs.items()}
        self._in_stage = self._out_stage = self._out_full = None
        self._mk = mk
        if self.rank == src:
            self._out_full = {k: mk((self.n, *shp), dt) for k, (shp, dt) in self.out_specs.items()}

    def _stages(self):
        """staging buffers of the COLLECTIVE form for ragged blocks, allocated at its first use (the point-to-point form
        needs none)"""
        if self._in_stage is None and self.rank == self.src and not self.even:
            self._in_stage = {k: self._mk((self.world, self.width, *shp), dt) for k, (shp, dt) in self.in_specs.items()}
            self._out_stage = {k: self._mk((self.world, self.width, *shp), dt) for k, (shp, dt) in self.out_specs.items()}

    @property
    def n_local(self) -> int:
        return self.hi - self.lo

    def scatter_bytes(self) -> int:
        return sum(self.n * int(torch.tensor([], dtype=dt).element_size()) * _numel(shp) for shp, dt in self.in_specs.values())

    def gather_bytes(self) -> int:
        return sum(self.n * int(torch.tensor([], dtype=dt).element_size()) * _numel(shp) for shp, dt in self.out_specs.values())

    # ---- point-to-point form: exact (unpadded) blocks, any number of plans coalesced into ONE NCCL group ------------------
    def scatter_ops(self, inputs: Optional[Dict[str, torch.Tensor]]):
        """-> (P2POp list, this rank's blocks).  Rank src sends rows [lo_r, hi_r) of every input tensor straight out of the
        caller's tensors (views, no staging, no padding); rank r receives into its pre-sized block buffers.  Hand the ops
        of several plans to run_p2p() together: NCCL then issues them as one group."""
        ops, out = [], {}
        for k, (shp, dt) in self.in_specs.items():
            if self.world == 1:
                out[k] = inputs[k]
                continue
            blk = self._in_block[k][: self.n_local]
            if self.rank == self.src:
                full = inputs[k]
                if tuple(full.shape) != (self.n, *shp) or full.dtype != dt:
                    raise ValueError(f"{k}: expected [{self.n}, {shp}] {dt}, got {tuple(full.shape)} {full.dtype}")
                full = full.contiguous()
                for r, (l, h) in enumerate(self.bounds):
                    if r == self.src:
                        out[k] = full[l:h]                               # own rows: a view, no copy
                    elif h > l:
                        ops.append(dist.P2POp(dist.isend, full[l:h], r, self.group))
            else:
                if self.n_local > 0:
                    ops.append(dist.P2POp(dist.irecv, blk, self.src, self.group))
                out[k] = blk
        return ops, out

    def gather_ops(self, outputs: Dict[str, torch.Tensor]):
        """-> (P2POp list, full outputs on rank src or None).  Every rank sends its rows, rank src receives them in place into
        its [n_agents, ...] buffers (valid until the next gather)."""
        ops, res = [], {}
        for k, (shp, dt) in self.out_specs.items():
            mine = outputs[k]
            if self.world == 1:
                res[k] = mine
                continue
            if self.rank == self.src:
                full = self._out_full[k]
                for r, (l, h) in enumerate(self.bounds):
                    if r == self.src:
                        full[l:h].copy_(mine)
                    elif h > l:
                        ops.append(dist.P2POp(dist.irecv, full[l:h], r, self.group))
                res[k] = full
            elif self.n_local > 0:
                ops.append(dist.P2POp(dist.isend, mine.contiguous(), self.src, self.group))
        return ops, (res if self.rank == self.src else None)

    def scatter(self, inputs: Optional[Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
        """`inputs[k]` = [n_agents, *shape] on rank src (device tensors; ignored elsewhere) -> this rank's rows."""
        out = {}
        self._stages()
        for k, (shp, dt) in self.in_specs.items():
            buf = self._in_block[k]
            if self.world == 1:
                out[k] = inputs[k]
                continue
            if self.rank == self.src:
                full = inputs[k]
                if tuple(full.shape) != (self.n, *shp) or full.dtype != dt:
                    raise ValueError(f"{k}: expected [{self.n}, {shp}] {dt}, got {tuple(full.shape)} {full.dtype}")
                if self.even:
                    chunks = list(full.contiguous().view(self.world, self.width, *shp).unbind(0))
                else:
                    st = self._in_stage[k]
                    for r, (l, h) in enumerate(self.bounds):
                        st[r, : h - l].copy_(full[l:h])
                    chunks = list(st.unbind(0))
                dist.scatter(buf, chunks, src=self.src, group=self.group)
            else:
                dist.scatter(buf, None, src=self.src, group=self.group)
            out[k] = buf[: self.n_local]
        return out

    def gather(self, outputs: Dict[str, torch.Tensor]) -> Optional[Dict[str, torch.Tensor]]:
        """`outputs[k]` = this rank's [n_r, *shape] -> [n_agents, *shape] on rank src (None elsewhere).  The
        returned tensors are the plan's own buffers: valid until the next gather()."""
        res = {}
        self._stages()
        for k, (shp, dt) in self.out_specs.items():
            mine = outputs[k]
            if self.world == 1:
                res[k] = mine
                continue
            if mine.shape[0] == self.width and mine.is_contiguous():
                send = mine
            else:
                send = self._out_block[k]
                send[: mine.shape[0]].copy_(mine)
            if self.rank == self.src:
                full = self._out_full[k]
                if self.even:
                    parts = list(full.view(self.world, self.width, *shp).unbind(0))
                    dist.gather(send, parts, dst=self.src, group=self.group)
                else:
                    st = self._out_stage[k]
                    dist.gather(send, list(st.unbind(0)), dst=self.src, group=self.group)
                    for r, (l, h) in enumerate(self.bounds):
                        full[l:h].copy_(st[r, : h - l])
                res[k] = full
            else:
                dist.gather(send, None, dst=self.src, group=self.group)
        return res if self.rank == self.src else None


def run_p2p(ops) -> None:
    """Issue a list of P2POps as one coalesced group (NCCL: a single ncclGroupStart/End) and wait for them."""
    if not ops:
        return
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def _numel(shape) -> int:
    n = 1
    for s in shape:
        n *= int(s)
    return n


def solve_sharded(solve_block: Callable[[Dict[str, torch.Tensor]], Dict[str, torch.Tensor]],
                  inputs: Optional[Dict[str, torch.Tensor]], specs: Spec,
                  n_agents: int, device, src: int = 0, group=None, out_specs: Optional[Spec] = None,
                  plan: Optional[ShardPlan] = None) -> Optional[Dict[str, torch.Tensor]]:
    """scatter -> per-rank solve on its block -> gather, in one call.  `inputs` only needs to exist on rank `src`;
    `specs` maps input name -> (row shape, dtype) so the other ranks can allocate.  Shared (per-scene) tensors are
    broadcast by the caller.  Returns the gathered outputs on rank `src`.  Pass a `plan` (or keep the one this
    returns through ShardPlan directly) to reuse the buffers across steps; without `out_specs` the output
    shapes are taken from the first solve."""
    if plan is None:
        plan = ShardPlan(n_agents, specs, out_specs or {}, device, src, group)
    block = plan.scatter(inputs)
    out = solve_block(block)
    if not plan.out_specs:                      # one-shot use: learn the output rows from the solve itself
        ospecs = {k: (tuple(v.shape[1:]), v.dtype) for k, v in out.items()}
        plan2 = ShardPlan(n_agents, {}, ospecs, device, src, group)
        return plan2.gather(out)
    return plan.gather(out)
