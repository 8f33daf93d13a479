"""Heterogeneous batches (BASELINE config 5: DynamicUnicycle2D / KinematicBicycle2D / Quad3D agents in one
job): agents are grouped by model on the host, each group is ONE launch of that model's kernel template, and
the groups run concurrently on separate CUDA streams.  The reference has no notion of a batch at all
(examples/test_multi_robot.py:77-85 loops over controller objects), so there is nothing to mirror but the
per-agent semantics, which are those of BatchedMPCCBF."""
from typing import Dict, List, Sequence

import torch

from .batched import BatchedMPCCBF


class MixedMPCCBF:
    def __init__(self, robot_specs: Sequence[dict], num_obs: int, horizon: int, dt: float = 0.05):
        self.groups: List[BatchedMPCCBF] = [BatchedMPCCBF(s, num_obs=num_obs, dt=dt, horizon=horizon) for s in robot_specs]
        self.streams = None
        # launch order: the group whose single hardest agent runs longest goes first, so that its tail overlaps the other
        # groups' bulk instead of ending the step alone (measured per-launch times at config-5 shapes: KinematicBicycle2D 17 ms,
        # Quad3D 12 ms, DynamicUnicycle2D 9.5 ms for 2731 agents; results do not depend on the order)
        cost = {"KinematicBicycle2D": 3, "KinematicBicycle2D_C3BF": 4, "KinematicBicycle2D_DPCBF": 4, "VTOL2D": 5, "Quad3D": 2}
        self.launch_order = sorted(range(len(self.groups)), key=lambda g: -cost.get(self.groups[g].model, 1))

    @property
    def launches(self):
        return sum(g.launches for g in self.groups)

    def solve(self, inputs: Sequence[Dict[str, torch.Tensor]], want_active: bool = False):
        """inputs[g] = dict(X, goal, u_prev, OBS, nobs) for group g -> list of BatchedMPCCBF.solve() dicts.
        Every group is enqueued on its own stream; the caller's current stream waits for all of them."""
        if self.streams is None:
            self.streams = [torch.cuda.Stream() for _ in self.groups]
        cur = torch.cuda.current_stream()
        outs = [None] * len(self.groups)
        for i in self.launch_order:
            g, st, a = self.groups[i], self.streams[i], inputs[i]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                outs[i] = g.solve(a["X"], a["goal"], a["u_prev"], a["OBS"], a.get("nobs"), want_active=want_active)
        for st in self.streams:
            cur.wait_stream(st)
        return outs


def split_counts(n_agents: int, n_groups: int) -> List[int]:
    base, extra = divmod(n_agents, n_groups)
    return [base + (1 if g < extra else 0) for g in range(n_groups)]


class ShardedMixedMPCCBF:
    """BASELINE config 5 as north_star describes it: rank `src` holds the whole heterogeneous batch, NCCL scatters
    each model group's rows over the ranks (contiguous blocks per group, so every GPU gets the same model mix --
    a Quad3D solve costs several DynamicUnicycle2D solves), every rank solves its blocks with one launch per group
    on concurrent streams (MixedMPCCBF), and NCCL gathers U / status / iters (/ active) back to `src`
    (SURVEY.md section 8e).  No communication inside the solve.  Every buffer is allocated once (ShardPlan).

    counts[g] = global number of agents of group g.  solve(inputs) takes, on `src`, inputs[g] = dict(X, goal, u_prev,
    OBS [n_g, M, 7], nobs) of device tensors, and returns there a list of dict(U, status, iters[, active])
    with n_g rows each (None on the other ranks)."""

    def __init__(self, robot_specs: Sequence[dict], counts: Sequence[int], num_obs: int, horizon: int, device,
                 dt: float = 0.05, want_active: bool = False, src: int = 0, group=None):
        from .sharding import ShardPlan
        self.mixed = MixedMPCCBF(robot_specs, num_obs, horizon, dt)
        self.want_active = bool(want_active)
        self.plans = []
        F64, I32 = torch.float64, torch.int32
        for g, n in zip(self.mixed.groups, counts):
            ins = {"X": ((g.nx,), F64), "goal": ((g.ngoal,), F64), "u_prev": ((g.nu,), F64),
                   "OBS": ((num_obs, 7), F64), "nobs": ((), I32)}
            outs = {"U": ((g.nu,), F64), "status": ((), I32), "iters": ((), I32)}
            if self.want_active:
                outs["active"] = ((g.active_words,), torch.int64)
            self.plans.append(ShardPlan(int(n), ins, outs, device, src, group))
        self.rank = self.plans[0].rank if self.plans else 0
        self.src = src

    @property
    def launches(self):
        return self.mixed.launches

    def scatter(self, inputs):
        """all 15 tensors (5 per model group) leave rank src in ONE coalesced NCCL group of point-to-point sends: exact
        blocks straight out of the caller's tensors, no padding, no staging copy"""
        from .sharding import run_p2p
        ops, blocks = [], []
        for g, pl in enumerate(self.plans):
            o, b = pl.scatter_ops(inputs[g] if inputs is not None else None)
            ops += o; blocks.append(b)
        run_p2p(ops)
        return blocks

    def solve_local(self, blocks):
        return self.mixed.solve(blocks, want_active=self.want_active)

    def gather(self, outs):
        from .sharding import run_p2p
        ops, res = [], []
        for pl, o in zip(self.plans, outs):
            p, r = pl.gather_ops({k: o[k] for k in pl.out_specs})
            ops += p; res.append(r)
        run_p2p(ops)
        return res if self.rank == self.src else None

    def solve(self, inputs):
        return self.gather(self.solve_local(self.scatter(inputs)))
