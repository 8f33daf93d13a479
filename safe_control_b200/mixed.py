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

    @property
    def launches(self):
        return sum(g.launches for g in self.groups)

    def solve(self, inputs: Sequence[Dict[str, torch.Tensor]]):
        """inputs[g] = dict(X, goal, u_prev, OBS, nobs) for group g -> list of BatchedMPCCBF.solve() dicts.
        Every group is enqueued on its own stream; the caller's current stream waits for all of them."""
        if self.streams is None:
            self.streams = [torch.cuda.Stream() for _ in self.groups]
        cur = torch.cuda.current_stream()
        outs = []
        for g, st, a in zip(self.groups, self.streams, inputs):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                outs.append(g.solve(a["X"], a["goal"], a["u_prev"], a["OBS"], a.get("nobs")))
        for st in self.streams:
            cur.wait_stream(st)
        return outs


def split_counts(n_agents: int, n_groups: int) -> List[int]:
    base, extra = divmod(n_agents, n_groups)
    return [base + (1 if g < extra else 0) for g in range(n_groups)]
