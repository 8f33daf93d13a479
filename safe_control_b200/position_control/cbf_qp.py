"""CBFQP -- same surface as the reference's position_control/cbf_qp.py:4-199, B200 backend.

    ctrl = CBFQP(robot, robot_spec, num_obs=10)
    u = ctrl.solve_control_problem(robot.X, {'u_ref': u_ref, ...}, obs_list)   # (nu, 1) ndarray
    ctrl.status == 'optimal'                                                   # tracking.py:628

Differences, all deliberate (DESIGN.md "Boundary"): an infeasible QP returns the last iterate
clipped to the input box instead of None (status still says 'infeasible', which is what the caller
checks); a batch of agents goes through safe_control_b200.BatchedCBFQP instead of a Python loop.
"""
import numpy as np

from ..params import resolve_params, cbf_param_dict
from ._common import host_ctx, obs_rows, status_string, PinnedIO


class CBFQP:
    def __init__(self, robot, robot_spec, num_obs=1, device=0):
        self.robot = robot
        self.robot_spec = robot_spec
        self.num_obs = int(num_obs)
        self.device = device
        self.params, self._spec = resolve_params(robot_spec, "cbf_qp", dt=getattr(robot, "dt", 0.05))
        self.cbf_param = cbf_param_dict(self.params, "cbf_qp", robot_spec["model"])
        self.status = "optimal"
        self.active = None
        self._io = None                 # pinned host buffers (first solve)

    def solve_control_problem(self, robot_state, control_ref, obs_list):
        u_ref = np.ascontiguousarray(np.asarray(control_ref["u_ref"], dtype=np.float64).reshape(1, -1))
        if obs_list is None:                       # cbf_qp.py:113-118: u_ref back, unclipped
            self.status = "optimal"
            return np.asarray(control_ref["u_ref"], dtype=np.float64).reshape(-1, 1)
        # the reference reads robot.X through the facade and ignores robot_state (cbf_qp.py:156-183)
        X = np.ascontiguousarray(np.asarray(getattr(self.robot, "X", robot_state), dtype=np.float64).reshape(1, -1))
        OBS, nobs = obs_rows(obs_list, self.num_obs)
        if self._io is None:
            import torch
            nx, nu, M = self.params.nx, self.params.nu, max(self.num_obs, 1)
            words = (self.num_obs + 2 * nu + 63) // 64
            self._io = PinnedIO(X=((1, nx), torch.float64), U_ref=((1, nu), torch.float64), OBS=((1, M, 7), torch.float64),
                                nobs=((1,), torch.int32), U=((1, nu), torch.float64), status=((1,), torch.int32),
                                active=((1, words), torch.int64))
        io = self._io
        out = (io["U"], io["status"], io["active"].view(np.uint64))
        U, st, act = host_ctx(self.device).cbfqp_solve(self.params, self.num_obs, io.put("X", X[:, : self.params.nx]),
                                                       io.put("U_ref", u_ref), io.put("OBS", OBS), io.put("nobs", nobs), out=out)
        self.status = status_string(st[0])
        self.active = act[0].copy()
        return U.copy().reshape(-1, 1)
