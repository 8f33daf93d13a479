"""OptimalDecayCBFQP -- surface of position_control/optimal_decay_cbf_qp.py:13-159, B200 backend.

The single CBF row is built from ONE obstacle: the nearest of the rows passed in
(tracking.py:585-586 `nearest_multi_obs[0]`; passing the 2-D list to the reference's own class
crashes at HEAD).  `.omega` holds the optimal decay variables."""
import numpy as np

from ..params import resolve_params, cbf_param_dict, NotCompatibleError  # noqa: F401
from ._common import host_ctx, obs_rows, status_string


class OptimalDecayCBFQP:
    def __init__(self, robot, robot_spec, num_obs=10, device=0):
        self.robot = robot
        self.robot_spec = robot_spec
        self.num_obs = int(num_obs)
        self.device = device
        self.params, self._spec = resolve_params(robot_spec, "optimal_decay_cbf_qp", dt=getattr(robot, "dt", 0.05))
        self.cbf_param = cbf_param_dict(self.params, "optimal_decay_cbf_qp", robot_spec["model"])
        self.status = "optimal"
        self.omega = None

    def solve_control_problem(self, robot_state, control_ref, nearest_obs):
        u_ref = np.ascontiguousarray(np.asarray(control_ref["u_ref"], dtype=np.float64).reshape(1, -1))
        X = np.ascontiguousarray(np.asarray(getattr(self.robot, "X", robot_state), dtype=np.float64).reshape(1, -1))
        OBS, nobs = obs_rows(nearest_obs, self.num_obs)
        nobs = np.maximum(nobs, 0).astype(np.int32)        # None -> zero row (optimal_decay_cbf_qp.py:134-138)
        U, om, sel, st, act = host_ctx(self.device).odcbf_solve(self.params, self.num_obs,
                                                                np.ascontiguousarray(X[:, :self.params.nx]), u_ref, OBS, nobs)
        self.status = status_string(st[0])
        self.omega = om[0]
        return U.reshape(-1, 1)
