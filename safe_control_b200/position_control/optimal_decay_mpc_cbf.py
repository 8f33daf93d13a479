"""OptimalDecayMPCCBF -- surface of position_control/optimal_decay_mpc_cbf.py:15-397, B200 backend.

Constructor `(robot, robot_spec)` and `solve_control_problem(robot_state, control_ref, nearest_obs)` as in the reference
(tracking.py:149-151 constructs it without num_obs: 5 obstacle slots, :125, 271-280).  Cold start every step with the
previous solution's first stage (do-mpc's u0: inputs AND omegas, zeros at the first call), `state_machine != 'track'`
returns u_ref (:367-369), `.status` stays 'optimal' like the reference (:22; the true solver status is `.solver_status`),
the return value is `u[:n_controls]` (:396) and `.omega1 / .omega2` hold the optimal decay variables of stage 0."""
import numpy as np

from ..params import resolve_params, cbf_param_dict, NotCompatibleError  # noqa: F401
from ._common import host_ctx, obs_rows, status_string


class OptimalDecayMPCCBF:
    def __init__(self, robot, robot_spec, device=0):
        self.robot = robot
        self.robot_spec = robot_spec
        self.device = device
        self.num_obs = 5
        self.dt = getattr(robot, "dt", 0.05)
        self.params, self._spec = resolve_params(robot_spec, "optimal_decay_mpc_cbf", dt=self.dt)
        self.horizon = int(self._spec["mpc_horizon"])
        self.cbf_param = cbf_param_dict(self.params, "optimal_decay_mpc_cbf", robot_spec["model"])
        self.n_states, self.n_controls = self.params.nx, self.params.nu
        self.status = "optimal"
        self.solver_status = "optimal"
        self.u_prev = np.zeros((1, self.n_controls + 2))
        self.omega1 = self.omega2 = None
        self.pred_x = self.pred_u = None

    def solve_control_problem(self, robot_state, control_ref, nearest_obs):
        if control_ref["state_machine"] != "track":
            return control_ref["u_ref"]
        X = np.ascontiguousarray(np.asarray(robot_state, dtype=np.float64).reshape(1, -1))
        goal = np.zeros((1, 2))
        g = np.asarray(control_ref["goal"], dtype=np.float64).reshape(-1)
        goal[0, : min(2, g.size)] = g[:2]
        OBS, nobs = obs_rows(nearest_obs, self.num_obs)
        nobs = np.maximum(nobs, 0).astype(np.int32)        # None -> all dummy obstacles (:373-375)
        out = host_ctx(self.device).mpccbf_solve(self.params, self.num_obs, self.horizon, X, goal,
                                                 np.ascontiguousarray(self.u_prev), OBS, nobs, want_pred=True)
        self.solver_status = status_string(out["status"][0])
        self.pred_x, self.pred_u = out["pred_x"][0], out["pred_u"][0]
        self.u_prev = out["U"].copy()
        self.omega1, self.omega2 = float(out["U"][0, -2]), float(out["U"][0, -1])
        return out["U"][0, : self.n_controls].reshape(-1, 1)
