"""MPCCBF -- surface of position_control/mpc_cbf.py:6-402, B200 backend.

Keeps the reference's behaviours: cold start every step, u_prev = last applied input (0 first),
`state_machine != 'track'` returns u_ref without solving (:379-381), `.status` is always 'optimal'
like the reference (:10, :400) -- the true solver status is in `.solver_status`.  Predictions
(`.pred_x [H+1, nx]`, `.pred_u [H, nu]`) replace `mpc.opt_x_num` for consumers such as
attitude_control/gatekeeper_attitude.py:172-173."""
import numpy as np

from ..params import resolve_params, cbf_param_dict
from ._common import host_ctx, obs_rows, status_string


class MPCCBF:
    def __init__(self, robot, robot_spec, show_mpc_traj=False, num_obs=5, device=0):
        self.robot = robot
        self.robot_spec = robot_spec
        self.num_obs = int(num_obs)
        self.device = device
        self.show_mpc_traj = show_mpc_traj
        self.horizon = int(robot_spec.get("mpc_horizon", 10))
        self.dt = getattr(robot, "dt", 0.05)
        self.params, self._spec = resolve_params(robot_spec, "mpc_cbf", dt=self.dt)
        self.cbf_param = cbf_param_dict(self.params, "mpc_cbf", robot_spec["model"])
        self.n_states, self.n_controls = self.params.nx, self.params.nu
        self.status = "optimal"
        self.solver_status = "optimal"
        self.u_prev = np.zeros((1, self.n_controls))
        self.pred_x = self.pred_u = None

    def solve_control_problem(self, robot_state, control_ref, nearest_obs):
        if control_ref["state_machine"] != "track":
            return control_ref["u_ref"]
        X = np.ascontiguousarray(np.asarray(robot_state, dtype=np.float64).reshape(1, -1))
        ng = 3 if self.robot_spec["model"] == "Quad3D" else 2
        goal = np.zeros((1, ng))
        g = np.asarray(control_ref["goal"], dtype=np.float64).reshape(-1)
        goal[0, : min(ng, g.size)] = g[:ng]
        OBS, nobs = obs_rows(nearest_obs, self.num_obs)
        nobs = np.maximum(nobs, 0).astype(np.int32)        # None -> all dummy obstacles (mpc_cbf.py:341-343)
        if self.robot_spec["model"] in ("SingleIntegrator2D", "DynamicUnicycle2D", "DoubleIntegrator2D"):
            # superellipsoid rows (flag 1) take the general-row kernel (their agent_barrier_dt's if_else branch)
            self.params.mpc_superellipsoid = int(bool((OBS[0, : int(nobs[0]), 6] >= 0.5).any()))
        out = host_ctx(self.device).mpccbf_solve(self.params, self.num_obs, self.horizon, X, goal,
                                                 np.ascontiguousarray(self.u_prev), OBS, nobs, want_pred=True)
        self.solver_status = status_string(out["status"][0])
        self.pred_x, self.pred_u = out["pred_x"][0], out["pred_u"][0]
        self.u_prev = out["U"].copy()
        return out["U"].reshape(-1, 1)
