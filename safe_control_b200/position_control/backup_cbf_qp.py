"""BackupCBF -- same surface as the reference's position_control/backup_cbf_qp.py:33-826, B200 backend.

    shielding = BackupCBF(robot=dynamics, robot_spec=robot_spec, dt=0.1, backup_horizon=12.0)
    shielding.set_backup_controller(EvadeBackupController(...))          # backup_controller.py:420
    shielding.set_environment(EvadeEnv(...))                             # envs/evade_env.py:19
    shielding.set_moving_obstacles(get_obstacles)                        # callable t -> dict / list of dicts / None
    shielding.set_nominal_trajectory(nom_x, nom_u)
    u = shielding.solve_control_problem(state)                           # (2, 1) ndarray
    shielding.is_using_backup(); shielding.get_status()

Covered: the double integrator in the evade scene (examples/evade/test_evade.py --algo backupcbf), the only non-drift-car
use of the class.  Anything else (DriftingCar, Quad3D, other environments or backup policies) raises NotImplementedError:
there is no CPU path behind this class.

The scene is read ONCE per solve from the objects the caller handed over (environment geometry, policy gains and bounds,
robot_spec), so changing them between steps behaves as in the reference.  Moving obstacles: the callable is sampled at
t = 0 (position, size, 'vx' / 'vy' when the dict has them, else a finite difference of the callable over 1 s) and advanced
at constant velocity on the device -- exactly what the evade callable does (test_evade.py:373-385).

Differences (DESIGN.md "Boundary"): the QP is solved exactly (the reference uses OSQP, ~1e-4 accurate); visualisation
handles (`ax`) are accepted and ignored.
"""
import numpy as np

from .. import backup as _bk
from ._common import host_ctx


def _obstacle_rows(moving_obstacles):
    """-> [K, 8] rows (backup.py) from the reference's obstacle description (backup_cbf_qp.py:320-339, 418-442)."""
    if moving_obstacles is None:
        return None

    def sample(t):
        if callable(moving_obstacles):
            try:
                st = moving_obstacles(t)
            except TypeError:
                st = moving_obstacles()
        else:
            st = moving_obstacles
        if st is None:
            return []
        if isinstance(st, (list, tuple)):
            return [o for o in st if o is not None]
        return [st]

    now, later = sample(0.0), None
    rows = []
    for k, o in enumerate(now):
        if not o.get("active", True):
            continue
        x, y = float(o.get("x", 0)), float(o.get("y", 0))
        if "vx" in o or "vy" in o:
            vx, vy = float(o.get("vx", 0.0)), float(o.get("vy", 0.0))
        else:
            if later is None:
                later = sample(1.0)
            vx = float(later[k].get("x", 0)) - x if k < len(later) else 0.0
            vy = float(later[k].get("y", 0)) - y if k < len(later) else 0.0
        if "length" in o and "width" in o:
            rows.append([x, y, vx, vy, float(o["length"]), float(o["width"]), 0.0, float(_bk.KIND_RECT)])
        else:
            rows.append([x, y, vx, vy, 0.0, 0.0, float(o.get("radius", 1.0)), float(_bk.KIND_CIRCLE)])
    return np.array(rows, dtype=np.float64).reshape(-1, _bk.MOV_COLS)


class BackupCBF:
    def __init__(self, robot, robot_spec, dt=0.05, backup_horizon=2.0, ax=None, device=0):
        model = robot_spec.get("model", "DoubleIntegrator2D")
        if model not in ("DoubleIntegrator2D", "double_integrator"):
            raise NotImplementedError(f"BackupCBF on B200 covers DoubleIntegrator2D in the evade scene, not {model!r}")
        self.robot, self.robot_spec = robot, robot_spec
        self.dt, self.backup_horizon = dt, backup_horizon
        self.N = int(backup_horizon / dt)                                    # backup_cbf_qp.py:56
        self.n_states, self.n_controls = 4, 2
        self.nominal_controller = self.backup_controller = self.backup_target = None
        self.env = self.moving_obstacles = None
        self.nominal_x_traj = self.nominal_u_traj = None
        self.alpha, self.alpha_terminal = 1.0, 2.0                           # :91-92
        self.safety_margin = robot_spec.get("safety_margin", 0.0)            # :96-99
        self.Q_u = np.array([1.0, 1.0])
        self.ax, self.visualize_backup, self.backup_trajs, self.save_every_N, self.curr_step = ax, False, [], 5, 0
        self._using_backup = self._last_intervention = False
        self._last_h_min, self.global_min_h = 1.0, float("inf")
        self.latest_backup_trajectory = None
        self.status = "optimal"
        self.device = device

    # ---- configuration (backup_cbf_qp.py:138-171) ----------------------------------------------------------------------------
    def set_nominal_controller(self, nominal_controller):
        self.nominal_controller = nominal_controller

    def set_backup_controller(self, backup_controller, target=None):
        self.backup_controller, self.backup_target = backup_controller, target

    def set_environment(self, env):
        self.env = env

    def set_nominal_trajectory(self, nominal_x_traj, nominal_u_traj):
        if nominal_x_traj is not None:
            if nominal_x_traj.ndim == 2 and nominal_x_traj.shape[0] < nominal_x_traj.shape[1]:
                nominal_x_traj = nominal_x_traj.T
            self.nominal_x_traj = np.array(nominal_x_traj)
        if nominal_u_traj is not None:
            if nominal_u_traj.ndim == 2 and nominal_u_traj.shape[0] < nominal_u_traj.shape[1]:
                nominal_u_traj = nominal_u_traj.T
            self.nominal_u_traj = np.array(nominal_u_traj)

    def set_moving_obstacles(self, obstacles):
        self.moving_obstacles = obstacles

    def _get_nominal_control(self, state):                                   # :173-180
        if self.nominal_u_traj is not None and len(self.nominal_u_traj) > 0:
            return self.nominal_u_traj[0].flatten()
        if self.nominal_controller is not None:
            return np.array(self.nominal_controller(state.reshape(-1, 1))).flatten()
        return np.zeros(self.n_controls)

    def _scene(self):
        env, bc = self.env, self.backup_controller
        need = ("half_width", "hallway_length", "pocket_x_min", "pocket_x_max", "pocket_y_max", "get_pocket_bounds")
        if env is None or not all(hasattr(env, k) for k in need):
            raise NotImplementedError("BackupCBF on B200 needs the evade environment (envs/evade_env.py: hallway + pocket)")
        if bc is None or not all(hasattr(bc, k) for k in ("safe_center", "safe_bounds", "Kp", "Kd")):
            raise NotImplementedError("BackupCBF on B200 needs an EvadeBackupController-like backup policy "
                                      "(safe_center, safe_bounds, goal_bounds, Kp, Kd)")
        spec = self.robot_spec
        p = _bk.EvadeSceneParams(
            hallway_length=env.hallway_length, hallway_width=2 * env.half_width, radius=spec.get("radius", 0.5),
            a_max=spec.get("a_max", 2.0), v_max=spec.get("v_max", 1.5), safety_margin=self.safety_margin,
            dt=self.dt, backup_horizon=self.backup_horizon, alpha=self.alpha, alpha_terminal=self.alpha_terminal,
            Kp=bc.Kp, Kd=bc.Kd, goal_bounds=getattr(bc, "goal_bounds", None), use_goal=False,
            pocket_center=np.asarray(bc.safe_center).flatten(), pocket_bounds=bc.safe_bounds)
        # the barrier reads the ENVIRONMENT's pocket, the policy its own copy (identical in the example); the kernel has one
        pb = env.get_pocket_bounds()
        if any(abs(pb[k] - bc.safe_bounds[k]) > 0 for k in ("x_min", "x_max", "y_min", "y_max")):
            raise NotImplementedError("environment pocket and backup-policy pocket differ")
        p.half_width = env.half_width
        p.n_backup = self.N
        p.q0, p.q1 = float(self.Q_u[0]), float(self.Q_u[1])
        return p

    # ---- the solve (backup_cbf_qp.py:563-794) ------------------------------------------------------------------------------
    def solve_control_problem(self, robot_state, friction=None):
        x = np.ascontiguousarray(np.array(robot_state, dtype=np.float64).reshape(1, -1)[:, :4])
        u_ref = np.ascontiguousarray(np.asarray(self._get_nominal_control(x.reshape(-1)), dtype=np.float64).reshape(1, 2))
        mov = _obstacle_rows(self.moving_obstacles)
        if mov is not None and mov.shape[0] == 0:
            mov = None
        out = _bk.host_solve(host_ctx(self.device), self._scene(), x, u_ref, None if mov is None else mov[None].copy(),
                             want_phi=True)
        phi = out["phi"][0]
        self._last_h_min = float(out["h_min"][0])
        self.global_min_h = min(self.global_min_h, self._last_h_min)
        if self.visualize_backup and self.curr_step % self.save_every_N == 0:
            self.backup_trajs.append(phi.copy())
        self.latest_backup_trajectory = phi.copy()
        self.curr_step += 1
        self._using_backup = self._last_intervention = bool(out["intervene"][0])
        self.status = "optimal" if int(out["status"][0]) == 0 else "infeasible"
        return out["U"][0].reshape(-1, 1).copy()

    def is_using_backup(self):
        return self._using_backup

    def get_status(self):
        return {"using_backup": self._using_backup, "last_intervention": self._last_intervention,
                "backup_horizon": self.backup_horizon, "h_min": self._last_h_min, "global_min_h": self.global_min_h,
                "num_constraints": self.N}

    def get_backup_trajectories(self):
        return self.backup_trajs.copy() if self.visualize_backup else []

    def clear_trajectories(self):
        self.backup_trajs.clear()
