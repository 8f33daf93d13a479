"""Oracle restatement of the reference's closed loop around the solve (TEST INFRASTRUCTURE).

One agent at a time, numpy float64, following

  tracking.py:37-182   LocalTrackingController.__init__ (X0 padding 60-99, num_constraints 134-138)
  tracking.py:216-260  set_waypoints / filter_waypoints
  tracking.py:262-267  goal_reached
  tracking.py:345-403  get_nearest_unpassed_obs (oracle/controllers.py:nearest_unpassed_obs)
  tracking.py:445-495  is_collide_unknown (known obstacles)
  tracking.py:497-535  update_goal
  tracking.py:559-668  control_step
  robots/robot.py:401-453, 854-872  nominal_input / stop / has_stopped / rotate_to / step / is_in_fov
  robots/<model>.py    stop, has_stopped, rotate_to (the rest is in oracle/models.py)
  attitude_control/velocity_tracking_yaw.py:35-62
  dynamic_env/main.py:54-58, 152  step_dyn_obs

Pinned against the reference's own LocalTrackingController run through oracle/refshim
(tests/golden/gen_tracking_from_reference.py -> tests/golden/ref_tracking.npz).
"""
import numpy as np

from .controllers import OracleCBFQP, OracleOptimalDecayCBFQP, nearest_unpassed_obs
from .models import angle_normalize, make_model


class OracleTrackingController:
    def __init__(self, X0, robot_spec, controller="cbf_qp", dt=0.05, enable_rotation=True, obs=None,
                 dynamic_obs=False, mpc_horizon=None, mpc_solver=None):
        self.spec = dict(robot_spec)
        self.name = self.spec["model"]
        self.dt = dt
        self.controller = controller
        self.enable_rotation = enable_rotation
        self.dynamic_obs = dynamic_obs
        self.model = make_model(self.spec, dt)
        self.spec = self.model.spec
        self.radius = self.model.radius
        self.state_machine = "idle"
        self.rotation_threshold = 0.1
        self.current_goal_index = 0
        self.reached_threshold = self.spec.get("reached_threshold", 0.3)
        X0 = np.asarray(X0, float).reshape(-1)
        # X0 padding (tracking.py:60-99) and yaw bookkeeping (robots/robot.py:65-131)
        if self.name == "SingleIntegrator2D":
            if X0.size == 2:
                X0 = np.array([X0[0], X0[1], 0.0])
            self.yaw = float(X0[2]); X0 = X0[:2]
        elif self.name == "DoubleIntegrator2D":                 # tracking.py:70-77, robots/robot.py:80-82
            if X0.size == 3:
                X0 = np.array([X0[0], X0[1], 0.0, 0.0, X0[2]])
            elif X0.size == 2:
                X0 = np.array([X0[0], X0[1], 0.0, 0.0, 0.0])
            self.yaw = float(X0[4]); X0 = X0[:4]
        elif self.name == "Quad3D":
            if X0.size == 2:
                X0 = np.concatenate([X0, np.zeros(10)])
            elif X0.size == 3:
                X0 = np.array([X0[0], X0[1], 0, 0, 0, X0[2], 0, 0, 0, 0, 0, 0], float)
            elif X0.size == 4:
                X0 = np.array([X0[0], X0[1], X0[2], 0, 0, X0[3], 0, 0, 0, 0, 0, 0], float)
            self.yaw = float(X0[5])
        elif self.name == "Unicycle2D":                         # no padding (robots/robot.py:84-90)
            self.yaw = float(X0[2])
        elif self.name == "Quad2D":                             # tracking.py:81-85: [x, z, theta, x_dot, z_dot, theta_dot]
            if X0.size in (2, 3):
                X0 = np.array([X0[0], X0[1], 0.0, 0.0, 0.0, 0.0])
            self.yaw = float(X0[2])
        else:
            if X0.size == 3:
                X0 = np.append(X0, 0.0)
            self.yaw = float(X0[2])
        self.X = X0.copy()
        self.fov_angle = np.deg2rad(float(self.spec.get("fov_angle", 70.0)))       # robots/robot.py:52-53
        self.num_constraints = int(self.spec.get("num_constraints", 10))
        self.obs = np.zeros((0, 7)) if obs is None else np.array(obs, float)
        self.u_att = None
        self.goal = None
        self.u_pos = None
        self.att = enable_rotation and self.name in ("SingleIntegrator2D", "DoubleIntegrator2D")
        if controller == "cbf_qp":
            self.pos = OracleCBFQP(self.spec, num_obs=self.num_constraints, dt=dt)
        elif controller == "optimal_decay_cbf_qp":
            self.pos = OracleOptimalDecayCBFQP(self.spec, dt=dt)
        elif controller == "mpc_cbf":
            from .mpc_cbf import OracleMPCCBF
            self.pos = OracleMPCCBF(self.spec, num_obs=self.num_constraints, horizon=mpc_horizon, dt=dt)
            self.u_prev = np.zeros(self.model.nu)
            self.mpc_solver = mpc_solver
        else:
            raise ValueError(controller)
        self.status = "optimal"

    # ---- robots/robot.py facade pieces --------------------------------------------------------
    def is_in_fov(self, point):
        if self.name == "Quad2D":                               # robots/robot.py:858-860: always in view
            return True
        to_point = np.asarray(point[:2], float) - self.X[:2]
        ang = np.arctan2(to_point[1], to_point[0])
        return abs(angle_normalize(ang - self.yaw)) <= self.fov_angle / 2

    def stop(self):
        n, X, s = self.name, self.X, self.spec
        if n == "SingleIntegrator2D":
            return np.zeros(2)
        if n == "DoubleIntegrator2D":                           # double_integrator2D.py:147-153
            k_a = s.get("nominal_k_a", 1.0)
            return np.array([k_a * (0.0 - X[2]), k_a * (0.0 - X[3])])
        if n == "Unicycle2D":                                   # unicycle2D.py:88-89
            return np.zeros(2)
        if n == "Quad2D":                                       # quad2D.py:152-161: nominal input towards the current position
            return self.model.nominal_input(X, X[0:2])
        if n == "DynamicUnicycle2D":
            return np.array([s.get("nominal_k_a", 1.0) * (0.0 - X[3]), 0.0])
        if n.startswith("KinematicBicycle2D"):
            return np.zeros(2)
        m = self.model                                          # Quad3D (quad3D.py:208-236)
        k = 1.0
        ax, ay, az = -k * X[6], -k * X[7], -k * X[8]
        th, ph, F = ax / m.gravity, -ay / m.gravity, s["mass"] * az
        w = np.array([F, s["Iy"] * k * (th - X[3] - X[9] / k), s["Ix"] * k * (ph - X[4] - X[10] / k),
                      s["Iz"] * k * (0 - X[5] - X[11] / k)])
        return np.clip(np.linalg.pinv(m.B2) @ w, s["u_min"], s["u_max"])

    def has_stopped(self):
        n, X = self.name, self.X
        if n in ("SingleIntegrator2D", "Unicycle2D"):
            return True
        if n == "DoubleIntegrator2D":                           # :155-156
            return np.linalg.norm(X[2:4]) < 0.05
        if n == "Quad2D":                                       # quad2D.py:163-165
            return np.linalg.norm(X[3:5]) < 0.05
        if n == "Quad3D":
            return np.linalg.norm(X[6:9]) < 0.05 and np.linalg.norm(X[9:12]) < 0.05
        return abs(X[3]) < 0.05

    def rotate_to(self, theta):
        """-> (u_ref, u_att) as tracking.py:589-597 uses them."""
        n, X, s = self.name, self.X, self.spec
        if n in ("SingleIntegrator2D", "DoubleIntegrator2D"):   # rotate_to(yaw, theta); u_ref = stop() (tracking.py:592-594)
            w = np.clip(2.0 * angle_normalize(theta - self.yaw), -s["w_max"], s["w_max"])
            return self.stop(), float(w)
        if n == "Quad3D":
            m = self.model; k = 2.0
            w = np.array([s["mass"] * m.gravity, s["Iy"] * k * (0 - X[3] - X[9] / k), s["Ix"] * k * (0 - X[4] - X[10] / k),
                          s["Iz"] * k * (theta - X[5] - X[11] / k)])
            return np.clip(np.linalg.pinv(m.B2) @ w, s["u_min"], s["u_max"]), self.u_att
        return np.array([0.0, 2.0 * angle_normalize(theta - X[2])]), self.u_att

    def nominal_input(self, goal):
        od = self.controller == "optimal_decay_cbf_qp"
        k_omega, k_a, k_v = (3.0, 0.5, 0.5) if od else (2.0, 1.0, 1.0)
        n, m = self.name, self.model
        if n == "SingleIntegrator2D":
            return m.nominal_input(self.X, goal, 0.05, k_v)
        if n == "Unicycle2D":                                   # facade: (X, goal, d_min, k_omega, k_v) (robots/robot.py:404-405)
            return m.nominal_input(self.X, goal, 0.05, k_omega, k_v)
        if n == "DoubleIntegrator2D":                           # facade: (X, goal, d_min, k_v, k_a) (robots/robot.py:408-409)
            return m.nominal_input(self.X, goal, 0.05, k_v, k_a)
        if n == "DynamicUnicycle2D":
            s = self.spec
            return m.nominal_input(self.X, goal, 0.05, s.get("nominal_k_omega", k_omega), s.get("nominal_k_a", k_a),
                                   s.get("nominal_k_v", k_v))
        if n.startswith("KinematicBicycle2D"):
            return m.nominal_input(self.X, goal, 0.05, k_omega, k_a, k_v)
        return m.nominal_input(self.X, goal)

    # ---- tracking.py ----------------------------------------------------------------------------
    def set_waypoints(self, waypoints):
        waypoints = np.array(waypoints, float)
        self.waypoints = self.filter_waypoints(waypoints)
        self.current_goal_index = 0
        self.goal = self.update_goal()
        if self.goal is not None:
            if not self.is_in_fov(self.goal):
                if self.spec.get("exploration", False):
                    self.state_machine = "rotate"
                else:
                    self.state_machine = "stop"
                    self.goal = None
            else:
                self.state_machine = "track"

    def filter_waypoints(self, waypoints):
        if len(waypoints) < 2:
            return waypoints
        if self.name == "Quad3D":
            n_pos = 3; robot_pos = self.X[0:3]
        else:
            n_pos = 2; robot_pos = self.X[0:2]
        aug = np.vstack((robot_pos, waypoints[:, :n_pos]))
        d = np.linalg.norm(np.diff(aug, axis=0), axis=1)
        mask = np.concatenate(([False], d >= self.reached_threshold))
        return aug[mask]

    def update_goal(self):
        n_pos = 3 if self.name == "Quad3D" else 2
        if self.state_machine == "rotate":
            rotate_goal = self.waypoints[self.current_goal_index]
            goal_angle = np.arctan2(rotate_goal[1] - self.X[1], rotate_goal[0] - self.X[0])
            if self.name == "Quad2D":                           # tracking.py:512-513: skips the 'rotate' state
                self.state_machine = "track"
            if not self.enable_rotation:
                self.state_machine = "track"
            if abs(self.yaw - goal_angle) > self.rotation_threshold:
                return rotate_goal[:n_pos]
            self.state_machine = "track"
            self.u_att = None
        if self.current_goal_index >= len(self.waypoints):
            return None
        wp = self.waypoints[self.current_goal_index]
        if np.linalg.norm(self.X[:2] - wp[:2]) < self.reached_threshold:
            self.current_goal_index += 1
            if self.current_goal_index >= len(self.waypoints):
                self.state_machine = "idle"
                return None
        return np.array(self.waypoints[self.current_goal_index][0:n_pos])

    def is_collide(self):
        pos = self.X[:2]
        for o in np.atleast_2d(self.obs) if len(self.obs) else []:
            se = np.isclose(o[6], 1.0) and not np.isclose(o[6], 0.0) and o[4] >= 2.0
            if not se:
                if np.linalg.norm(pos - o[:2]) < o[2] + self.radius:
                    return True
            else:
                ct, st = np.cos(o[5]), np.sin(o[5])
                xp = ct * (pos[0] - o[0]) + st * (pos[1] - o[1])
                yp = -st * (pos[0] - o[0]) + ct * (pos[1] - o[1])
                if (xp / (o[2] + self.radius)) ** o[4] + (yp / (o[3] + self.radius)) ** o[4] - 1 <= 0:
                    return True
        return False

    def control_step(self):
        """-> return code of tracking.py:559-668; self.info holds the per-step record."""
        if self.state_machine == "stop":
            if self.has_stopped():
                self.state_machine = "rotate" if self.enable_rotation else "track"
                self.goal = self.update_goal()
        else:
            self.goal = self.update_goal()

        if len(self.obs) == 0:
            sel, sel_idx = None, None
        else:
            sel, sel_idx = nearest_unpassed_obs(self.name, self.X[:2], self.yaw, self.obs, self.num_constraints)
        if self.dynamic_obs and len(self.obs):                 # dynamic_env/main.py:152 (after the selection copy)
            self.obs[:, 0] += self.obs[:, 3] * self.dt
            self.obs[:, 1] += self.obs[:, 4] * self.dt

        if self.state_machine == "rotate":
            goal_angle = np.arctan2(self.goal[1] - self.X[1], self.goal[0] - self.X[0])
            u_ref, self.u_att = self.rotate_to(goal_angle)
        elif self.goal is None:
            u_ref = self.stop()
        else:
            u_ref = self.nominal_input(self.goal)
        u_ref = np.asarray(u_ref, float).reshape(-1)

        info = dict(u_ref=u_ref.copy(), sel_idx=sel_idx, state_machine=self.state_machine,
                    goal=None if self.goal is None else np.array(self.goal, float))
        if self.controller == "cbf_qp":
            u, res = self.pos.solve(self.X, u_ref, sel)
            self.status = self.pos.status
        elif self.controller == "optimal_decay_cbf_qp":
            u, om, res = self.pos.solve(self.X, u_ref, None if sel is None else sel[0])
            self.status = self.pos.status
        else:
            if self.state_machine != "track":                  # mpc_cbf.py:379-381
                u = u_ref.copy()
            else:
                kw = {} if self.mpc_solver is None else dict(method=self.mpc_solver)
                u, _ = self.pos.solve(self.X, self.goal, self.u_prev, sel if sel is not None else np.zeros((0, 7)), **kw)
                self.u_prev = np.asarray(u, float).copy()
            self.status = "optimal"                            # hard-wired in the reference (mpc_cbf.py:10,400)
        info["u"] = None if u is None else np.asarray(u, float).copy()
        self.info = info

        if self.att and self.state_machine == "track":         # velocity_tracking_yaw.py:35-62
            if self.name == "DoubleIntegrator2D":              # heading follows the STATE velocity (:44-50, preview_time 0)
                vel = self.X[2:4]
            else:
                vel = u
            speed = np.hypot(vel[0], vel[1]) if vel is not None else 0.0
            if vel is None or speed < 1e-2:
                self.u_att = 0.0
            else:
                kp = float(self.spec.get("velocity_tracking_yaw_kp", 1.5))
                w_max = self.spec.get("w_max", 0.5)
                self.u_att = float(np.clip(kp * angle_normalize(np.arctan2(vel[1], vel[0]) - self.yaw), -w_max, w_max))

        collide = self.is_collide()
        if self.status != "optimal" or collide:
            return -2
        self.X = self.model.step(self.X, np.asarray(u, float))
        self.u_pos = np.asarray(u, float)
        if self.name in ("SingleIntegrator2D", "DoubleIntegrator2D"):
            if self.u_att is not None:
                self.yaw = float(angle_normalize(self.yaw + self.u_att * self.dt))
        elif self.name == "Quad3D":
            self.yaw = float(self.X[5])
        else:
            self.yaw = float(self.X[2])
        if self.is_collide():
            return -2
        if self.goal is None and self.state_machine != "stop":
            return -1
        return 0

    def run_all_steps(self, tf=30):
        total = 0
        for _ in range(int(tf / self.dt)):
            ret = self.control_step()
            total += ret
            if ret in (-1, -2):
                break
        return total
