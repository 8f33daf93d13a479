"""BatchedTrackingController -- LocalTrackingController (tracking.py:36-747) for N agents on one GPU.

    tc = BatchedTrackingController(X0, robot_spec, controller_type={'pos': 'cbf_qp'}, obs=known_obs)
    tc.set_waypoints(waypoints)              # [W, 3] shared, or [N, W, 3] / list of per-agent arrays
    ret = tc.control_step()                  # int32 [N]: 0 / -1 (all waypoints reached) / -2 (infeasible or collision)
    tc.run_all_steps(tf=30)                  # the whole closed loop stays on the device

Every control step is scb_control_step (include/scb.h): waypoint state machine, obstacle selection,
nominal input, the controller's solve, collision checks and robot.step as CUDA kernels on the tracker
state held in device tensors.  What runs on the host, once, with numpy: X0 padding (tracking.py:60-99),
filter_waypoints and the initial state machine of set_waypoints (tracking.py:216-260).
Unknown-obstacle sensing (shapely footprints) and rendering are out of scope.
"""
import ctypes as C
import math

import numpy as np

from . import _abi
from .params import resolve_params, cbf_param_dict

_NPOS = {"Quad3D": 3}


def _angle_normalize(x):
    return ((x + np.pi) % (2 * np.pi)) - np.pi


def pad_initial_state(model, X0):
    """tracking.py:60-99 + robots/robot.py:65-131 -> (X [N, nx], yaw [N])."""
    X0 = np.array(np.atleast_2d(np.asarray(X0, dtype=np.float64)))       # always a private copy
    N, n = X0.shape
    if model == "SingleIntegrator2D":
        if n == 2:
            X0 = np.hstack([X0, np.zeros((N, 1))])
        elif n != 3:
            raise ValueError("Invalid initial state dimension for SingleIntegrator2D")
        return np.ascontiguousarray(X0[:, :2]), np.ascontiguousarray(X0[:, 2])
    if model == "Unicycle2D":                                    # no padding in tracking.py; robots/robot.py:84-90
        if n != 3:
            raise ValueError("Invalid initial state dimension for Unicycle2D")
        return np.ascontiguousarray(X0), np.ascontiguousarray(X0[:, 2])
    if model == "Quad2D":                                        # tracking.py:81-85
        if n in (2, 3):
            X0 = np.hstack([X0[:, 0:2], np.zeros((N, 4))])
        elif n != 6:
            raise ValueError("Invalid initial state dimension for Quad2D")
        return np.ascontiguousarray(X0), np.ascontiguousarray(X0[:, 2])
    if model == "DoubleIntegrator2D":                            # tracking.py:70-77; robots/robot.py:80-82: [x, y, vx, vy, theta]
        if n == 3:
            X0 = np.hstack([X0[:, 0:2], np.zeros((N, 2)), X0[:, 2:3]])
        elif n == 2:
            X0 = np.hstack([X0, np.zeros((N, 3))])
        elif n != 5:
            raise ValueError("Invalid initial state dimension for DoubleIntegrator2D")
        return np.ascontiguousarray(X0[:, :4]), np.ascontiguousarray(X0[:, 4])
    if model == "Quad3D":
        X = np.zeros((N, 12))
        if n == 2:
            X[:, 0:2] = X0
        elif n == 3:
            X[:, 0:2] = X0[:, 0:2]; X[:, 5] = X0[:, 2]
        elif n == 4:
            X[:, 0:3] = X0[:, 0:3]; X[:, 5] = X0[:, 3]
        elif n == 12:
            X = X0.copy()
        else:
            raise ValueError("Invalid initial state dimension for Quad3D")
        return np.ascontiguousarray(X), np.ascontiguousarray(X[:, 5])
    if n == 3:                                                   # DU / KB*: initial velocity 0
        X0 = np.hstack([X0, np.zeros((N, 1))])
    if X0.shape[1] != 4:
        raise ValueError(f"Invalid initial state dimension for {model}")
    return np.ascontiguousarray(X0), np.ascontiguousarray(X0[:, 2])


def pad_obstacles(obs):
    """tracking.py:277-291: rows of 3 / 5 / 7 columns -> [K, 7]."""
    if obs is None:
        return np.zeros((0, 7))
    obs = np.asarray(obs, dtype=np.float64)
    if obs.size == 0:
        return np.zeros((0, 7))
    if obs.ndim == 1:
        obs = obs.reshape(1, -1)
    if obs.shape[1] < 7:
        obs = np.hstack([obs, np.zeros((obs.shape[0], 7 - obs.shape[1]))])
    return np.array(obs[:, :7], order="C")                              # private copy (stepped in place when dynamic)


class TrackerHostState:
    """Host-side (numpy) preparation of the tracker arrays + the scb_track configuration.
    Shared by BatchedTrackingController (uploads them) and the CPU host-sim tests (uses them in place)."""

    def __init__(self, X0, robot_spec, controller="cbf_qp", dt=0.05, enable_rotation=True, obs=None,
                 dynamic_obs=False, lib=None):
        self.controller = controller
        self.params, self.spec = resolve_params(robot_spec, controller, dt, lib=lib)
        self.model = self.spec["model"]
        self.dt = dt
        self.nx, self.nu = self.params.nx, self.params.nu
        self.npos = _NPOS.get(self.model, 2)
        self.enable_rotation = bool(enable_rotation)
        self.dynamic_obs = bool(dynamic_obs)
        self.X, self.yaw = pad_initial_state(self.model, X0)
        self.N = self.X.shape[0]
        self.M = int(self.spec.get("num_constraints", 10))              # tracking.py:134-138
        self.H = int(self.spec.get("mpc_horizon", 10))                  # mpc_cbf.py:15
        self.reached_threshold = float(self.spec.get("reached_threshold", 0.3))
        self.fov_angle = math.radians(float(self.spec.get("fov_angle", 70.0)))     # robots/robot.py:52-53
        self.scene = pad_obstacles(obs)
        N, nu = self.N, self.nu
        self.sm = np.zeros(N, np.int32)
        self.wp_idx = np.zeros(N, np.int32)
        self.WP = np.zeros((N, 1, 3)); self.nwp = np.zeros(N, np.int32)
        self.goal = np.zeros((N, self.npos)); self.has_goal = np.zeros(N, np.int32)
        self.u_att = np.full(N, np.nan)
        self.u_prev = np.zeros((N, nu))
        self.ret = np.zeros(N, np.int32); self.done = np.zeros(N, np.int32); self.nsteps = np.zeros(N, np.int32)
        self.words = (self.M + 2 * nu + 63) // 64

    # ---- set_waypoints (tracking.py:216-260), vectorised over agents ------------------------------
    def set_waypoints(self, waypoints):
        N = self.N
        if isinstance(waypoints, (list, tuple)) and len(waypoints) == N and np.ndim(waypoints[0]) == 2:
            per_agent = [np.asarray(w, dtype=np.float64) for w in waypoints]
        else:
            w = np.asarray(waypoints, dtype=np.float64)
            per_agent = [w] * N if w.ndim == 2 else [w[i] for i in range(N)]
        filtered = [self._filter(i, w) for i, w in enumerate(per_agent)]
        W = max(1, max(len(w) for w in filtered))
        self.WP = np.zeros((N, W, 3)); self.nwp = np.zeros(N, np.int32)
        for i, w in enumerate(filtered):
            self.nwp[i] = len(w)
            if len(w):
                self.WP[i, : len(w), : w.shape[1]] = w[:, :3]
        self.wp_idx[:] = 0
        # goal = update_goal() (not in 'rotate' here), then the FOV test picks 'stop' or 'track' (:222-235)
        for i in range(N):
            g = self._update_goal_host(i)
            if g is None:
                self.has_goal[i] = 0
                continue
            to = g[:2] - self.X[i, :2]
            in_fov = abs(_angle_normalize(math.atan2(to[1], to[0]) - self.yaw[i])) <= self.fov_angle / 2
            if self.model == "Quad2D":                                   # robots/robot.py:858-860: always in view
                in_fov = True
            if not in_fov:
                if self.spec.get("exploration", False):
                    self.sm[i] = _abi.SM_ROTATE; self.has_goal[i] = 1; self.goal[i] = g[: self.npos]
                else:
                    self.sm[i] = _abi.SM_STOP; self.has_goal[i] = 0      # let the robot stop then rotate
            else:
                self.sm[i] = _abi.SM_TRACK; self.has_goal[i] = 1; self.goal[i] = g[: self.npos]
        self.done[:] = 0; self.ret[:] = 0

    def _filter(self, i, waypoints):
        """filter_waypoints (tracking.py:239-260) for agent i."""
        if len(waypoints) < 2:
            return waypoints
        n_pos = self.npos
        aug = np.vstack((self.X[i, :n_pos], waypoints[:, :n_pos]))
        d = np.linalg.norm(np.diff(aug, axis=0), axis=1)
        mask = np.concatenate(([False], d >= self.reached_threshold))
        return aug[mask]

    def _update_goal_host(self, i):
        """update_goal (tracking.py:497-535) outside the 'rotate' state (only used by set_waypoints)."""
        if self.wp_idx[i] >= self.nwp[i]:
            return None
        wp = self.WP[i, self.wp_idx[i]]
        if np.linalg.norm(self.X[i, :2] - wp[:2]) < self.reached_threshold:
            self.wp_idx[i] += 1
            if self.wp_idx[i] >= self.nwp[i]:
                self.sm[i] = _abi.SM_IDLE
                return None
        return self.WP[i, self.wp_idx[i]].copy()

    # ---- scb_track ----------------------------------------------------------------------------
    def config(self):
        """scb_track with the scalar configuration filled in (pointers still NULL)."""
        t = _abi.ScbTrack()
        s = self.spec
        od = self.controller == "optimal_decay_cbf_qp"
        k_omega, k_a, k_v = (3.0, 0.5, 0.5) if od else (2.0, 1.0, 1.0)          # tracking.py:599-604
        if self.model == "DynamicUnicycle2D":                                    # dynamic_unicycle2D.py:84-86
            k_omega = s.get("nominal_k_omega", k_omega); k_a = s.get("nominal_k_a", k_a); k_v = s.get("nominal_k_v", k_v)
        if self.model == "DoubleIntegrator2D":                                   # double_integrator2D.py:117-118
            k_a = s.get("nominal_k_a", k_a); k_v = s.get("nominal_k_v", k_v)
        t.controller = _abi.CONTROLLER_IDS[self.controller]
        t.N, t.K, t.M, t.W, t.H = self.N, self.scene.shape[0], self.M, self.WP.shape[1], self.H
        t.enable_rotation = int(self.enable_rotation)
        t.dynamic_obs = int(self.dynamic_obs)
        t.att_velocity_tracking = int(self.model in ("SingleIntegrator2D", "DoubleIntegrator2D") and self.enable_rotation)
        t.reached_threshold = self.reached_threshold
        t.rotation_threshold = 0.1                                               # tracking.py:50
        t.k_omega, t.k_a, t.k_v = float(k_omega), float(k_a), float(k_v)
        t.k_a_stop = float(s.get("nominal_k_a", 1.0))
        t.w_max = float(s.get("w_max", 0.5))
        t.att_kp = float(s.get("velocity_tracking_yaw_kp", 1.5))
        t.wheel_base = float(s.get("wheel_base", 0.4)); t.delta_max = float(s.get("delta_max", math.radians(32)))
        t.mpc_ws_bytes = 4 * (16 + 1024 + 2 * self.N)                            # scb_mpccbf_workspace_bytes(N)
        t.mpc_strict = int(bool(s.get("mpc_strict", False)))                     # ours: non-optimal MPC solve -> ret -2
        return t

    STATE_ARRAYS = ("X", "yaw", "sm", "wp_idx", "WP", "nwp", "goal", "has_goal", "u_att", "u_prev", "ret", "done",
                    "nsteps")

    def solve_buffers(self):
        N, M, nu = self.N, self.M, self.nu
        return dict(Uref=np.zeros((N, nu)), OBS=np.zeros((N, max(M, 1), 7)), nobs=np.zeros(N, np.int32),
                    U=np.zeros((N, nu)), status=np.zeros(N, np.int32), active=np.zeros((N, self.words), np.uint64),
                    track_flag=np.zeros(N, np.int32), mpc_iters=np.zeros(N, np.int32),
                    mpc_ws=np.zeros(16 + 1024 + 2 * N, np.int32), mpc_fail=np.zeros(N, np.int32))


class BatchedTrackingController:
    """N independent LocalTrackingControllers, state resident in HBM."""

    def __init__(self, X0, robot_spec, controller_type=None, dt=0.05, enable_rotation=True, obs=None,
                 dynamic_obs=False, device=None):
        import torch
        from ._lib import lib, require_cuda
        require_cuda()
        self._torch = torch
        self._lib = lib()
        pos = (controller_type or {}).get("pos", "cbf_qp") if isinstance(controller_type, (dict, type(None))) \
            else controller_type
        self.pos_controller_type = pos
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.host = TrackerHostState(X0, robot_spec, pos, dt, enable_rotation, obs, dynamic_obs)
        self.params, self.robot_spec = self.host.params, self.host.spec
        self.cbf_param = cbf_param_dict(self.params, pos, self.robot_spec["model"])
        self.N, self.dt = self.host.N, dt
        self.launches = 0
        self._t = None
        self._bufs = {}
        self._upload()

    # ---- device state -----------------------------------------------------------------------------
    def _upload(self):
        torch, h = self._torch, self.host
        d = {k: torch.from_numpy(np.ascontiguousarray(getattr(h, k))).to(self.device) for k in h.STATE_ARRAYS}
        d["SCENE"] = torch.from_numpy(h.scene.copy()).to(self.device)
        for k, v in h.solve_buffers().items():
            d[k] = torch.from_numpy(v.view(np.int64) if v.dtype == np.uint64 else v).to(self.device)
        self._bufs = d
        t = h.config()
        for k, v in d.items():
            setattr(t, k, v.data_ptr())
        self._t = t

    def set_waypoints(self, waypoints):
        self._sync_host()
        self.host.set_waypoints(waypoints)
        self._upload()

    def set_obs(self, obs):
        """tracking_controller.obs = known_obs (examples/test_tracking.py:168)."""
        self._sync_host()
        self.host.scene = pad_obstacles(obs)
        self._upload()

    def _sync_host(self):
        if not self._bufs:
            return
        for k in self.host.STATE_ARRAYS:
            setattr(self.host, k, self._bufs[k].cpu().numpy())
        self.host.scene = self._bufs["SCENE"].cpu().numpy()

    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream().cuda_stream)

    def _check(self, rc, what):
        from ._lib import check
        check(rc, what)

    # ---- the loop ---------------------------------------------------------------------------------
    def control_step(self):
        """One control_step() for every agent still running -> ret [N] int32 (device tensor)."""
        self._check(self._lib.scb_control_step(self.params, self._t, self._stream()), "scb_control_step")
        solve = 1
        if self.pos_controller_type == "mpc_cbf":              # key + counting sort + solve when the schedule applies
            solve = max(1, int(self._lib.scb_mpccbf_launch_count(self.params, self.N, self.host.M, self.host.H, 1)))
        self.launches += 2 + solve + int(self.host.dynamic_obs and self.host.scene.shape[0] > 0)
        return self._bufs["ret"]

    def run_steps(self, n_steps):
        self._check(self._lib.scb_run_all_steps(self.params, self._t, int(n_steps), self._stream()), "scb_run_all_steps")
        self.launches += int(self._lib.scb_run_all_steps_launches(self.params, self._t, int(n_steps)))

    def run_all_steps(self, tf=30, chunk=64):
        """tracking.py:711-747: int(tf/dt) steps, each agent stops at its first -1 / -2.
        -> ret [N] (numpy): the last return code of every agent."""
        total = int(tf / self.dt)
        k = 0
        while k < total:
            n = min(chunk, total - k)
            self.run_steps(n)
            k += n
            if bool(self._bufs["done"].all().item()):
                break
        return self._bufs["ret"].cpu().numpy()

    # ---- views --------------------------------------------------------------------------------------
    @property
    def X(self): return self._bufs["X"]
    @property
    def yaw(self): return self._bufs["yaw"]
    @property
    def state_machine(self): return self._bufs["sm"]
    @property
    def goal(self): return self._bufs["goal"]
    @property
    def status(self): return self._bufs["status"]
    @property
    def obs(self): return self._bufs["SCENE"]
    @property
    def nearest_multi_obs(self): return self._bufs["OBS"], self._bufs["nobs"]
    @property
    def done(self): return self._bufs["done"]
    @property
    def nsteps(self): return self._bufs["nsteps"]
    @property
    def mpc_fail(self):
        """mpc_cbf: per agent, the number of control steps whose MPC solve did not end 'optimal' (the reference's
        MPCCBF.status is hard-wired 'optimal', so this is extra visibility; robot_spec['mpc_strict'] = True makes such a
        step return -2 like a failed QP)."""
        return self._bufs["mpc_fail"]

    def get_control_input(self):
        return self._bufs["U"]

    def buffers(self):
        return self._bufs
