"""Gatekeeper / MPS trajectory-rollout shields (shielding/gatekeeper.py, shielding/mps.py) for double-integrator agents in
the evade scene -- batched, with the shields' state resident on the device.

    sh = BatchedShield(n_agents, mode="gatekeeper", scene=EvadeSceneParams(...), nominal_steps=100, device="cuda:0")
    out = sh.step(X, NOMX, NOMU, MOV, STAT)          # one control step of every agent: one or two kernel launches
    out["U"], out["using_backup"]; sh.committed_horizon(), sh.current_time_idx, ...

and the drop-in classes `Gatekeeper` / `MPS` with the reference's constructor, set_* methods, solve_control_problem,
is_using_backup, get_status, get_committed_trajectory (N = 1, numpy in / out; examples/evade/test_evade.py:326-371 usage).

Per agent and step the kernel (csrc/scb_shield.cuh) evaluates every candidate of the reference's backward search at once
-- one candidate per lane: nominal prefix + 120-step backup rollout, each state checked against the walls, the bullet's
current hitbox and the moving obstacles at their predicted positions -- and commits the first valid one.
"""
import ctypes as C

import numpy as np
import torch

from . import _abi
from ._lib import lib, check, require_cuda
from .backup import EvadeSceneParams, MOV_COLS, KIND_RECT, KIND_CIRCLE  # noqa: F401
from .batched import _dev_f64, _dev_i32, _ptr, _stream, F64, I32

MODES = {"gatekeeper": 0, "mps": 1}


def shield_params(scene=None, mode="gatekeeper", event_offset=0.5, horizon_discount=None, nominal_steps=100):
    """-> _abi.ScbShieldParams.  Defaults follow Gatekeeper.__init__ (gatekeeper.py:43-69): horizon_discount = 5 dt."""
    sp = _abi.ScbShieldParams()
    if scene is None:
        scene = _abi.ScbBackupParams()
        lib().scb_backup_params_default(C.byref(scene))
    sp.scene = scene
    sp.event_offset = event_offset
    sp.mode = MODES[mode]
    hd = horizon_discount if horizon_discount is not None else 5 * scene.dt
    sp.discount_steps = max(1, int(hd / scene.dt))                         # gatekeeper.py:601
    sp.nom_cap = int(nominal_steps)
    return sp


class BatchedShield:
    def __init__(self, n_agents, mode="gatekeeper", scene=None, event_offset=0.5, horizon_discount=None, nominal_steps=100,
                 device="cuda", keep_states=False):
        require_cuda()
        self.params = shield_params(scene, mode, event_offset, horizon_discount, nominal_steps)
        self.N, self.T, self.Nb = int(n_agents), int(nominal_steps), int(self.params.scene.n_backup)
        self.mode = mode
        dev = torch.device(device)
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        L = self.T + self.Nb
        # committed trajectories are double-buffered on the device (a new commitment is built in the spare buffer, then swapped)
        self._CU2 = torch.zeros((self.N, 2, L, 2), dtype=F64, device=dev)
        self._CX2 = torch.zeros((self.N, 2, L + 1, 4), dtype=F64, device=dev) if keep_states else None
        self.cbuf = torch.zeros((self.N,), dtype=I32, device=dev)
        self._work = torch.zeros((self.N + 1,), dtype=I32, device=dev)        # pending list of the two-launch search
        self.clen = torch.full((self.N,), -1, dtype=I32, device=dev)
        self.cidx = torch.zeros((self.N,), dtype=I32, device=dev)
        self.nsteps = torch.zeros((self.N,), dtype=I32, device=dev)
        self.next_event = torch.zeros((self.N,), dtype=F64, device=dev)
        self._state = _abi.ScbShieldState(self._CU2.data_ptr(), self._CX2.data_ptr() if keep_states else None, self.clen.data_ptr(),
                                          self.cidx.data_ptr(), self.nsteps.data_ptr(), self.next_event.data_ptr(),
                                          self.cbuf.data_ptr(), self._work.data_ptr())
        self.device = dev
        self.launches = 0

    @property
    def CU(self):
        """[N, T + n_backup, 2]: every agent's committed input trajectory (rows beyond clen are stale)"""
        idx = self.cbuf.long().view(-1, 1, 1, 1).expand(-1, 1, *self._CU2.shape[2:])
        return self._CU2.gather(1, idx).squeeze(1)

    @property
    def CX(self):
        if self._CX2 is None:
            return None
        idx = self.cbuf.long().view(-1, 1, 1, 1).expand(-1, 1, *self._CX2.shape[2:])
        return self._CX2.gather(1, idx).squeeze(1)

    def reset(self):
        """forget every committed trajectory (the state of a freshly constructed Gatekeeper, gatekeeper.py:103-112)"""
        self.clen.fill_(-1); self.cidx.zero_(); self.nsteps.zero_(); self.next_event.zero_(); self.cbuf.zero_()

    def step(self, X, NOMX, NOMU, MOV=None, STAT=None, nom_len=None):
        """X [N,4]; NOMX [N,T+1,4], NOMU [N,T,2] nominal trajectories (NOMX[:,0] = start state); MOV [N,K,8] / [K,8] moving
        obstacles; STAT [N,5] current hitbox (x_min, x_max, y_min, y_max, active) or None; nom_len [N] i32 states available.
        -> dict(U [N,2], using_backup [N] i32)"""
        N, T = self.N, self.T
        X = _dev_f64(X, (N, 4), "X")
        NOMX = _dev_f64(NOMX, (N, T + 1, 4), "NOMX")
        NOMU = _dev_f64(NOMU, (N, T, 2), "NOMU")
        K, stride = 0, 0
        if MOV is not None:
            if MOV.dim() == 2:
                K = MOV.shape[0]; MOV = _dev_f64(MOV, (K, MOV_COLS), "MOV")
            else:
                K = MOV.shape[1]; MOV = _dev_f64(MOV, (N, K, MOV_COLS), "MOV"); stride = K * MOV_COLS
        if STAT is not None:
            STAT = _dev_f64(STAT, (N, 5), "STAT")
        nom_len = _dev_i32(nom_len, N, "nom_len")
        for t in (X, NOMX, NOMU, MOV, STAT, nom_len):
            if t is not None and t.device != self.device:
                raise ValueError(f"inputs must live on {self.device}")
        U = torch.empty((N, 2), dtype=F64, device=self.device)
        ub = torch.empty((N,), dtype=I32, device=self.device)
        check(lib().scb_shield_step(C.byref(self.params), C.byref(self._state), N, K, _ptr(X), _ptr(NOMX), _ptr(NOMU),
                                    _ptr(nom_len), _ptr(MOV) if K else None, stride, _ptr(STAT), _ptr(U), _ptr(ub), _stream()),
              "scb_shield_step")
        self.launches += 2 if (self.mode == "gatekeeper" and N >= 1024) else 1     # (csrc/scb_api.cu: two-launch search)
        return dict(U=U, using_backup=ub)

    def committed_horizon(self):
        """[N] seconds: committed_horizon of every agent (gatekeeper.py:538)"""
        return self.nsteps.to(F64) * self.params.scene.dt

    # ---- numpy in / out (what the N = 1 drop-in classes use) ----
    def step_numpy(self, X, NOMX, NOMU, MOV=None, STAT=None, nom_len=None):
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        out = self.step(t(X), t(NOMX), t(NOMU), t(MOV), t(STAT), t(nom_len))
        return out["U"].cpu().numpy(), out["using_backup"].cpu().numpy()

    def state_numpy(self, agent=0):
        """scalar state + committed trajectories of one agent as numpy"""
        clen = int(self.clen[agent].cpu())
        d = dict(clen=clen, cidx=int(self.cidx[agent].cpu()), nsteps=int(self.nsteps[agent].cpu()),
                 next_event=float(self.next_event[agent].cpu()), CU=None, CX=None)
        if clen >= 0:
            cb = int(self.cbuf[agent].cpu())
            d["CU"] = self._CU2[agent, cb, :clen].cpu().numpy()
            if self._CX2 is not None:
                d["CX"] = self._CX2[agent, cb, : clen + 1].cpu().numpy()
        return d


# ---------------------------------------------------------------------------------------------------------------- drop-in classes
def _new_shield(mode, scene, event_offset, horizon_discount, nominal_steps, device):
    """the N = 1 shield behind a drop-in object (one place to stand a test double in)"""
    return BatchedShield(1, mode, scene, event_offset, horizon_discount, nominal_steps, device=torch.device("cuda", device),
                         keep_states=True)


def _obstacle_rows(moving_obstacles):
    from .position_control.backup_cbf_qp import _obstacle_rows as rows
    return rows(moving_obstacles)


class Gatekeeper:
    """shielding/gatekeeper.py:35-754 surface for DoubleIntegrator2D in the evade environment (N = 1)."""
    _mode = "gatekeeper"

    def __init__(self, robot, robot_spec, dt=0.05, backup_horizon=2.0, event_offset=0.5, ax=None, nominal_horizon=None,
                 horizon_discount=None, safety_margin=1.0, device=0):
        model = robot_spec.get("model", "DynamicBicycle2D")
        if model not in ("DoubleIntegrator2D", "double_integrator"):
            raise NotImplementedError(f"{type(self).__name__} on B200 covers DoubleIntegrator2D in the evade scene, not {model!r}")
        self.robot, self.robot_spec, self.dt = robot, robot_spec, dt
        self.backup_horizon, self.event_offset = backup_horizon, event_offset
        self.horizon_discount = horizon_discount if horizon_discount is not None else 5 * dt
        self.safety_margin = safety_margin
        self.nominal_horizon = nominal_horizon if nominal_horizon is not None else backup_horizon
        self.n_states, self.n_controls = 4, 2
        self.nominal_controller = self.backup_controller = self.backup_target = None
        self.env = self.moving_obstacles = None
        self.nominal_x_traj = self.nominal_u_traj = None
        self.ax, self.visualize_backup, self.backup_trajs = ax, False, []
        self.device = device
        self._sh = None
        self._using_backup = True

    def set_nominal_controller(self, nominal_controller):
        """A nominal CONTROLLER (callable state (n, 1) -> input) instead of an external trajectory: its closed-loop rollout
        with `robot.step` (gatekeeper.py:235-269) is done here on the host, once per control step, and handed to the kernel as
        the nominal trajectory -- every candidate's nominal leg is a prefix of that one rollout."""
        self.nominal_controller = nominal_controller

    def _rollout_nominal(self, x):
        if self.robot is None or not hasattr(self.robot, "step"):
            raise NotImplementedError("a nominal controller needs robot.step to roll its plan out")
        steps = (len(self.nominal_x_traj) - 1) if self.nominal_x_traj is not None else int(self.nominal_horizon / self.dt)   # :592-599
        xs, us = [np.array(x, dtype=np.float64).reshape(-1)], []
        for _ in range(max(steps, 0)):
            u = np.array(self.nominal_controller(xs[-1].reshape(-1, 1)), dtype=np.float64).reshape(-1)
            us.append(u)
            xs.append(np.array(self.robot.step(xs[-1].reshape(-1, 1), u.reshape(-1, 1)), dtype=np.float64).reshape(-1))
        return np.array(xs), (np.array(us) if us else np.zeros((0, 2)))

    def set_backup_controller(self, backup_controller, target=None):
        self.backup_controller, self.backup_target = backup_controller, target

    def set_environment(self, env):
        self.env = env

    def set_nominal_trajectory(self, nominal_x_traj, nominal_u_traj):                 # gatekeeper.py:188-205
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

    def _scene(self):
        from .position_control.backup_cbf_qp import BackupCBF
        probe = BackupCBF(self.robot, dict(self.robot_spec, safety_margin=self.safety_margin), self.dt, self.backup_horizon)
        probe.set_environment(self.env); probe.set_backup_controller(self.backup_controller, self.backup_target)
        return probe._scene()

    def _static_rect(self):
        """the hitbox env.check_obstacle_collision tests (evade_env.py:454-485): the bullet where it is now"""
        e = self.env
        if not all(hasattr(e, k) for k in ("bullet_x", "bullet_y", "bullet_length", "bullet_width", "bullet_active")):
            return None
        return np.array([[e.bullet_x - e.bullet_length / 2, e.bullet_x + e.bullet_length / 2 + e.bullet_length / 3,
                          e.bullet_y - e.bullet_width / 2, e.bullet_y + e.bullet_width / 2, 1.0 if e.bullet_active else 0.0]])

    def solve_control_problem(self, robot_state, friction=None):
        x = np.ascontiguousarray(np.array(robot_state, dtype=np.float64).reshape(1, -1)[:, :4])
        if self.nominal_controller is not None:
            nx, nu = self._rollout_nominal(x[0])
        else:
            nx = self.nominal_x_traj if self.nominal_x_traj is not None else np.zeros((0, 4))
            nu = self.nominal_u_traj if self.nominal_u_traj is not None else np.zeros((0, 2))
        T = max(int(self.nominal_horizon / self.dt), len(nx) - 1, 1)
        if self._sh is None or self._sh.T < T:
            if self._sh is not None:
                raise NotImplementedError("nominal trajectory longer than the first one handed over")
            self._sh = _new_shield(self._mode, self._scene(), self.event_offset, self.horizon_discount, T, self.device)
        sh = self._sh
        sh.params.scene = self._scene()                           # (the scene objects may have changed between steps)
        NOMX = np.zeros((1, sh.T + 1, 4)); NOMU = np.zeros((1, sh.T, 2))
        n = min(len(nx), sh.T + 1)
        NOMX[0, :n] = nx[:n]; NOMU[0, : max(n - 1, 0)] = nu[: max(n - 1, 0)]
        mov = _obstacle_rows(self.moving_obstacles)
        U, ub = sh.step_numpy(x, NOMX, NOMU, None if mov is None or mov.shape[0] == 0 else mov[None], self._static_rect(),
                              np.array([n], np.int32))
        self._using_backup = bool(ub[0])
        return U.reshape(-1, 1)

    # ---- state queries (gatekeeper.py:721-754) ----
    def _st(self):
        return self._sh.state_numpy(0) if self._sh is not None else None

    @property
    def current_time_idx(self):
        return self._st()["cidx"] if self._sh is not None else int(self.backup_horizon / self.dt)

    @property
    def committed_horizon(self):
        return self._st()["nsteps"] * self.dt if self._sh is not None else 0.0

    @property
    def next_event_time(self):
        return self._st()["next_event"] if self._sh is not None else 0.0

    @property
    def committed_u_traj(self):
        return self._st()["CU"] if self._sh is not None else None

    @property
    def committed_x_traj(self):
        return self._st()["CX"] if self._sh is not None else None

    def get_committed_trajectory(self):
        return self.committed_x_traj, self.committed_u_traj

    def get_committed_horizon(self):
        return self.committed_horizon

    def get_backup_trajectories(self):
        return []

    def clear_trajectories(self):
        self.backup_trajs.clear()

    def is_using_backup(self):
        return self._using_backup

    def get_status(self):
        cu = self.committed_u_traj
        return {"current_time_idx": self.current_time_idx, "committed_horizon": self.committed_horizon,
                "next_event_time": self.next_event_time, "using_backup": self.is_using_backup(),
                "committed_length": len(cu) if cu is not None else 0}


class MPS(Gatekeeper):
    """shielding/mps.py:28-166 surface."""
    _mode = "mps"

    def __init__(self, robot, robot_spec, dt=0.05, backup_horizon=2.0, event_offset=0.5, ax=None, safety_margin=1.0, device=0):
        super().__init__(robot, robot_spec, dt, backup_horizon, event_offset, ax, safety_margin=safety_margin, device=device)
