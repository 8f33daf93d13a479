"""Backup-CBF QP (position_control/backup_cbf_qp.py) for double-integrator agents in the evade scene, batched.

    ctrl = BatchedBackupCBF(EvadeSceneParams(...))               # or BatchedBackupCBF() for the example's defaults
    out = ctrl.solve(X, U_ref, MOV)                              # CUDA float64 tensors: [N,4], [N,2], [N,K,8] or [K,8]
    out["U"], out["status"], out["intervene"], out["h_min"]      # + phi / rows / active on request

Two kernel launches per call (csrc/scb_backup.cuh): rollout of the backup policy with forward-difference sensitivities +
the CBF rows, then the exact QP and the reference's fall-backs, one lane group per agent each.  Moving obstacles are rows
[x, y, vx, vy, length, width, radius, kind] (kind 0 absent, 1 rectangle, 2 circle), advanced at constant velocity -- what
the reference's callable does (examples/evade/test_evade.py:373-385).
"""
import ctypes as C

import numpy as np
import torch

from . import _abi
from ._lib import lib, check, require_cuda
from .batched import _dev_f64, _ptr, _stream, F64, I32, HostContext

MOV_COLS = 8
KIND_NONE, KIND_RECT, KIND_CIRCLE = 0, 1, 2


def EvadeSceneParams(hallway_length=60.0, hallway_width=4.0, pocket_x=25.0, pocket_length=10.0, pocket_width=4.0,
                     goal_length=5.0, radius=0.5, a_max=2.0, v_max=1.5, safety_margin=0.5, use_goal=True, dt=0.1,
                     backup_horizon=12.0, alpha=1.0, alpha_terminal=2.0, Kp=2.0, Kd=2.0, goal_bounds=None,
                     pocket_center=None, pocket_bounds=None):
    """-> _abi.ScbBackupParams from the numbers EvadeEnv (envs/evade_env.py:30-85), EvadeBackupController
    (backup_controller.py:431-454) and BackupCBF (backup_cbf_qp.py:41-110) hold."""
    p = _abi.ScbBackupParams()
    p.hallway_length = hallway_length
    p.half_width = hallway_width / 2
    if pocket_bounds is None:
        pocket_bounds = dict(x_min=pocket_x, x_max=pocket_x + pocket_length, y_min=p.half_width, y_max=p.half_width + pocket_width)
    p.pocket_x_min, p.pocket_x_max = pocket_bounds["x_min"], pocket_bounds["x_max"]
    p.pocket_y_min, p.pocket_y_max = pocket_bounds["y_min"], pocket_bounds["y_max"]
    if pocket_center is None:
        pocket_center = ((p.pocket_x_min + p.pocket_x_max) / 2, (p.pocket_y_min + p.pocket_y_max) / 2)
    p.center_x, p.center_y = float(pocket_center[0]), float(pocket_center[1])
    if goal_bounds is None and use_goal:
        goal_bounds = dict(x_min=hallway_length - goal_length, x_max=hallway_length, y_min=-p.half_width, y_max=p.half_width)
    p.use_goal = int(goal_bounds is not None)
    if goal_bounds is not None:
        p.goal_x_min, p.goal_x_max = goal_bounds["x_min"], goal_bounds["x_max"]
        p.goal_y_min, p.goal_y_max = goal_bounds["y_min"], goal_bounds["y_max"]
    p.radius, p.a_max, p.v_max, p.safety_margin = radius, a_max, v_max, safety_margin
    p.Kp, p.Kd = Kp, Kd
    p.dt, p.backup_horizon = dt, backup_horizon
    p.n_backup = int(backup_horizon / dt)                       # backup_cbf_qp.py:56
    p.alpha, p.alpha_terminal = alpha, alpha_terminal
    p.q0, p.q1 = 1.0, 1.0                                       # Q_u of the double integrator (backup_cbf_qp.py:110)
    return p


def bullet_row(bullet_x, bullet_length=3.0, bullet_width=4.0, bullet_speed=3.0, bullet_y=0.0, active=True):
    """EvadeEnv.get_bullet_state (envs/evade_env.py:386-406) as a moving-obstacle row."""
    return np.array([bullet_x + (bullet_length / 6), bullet_y, bullet_speed, 0.0, bullet_length * (1 + 1 / 3), bullet_width,
                     0.0, float(KIND_RECT if active else KIND_NONE)])


class BatchedBackupCBF:
    def __init__(self, params=None):
        if params is None:
            params = _abi.ScbBackupParams()
            lib().scb_backup_params_default(C.byref(params))
        self.params = params
        self.n_backup = int(params.n_backup)
        self.active_words = int(lib().scb_backup_active_words(self.n_backup))
        if self.active_words < 1:
            raise ValueError("n_backup must be >= 1")
        self.launches = 0
        self._rows = None               # [N, n_backup, 3] row buffer, kept between calls (with it the solve is two launches)

    def solve(self, X, U_ref, MOV=None, want_phi=False, want_rows=False, want_active=False):
        """-> dict(U [N,2], status [N] i32 (0 optimal, 1 = QP infeasible and U is the reference's fall-back), intervene [N] i32,
        h_min [N][, phi [N,n_backup,4]][, rows [N,n_backup,3]][, active [N,words] int64 bit pattern])"""
        require_cuda()
        N = X.shape[0]
        X = _dev_f64(X, (N, 4), "X")
        U_ref = _dev_f64(U_ref, (N, 2), "U_ref")
        K, stride = 0, 0
        if MOV is not None:
            if MOV.dim() == 2:
                K = MOV.shape[0]; MOV = _dev_f64(MOV, (K, MOV_COLS), "MOV")
            else:
                K = MOV.shape[1]; MOV = _dev_f64(MOV, (N, K, MOV_COLS), "MOV"); stride = K * MOV_COLS
        dev, nb = X.device, self.n_backup
        U = torch.empty((N, 2), dtype=F64, device=dev)
        status = torch.empty((N,), dtype=I32, device=dev)
        intervene = torch.empty((N,), dtype=I32, device=dev)
        h_min = torch.empty((N,), dtype=F64, device=dev)
        phi = torch.empty((N, nb, 4), dtype=F64, device=dev) if want_phi else None
        if self._rows is None or self._rows.shape[0] < N or self._rows.device != dev:
            self._rows = torch.empty((N, nb, 3), dtype=F64, device=dev)
        rows = self._rows[:N]
        active = torch.empty((N, self.active_words), dtype=torch.int64, device=dev) if want_active else None
        check(lib().scb_backupcbf_solve(C.byref(self.params), N, K, _ptr(X), _ptr(U_ref), _ptr(MOV) if K else None, stride,
                                        _ptr(U), _ptr(status), _ptr(intervene), _ptr(h_min), _ptr(phi), _ptr(rows),
                                        _ptr(active), _stream()), "scb_backupcbf_solve")
        self.launches += 2               # rollout + QP (csrc/scb_api.cu: scb_backupcbf_solve with a row buffer)
        out = dict(U=U, status=status, intervene=intervene, h_min=h_min)
        if want_phi:
            out["phi"] = phi
        if want_rows:
            out["rows"] = rows.clone()
        if want_active:
            out["active"] = active
        return out


def host_solve(ctx: HostContext, params, X, U_ref, MOV=None, want_phi=False, want_rows=False, want_active=False):
    """The same call on numpy host buffers (H2D + kernel + D2H inside scb_backupcbf_solve_host)."""
    N = X.shape[0]
    X = ctx._np(X, np.float64, "X", (N, 4)); U_ref = ctx._np(U_ref, np.float64, "U_ref", (N, 2))
    K, stride = 0, 0
    if MOV is not None:
        if MOV.ndim == 2:
            K = MOV.shape[0]; MOV = ctx._np(MOV, np.float64, "MOV", (K, MOV_COLS))
        else:
            K = MOV.shape[1]; MOV = ctx._np(MOV, np.float64, "MOV", (N, K, MOV_COLS)); stride = K * MOV_COLS
    nb = int(params.n_backup)
    words = int(lib().scb_backup_active_words(nb))
    U = np.empty((N, 2)); status = np.empty(N, np.int32); intervene = np.empty(N, np.int32); h_min = np.empty(N)
    phi = np.empty((N, nb, 4)) if want_phi else None
    rows = np.empty((N, nb, 3)) if want_rows else None
    active = np.empty((N, words), np.uint64) if want_active else None
    check(lib().scb_backupcbf_solve_host(ctx._h, C.byref(params), N, K, ctx._p(X), ctx._p(U_ref), ctx._p(MOV) if K else None,
                                         stride, ctx._p(U), ctx._p(status), ctx._p(intervene), ctx._p(h_min), ctx._p(phi),
                                         ctx._p(rows), ctx._p(active)), "scb_backupcbf_solve_host")
    out = dict(U=U, status=status, intervene=intervene, h_min=h_min)
    if want_phi:
        out["phi"] = phi
    if want_rows:
        out["rows"] = rows
    if want_active:
        out["active"] = active
    return out


class ShardedBackupCBF:
    """Backup-CBF QPs of a batch held by rank 0, solved by all ranks (agents are independent: contiguous blocks, no collective
    inside the solve; sharding.ShardPlan moves X / U_ref / MOV out and U / status / intervene / h_min back).

        sh = ShardedBackupCBF(n_agents, k_mov, params, device)           # every rank, once
        out = sh.solve(dict(X=.., U_ref=.., MOV=..) if rank == 0 else None)   # rank 0: dict of [n_agents, ...]; others: None

    `solve_block` (default: BatchedBackupCBF on this rank's device) is the per-rank solve; the gloo test substitutes the CPU
    build of the kernel body."""

    def __init__(self, n_agents, k_mov, params=None, device="cuda", src=0, group=None, solve_block=None):
        from .sharding import ShardPlan
        self.k_mov = int(k_mov)
        ins = {"X": ((4,), F64), "U_ref": ((2,), F64)}
        if self.k_mov > 0:
            ins["MOV"] = ((self.k_mov, MOV_COLS), F64)
        outs = {"U": ((2,), F64), "status": ((), I32), "intervene": ((), I32), "h_min": ((), F64)}
        self.plan = ShardPlan(n_agents, ins, outs, device, src, group)
        if solve_block is None:
            ctrl = BatchedBackupCBF(params)
            solve_block = lambda b: ctrl.solve(b["X"], b["U_ref"], b.get("MOV"))
        self._solve_block = solve_block

    def solve(self, inputs):
        from .sharding import run_p2p
        ops, block = self.plan.scatter_ops(inputs)
        run_p2p(ops)
        if self.plan.n_local > 0:
            out = self._solve_block(block)
            out = {k: out[k] for k in self.plan.out_specs}
        else:
            out = {k: torch.empty((0, *shp), dtype=dt, device=self.plan.device) for k, (shp, dt) in self.plan.out_specs.items()}
        ops, res = self.plan.gather_ops(out)
        run_p2p(ops)
        return res
