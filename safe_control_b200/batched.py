"""Batched solves on torch CUDA tensors (float64), plus the host-buffer (numpy) path.

    ctrl = BatchedCBFQP(robot_spec, num_obs=16)
    U, status, active = ctrl.solve(X, U_ref, OBS, nobs=None)       # device tensors in, device tensors out

Semantics per agent are those of the reference's controllers
(position_control/cbf_qp.py:108-199, optimal_decay_cbf_qp.py:132-159, mpc_cbf.py:366-402);
torch is only used for device memory and streams.
"""
import ctypes as C

import numpy as np
import torch

from . import _abi
from ._lib import lib, check, require_cuda, ScbError
from .params import resolve_params, cbf_param_dict

F64, I32 = torch.float64, torch.int32


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_f64(t, shape, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == F64):
        raise TypeError(f"{name} must be a CUDA float64 tensor")
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    return t.contiguous()


def _dev_out(t, shape, dtype, name, dev):
    """caller-supplied output buffer: CUDA, contiguous, exact shape / dtype, same device as the inputs"""
    if t is None:
        return None
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise TypeError(f"out {name} must be a contiguous CUDA {dtype} tensor")
    if tuple(t.shape) != tuple(shape) or t.device != dev:
        raise ValueError(f"out {name} has shape {tuple(t.shape)} on {t.device}, expected {tuple(shape)} on {dev}")
    return t


def _dev_i32(t, N, name):
    if t is None:
        return None
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == I32 and tuple(t.shape) == (N,)):
        raise TypeError(f"{name} must be a CUDA int32 tensor of shape ({N},)")
    return t.contiguous()


class _Base:
    controller = None

    def __init__(self, robot_spec, num_obs, dt=0.05):
        self.params, self.robot_spec = resolve_params(robot_spec, self.controller, dt)
        self.model = self.robot_spec["model"]
        self.nx, self.nu = self.params.nx, self.params.nu
        self.num_obs = int(num_obs)
        self.dt = dt
        self.cbf_param = cbf_param_dict(self.params, self.controller, self.model)
        self.launches = 0

    def _obs_args(self, OBS, N):
        """OBS [N, M, 7] (per agent) or [M, 7] (shared) -> (tensor, stride in doubles)."""
        M = self.num_obs
        if OBS.dim() == 2:
            return _dev_f64(OBS, (M, 7), "OBS"), 0
        return _dev_f64(OBS, (N, M, 7), "OBS"), 7 * M

    @staticmethod
    def _nobs(nobs, N, dev):
        return _dev_i32(nobs, N, "nobs")


class BatchedCBFQP(_Base):
    """N independent CBF-QPs per call: min ||u - u_ref||^2 s.t. A u + b >= 0, box."""
    controller = "cbf_qp"

    def __init__(self, robot_spec, num_obs=1, dt=0.05):
        super().__init__(robot_spec, num_obs, dt)
        self.words = (self.num_obs + 2 * self.nu + 63) // 64

    def rows(self, X, OBS, nobs=None):
        """Constraint rows only: A [N, M, nu], b [N, M]."""
        require_cuda()
        N = X.shape[0]
        X = _dev_f64(X, (N, self.nx), "X")
        OBS, stride = self._obs_args(OBS, N)
        nobs = self._nobs(nobs, N, X.device)
        A = torch.empty((N, self.num_obs, self.nu), dtype=F64, device=X.device)
        b = torch.empty((N, self.num_obs), dtype=F64, device=X.device)
        check(lib().scb_cbfqp_rows(self.params, N, self.num_obs, _ptr(X), _ptr(OBS), stride, _ptr(nobs),
                                   _ptr(A), _ptr(b), _stream()), "scb_cbfqp_rows")
        self.launches += 1
        return A, b

    def solve(self, X, U_ref, OBS, nobs=None, out=None, want_active=True):
        """-> U [N, nu] f64, status [N] i32, active [N, words] int64 (bit pattern of u64) or None."""
        require_cuda()
        N = X.shape[0]
        X = _dev_f64(X, (N, self.nx), "X")
        U_ref = _dev_f64(U_ref, (N, self.nu), "U_ref")
        OBS, stride = self._obs_args(OBS, N)
        nobs = self._nobs(nobs, N, X.device)
        if out is None:
            U = torch.empty((N, self.nu), dtype=F64, device=X.device)
            status = torch.empty((N,), dtype=I32, device=X.device)
            active = torch.empty((N, self.words), dtype=torch.int64, device=X.device) if want_active else None
        else:
            U, status, active = out
            U = _dev_out(U, (N, self.nu), F64, "U", X.device)
            status = _dev_out(status, (N,), I32, "status", X.device)
            active = _dev_out(active, (N, self.words), torch.int64, "active", X.device)
            if U is None or status is None:
                raise TypeError("out must be (U, status, active-or-None)")
        check(lib().scb_cbfqp_solve(self.params, N, self.num_obs, _ptr(X), _ptr(U_ref), _ptr(OBS), stride,
                                    _ptr(nobs), _ptr(U), _ptr(status), _ptr(active), _stream()), "scb_cbfqp_solve")
        self.launches += 1
        return U, status, active


class BatchedOptimalDecayCBFQP(_Base):
    """N optimal-decay CBF-QPs; the single CBF row comes from each agent's nearest valid obstacle."""
    controller = "optimal_decay_cbf_qp"

    def __init__(self, robot_spec, num_obs=1, dt=0.05):
        super().__init__(robot_spec, num_obs, dt)

    def solve(self, X, U_ref, OBS, nobs=None):
        """-> U [N,2], omega [N,2], sel [N] i32, status [N] i32, active [N] int64"""
        require_cuda()
        N = X.shape[0]
        X = _dev_f64(X, (N, self.nx), "X")
        U_ref = _dev_f64(U_ref, (N, self.nu), "U_ref")
        OBS, stride = self._obs_args(OBS, N)
        nobs = self._nobs(nobs, N, X.device)
        dev = X.device
        U = torch.empty((N, 2), dtype=F64, device=dev)
        omega = torch.empty((N, 2), dtype=F64, device=dev)
        sel = torch.empty((N,), dtype=I32, device=dev)
        status = torch.empty((N,), dtype=I32, device=dev)
        active = torch.empty((N,), dtype=torch.int64, device=dev)
        check(lib().scb_odcbf_solve(self.params, N, self.num_obs, _ptr(X), _ptr(U_ref), _ptr(OBS), stride, _ptr(nobs),
                                    _ptr(U), _ptr(omega), _ptr(sel), _ptr(status), _ptr(active), _stream()),
              "scb_odcbf_solve")
        self.launches += 1
        return U, omega, sel, status, active


class BatchedMPCCBF(_Base):
    """N nonlinear MPC-CBF problems (horizon H) per call."""
    controller = "mpc_cbf"

    def __init__(self, robot_spec, num_obs=5, dt=0.05, horizon=None):
        super().__init__(robot_spec, num_obs, dt)
        self.horizon = int(horizon if horizon is not None else self.robot_spec.get("mpc_horizon", 10))
        self.ngoal = 3 if self.model == "Quad3D" else 2
        self._ws = None                 # scheduling scratch (scb_mpccbf_workspace_bytes), grown on demand
        self.schedule = True            # False: index order (scb_mpccbf_solve)
        self.active_words = int(lib().scb_mpc_active_words(self.params, self.num_obs, self.horizon))

    def _workspace(self, N, dev):
        need = int(lib().scb_mpccbf_workspace_bytes(N))
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty((need,), dtype=torch.uint8, device=dev)
        return self._ws, need

    def solve(self, X, goal, u_prev, OBS, nobs=None, U_ref=None, track=None, want_pred=False, want_active=False):
        """-> dict(U, status, iters, kkt[, pred_x, pred_u][, active [N, active_words] int64: bit pattern of the u64 mask,
        bit k*M + j = CBF row (stage k, obstacle slot j), then the input / velocity bounds, include/scb.h])"""
        require_cuda()
        N, H = X.shape[0], self.horizon
        X = _dev_f64(X, (N, self.nx), "X")
        goal = _dev_f64(goal, (N, self.ngoal), "goal")
        u_prev = _dev_f64(u_prev, (N, self.nu), "u_prev")
        if U_ref is not None:
            U_ref = _dev_f64(U_ref, (N, self.nu), "U_ref")
        OBS, stride = self._obs_args(OBS, N)
        nobs = self._nobs(nobs, N, X.device)
        track = _dev_i32(track, N, "track")
        if track is not None and U_ref is None:
            raise ValueError("track needs U_ref (agents with track == 0 get U_ref back)")
        dev = X.device
        U = torch.empty((N, self.nu), dtype=F64, device=dev)
        status = torch.empty((N,), dtype=I32, device=dev)
        iters = torch.empty((N,), dtype=I32, device=dev)
        active = torch.empty((N, self.active_words), dtype=torch.int64, device=dev) if want_active else None
        kkt = torch.empty((N,), dtype=F64, device=dev)
        px = torch.empty((N, H + 1, self.nx), dtype=F64, device=dev) if want_pred else None
        pu = torch.empty((N, H, self.nu), dtype=F64, device=dev) if want_pred else None
        ws, ws_bytes = self._workspace(N, dev) if self.schedule else (None, 0)
        check(lib().scb_mpccbf_solve_ws(self.params, N, self.num_obs, H, _ptr(X), _ptr(U_ref), _ptr(goal), _ptr(u_prev),
                                        _ptr(track), _ptr(OBS), stride, _ptr(nobs), _ptr(U), _ptr(status), _ptr(px),
                                        _ptr(pu), _ptr(iters), _ptr(kkt), _ptr(active), _ptr(ws), ws_bytes, _stream()),
              "scb_mpccbf_solve_ws")
        self.launches += int(lib().scb_mpccbf_launch_count(self.params, N, self.num_obs, H, int(self.schedule)))
        out = dict(U=U, status=status, iters=iters, kkt=kkt)
        if want_pred:
            out.update(pred_x=px, pred_u=pu)
        if want_active:
            out["active"] = active
        return out


class BatchedOptimalDecayMPCCBF(BatchedMPCCBF):
    """N optimal-decay MPC-CBF problems (position_control/optimal_decay_mpc_cbf.py): omega1, omega2 are two extra inputs of
    every stage, so u_prev / U_ref / U / pred_u carry nu + 2 columns [u, omega1, omega2] (`self.nu` is that width,
    `self.nu_model` the robot's own).  The reference fixes 5 obstacle slots (:125, 271-280) and horizon 10 (30 for VTOL2D)."""
    controller = "optimal_decay_mpc_cbf"

    def __init__(self, robot_spec, num_obs=5, dt=0.05, horizon=None):
        super().__init__(robot_spec, num_obs, dt, horizon)
        self.nu_model = self.nu
        self.nu = self.nu_model + 2


class HostContext:
    """Host-buffer (numpy) entry points: H2D + kernel + D2H inside one C call.
    This is the path a reference-side binding uses and what bench.py's e2e times.
    When every array is page-locked (e.g. `torch.from_numpy(a).pin_memory().numpy()`) the QP calls run
    zero-copy: the kernel reads / writes the host buffers over PCIe directly (env SCB_HOST_PATH=staged|mapped
    overrides the automatic choice, see csrc/scb_api.cu)."""

    def __init__(self, device=0):
        require_cuda()
        self._h = C.c_void_p()
        check(lib().scb_ctx_create(C.byref(self._h), int(device)), "scb_ctx_create")

    def close(self):
        if self._h:
            lib().scb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(lib().scb_ctx_launches(self._h))

    @staticmethod
    def _np(a, dtype, name, shape=None):
        """C-contiguous numpy array of `dtype` and (when given) exactly `shape`: the C call reads / writes
        prod(shape) elements through the raw pointer, so a mismatch must be a Python error, not a fault."""
        if a is None:
            return None
        if not (isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous):
            raise TypeError(f"{name} must be a C-contiguous numpy array of {dtype}")
        if shape is not None and tuple(a.shape) != tuple(shape):
            raise ValueError(f"{name} has shape {tuple(a.shape)}, expected {tuple(shape)}")
        return a

    @classmethod
    def _obs(cls, OBS, N, M):
        if not isinstance(OBS, np.ndarray) or OBS.ndim not in (2, 3):
            raise TypeError("OBS must be a numpy array [N, M, 7] or [M, 7]")
        if OBS.ndim == 2:
            return cls._np(OBS, np.float64, "OBS", (M, 7)), 0
        return cls._np(OBS, np.float64, "OBS", (N, M, 7)), 7 * M

    @staticmethod
    def _dims(params):
        nx, nu = _abi.MODEL_DIMS[int(params.model)]
        return nx, nu

    @staticmethod
    def _p(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    def cbfqp_solve(self, params, M, X, U_ref, OBS, nobs=None, out=None, want_active=True):
        N = X.shape[0]
        nx, nu = self._dims(params)
        X = self._np(X, np.float64, "X", (N, nx)); U_ref = self._np(U_ref, np.float64, "U_ref", (N, nu))
        OBS, stride = self._obs(OBS, N, M); nobs = self._np(nobs, np.int32, "nobs", (N,))
        words = (M + 2 * nu + 63) // 64
        if out is None:
            U = np.empty((N, nu)); status = np.empty(N, np.int32)
            active = np.empty((N, words), np.uint64) if want_active else None
        else:
            U, status, active = out
            U = self._np(U, np.float64, "out U", (N, nu)); status = self._np(status, np.int32, "out status", (N,))
            active = self._np(active, np.uint64, "out active", (N, words))
            if U is None or status is None:
                raise TypeError("out must be (U, status, active-or-None)")
        check(lib().scb_cbfqp_solve_host(self._h, params, N, M, self._p(X), self._p(U_ref), self._p(OBS), stride,
                                         self._p(nobs), self._p(U), self._p(status), self._p(active)),
              "scb_cbfqp_solve_host")
        return U, status, active

    def odcbf_solve(self, params, M, X, U_ref, OBS, nobs=None, out=None):
        N = X.shape[0]
        nx, nu = self._dims(params)
        X = self._np(X, np.float64, "X", (N, nx)); U_ref = self._np(U_ref, np.float64, "U_ref", (N, 2))
        OBS, stride = self._obs(OBS, N, M); nobs = self._np(nobs, np.int32, "nobs", (N,))
        if out is None:
            U = np.empty((N, 2)); omega = np.empty((N, 2)); sel = np.empty(N, np.int32)
            status = np.empty(N, np.int32); active = np.empty(N, np.uint64)
        else:
            U, omega, sel, status, active = out
            U = self._np(U, np.float64, "out U", (N, 2)); omega = self._np(omega, np.float64, "out omega", (N, 2))
            sel = self._np(sel, np.int32, "out sel", (N,)); status = self._np(status, np.int32, "out status", (N,))
            active = self._np(active, np.uint64, "out active", (N,))
            if U is None or status is None:
                raise TypeError("out must be (U, omega, sel, status, active)")
        check(lib().scb_odcbf_solve_host(self._h, params, N, M, self._p(X), self._p(U_ref), self._p(OBS), stride,
                                         self._p(nobs), self._p(U), self._p(omega), self._p(sel), self._p(status),
                                         self._p(active)), "scb_odcbf_solve_host")
        return U, omega, sel, status, active

    def mpccbf_solve(self, params, M, H, X, goal, u_prev, OBS, nobs=None, U_ref=None, track=None, want_pred=False,
                     want_active=False):
        N = X.shape[0]
        nx, nu = self._dims(params)
        if int(params.od_mpc):
            nu += 2                                             # optimal decay: [u, omega1, omega2]
        ng = 3 if int(params.model) == _abi.MODEL_IDS["Quad3D"] else 2
        X = self._np(X, np.float64, "X", (N, nx)); goal = self._np(goal, np.float64, "goal", (N, ng))
        u_prev = self._np(u_prev, np.float64, "u_prev", (N, nu)); OBS, stride = self._obs(OBS, N, M)
        nobs = self._np(nobs, np.int32, "nobs", (N,)); U_ref = self._np(U_ref, np.float64, "U_ref", (N, nu))
        track = self._np(track, np.int32, "track", (N,))
        if track is not None and U_ref is None:
            raise ValueError("track needs U_ref")
        U = np.empty((N, nu)); status = np.empty(N, np.int32); iters = np.empty(N, np.int32); kkt = np.empty(N)
        px = np.empty((N, H + 1, nx)) if want_pred else None
        pu = np.empty((N, H, nu)) if want_pred else None
        active = np.empty((N, int(lib().scb_mpc_active_words(params, M, H))), np.uint64) if want_active else None
        check(lib().scb_mpccbf_solve_host(self._h, params, N, M, H, self._p(X), self._p(U_ref), self._p(goal),
                                          self._p(u_prev), self._p(track), self._p(OBS), stride, self._p(nobs),
                                          self._p(U), self._p(status), self._p(px), self._p(pu), self._p(iters),
                                          self._p(kkt), self._p(active)), "scb_mpccbf_solve_host")
        out = dict(U=U, status=status, iters=iters, kkt=kkt)
        if want_pred:
            out.update(pred_x=px, pred_u=pu)
        if want_active:
            out["active"] = active
        return out
