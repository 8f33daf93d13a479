"""Loader for the CPU host-sim of the kernel bodies (test aid, see tests/_hostsim/hostsim.cpp)."""
import ctypes as C
import importlib.util
import os

import numpy as np

from safe_control_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}
_fma = False          # which build hostsim() returns; tests flip it with use_fma()


def use_fma(flag):
    global _fma
    _fma = bool(flag)


def _has_fma():
    try:
        with open("/proc/cpuinfo") as f:
            return " fma " in f.read().replace("\n", " ")
    except OSError:
        return False


def hostsim(fma=None):
    fma = _fma if fma is None else fma
    if fma and not _has_fma():
        fma = False
    if fma not in _libs:
        spec = importlib.util.spec_from_file_location("hostsim_build", os.path.join(_HERE, "_hostsim", "build.py"))
        mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
        lib = C.CDLL(mod.build(fma=fma))
        _abi.bind(lib, names=("scb_params_default", "scb_strerror", "scb_model_dims", "scb_active_words", "scb_version"))
        _libs[fma] = lib
    return _libs[fma]


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def hs_cbfqp_rows(p, X, OBS, nobs=None):
    lib = hostsim()
    N, M = OBS.shape[0], OBS.shape[1]
    X, OBS = f64(X), f64(OBS)
    A = np.zeros((N, M, p.nu)); b = np.zeros((N, M))
    no = None if nobs is None else np.ascontiguousarray(nobs, dtype=np.int32)
    rc = lib.hostsim_cbfqp_rows(C.byref(p), N, M, ptr(X), ptr(OBS), C.c_long(7 * M), ptr(no), ptr(A), ptr(b))
    assert rc == 0, rc
    return A, b


def hs_cbfqp_solve(p, X, Uref, OBS, nobs=None):
    lib = hostsim()
    N, M = OBS.shape[0], OBS.shape[1]
    X, Uref, OBS = f64(X), f64(Uref), f64(OBS)
    words = (M + 2 * p.nu + 63) // 64
    U = np.zeros((N, p.nu)); st = np.zeros(N, np.int32); act = np.zeros((N, words), np.uint64)
    no = None if nobs is None else np.ascontiguousarray(nobs, dtype=np.int32)
    rc = lib.hostsim_cbfqp_solve(C.byref(p), N, M, ptr(X), ptr(Uref), ptr(OBS), C.c_long(7 * M), ptr(no),
                                 ptr(U), ptr(st), ptr(act))
    assert rc == 0, rc
    return U, st, act


def hs_odcbf_solve(p, X, Uref, OBS, nobs=None):
    lib = hostsim()
    N, M = OBS.shape[0], OBS.shape[1]
    X, Uref, OBS = f64(X), f64(Uref), f64(OBS)
    U = np.zeros((N, 2)); om = np.zeros((N, 2)); sel = np.zeros(N, np.int32)
    st = np.zeros(N, np.int32); act = np.zeros(N, np.uint64)
    no = None if nobs is None else np.ascontiguousarray(nobs, dtype=np.int32)
    rc = lib.hostsim_odcbf_solve(C.byref(p), N, M, ptr(X), ptr(Uref), ptr(OBS), C.c_long(7 * M), ptr(no),
                                 ptr(U), ptr(om), ptr(sel), ptr(st), ptr(act))
    assert rc == 0, rc
    return U, om, sel, st, act


def hs_mpccbf_solve(p, H, X, goal, u_prev, OBS, nobs=None, want_active=False):
    lib = hostsim()
    N, M = OBS.shape[0], OBS.shape[1]
    X, goal, u_prev, OBS = f64(X), f64(goal), f64(u_prev), f64(OBS)
    nu = p.nu + (2 if p.od_mpc else 0)                  # optimal-decay MPC: [u, omega1, omega2] per stage
    assert u_prev.shape == (N, nu), (u_prev.shape, nu)
    U = np.zeros((N, nu)); st = np.zeros(N, np.int32); it = np.zeros(N, np.int32); kkt = np.zeros(N)
    px = np.zeros((N, H + 1, p.nx)); pu = np.zeros((N, H, nu))
    no = None if nobs is None else np.ascontiguousarray(nobs, dtype=np.int32)
    lib.scb_mpc_active_words.restype = C.c_int
    act = np.zeros((N, int(lib.scb_mpc_active_words(C.byref(p), M, H))), np.uint64) if want_active else None
    rc = lib.hostsim_mpccbf_solve(C.byref(p), N, M, H, ptr(X), ptr(goal), ptr(u_prev), ptr(OBS), C.c_long(7 * M), ptr(no),
                                  ptr(U), ptr(st), ptr(px), ptr(pu), ptr(it), ptr(kkt), ptr(act))
    assert rc == 0, rc
    out = dict(U=U, status=st, iters=it, kkt=kkt, pred_x=px, pred_u=pu)
    if want_active:
        out["active"] = act
    return out


# ---- closed loop (scb_track.cuh) ------------------------------------------------------------------
def hs_select_obstacles(p, X, scene, M, yaw=None):
    lib = hostsim()
    X, scene = f64(X), f64(scene)
    N, K = X.shape[0], scene.shape[0]
    OBS = np.zeros((N, M, 7)); nobs = np.zeros(N, np.int32); idx = np.zeros((N, M), np.int32)
    yw = None if yaw is None else f64(yaw)
    rc = lib.hostsim_select_obstacles(C.byref(p), N, K, M, ptr(X), ptr(yw), ptr(scene), C.c_long(0), ptr(OBS), ptr(nobs),
                                      ptr(idx))
    assert rc == 0, rc
    return OBS, nobs, idx


class HostSimTracker:
    """Drives hostsim_control_step on numpy arrays prepared by safe_control_b200.tracking.TrackerHostState."""

    def __init__(self, X0, robot_spec, controller="cbf_qp", dt=0.05, enable_rotation=True, obs=None, dynamic_obs=False):
        from safe_control_b200.tracking import TrackerHostState
        self.lib = hostsim()
        self.host = TrackerHostState(X0, robot_spec, controller, dt, enable_rotation, obs, dynamic_obs, lib=self.lib)
        self.params = self.host.params
        self.bufs = None

    def set_waypoints(self, wp):
        self.host.set_waypoints(wp)
        self._bind()

    def _bind(self):
        h = self.host
        assert self.lib.hostsim_track_sizeof() == C.sizeof(_abi.ScbTrack)
        self.bufs = {k: np.ascontiguousarray(getattr(h, k)) for k in h.STATE_ARRAYS}
        self.bufs["SCENE"] = np.ascontiguousarray(h.scene.copy())
        self.bufs.update(h.solve_buffers())
        t = h.config()
        for k, v in self.bufs.items():
            setattr(t, k, v.ctypes.data)
        self.t = t

    def load_state(self, **arrays):
        """Overwrite tracker state arrays in place (teacher forcing)."""
        for k, v in arrays.items():
            self.bufs[k][...] = v

    def control_step(self):
        self.lib.hostsim_control_step.restype = C.c_int
        rc = self.lib.hostsim_control_step(C.byref(self.params), C.byref(self.t))
        assert rc == 0, rc
        return self.bufs["ret"]
