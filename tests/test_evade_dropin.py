"""CPU-only integration test of the evade-scene drop-in classes against THE REFERENCE'S OWN OBJECTS.

examples/evade/test_evade.py:265-470 is replayed with the reference's unmodified EvadeEnv, DoubleIntegrator2D and
EvadeBackupController (imported through oracle/refshim) but with OUR BackupCBF / Gatekeeper / MPS classes in the place of
the reference's -- the classes read the scene from those real objects (geometry, gains, bounds, the bullet callable), so this
pins the host logic of the drop-in boundary (scene extraction, obstacle sampling, state queries) end to end.  The device is
replaced by the CPU build of the kernel bodies (tests/_hostsim): on the GPU box tests/test_gpu_backup.py and
test_gpu_shield.py run the same classes on the real kernels.  Every step is compared with the recorded runs of the
reference's own classes (tests/golden/ref_backupcbf.npz, ref_shield.npz).
"""
import ctypes as C
import os
import sys
from unittest import mock

import numpy as np
import pytest

from conftest import ROOT
from oracle import refshim
from safe_control_b200 import _abi
import hostsim_util as H

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference checkout not present")


class HostShieldTwin:
    """what safe_control_b200.shield._new_shield returns, on the host build of csrc/scb_shield.cuh"""

    def __init__(self, mode, scene, event_offset, horizon_discount, nominal_steps, device):
        lib = H.hostsim(False)
        sp = _abi.ScbShieldParams()
        sp.scene = scene; sp.event_offset = event_offset; sp.mode = {"gatekeeper": 0, "mps": 1}[mode]
        hd = horizon_discount if horizon_discount is not None else 5 * scene.dt
        sp.discount_steps = max(1, int(hd / scene.dt)); sp.nom_cap = int(nominal_steps)
        self.params, self.T, self.Nb = sp, int(nominal_steps), int(scene.n_backup)
        L = self.T + self.Nb
        self.CU2 = np.zeros((1, 2, L, 2)); self.CX2 = np.zeros((1, 2, L + 1, 4))
        self.clen = np.full(1, -1, np.int32); self.cidx = np.zeros(1, np.int32); self.nsteps = np.zeros(1, np.int32)
        self.next_event = np.zeros(1); self.cbuf = np.zeros(1, np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self.st = _abi.ScbShieldState(p(self.CU2), p(self.CX2), p(self.clen), p(self.cidx), p(self.nsteps), p(self.next_event),
                                      p(self.cbuf), None)
        f = lib.hostsim_shield_step
        f.restype = C.c_int
        f.argtypes = [C.POINTER(_abi.ScbShieldParams), C.POINTER(_abi.ScbShieldState), C.c_int, C.c_int] + [C.c_void_p] * 5 + \
                     [C.c_long] + [C.c_void_p] * 3
        self.f = f

    def step_numpy(self, X, NOMX, NOMU, MOV=None, STAT=None, nom_len=None):
        keep = [None if a is None else np.ascontiguousarray(a) for a in (X, NOMX, NOMU, nom_len, MOV, STAT)]
        ptrs = [None if a is None else a.ctypes.data_as(C.c_void_p) for a in keep]
        K = 0 if MOV is None else MOV.shape[1]
        U = np.zeros((1, 2)); ub = np.zeros(1, np.int32)
        assert self.f(C.byref(self.params), C.byref(self.st), 1, K, ptrs[0], ptrs[1], ptrs[2], ptrs[3], ptrs[4], K * 8, ptrs[5],
                      U.ctypes.data_as(C.c_void_p), ub.ctypes.data_as(C.c_void_p)) == 0
        return U, ub

    def state_numpy(self, agent=0):
        clen, cb = int(self.clen[0]), int(self.cbuf[0])
        return dict(clen=clen, cidx=int(self.cidx[0]), nsteps=int(self.nsteps[0]), next_event=float(self.next_event[0]),
                    CU=self.CU2[0, cb, :clen].copy() if clen >= 0 else None, CX=self.CX2[0, cb, : clen + 1].copy() if clen >= 0 else None)


def host_backup_solve(ctx, params, X, U_ref, MOV=None, want_phi=False, want_rows=False, want_active=False):
    """safe_control_b200.backup.host_solve on the host build of csrc/scb_backup.cuh"""
    lib = H.hostsim(False)
    N, nb = X.shape[0], int(params.n_backup)
    K = 0 if MOV is None else MOV.shape[1]
    U = np.zeros((N, 2)); st = np.zeros(N, np.int32); iv = np.zeros(N, np.int32); hm = np.zeros(N)
    phi = np.zeros((N, nb, 4)); rows = np.zeros((N, nb, 3)); act = np.zeros((N, (nb + 4 + 63) // 64), np.uint64)
    f = lib.hostsim_backupcbf_solve
    f.restype = C.c_int
    f.argtypes = [C.POINTER(_abi.ScbBackupParams), C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_long] + [C.c_void_p] * 7
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    X = np.ascontiguousarray(X); U_ref = np.ascontiguousarray(U_ref); MOV = None if MOV is None else np.ascontiguousarray(MOV)
    assert f(C.byref(params), N, K, ptr(X), ptr(U_ref), ptr(MOV), K * 8, ptr(U), ptr(st), ptr(iv), ptr(hm), ptr(phi), ptr(rows), ptr(act)) == 0
    return dict(U=U, status=st, intervene=iv, h_min=hm, phi=phi, rows=rows, active=act)


@pytest.fixture()
def evade(monkeypatch):
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.startswith(("safe_control.", "matplotlib")) or k == "safe_control"}
    sys.modules.setdefault("matplotlib.collections", mock.MagicMock(name="matplotlib.collections"))
    refshim.install()
    from safe_control.envs.evade_env import EvadeEnv
    from safe_control.robots.double_integrator2D import DoubleIntegrator2D
    from safe_control.position_control.backup_controller import EvadeBackupController
    import safe_control_b200.shield as shield_mod
    import safe_control_b200.position_control.backup_cbf_qp as bk_mod
    monkeypatch.setattr(shield_mod, "_new_shield", HostShieldTwin)
    monkeypatch.setattr(bk_mod._bk, "host_solve", host_backup_solve)
    monkeypatch.setattr(bk_mod, "host_ctx", lambda device=0: None)

    def make(kind, bullet_x0=-10.0):
        """test_evade.py:272-386 with our class in the place of the reference's"""
        env = EvadeEnv(hallway_length=60.0, hallway_width=4.0, pocket_x=25.0, pocket_length=10.0, pocket_width=4.0,
                       goal_length=5.0, bullet_speed=3.0, bullet_length=3.0, bullet_start_x=-10.0)
        env._draw_bullet_bill = lambda: None
        env.bullet_x = bullet_x0
        spec = {"radius": 0.5, "a_max": 2.0, "v_max": 1.5, "model": "DoubleIntegrator2D", "safety_margin": 0.5}
        goal_bounds = {"x_min": env.goal_x_min, "x_max": env.goal_x_max, "y_min": -env.half_width, "y_max": env.half_width}
        backup = EvadeBackupController(spec, 0.1, env.get_pocket_center(), env.get_pocket_bounds(), goal_bounds)
        dyn = DoubleIntegrator2D(0.1, spec)
        if kind == "backupcbf":
            sh = bk_mod.BackupCBF(robot=dyn, robot_spec=spec, dt=0.1, backup_horizon=12.0, ax=None)
        elif kind == "mps":
            sh = shield_mod.MPS(robot=dyn, robot_spec=spec, dt=0.1, backup_horizon=12.0, event_offset=0.05, ax=None, safety_margin=0.5)
        else:
            sh = shield_mod.Gatekeeper(robot=dyn, robot_spec=spec, dt=0.1, backup_horizon=12.0, nominal_horizon=10.0,
                                       event_offset=0.05, ax=None, safety_margin=0.5)
        sh.set_backup_controller(backup)
        sh.set_environment(env)

        def get_obstacles(t=0.0):
            st = env.get_bullet_state()
            if not st["active"]:
                return None
            fut = st.copy()
            fut["x"] = st["x"] + st["vx"] * t
            return fut

        sh.set_moving_obstacles(get_obstacles)
        return env, spec, dyn, sh

    try:
        yield make
    finally:
        for k in [k for k in sys.modules if k.startswith("safe_control.") or k == "safe_control"]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


def nominal(spec, state):            # EvadeNominalController (test_evade.py:141-168), as in the fixture generators
    x, y, vx, vy = np.asarray(state, float).flatten()
    ax = 2.0 * (spec["v_max"] - vx)
    ay = 2.0 * (0.0 - y) + 2.0 * (0.0 - vy)
    a = np.sqrt(ax ** 2 + ay ** 2)
    if a > spec["a_max"]:
        ax, ay = ax * spec["a_max"] / a, ay * spec["a_max"] / a
    return np.array([[ax], [ay]])


def advance(env, spec, dyn, state, u):
    """test_evade.py:448-459"""
    state = dyn.step(state, u)
    vx, vy = state[2, 0], state[3, 0]
    vm = np.sqrt(vx ** 2 + vy ** 2)
    if vm > spec["v_max"]:
        state[2, 0] = vx * spec["v_max"] / vm
        state[3, 0] = vy * spec["v_max"] / vm
    env.step_bullet(0.1)
    return state


def test_backupcbf_closed_loop_with_the_reference_objects(evade):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_backupcbf.npz"))
    env, spec, dyn, sh = evade("backupcbf")
    state = np.array([20.0, 0.0, 0.0, 0.0]).reshape(-1, 1)
    n_backup = 0
    for k in range(160):
        # (closed loop on our own inputs: the 1e-10 differences of the finite-differenced rows feed back through the state)
        assert np.abs(state.flatten() - gold["loop_state"][k]).max() < 1e-6, k
        assert abs(env.bullet_x - gold["loop_bullet_x"][k]) < 1e-12
        uref = nominal(spec, state).reshape(1, 2)
        sh.set_nominal_trajectory(np.tile(state.reshape(1, -1), (4, 1)), np.tile(uref, (3, 1)))
        u = sh.solve_control_problem(state)
        assert u.shape == (2, 1) and np.abs(u.flatten() - gold["loop_u"][k]).max() < 1e-6, k
        assert sh.is_using_backup() == bool(gold["loop_using_backup"][k])
        assert abs(sh.get_status()["h_min"] - gold["loop_h_min"][k]) < 1e-6
        assert np.abs(sh.latest_backup_trajectory - gold["loop_phi"][k]).max() < 1e-5
        assert (sh.status == "optimal") == (gold["loop_qp_status"][k] == 0)
        n_backup += sh.is_using_backup()
        state = advance(env, spec, dyn, state, u)
    assert n_backup > 20 and sh.get_status()["num_constraints"] == 120


@pytest.mark.parametrize("algo,steps", [("gatekeeper", 320), ("mps", 220)])
def test_shields_closed_loop_with_the_reference_objects(evade, algo, steps):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_shield.npz"))
    g = lambda k: gold[f"{algo}_scenario_{k}"]
    env, spec, dyn, sh = evade(algo)
    assert sh.is_using_backup() and sh.get_status()["committed_length"] == 0
    state = np.array([20.0, 0.0, 0.0, 0.0]).reshape(-1, 1)
    steps = min(steps, g("u").shape[0])
    for k in range(steps):
        assert np.abs(state.flatten() - g("state")[k]).max() < 1e-9, k
        xs, us = [state.flatten()], []                 # rollout_nominal (test_evade.py:387-408)
        cur = state
        for _ in range(100):
            u_n = nominal(spec, cur)
            cur = dyn.step(cur, u_n)
            xs.append(cur.flatten()); us.append(u_n.flatten())
        sh.set_nominal_trajectory(np.array(xs), np.array(us))
        u = sh.solve_control_problem(state)
        assert u.shape == (2, 1) and np.abs(u.flatten() - g("u")[k]).max() < 1e-12, k
        assert sh.is_using_backup() == bool(g("using_backup")[k]), k
        st = sh.get_status()
        assert st["current_time_idx"] == g("idx")[k] and st["committed_length"] == g("clen")[k]
        assert abs(st["committed_horizon"] - g("horizon")[k]) < 1e-9 and abs(st["next_event_time"] - g("next_event")[k]) < 1e-12
        if k in g("snap_at"):
            j = list(g("snap_at")).index(k)
            cx, cu = sh.get_committed_trajectory()
            assert np.abs(cu - g("snap_cu")[j][: len(cu)]).max() < 1e-12 and cx.shape == (len(cu) + 1, 4)
        state = advance(env, spec, dyn, state, u)
    if algo == "gatekeeper":
        assert env.check_goal_reached(state[:2, 0])        # the example's pass criterion
