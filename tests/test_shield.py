"""CPU-only checks of the gatekeeper / MPS path (shielding/gatekeeper.py:553-672, shielding/mps.py:59-160):

* the oracle restatement (oracle/shielding.py) against tests/golden/ref_shield.npz -- closed-loop runs of the REFERENCE'S
  OWN Gatekeeper / MPS in the evade scenario (tests/golden/gen_shield_from_reference.py), step by step: input, backup flag,
  current_time_idx, committed horizon / length, next event time, committed input trajectories;
* the kernel body (csrc/scb_shield.cuh, host build, with and without FMA contraction) against the same runs, and against
  the oracle on seeded multi-step batches (moving + current-hitbox obstacles, short and empty nominal trajectories);
* the shadow route: the reference's own import paths resolve to the B200 classes.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import ROOT
from oracle import backup_cbf as B, shielding as S
from safe_control_b200 import _abi
import hostsim_util as H
from test_backupcbf import c_params

GOLD = os.path.join(ROOT, "tests", "golden", "ref_shield.npz")
RUNS = [("gatekeeper", "scenario"), ("gatekeeper", "cornered"), ("mps", "scenario"), ("mps", "cornered")]
T_NOM = 100


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def shield_c_params(sc, mode, T, event_offset=0.05, disc=5):
    sp = _abi.ScbShieldParams()
    sp.scene = c_params(sc); sp.event_offset = event_offset; sp.mode = int(mode == "mps"); sp.discount_steps = disc; sp.nom_cap = T
    return sp


class HostShield:
    """numpy twin of safe_control_b200.shield.BatchedShield on the host build of the kernel body"""

    def __init__(self, lib, sc, mode, n, T=T_NOM, **kw):
        self.lib, self.N, self.T, self.Nb = lib, n, T, sc.N
        self.sp = shield_c_params(sc, mode, T, **kw)
        self.CU2 = np.zeros((n, 2, T + self.Nb, 2)); self.CX2 = np.zeros((n, 2, T + self.Nb + 1, 4))
        self.clen = np.full(n, -1, np.int32); self.cidx = np.zeros(n, np.int32); self.nsteps = np.zeros(n, np.int32)
        self.next_event = np.zeros(n); self.cbuf = np.zeros(n, np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self.st = _abi.ScbShieldState(p(self.CU2), p(self.CX2), p(self.clen), p(self.cidx), p(self.nsteps), p(self.next_event),
                                      p(self.cbuf), None)
        f = lib.hostsim_shield_step
        f.restype = C.c_int
        f.argtypes = [C.POINTER(_abi.ScbShieldParams), C.POINTER(_abi.ScbShieldState), C.c_int, C.c_int] + [C.c_void_p] * 5 + \
                     [C.c_long] + [C.c_void_p] * 3
        self.f = f

    @property
    def CU(self):
        return self.CU2[np.arange(self.N), self.cbuf]

    @property
    def CX(self):
        return self.CX2[np.arange(self.N), self.cbuf]

    def step(self, X, NOMX, NOMU, MOV, STAT, nom_len=None):
        p = lambda a: None if a is None else np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
        keep = [np.ascontiguousarray(a) if a is not None else None for a in (X, NOMX, NOMU, nom_len, MOV, STAT)]
        K = 0 if MOV is None else MOV.shape[1]
        U = np.zeros((self.N, 2)); ub = np.zeros(self.N, np.int32)
        ptrs = [None if a is None else a.ctypes.data_as(C.c_void_p) for a in keep]
        assert self.f(C.byref(self.sp), C.byref(self.st), self.N, K, ptrs[0], ptrs[1], ptrs[2], ptrs[3], ptrs[4], K * 8, ptrs[5],
                      U.ctypes.data_as(C.c_void_p), ub.ctypes.data_as(C.c_void_p)) == 0
        return U, ub


def replay(gold, algo, tag, make_step):
    """feed a recorded run's states / bullet positions to a shield, compare everything the reference exposes per step"""
    g = lambda k: gold[f"{algo}_{tag}_{k}"]
    sc = B.EvadeScene()
    step, probe = make_step(sc)
    n = g("u").shape[0]
    for k in range(n):
        x = g("state")[k]
        nx, nu = S.nominal_rollout(sc, x)
        assert abs(nx.sum() + nu.sum() - g("nom_sum")[k]) < 1e-9          # the nominal plan the reference was handed
        act = bool(g("bullet_active")[k])
        mov = B.bullet_row(g("bullet_x")[k], active=act)[None]
        stat = S.bullet_static_rect(g("bullet_x")[k], active=act)
        u, ub = step(x, nx, nu, mov, stat)
        assert np.abs(u - g("u")[k]).max() < 1e-12, (algo, tag, k)
        st = probe()
        assert ub == bool(g("using_backup")[k]), (algo, tag, k)
        assert st["idx"] == g("idx")[k] and st["clen"] == g("clen")[k]
        assert abs(st["horizon"] - g("horizon")[k]) < 1e-9 and abs(st["next_event"] - g("next_event")[k]) < 1e-12
        if k in g("snap_at"):
            j = list(g("snap_at")).index(k)
            assert np.abs(st["cu"] - g("snap_cu")[j][: st["clen"]]).max() < 1e-12


@pytest.mark.parametrize("algo,tag", RUNS)
def test_oracle_reproduces_the_reference(gold, algo, tag):
    def make(sc):
        sh = S.OracleShield(sc, mode=algo)
        step = lambda x, nx, nu, mov, stat: (sh.solve(x, nx, nu, mov, stat), sh.is_using_backup())
        probe = lambda: dict(idx=sh.current_time_idx, clen=len(sh.committed_u), horizon=sh.committed_horizon,
                             next_event=sh.next_event_time, cu=sh.committed_u)
        return step, probe
    replay(gold, algo, tag, make)
    if (algo, tag) == ("gatekeeper", "scenario"):
        assert bool(gold["gatekeeper_scenario_reached_goal"])            # the example's pass criterion (test_evade.py:539-541)


@pytest.mark.parametrize("fma", [False, True])
@pytest.mark.parametrize("algo,tag", RUNS)
def test_kernel_body_reproduces_the_reference(gold, algo, tag, fma):
    def make(sc):
        hs = HostShield(H.hostsim(fma), sc, algo, 1)

        def step(x, nx, nu, mov, stat):
            U, ub = hs.step(x[None], nx[None], nu[None], mov[None], stat[None])
            return U[0], bool(ub[0])
        probe = lambda: dict(idx=int(hs.cidx[0]), clen=int(hs.clen[0]), horizon=hs.nsteps[0] * sc.dt, next_event=float(hs.next_event[0]),
                             cu=hs.CU[0, : hs.clen[0]])
        return step, probe
    replay(gold, algo, tag, make)


def shield_batch(sc, n, seed, T):
    """seeded agents around the hallway / pocket with their nominal plans, a bullet each (some inactive) and a slow disc"""
    rng = np.random.default_rng(seed)
    X = np.zeros((n, 4))
    X[:, 0] = rng.uniform(2.0, 58.0, n)
    X[:, 1] = np.where(rng.random(n) < 0.25, rng.uniform(0.0, 5.0, n), rng.uniform(-1.3, 1.3, n))
    X[:, 0] = np.where(X[:, 1] > 1.4, rng.uniform(26.0, 34.0, n), X[:, 0])
    X[:, 2:] = rng.uniform(-0.8, 0.8, (n, 2))
    bullet = rng.uniform(-10.0, 60.0, n); active = rng.random(n) < 0.85
    disc = np.zeros((n, 8))
    disc[:, 0] = rng.uniform(0, 60, n); disc[:, 1] = rng.uniform(-1.5, 1.5, n); disc[:, 2] = rng.uniform(-0.5, 0.5, n)
    disc[:, 6] = rng.uniform(0.2, 0.6, n); disc[:, 7] = np.where(rng.random(n) < 0.3, 2.0, 0.0)
    nom_len = np.where(rng.random(n) < 0.15, rng.integers(0, 12, n), T + 1).astype(np.int32)
    return X, bullet, active, disc, nom_len


@pytest.mark.parametrize("mode", ["gatekeeper", "mps"])
def test_kernel_body_vs_oracle_multi_step(mode):
    sc = B.EvadeScene(dt=0.1, backup_horizon=6.0)
    T, n, steps = 40, 24, 12
    X, bullet, active, disc, nom_len = shield_batch(sc, n, seed=7, T=T)
    hs = HostShield(H.hostsim(False), sc, mode, n, T=T, event_offset=0.25, disc=4)
    orc = [S.OracleShield(sc, mode=mode, event_offset=0.25, horizon_discount=4 * sc.dt) for _ in range(n)]
    flags = set()
    for k in range(steps):
        NOMX = np.zeros((n, T + 1, 4)); NOMU = np.zeros((n, T, 2)); MOV = np.zeros((n, 2, 8)); STAT = np.zeros((n, 5))
        plans = []
        for a in range(n):
            nx, nu = S.nominal_rollout(sc, X[a], horizon_time=T * sc.dt)
            L = int(nom_len[a]); nx, nu = nx[:L], nu[: max(L - 1, 0)]
            NOMX[a, :L] = nx; NOMU[a, : max(L - 1, 0)] = nu
            MOV[a, 0] = B.bullet_row(bullet[a], active=bool(active[a])); MOV[a, 1] = disc[a]
            STAT[a] = S.bullet_static_rect(bullet[a], active=bool(active[a]))
            plans.append((nx, nu))
        U, ub = hs.step(X, NOMX, NOMU, MOV, STAT, nom_len)
        for a in range(n):
            u = orc[a].solve(X[a], plans[a][0], plans[a][1], MOV[a], STAT[a])
            assert np.abs(U[a] - u).max() < 1e-12, (k, a)
            assert bool(ub[a]) == orc[a].is_using_backup(), (k, a)
            assert hs.cidx[a] == orc[a].current_time_idx and hs.clen[a] == len(orc[a].committed_u) and hs.nsteps[a] == orc[a].actual_nominal_steps
            assert abs(hs.next_event[a] - orc[a].next_event_time) < 1e-12
            assert np.abs(hs.CU[a, : hs.clen[a]] - orc[a].committed_u).max() < 1e-12
            assert np.abs(hs.CX[a, : hs.clen[a] + 1] - orc[a].committed_x).max() < 1e-12
            flags.add((bool(ub[a]), int(hs.nsteps[a]) > 0))
            X[a] = B.di_step(sc, X[a], U[a])
        bullet = bullet + 3.0 * sc.dt
        disc[:, 0] += disc[:, 2] * sc.dt
    assert len(flags) >= 3          # committed nominal legs and pure-backup commitments both occurred


@pytest.mark.parametrize("mode", ["gatekeeper", "mps"])
def test_other_geometries_and_parameters(mode):
    """other hallway / pocket sizes, radii, limits, gains, margins, event offsets and discounts against the oracle"""
    rng = np.random.default_rng(33)
    for trial in range(3):
        sc = B.EvadeScene(hallway_length=rng.uniform(40, 80), hallway_width=rng.uniform(3, 6), pocket_x=rng.uniform(10, 30),
                          pocket_length=rng.uniform(6, 12), pocket_width=rng.uniform(3, 5), goal_length=rng.uniform(3, 8),
                          radius=rng.uniform(0.3, 0.6), a_max=rng.uniform(1.0, 3.0), v_max=rng.uniform(1.0, 2.5),
                          safety_margin=rng.uniform(0.0, 0.8), use_goal=bool(trial % 2), dt=0.1, backup_horizon=float(rng.choice([3.0, 5.0])))
        sc.Kp, sc.Kd = rng.uniform(1.0, 3.0), rng.uniform(1.0, 3.0)
        T, n, steps = 30, 12, 8
        off, disc = float(rng.choice([0.05, 0.15, 0.3])), int(rng.choice([2, 5, 7]))
        X = np.zeros((n, 4))
        X[:, 0] = rng.uniform(1.0, sc.hallway_length - 1.0, n)
        X[:, 1] = rng.uniform(-sc.half_width + 0.8, sc.half_width - 0.8, n)
        X[:, 2:] = rng.uniform(-0.8, 0.8, (n, 2))
        bullet = rng.uniform(-10.0, sc.hallway_length, n); speed = rng.uniform(1.0, 4.0)
        hs = HostShield(H.hostsim(False), sc, mode, n, T=T, event_offset=off, disc=disc)
        orc = [S.OracleShield(sc, mode=mode, event_offset=off, horizon_discount=disc * sc.dt) for _ in range(n)]
        for k in range(steps):
            NOMX = np.zeros((n, T + 1, 4)); NOMU = np.zeros((n, T, 2)); MOV = np.zeros((n, 1, 8)); STAT = np.zeros((n, 5))
            plans = []
            for a in range(n):
                nx, nu = S.nominal_rollout(sc, X[a], horizon_time=T * sc.dt)
                NOMX[a], NOMU[a] = nx, nu
                MOV[a, 0] = B.bullet_row(bullet[a], bullet_width=2 * sc.half_width, bullet_speed=speed)
                STAT[a] = S.bullet_static_rect(bullet[a], bullet_width=2 * sc.half_width)
                plans.append((nx, nu))
            U, ub = hs.step(X, NOMX, NOMU, MOV, STAT)
            for a in range(n):
                u = orc[a].solve(X[a], plans[a][0], plans[a][1], MOV[a], STAT[a])
                assert np.abs(U[a] - u).max() < 1e-12, (trial, k, a)
                assert bool(ub[a]) == orc[a].is_using_backup()
                assert hs.cidx[a] == orc[a].current_time_idx and hs.clen[a] == len(orc[a].committed_u) and hs.nsteps[a] == orc[a].actual_nominal_steps
                assert abs(hs.next_event[a] - orc[a].next_event_time) < 1e-12
                X[a] = B.di_step(sc, X[a], U[a])
            bullet = bullet + speed * sc.dt


def test_shadow_route_for_the_shields():
    from unittest import mock
    from oracle import refshim
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.startswith(("safe_control.", "matplotlib")) or k == "safe_control"}
    sys.modules.setdefault("matplotlib.collections", mock.MagicMock(name="matplotlib.collections"))
    if not refshim.available():
        pytest.skip("reference checkout not present")
    refshim.install()
    import safe_control_b200.shadow as shadow
    import safe_control_b200.shield as ours
    try:
        sp = shadow.install_shielding()
        from safe_control.shielding.gatekeeper import Gatekeeper
        from safe_control.shielding.mps import MPS
        from safe_control.shielding import Gatekeeper as G2
        assert Gatekeeper is ours.Gatekeeper and MPS is ours.MPS and G2 is ours.Gatekeeper and sp.MPS is ours.MPS
        spec = {"model": "DoubleIntegrator2D", "radius": 0.5, "a_max": 2.0, "v_max": 1.5}
        gk = Gatekeeper(None, spec, dt=0.1, backup_horizon=12.0, nominal_horizon=10.0, event_offset=0.05, safety_margin=0.5)
        assert gk.is_using_backup() and gk.get_status()["committed_length"] == 0 and gk.current_time_idx == 120
        with pytest.raises(NotImplementedError):
            MPS(None, {"model": "DriftingCar"})
        shadow.uninstall_shielding()
        from safe_control.shielding.gatekeeper import Gatekeeper as Ref
        assert Ref is not ours.Gatekeeper and Ref.__module__ == "safe_control.shielding.gatekeeper"
    finally:
        for k in [k for k in sys.modules if k.startswith("safe_control.") or k == "safe_control"]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


def test_abi_struct_mirror():
    from safe_control_b200 import build
    build.build()
    from safe_control_b200._lib import lib
    assert lib().scb_shield_params_sizeof() == C.sizeof(_abi.ScbShieldParams)
