"""Parity tests proper for the gatekeeper / MPS path: the CUDA kernel behind scb_shield_step against the
recorded runs of the reference's own classes (tests/golden/ref_shield.npz), against the oracle on seeded multi-step
batches, across launch geometries at 65 536 agents, and the drop-in classes driven like examples/evade/test_evade.py."""
import os

import numpy as np
import pytest
import torch

from oracle import backup_cbf as B, shielding as S
from test_backupcbf import c_params
from test_shield import GOLD, RUNS, replay, shield_batch

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


class Lanes:
    """SCB_SHIELD_LANES = 1 / 8 / 32 and SCB_SHIELD_TWO_PHASE = 0 / 1 for the calls inside the block (A/B switches of the
    dispatch in csrc/scb_api.cu, read per call)"""

    def __init__(self, lanes, two_phase=None):
        self.lanes, self.two = lanes, two_phase

    def __enter__(self):
        self.old = (os.environ.pop("SCB_SHIELD_LANES", None), os.environ.pop("SCB_SHIELD_TWO_PHASE", None))
        if self.lanes:
            os.environ["SCB_SHIELD_LANES"] = str(self.lanes)
        if self.two is not None:
            os.environ["SCB_SHIELD_TWO_PHASE"] = str(int(self.two))

    def __exit__(self, *a):
        os.environ.pop("SCB_SHIELD_LANES", None); os.environ.pop("SCB_SHIELD_TWO_PHASE", None)
        if self.old[0] is not None:
            os.environ["SCB_SHIELD_LANES"] = self.old[0]
        if self.old[1] is not None:
            os.environ["SCB_SHIELD_TWO_PHASE"] = self.old[1]


def make_shield(sc, mode, n, T, event_offset=0.05, disc_steps=5, keep_states=True):
    from safe_control_b200 import BatchedShield
    return BatchedShield(n, mode, c_params(sc), event_offset, disc_steps * sc.dt, T, device="cuda", keep_states=keep_states)


@pytest.mark.parametrize("lanes", [None, 8, 1, "two-phase"])
@pytest.mark.parametrize("algo,tag", RUNS)
def test_reference_runs(algo, tag, lanes):
    if lanes is not None and tag == "scenario":
        pytest.skip("geometry variants replay the short runs")
    two = None
    if lanes == "two-phase":                       # candidate 0 with a thread per agent, the rest with a lane group (forced: N = 1)
        lanes, two = None, True
    gold = np.load(GOLD)

    def make(sc):
        sh = make_shield(sc, algo, 1, 100)

        def step(x, nx, nu, mov, stat):
            with Lanes(lanes, two):
                o = sh.step(dev(x[None]), dev(nx[None]), dev(nu[None]), dev(mov[None]), dev(stat[None]))
            return o["U"].cpu().numpy()[0], bool(o["using_backup"].cpu()[0])

        def probe():
            clen = int(sh.clen.cpu()[0])
            return dict(idx=int(sh.cidx.cpu()[0]), clen=clen, horizon=float(sh.committed_horizon().cpu()[0]),
                        next_event=float(sh.next_event.cpu()[0]), cu=sh.CU[0, :clen].cpu().numpy())
        return step, probe
    replay(gold, algo, tag, make)


@pytest.mark.parametrize("mode", ["gatekeeper", "mps"])
def test_vs_oracle_multi_step(mode):
    sc = B.EvadeScene(dt=0.1, backup_horizon=6.0)
    T, n, steps = 40, 48, 10
    X, bullet, active, disc, nom_len = shield_batch(sc, n, seed=17, T=T)
    sh = make_shield(sc, mode, n, T, event_offset=0.25, disc_steps=4)
    orc = [S.OracleShield(sc, mode=mode, event_offset=0.25, horizon_discount=4 * sc.dt) for _ in range(n)]
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
        o = sh.step(dev(X), dev(NOMX), dev(NOMU), dev(MOV), dev(STAT), dev(nom_len))
        U, ub = o["U"].cpu().numpy(), o["using_backup"].cpu().numpy()
        cidx, clen, ns, ne = (t.cpu().numpy() for t in (sh.cidx, sh.clen, sh.nsteps, sh.next_event))
        CU, CX = sh.CU.cpu().numpy(), sh.CX.cpu().numpy()
        for a in range(n):
            u = orc[a].solve(X[a], plans[a][0], plans[a][1], MOV[a], STAT[a])
            assert np.abs(U[a] - u).max() < 1e-12, (k, a)
            assert bool(ub[a]) == orc[a].is_using_backup(), (k, a)
            assert cidx[a] == orc[a].current_time_idx and clen[a] == len(orc[a].committed_u) and ns[a] == orc[a].actual_nominal_steps
            assert abs(ne[a] - orc[a].next_event_time) < 1e-12
            assert np.abs(CU[a, : clen[a]] - orc[a].committed_u).max() < 1e-12
            assert np.abs(CX[a, : clen[a] + 1] - orc[a].committed_x).max() < 1e-12
            X[a] = B.di_step(sc, X[a], U[a])
        bullet = bullet + 3.0 * sc.dt
        disc[:, 0] += disc[:, 2] * sc.dt


def full_batch(n, T, seed):
    """n agents with their nominal plans (constant nominal acceleration towards v_max along the hallway: cheap to build)"""
    from safe_control_b200 import scenes
    sc = B.EvadeScene()
    X, _, MOV = scenes.make_evade_batch(n, seed=seed)
    NOMX = np.zeros((n, T + 1, 4)); NOMU = np.zeros((n, T, 2))
    s = X.copy(); NOMX[:, 0] = s
    for k in range(T):                                   # vectorised EvadeNominalController + DoubleIntegrator2D.step
        ax = 2.0 * (sc.v_max - s[:, 2]); ay = 2.0 * (0.0 - s[:, 1]) + 2.0 * (0.0 - s[:, 3])
        am = np.sqrt(ax ** 2 + ay ** 2); f = np.where(am > sc.a_max, sc.a_max / np.maximum(am, 1e-300), 1.0)
        ax, ay = ax * f, ay * f
        n2 = np.stack([s[:, 0] + s[:, 2] * sc.dt, s[:, 1] + s[:, 3] * sc.dt, s[:, 2] + ax * sc.dt, s[:, 3] + ay * sc.dt], axis=1)
        vm = np.sqrt(n2[:, 2] ** 2 + n2[:, 3] ** 2); g = np.where(vm > sc.v_max, sc.v_max / np.maximum(vm, 1e-300), 1.0)
        n2[:, 2] *= g; n2[:, 3] *= g
        NOMU[:, k, 0] = ax; NOMU[:, k, 1] = ay; NOMX[:, k + 1] = n2; s = n2
    STAT = np.zeros((n, 5))
    bx = MOV[:, 0, 0] - 0.5                              # bullet_row stores x + L / 6
    STAT[:, 0] = bx - 1.5; STAT[:, 1] = bx + 1.5 + 1.0; STAT[:, 2] = -2.0; STAT[:, 3] = 2.0; STAT[:, 4] = (MOV[:, 0, 7] != 0)
    return sc, X, NOMX, NOMU, MOV, STAT


@pytest.mark.parametrize("mode", ["gatekeeper", "mps"])
def test_full_size_geometries_and_properties(mode):
    """65 536 agents, two control steps: identical state and outputs across the launch geometries and for the same agents
    in a small batch; committed trajectories are what the scalar state says they are."""
    n, T = 65536, 100
    sc, X, NOMX, NOMU, MOV, STAT = full_batch(n, T, seed=5)
    d = [dev(v) for v in (X, NOMX, NOMU, MOV, STAT)]
    res = {}
    # (lanes, two-launch search): the default for this size is the two-launch search with 8 lanes in its second launch
    for cfg in (((32, True), (8, True), (8, False), (32, False)) if mode == "gatekeeper" else ((1, None), (32, None))):
        sh = make_shield(sc, mode, n, T, keep_states=False)
        with Lanes(*cfg):
            o1 = sh.step(*d); o2 = sh.step(*d)
        torch.cuda.synchronize()
        res[cfg] = [t.cpu().numpy() for t in (o1["U"], o2["U"], o2["using_backup"], sh.cidx, sh.clen, sh.nsteps, sh.next_event, sh.CU)]
        res[cfg][-1][np.arange(res[cfg][-1].shape[1])[None, :] >= res[cfg][4][:, None]] = 0.0      # rows beyond clen are stale
    a = list(res.values())[0]
    for b in list(res.values())[1:]:
        for u, v in zip(a, b):
            assert np.array_equal(u, v)
    U1, U2, ub, cidx, clen, ns, ne, CU = a
    sub = np.arange(0, n, 331)
    shs = make_shield(sc, mode, sub.size, T, keep_states=False)
    ds = [dev(v[sub]) for v in (X, NOMX, NOMU, MOV, STAT)]
    shs.step(*ds); o = shs.step(*ds)
    assert np.array_equal(o["U"].cpu().numpy(), U2[sub]) and np.array_equal(shs.clen.cpu().numpy(), clen[sub])
    # properties of the reference's bookkeeping
    assert np.all(clen == ns + sc.N) and np.all(ns >= 0) and np.all(ns <= (T if mode == "gatekeeper" else 1))
    assert np.all((cidx == 1) | (cidx == 2))                 # committed this step (idx 0 -> 1) or kept the first commitment (1 -> 2)
    fresh = cidx == 1
    assert fresh.sum() > 1000 and (~fresh).sum() > 100
    assert np.array_equal(U2[fresh], CU[fresh, 0])           # a fresh commitment plays its first input ...
    nom = fresh & (ns > 0)
    assert nom.sum() > 100 and np.array_equal(U2[nom], NOMU[nom, 0])        # ... which is the nominal one when a nominal leg was valid
    assert np.all(np.abs(CU[np.arange(n), np.maximum(clen - 1, 0)]) <= sc.a_max * (1 + 1e-12))


@pytest.mark.parametrize("algo", ["gatekeeper", "mps"])
def test_dropin_classes_closed_loop(algo):
    """examples/evade/test_evade.py:417-470 with the drop-in class, loop closed on OUR outputs: the recorded reference run
    is followed state by state (and the gatekeeper run reaches the goal like the reference's)."""
    from safe_control_b200.shield import Gatekeeper, MPS
    gold = np.load(GOLD)
    g = lambda k: gold[f"{algo}_scenario_{k}"]

    class Env:                      # the attributes of envs/evade_env.py::EvadeEnv the path reads
        hallway_length, half_width = 60.0, 2.0
        pocket_x_min, pocket_x_max, pocket_y_max = 25.0, 35.0, 6.0
        bullet_x, bullet_y, bullet_length, bullet_width, bullet_active = -10.0, 0.0, 3.0, 4.0, True

        def get_pocket_bounds(self):
            return dict(x_min=25.0, x_max=35.0, y_min=2.0, y_max=6.0)

        def get_bullet_state(self):              # evade_env.py:386-406
            return dict(x=self.bullet_x + 3.0 / 6, y=0.0, vx=3.0, vy=0.0, length=3.0 * (1 + 1 / 3), width=4.0, active=self.bullet_active)

    class Policy:                   # EvadeBackupController's attributes (backup_controller.py:431-454)
        safe_center, safe_bounds = np.array([30.0, 4.0]), dict(x_min=25.0, x_max=35.0, y_min=2.0, y_max=6.0)
        goal_bounds = dict(x_min=55.0, x_max=60.0, y_min=-2.0, y_max=2.0)
        Kp = Kd = 2.0

    env = Env()
    spec = {"model": "DoubleIntegrator2D", "radius": 0.5, "a_max": 2.0, "v_max": 1.5, "safety_margin": 0.5}
    if algo == "mps":
        sh = MPS(robot=None, robot_spec=spec, dt=0.1, backup_horizon=12.0, event_offset=0.05, safety_margin=0.5)
    else:
        sh = Gatekeeper(robot=None, robot_spec=spec, dt=0.1, backup_horizon=12.0, nominal_horizon=10.0, event_offset=0.05, safety_margin=0.5)
    sh.set_backup_controller(Policy()); sh.set_environment(env)

    def get_obstacles(t=0.0):                    # test_evade.py:373-385
        st = env.get_bullet_state()
        if not st["active"]:
            return None
        fut = st.copy(); fut["x"] = st["x"] + st["vx"] * t
        return fut

    sh.set_moving_obstacles(get_obstacles)
    sc = B.EvadeScene()
    state = g("state")[0].copy()
    n = min(g("u").shape[0], 330 if algo == "gatekeeper" else 200)
    for k in range(n):
        assert np.abs(state - g("state")[k]).max() < 1e-9, k
        env.bullet_x = float(g("bullet_x")[k]); env.bullet_active = bool(g("bullet_active")[k])
        nx, nu = S.nominal_rollout(sc, state)
        sh.set_nominal_trajectory(nx, nu)
        u = sh.solve_control_problem(state.reshape(-1, 1))
        assert u.shape == (2, 1) and np.abs(u.flatten() - g("u")[k]).max() < 1e-12
        assert sh.is_using_backup() == bool(g("using_backup")[k])
        st = sh.get_status()
        assert st["current_time_idx"] == g("idx")[k] and st["committed_length"] == g("clen")[k]
        assert abs(st["committed_horizon"] - g("horizon")[k]) < 1e-9
        state = B.di_step(sc, state, u.flatten())
    if algo == "gatekeeper":
        assert n == g("u").shape[0] and 55.0 <= state[0] <= 60.0 and abs(state[1]) <= 2.0        # goal zone (evade_env.py:487-500)


def test_dropin_gatekeeper_with_a_nominal_controller():
    """set_nominal_controller(callable) (gatekeeper.py:159-166, 235-269) instead of an external trajectory: same answers."""
    from safe_control_b200.shield import Gatekeeper
    sc = B.EvadeScene()

    class Env:
        hallway_length, half_width = 60.0, 2.0
        pocket_x_min, pocket_x_max, pocket_y_max = 25.0, 35.0, 6.0
        bullet_x, bullet_y, bullet_length, bullet_width, bullet_active = 8.0, 0.0, 3.0, 4.0, True

        def get_pocket_bounds(self):
            return dict(x_min=25.0, x_max=35.0, y_min=2.0, y_max=6.0)

    class Policy:
        safe_center, safe_bounds = np.array([30.0, 4.0]), dict(x_min=25.0, x_max=35.0, y_min=2.0, y_max=6.0)
        goal_bounds = dict(x_min=55.0, x_max=60.0, y_min=-2.0, y_max=2.0)
        Kp = Kd = 2.0

    class Robot:                    # DoubleIntegrator2D.step (robots/double_integrator2D.py:79-107)
        def step(self, X, U):
            return B.di_step(sc, np.asarray(X, float).reshape(-1), np.asarray(U, float).reshape(-1)).reshape(-1, 1)

    spec = {"model": "DoubleIntegrator2D", "radius": 0.5, "a_max": 2.0, "v_max": 1.5}
    env = Env()
    mov = lambda t=0.0: dict(x=env.bullet_x + 0.5 + 3.0 * t, y=0.0, vx=3.0, vy=0.0, length=4.0, width=4.0, active=True)
    shields = []
    for use_controller in (False, True):
        sh = Gatekeeper(Robot(), spec, dt=0.1, backup_horizon=12.0, nominal_horizon=10.0, event_offset=0.05, safety_margin=0.5)
        sh.set_backup_controller(Policy()); sh.set_environment(env); sh.set_moving_obstacles(mov)
        if use_controller:
            sh.set_nominal_controller(lambda s: B.nominal_control(sc, np.asarray(s).reshape(-1)).reshape(-1, 1))
        shields.append(sh)
    state = np.array([20.0, 0.0, 0.5, 0.0])
    for k in range(25):
        nx, nu = S.nominal_rollout(sc, state)
        shields[0].set_nominal_trajectory(nx, nu)
        u0 = shields[0].solve_control_problem(state.reshape(-1, 1))
        u1 = shields[1].solve_control_problem(state.reshape(-1, 1))
        assert np.array_equal(u0, u1), k
        assert shields[0].get_status() == shields[1].get_status()
        state = B.di_step(sc, state, u0.flatten())
        env.bullet_x += 0.3
