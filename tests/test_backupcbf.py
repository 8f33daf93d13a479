"""CPU-only checks of the Backup-CBF QP path (position_control/backup_cbf_qp.py:563-794):

* the oracle restatement (oracle/backup_cbf.py) against tests/golden/ref_backupcbf.npz -- outputs of the REFERENCE'S OWN
  BackupCBF / EvadeBackupController / DoubleIntegrator2D / EvadeEnv driven through oracle/refshim by
  tests/golden/gen_backupcbf_from_reference.py (closed loop of the evade scenario + seeded probes);
* its 2-variable exact QP against the general enumeration solver (oracle/qp_exact.py);
* the kernel bodies (csrc/scb_backup.cuh, host build, with and without FMA contraction) against the same fixture and
  against the oracle on fresh seeded batches;
* the host logic of the drop-in class (scene extraction, obstacle sampling) without a device.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import backup_cbf as B
from safe_control_b200 import _abi
import hostsim_util as H

GOLD = os.path.join(ROOT, "tests", "golden", "ref_backupcbf.npz")
SETS = ("loop", "probe", "short")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def scene_of(gold, tag):
    if tag == "loop":
        return B.EvadeScene()
    dt, hor, ug = gold[tag + "_cfg"]
    return B.EvadeScene(dt=dt, backup_horizon=hor, use_goal=bool(ug))


def movers_of(gold, tag):
    return np.stack([B.bullet_row(bx, active=bool(a))[None] for bx, a in zip(gold[tag + "_bullet_x"], gold[tag + "_bullet_active"])])


def c_params(sc):
    p = _abi.ScbBackupParams()
    for k, v in sc.as_vector().items():
        setattr(p, "n_backup" if k == "N" else k, v)
    return p


def hs_solve(lib, sc, X, Ur, MOV):
    p = c_params(sc)
    N, nb = X.shape[0], sc.N
    K = 0 if MOV is None else MOV.shape[1]
    U = np.zeros((N, 2)); st = np.zeros(N, np.int32); iv = np.zeros(N, np.int32); hm = np.zeros(N)
    phi = np.zeros((N, nb, 4)); rows = np.zeros((N, nb, 3)); act = np.zeros((N, (nb + 4 + 63) // 64), np.uint64)
    f = lib.hostsim_backupcbf_solve
    f.restype = C.c_int
    f.argtypes = [C.POINTER(_abi.ScbBackupParams), C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_long] + [C.c_void_p] * 7
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    X = np.ascontiguousarray(X); Ur = np.ascontiguousarray(Ur)
    MOV = None if MOV is None else np.ascontiguousarray(MOV)
    assert f(C.byref(p), N, K, ptr(X), ptr(Ur), ptr(MOV), K * 8, ptr(U), ptr(st), ptr(iv), ptr(hm), ptr(phi), ptr(rows), ptr(act)) == 0
    return dict(U=U, status=st, intervene=iv, h_min=hm, phi=phi, rows=rows, active=act)


def mask_of(rows_active, words):
    m = np.zeros(words, np.uint64)
    for r in rows_active:
        m[r >> 6] |= np.uint64(1) << np.uint64(r & 63)
    return m


# ---------------------------------------------------------------------------------------------------- oracle, pinned
@pytest.mark.parametrize("tag", SETS)
def test_oracle_reproduces_the_reference(gold, tag):
    sc = scene_of(gold, tag)
    mov = movers_of(gold, tag)
    n = gold[tag + "_u"].shape[0]
    seen = set()
    for k in (range(0, n, 9) if tag == "loop" else range(0, n, 2)):
        o = B.solve(sc, gold[tag + "_state"][k], gold[tag + "_u_ref"][k], mov[k])
        assert np.abs(o["u"] - gold[tag + "_u"][k]).max() < 1e-12, (tag, k)
        assert np.abs(o["phi"] - gold[tag + "_phi"][k]).max() < 1e-13
        assert abs(o["h_min"] - gold[tag + "_h_min"][k]) < 1e-13
        assert o["intervene"] == bool(gold[tag + "_using_backup"][k])
        m = int(gold[tag + "_qp_m"][k]); idx = np.where(o["keep"])[0]
        assert m == idx.size + 4                                   # the reference's ||lhs|| > 1e-6 filter, then 4 box rows
        assert np.abs(o["G"][idx] * sc.a_max - gold[tag + "_qp_A"][k][: m - 4]).max() < 1e-12
        assert np.abs(o["h"][idx] - gold[tag + "_qp_b"][k][: m - 4]).max() < 1e-12
        assert o["status"] == int(gold[tag + "_qp_status"][k])
        seen.add((o["status"], o["intervene"]))
    assert len(seen) >= (3 if tag != "short" else 2)   # optimal without / with intervention, infeasible fall-backs


def test_qp2_exact_agrees_with_the_general_solver():
    from oracle.qp_exact import solve_qp_exact
    rng = np.random.default_rng(5)
    n_inf = 0
    for _ in range(60):
        m = int(rng.integers(1, 9))
        A = rng.normal(size=(m, 2)); b = rng.normal(size=m) - 0.6
        A = np.vstack([A, np.eye(2), -np.eye(2)]); b = np.concatenate([b, -np.ones(4)])
        Q = rng.uniform(0.5, 2.0, 2); c = rng.normal(size=2)
        z, W, gap = B.qp2_exact(Q, c, A, b)
        res = solve_qp_exact(2.0 * np.diag(Q ** 2), -2.0 * Q ** 2 * c, -A, -b)
        if res["status"] != 0:
            assert z is None
            n_inf += 1
            continue
        assert np.abs(z - res["x"]).max() < 1e-9
    assert 0 < n_inf < 40


# ---------------------------------------------------------------------------------------------------- kernel bodies on the CPU
@pytest.mark.parametrize("fma", [False, True])
@pytest.mark.parametrize("tag", SETS)
def test_kernel_body_reproduces_the_reference(gold, tag, fma):
    sc = scene_of(gold, tag)
    o = hs_solve(H.hostsim(fma), sc, gold[tag + "_state"], gold[tag + "_u_ref"], movers_of(gold, tag))
    assert np.abs(o["U"] - gold[tag + "_u"]).max() < 1e-9
    assert np.abs(o["phi"] - gold[tag + "_phi"]).max() < 1e-13
    assert np.abs(o["h_min"] - gold[tag + "_h_min"]).max() < 1e-13
    assert np.array_equal(o["status"], gold[tag + "_qp_status"].astype(np.int32))
    assert np.array_equal(o["intervene"], gold[tag + "_using_backup"].astype(np.int32))


def random_batch(sc, n, seed, k_mov=2):
    rng = np.random.default_rng(seed)
    X = np.zeros((n, 4)); MOV = np.zeros((n, k_mov, 8))
    X[:, 0] = rng.uniform(0.3, 59.7, n)
    X[:, 1] = np.where(rng.random(n) < 0.3, rng.uniform(0.0, 5.4, n), rng.uniform(-1.6, 1.6, n))
    X[:, 2:] = rng.uniform(-1.0, 1.0, (n, 2)) * rng.uniform(0.0, 1.5, (n, 1))
    Ur = rng.uniform(-2.5, 2.5, (n, 2))
    for a in range(n):
        MOV[a, 0] = B.bullet_row(rng.uniform(-10.0, 62.0), active=bool(rng.random() < 0.9))
        if k_mov > 1:      # a slow disc somewhere in the hallway (the circle branch, backup_cbf_qp.py:436-440)
            MOV[a, 1] = [rng.uniform(0, 60), rng.uniform(-1.5, 1.5), rng.uniform(-0.5, 0.5), rng.uniform(-0.2, 0.2), 0, 0,
                         rng.uniform(0.2, 0.8), 2.0 if rng.random() < 0.5 else 0.0]
    return X, Ur, MOV


@pytest.mark.parametrize("fma", [False, True])
@pytest.mark.parametrize("dt,hor,goal", [(0.1, 12.0, True), (0.05, 2.0, False), (0.1, 6.0, True)])
def test_kernel_body_vs_oracle_random(dt, hor, goal, fma):
    sc = B.EvadeScene(dt=dt, backup_horizon=hor, use_goal=goal)
    X, Ur, MOV = random_batch(sc, 64, seed=int(hor * 10))
    o = hs_solve(H.hostsim(fma), sc, X, Ur, MOV)
    words = o["active"].shape[1]
    n_mask = 0
    for a in range(X.shape[0]):
        ref = B.solve(sc, X[a], Ur[a], MOV[a])
        assert np.abs(o["phi"][a] - ref["phi"]).max() < 1e-13
        assert abs(o["h_min"][a] - ref["h_min"]) < 1e-13
        assert np.abs(o["rows"][a][:, :2] - ref["G"]).max() < 1e-9 and np.abs(o["rows"][a][:, 2] - ref["h"]).max() < 1e-9
        assert o["status"][a] == ref["status"], a
        assert np.abs(o["U"][a] - ref["u"]).max() < 1e-9, a
        assert bool(o["intervene"][a]) == ref["intervene"]
        if ref["status"] == 0 and ref.get("gap", 0.0) > 1e-7:
            assert np.array_equal(o["active"][a], mask_of(ref["active"], words)), a
            n_mask += 1
    assert n_mask >= 4


def test_other_geometries_and_parameters():
    """nothing of the example's scene is baked into the kernel: other hallway / pocket sizes, radii, limits, gains, class-K
    gains and margins (all of them fields of scb_backup_params) against the oracle"""
    lib = H.hostsim(False)
    rng = np.random.default_rng(21)
    n_opt = 0
    for trial in range(6):
        sc = B.EvadeScene(hallway_length=rng.uniform(40, 80), hallway_width=rng.uniform(3, 6), pocket_x=rng.uniform(10, 30),
                          pocket_length=rng.uniform(6, 12), pocket_width=rng.uniform(3, 5), goal_length=rng.uniform(3, 8),
                          radius=rng.uniform(0.3, 0.7), a_max=rng.uniform(1.0, 3.0), v_max=rng.uniform(1.0, 2.5),
                          safety_margin=rng.uniform(0.0, 0.8), use_goal=bool(trial % 2), dt=0.1, backup_horizon=rng.choice([3.0, 5.0, 8.0]))
        sc.Kp, sc.Kd = rng.uniform(1.0, 3.0), rng.uniform(1.0, 3.0)
        sc.alpha, sc.alpha_terminal = rng.uniform(0.5, 2.0), rng.uniform(1.0, 3.0)
        n = 16
        X = np.zeros((n, 4))
        X[:, 0] = rng.uniform(0.8, sc.hallway_length - 0.8, n)
        X[:, 1] = rng.uniform(-sc.half_width + 0.8, sc.half_width - 0.8, n)
        inp = rng.random(n) < 0.3
        X[inp, 0] = rng.uniform(sc.pocket_x_min + 0.8, sc.pocket_x_max - 0.8, inp.sum())
        X[inp, 1] = rng.uniform(0.0, sc.pocket_y_max - 0.9, inp.sum())
        X[:, 2:] = rng.uniform(-1, 1, (n, 2)) * rng.uniform(0, sc.v_max, (n, 1))
        Ur = rng.uniform(-1.2 * sc.a_max, 1.2 * sc.a_max, (n, 2))
        MOV = np.zeros((n, 1, 8))
        for a in range(n):
            MOV[a, 0] = B.bullet_row(rng.uniform(-10, sc.hallway_length), bullet_length=rng.uniform(2, 4), bullet_width=2 * sc.half_width,
                                     bullet_speed=rng.uniform(1, 4))
        o = hs_solve(lib, sc, X, Ur, MOV)
        for a in range(n):
            ref = B.solve(sc, X[a], Ur[a], MOV[a])
            assert np.abs(o["phi"][a] - ref["phi"]).max() < 1e-12, (trial, a)
            assert abs(o["h_min"][a] - ref["h_min"]) < 1e-12
            assert o["status"][a] == ref["status"] and bool(o["intervene"][a]) == ref["intervene"], (trial, a)
            assert np.abs(o["U"][a] - ref["u"]).max() < 1e-8, (trial, a)
            n_opt += ref["status"] == 0
    assert n_opt >= 10


def test_no_obstacles_and_tiny_horizons():
    lib = H.hostsim(False)
    for hor in (0.1, 0.2, 0.5):
        sc = B.EvadeScene(dt=0.1, backup_horizon=hor)
        X, Ur, _ = random_batch(sc, 12, seed=3)
        o = hs_solve(lib, sc, X, Ur, None)
        for a in range(12):
            ref = B.solve(sc, X[a], Ur[a], None)
            assert o["status"][a] == ref["status"] and np.abs(o["U"][a] - ref["u"]).max() < 1e-9


# ---------------------------------------------------------------------------------------------------- host logic of the drop-in class
def test_dropin_scene_extraction_and_obstacle_rows():
    from safe_control_b200.position_control.backup_cbf_qp import BackupCBF, _obstacle_rows

    class Env:
        hallway_length, half_width = 60.0, 2.0
        pocket_x_min, pocket_x_max, pocket_y_max = 25.0, 35.0, 6.0

        def get_pocket_bounds(self):
            return dict(x_min=25.0, x_max=35.0, y_min=2.0, y_max=6.0)

    class Policy:
        safe_center, safe_bounds = np.array([30.0, 4.0]), dict(x_min=25.0, x_max=35.0, y_min=2.0, y_max=6.0)
        goal_bounds = dict(x_min=55.0, x_max=60.0, y_min=-2.0, y_max=2.0)
        Kp = Kd = 2.0

    spec = {"model": "DoubleIntegrator2D", "radius": 0.5, "a_max": 2.0, "v_max": 1.5, "safety_margin": 0.5}
    sh = BackupCBF(None, spec, dt=0.1, backup_horizon=12.0)
    sh.set_environment(Env()); sh.set_backup_controller(Policy())
    p = sh._scene()
    want = c_params(B.EvadeScene())
    for name, _ in _abi.ScbBackupParams._fields_:
        assert getattr(p, name) == getattr(want, name), name
    with pytest.raises(NotImplementedError):
        BackupCBF(None, {"model": "DriftingCar"})
    sh2 = BackupCBF(None, spec)
    with pytest.raises(NotImplementedError):
        sh2._scene()

    bullet = dict(x=3.5, y=0.0, vx=3.0, vy=0.0, length=4.0, width=4.0, active=True)
    rows = _obstacle_rows(lambda t=0.0: dict(bullet, x=bullet["x"] + 3.0 * t))
    assert rows.shape == (1, 8) and list(rows[0]) == [3.5, 0.0, 3.0, 0.0, 4.0, 4.0, 0.0, 1.0]
    assert _obstacle_rows(lambda t=0.0: None).shape == (0, 8) and _obstacle_rows(None) is None
    rows = _obstacle_rows([dict(x=1.0, y=2.0, radius=0.7), None, dict(x=5.0, y=0.0, radius=1.0, active=False)])
    assert rows.shape == (1, 8) and list(rows[0]) == [1.0, 2.0, 0.0, 0.0, 0.0, 0.0, 0.7, 2.0]
    rows = _obstacle_rows(lambda t: dict(x=1.0 + 0.5 * t, y=0.0, radius=1.0))        # no velocity fields: sampled
    assert list(rows[0][:4]) == [1.0, 0.0, 0.5, 0.0]


def test_abi_struct_mirror():
    from safe_control_b200 import build
    build.build()
    from safe_control_b200._lib import lib
    assert lib().scb_backup_params_sizeof() == C.sizeof(_abi.ScbBackupParams)
    p = _abi.ScbBackupParams()
    lib().scb_backup_params_default(C.byref(p))
    want = c_params(B.EvadeScene())
    for name, _ in _abi.ScbBackupParams._fields_:
        assert getattr(p, name) == getattr(want, name), name
    assert lib().scb_backup_active_words(120) == 2 and lib().scb_backup_active_words(60) == 1 and lib().scb_backup_active_words(0) < 0
