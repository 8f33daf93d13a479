"""Parity tests proper for the Backup-CBF QP path: the CUDA kernel, called through the C ABI
(scb_backupcbf_solve / _host, include/scb.h), against
  * tests/golden/ref_backupcbf.npz -- outputs of the reference's own BackupCBF in the evade scenario (closed loop + probes),
  * the oracle restatement (oracle/backup_cbf.py) on fresh seeded batches,
  * itself across lane-group geometries and batch sizes (size-independent properties at 65 536 agents),
and the drop-in class (position_control/backup_cbf_qp.py surface) driven exactly like examples/evade/test_evade.py."""
import os

import numpy as np
import pytest
import torch

from oracle import backup_cbf as B
from test_backupcbf import GOLD, SETS, scene_of, movers_of, c_params, random_batch, mask_of

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_solve(sc, X, Ur, MOV, lanes=None, fused=False, **want):
    """lanes / fused: A/B switches of the launch dispatch (csrc/scb_api.cu: SCB_BK_LANES, SCB_BK_FUSED), read per call"""
    from safe_control_b200 import BatchedBackupCBF
    old = os.environ.pop("SCB_BK_LANES", None)
    os.environ.pop("SCB_BK_FUSED", None)
    if lanes:
        os.environ["SCB_BK_LANES"] = str(lanes)
    if fused:
        os.environ["SCB_BK_FUSED"] = "1"
    try:
        ctrl = BatchedBackupCBF(c_params(sc))
        out = ctrl.solve(dev(X), dev(Ur), None if MOV is None else dev(MOV), **want)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("SCB_BK_LANES", None)
        os.environ.pop("SCB_BK_FUSED", None)
        if old is not None:
            os.environ["SCB_BK_LANES"] = old
    o = {k: v.cpu().numpy() for k, v in out.items()}
    if "active" in o:
        o["active"] = o["active"].view(np.uint64)
    return o


@pytest.mark.parametrize("lanes,fused", [(None, False), (5, False), (8, False), (32, False), (8, True), (32, True)])
@pytest.mark.parametrize("tag", SETS)
def test_reference_fixtures(tag, lanes, fused):
    gold = np.load(GOLD)
    sc = scene_of(gold, tag)
    o = gpu_solve(sc, gold[tag + "_state"], gold[tag + "_u_ref"], movers_of(gold, tag), lanes=lanes, fused=fused, want_phi=True)
    assert np.abs(o["U"] - gold[tag + "_u"]).max() < 1e-9                      # float tolerance of this path: 1e-9 absolute
    assert np.abs(o["phi"] - gold[tag + "_phi"]).max() < 1e-13
    assert np.abs(o["h_min"] - gold[tag + "_h_min"]).max() < 1e-13
    assert np.array_equal(o["status"], gold[tag + "_qp_status"].astype(np.int32))
    assert np.array_equal(o["intervene"], gold[tag + "_using_backup"].astype(np.int32))


@pytest.mark.parametrize("dt,hor,goal,n", [(0.1, 12.0, True, 96), (0.05, 2.0, False, 128), (0.1, 25.0, True, 24)])
def test_vs_oracle_random(dt, hor, goal, n):
    sc = B.EvadeScene(dt=dt, backup_horizon=hor, use_goal=goal)
    X, Ur, MOV = random_batch(sc, n, seed=int(hor * 7) + 1)
    o = gpu_solve(sc, X, Ur, MOV, want_phi=True, want_rows=True, want_active=True)
    words = o["active"].shape[1]
    n_mask = n_opt = 0
    for a in range(n):
        ref = B.solve(sc, X[a], Ur[a], MOV[a])
        assert np.abs(o["phi"][a] - ref["phi"]).max() < 1e-13
        assert abs(o["h_min"][a] - ref["h_min"]) < 1e-13
        assert np.abs(o["rows"][a][:, :2] - ref["G"]).max() < 1e-9 and np.abs(o["rows"][a][:, 2] - ref["h"]).max() < 1e-9
        assert o["status"][a] == ref["status"], a
        assert np.abs(o["U"][a] - ref["u"]).max() < 1e-9, a
        assert bool(o["intervene"][a]) == ref["intervene"]
        n_opt += ref["status"] == 0
        if ref["status"] == 0 and ref.get("gap", 0.0) > 1e-7:
            assert np.array_equal(o["active"][a], mask_of(ref["active"], words)), a
            n_mask += 1
    assert n_opt >= 3 and n_mask >= 3


def test_full_size_properties_and_geometries():
    """65 536 agents (BASELINE configs[4]'s batch size): every output bit-identical across the two lane-group geometries
    and to the same agents solved in a small batch; outputs inside the input box; fall-back semantics."""
    sc = B.EvadeScene()
    n = 65536
    X, Ur, MOV = random_batch(sc, n, seed=11, k_mov=1)
    a8 = gpu_solve(sc, X, Ur, MOV, lanes=8, want_active=True)
    a32 = gpu_solve(sc, X, Ur, MOV, lanes=32, want_active=True)
    f8 = gpu_solve(sc, X, Ur, MOV, lanes=8, fused=True, want_active=True)      # the one-launch variant
    a5 = gpu_solve(sc, X, Ur, MOV, lanes=5, want_active=True)                  # six agents per warp (the default for this size)
    for k in ("U", "status", "intervene", "h_min", "active"):
        assert np.array_equal(a8[k], a32[k]), k
        assert np.array_equal(a8[k], f8[k]), k
        assert np.array_equal(a8[k], a5[k]), k
    sub = np.arange(0, n, 257)
    small = gpu_solve(sc, X[sub], Ur[sub], MOV[sub], want_active=True)
    for k in ("U", "status", "intervene", "h_min", "active"):
        assert np.array_equal(a8[k][sub], small[k]), k
    st = a8["status"]
    assert set(np.unique(st)) <= {0, 1} and (st == 0).sum() > 1000 and (st == 1).sum() > 1000
    assert np.all(np.abs(a8["U"]) <= sc.a_max * (1 + 1e-12))                   # optimal: |z| <= 1; fall-backs: clipped / policy-clamped
    clipped = np.clip(Ur, -sc.a_max, sc.a_max)
    fb_nom = (st == 1) & (a8["h_min"] > 0.01)
    assert np.array_equal(a8["U"][fb_nom], clipped[fb_nom]) and not a8["intervene"][fb_nom].any()      # (may be empty)
    fb_bak = (st == 1) & (a8["h_min"] <= 0.01)
    assert fb_bak.any() and a8["intervene"][fb_bak].all()
    pol = np.array([B.backup_control(sc, X[a]) for a in np.where(fb_bak)[0][:64]])
    assert np.abs(a8["U"][np.where(fb_bak)[0][:64]] - pol).max() < 1e-12


def test_edge_cases_and_errors():
    from safe_control_b200 import BatchedBackupCBF
    from safe_control_b200._lib import ScbError
    sc = B.EvadeScene()
    X, Ur, MOV = random_batch(sc, 8, seed=2)
    o = gpu_solve(sc, X[:0], Ur[:0], MOV[:0])                                  # empty batch
    assert o["U"].shape == (0, 2)
    shared = gpu_solve(sc, X, Ur, MOV[0])                                      # one obstacle list shared by all agents
    per = gpu_solve(sc, X, Ur, np.tile(MOV[0][None], (8, 1, 1)))
    assert np.array_equal(shared["U"], per["U"])
    none = gpu_solve(sc, X, Ur, None)                                          # no moving obstacles at all
    off = MOV.copy(); off[:, :, 7] = 0.0                                       # all inactive == none
    assert np.array_equal(none["U"], gpu_solve(sc, X, Ur, off)["U"])
    for a in range(8):
        assert np.abs(none["U"][a] - B.solve(sc, X[a], Ur[a], None)["u"]).max() < 1e-9
    big = c_params(B.EvadeScene(dt=0.01, backup_horizon=12.0))                 # 1200 backup steps: refused, not truncated
    with pytest.raises(ScbError):
        BatchedBackupCBF(big).solve(dev(X), dev(Ur))
    with pytest.raises(ValueError):
        BatchedBackupCBF(c_params(sc)).solve(dev(X), dev(Ur[:, :1]))
    # an obstacle stride shorter than one agent's list is refused by the C entry point itself
    import ctypes as C
    from safe_control_b200._lib import lib
    p = c_params(sc)
    t = [dev(v) for v in (X, Ur, MOV)]
    U = torch.empty((8, 2), dtype=torch.float64, device="cuda"); st = torch.empty((8,), dtype=torch.int32, device="cuda")
    vp = lambda x: C.c_void_p(x.data_ptr())
    rc = lib().scb_backupcbf_solve(C.byref(p), 8, 2, vp(t[0]), vp(t[1]), vp(t[2]), 8, vp(U), vp(st), None, None, None, None, None, None)
    assert rc == -1


def test_host_path_matches_device_path():
    from safe_control_b200 import HostContext
    from safe_control_b200.backup import host_solve
    sc = B.EvadeScene(dt=0.1, backup_horizon=6.0)
    X, Ur, MOV = random_batch(sc, 300, seed=4)
    d = gpu_solve(sc, X, Ur, MOV, want_phi=True, want_rows=True, want_active=True)
    h = host_solve(HostContext(0), c_params(sc), X, Ur, MOV, want_phi=True, want_rows=True, want_active=True)
    for k in d:
        assert np.array_equal(d[k], h[k]), k


def test_dropin_evade_closed_loop_matches_reference_run():
    """examples/evade/test_evade.py:417-456 with the drop-in class: same inputs per step as the recorded reference run
    (its states, nominal inputs and bullet positions) -> same input, same backup flag, same h_min; then the loop closed
    on OUR outputs stays on the reference's trajectory."""
    from safe_control_b200.position_control.backup_cbf_qp import BackupCBF
    gold = np.load(GOLD)

    class Env:                      # the attributes of envs/evade_env.py::EvadeEnv the path reads
        hallway_length, half_width = 60.0, 2.0
        pocket_x_min, pocket_x_max, pocket_y_max = 25.0, 35.0, 6.0
        bullet_x, bullet_active = -10.0, True

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
    sh = BackupCBF(robot=None, robot_spec=spec, dt=0.1, backup_horizon=12.0)
    sh.set_backup_controller(Policy()); sh.set_environment(env)

    def get_obstacles(t=0.0):                    # test_evade.py:373-385
        st = env.get_bullet_state()
        if not st["active"]:
            return None
        fut = st.copy(); fut["x"] = st["x"] + st["vx"] * t
        return fut

    sh.set_moving_obstacles(get_obstacles)
    n = gold["loop_u"].shape[0]
    # (1) open loop on the recorded inputs
    for k in range(0, n, 3):
        env.bullet_x = float(gold["loop_bullet_x"][k])
        sh.set_nominal_trajectory(None, np.tile(gold["loop_u_ref"][k][None], (3, 1)))
        u = sh.solve_control_problem(gold["loop_state"][k].reshape(-1, 1))
        assert u.shape == (2, 1) and np.abs(u.flatten() - gold["loop_u"][k]).max() < 1e-9, k
        assert sh.is_using_backup() == bool(gold["loop_using_backup"][k])
        assert abs(sh.get_status()["h_min"] - gold["loop_h_min"][k]) < 1e-13
        assert np.abs(sh.latest_backup_trajectory - gold["loop_phi"][k]).max() < 1e-13
    # (2) closed loop on our own outputs for the first 150 steps (the bullet catches up, the robot hides in the pocket)
    sc = B.EvadeScene()
    state = gold["loop_state"][0].copy()
    env.bullet_x = float(gold["loop_bullet_x"][0])
    for k in range(150):
        assert np.abs(state - gold["loop_state"][k]).max() < 1e-6, k
        sh.set_nominal_trajectory(None, np.tile(B.nominal_control(sc, state)[None], (3, 1)))
        u = sh.solve_control_problem(state.reshape(-1, 1)).flatten()
        state = B.di_step(sc, state, u)                                        # dynamics.step + the example's clamp (:452-456, a no-op after step's own)
        env.bullet_x = float(gold["loop_bullet_x"][k + 1])
