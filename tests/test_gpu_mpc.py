"""Parity tests for the MPC-CBF kernel through the C ABI (device-pointer and host-pointer paths)."""
import numpy as np
import pytest
import torch

from parity_util import check_mpc, check_mpc_parallel

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def near_goal(sc, dist=3.0):
    X = sc["X"]
    th = X[:, 2] if X.shape[1] == 4 else np.zeros(len(X))
    g = X[:, :2] + dist * np.stack([np.cos(th), np.sin(th)], 1)
    if X.shape[1] == 12:
        g = np.concatenate([g, X[:, 2:3]], axis=1)
    return g


def solve(ctrl, sc, goal, **kw):
    out = ctrl.solve(dev(sc["X"]), dev(goal), dev(sc["u_prev"]), dev(sc["OBS"]), dev(sc["nobs"]), want_pred=True, **kw)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("model,N,H,M,near,n_check", [
    ("DynamicUnicycle2D", 4096, 8, 16, False, 256),    # BASELINE config 3 at full size, oracle on 256 agents (all host cores)
    ("DynamicUnicycle2D", 256, 8, 16, True, 32),       # goals nearby: interior optima, CBF rows active
    ("DynamicUnicycle2D", 2731, 10, 64, False, 256),   # config 5's three model groups at one GPU's share, 256 agents each
    ("KinematicBicycle2D", 2731, 10, 64, False, 256),
    ("Quad3D", 2730, 10, 64, False, 256),
    ("DynamicUnicycle2D", 64, 10, 64, True, 8),
    ("SingleIntegrator2D", 128, 10, 16, False, 12),
    ("Quad3D", 64, 8, 16, True, 8),
    ("DoubleIntegrator2D", 256, 10, 16, False, 16),    # the remaining circle-barrier MPC models
    ("Quad2D", 192, 8, 16, False, 12),
    ("Unicycle2D", 256, 10, 16, False, 16),
    ("KinematicBicycle2D_C3BF", 192, 8, 16, False, 12),     # general (non-quadratic) rows: collision cone / parabolic
    ("KinematicBicycle2D_DPCBF", 192, 8, 16, False, 12),
    ("VTOL2D", 128, 8, 8, False, 16),                        # 6 states, 4 inputs, 5 state-bound rows per node
])
def test_mpc_vs_oracle(model, N, H, M, near, n_check):
    from safe_control_b200 import BatchedMPCCBF, scenes
    sc = scenes.make_scene(model, N, M, seed=1234, dense=(model == "Quad3D" and near))
    goal = near_goal(sc) if near else sc["goal"]
    ctrl = BatchedMPCCBF(sc["spec"], num_obs=M, horizon=H)
    out = solve(ctrl, sc, goal, want_active=True)
    frac_ok = (out["status"] == 0).mean()
    # (16 moving obstacles x collision-cone rows: ~20 % of these random scenes are infeasible -- the oracle's SLSQP
    #  fails on the same agents)
    # (VTOL2D: 8 obstacles in an 11 m arena at 6-12 m/s leave about half of these random scenes without a feasible horizon)
    assert frac_ok > (0.7 if model.endswith("BF") else (0.45 if model == "VTOL2D" else 0.9)), (frac_ok, np.bincount(out["status"]))
    rng = np.random.default_rng(0)
    sample = rng.choice(N, n_check, replace=False)
    # u0 agreement with the oracle's SLSQP (misses classified by cost), bit-exact active masks on the non-degenerate
    # agreeing cases, trust-constr as the second oracle solver on a few of them
    stats = check_mpc_parallel(ctrl.robot_spec, M, H, sc["X"], goal, sc["u_prev"], sc["OBS"], sc["nobs"], out, sample=sample,
                               min_agree=0.9, second_solver=1 if (n_check >= 256 and model == "DynamicUnicycle2D" and M == 16) else 0)
    print(model, N, H, M, stats, "iters mean", out["iters"].mean(), "max", out["iters"].max(), "ok", frac_ok)
    assert stats["masks_compared"] >= 0.5 * stats["agree"], stats
    # size-independent properties on the WHOLE batch: predictions satisfy the Euler model, inputs in the box
    U, px, pu = out["U"], out["pred_x"], out["pred_u"]
    nu = ctrl.nu
    lb = np.array(list(ctrl.params.u_lb)[:nu]); ub = np.array(list(ctrl.params.u_ub)[:nu])
    assert ((U >= lb - 1e-12) & (U <= ub + 1e-12)).all()
    ok = out["status"] == 0
    np.testing.assert_allclose(px[:, 0], sc["X"], atol=0)
    assert np.abs(pu[ok][:, 0] - U[ok]).max() < 1e-12
    if model in ("DynamicUnicycle2D",) or model.startswith("KinematicBicycle2D"):   # |x_k[3]| <= v_max (mpc_cbf.py:194-195, 206-207)
        assert (np.abs(px[ok][:, :, 3]) <= ctrl.params.v_max + 1e-7).all()


@pytest.mark.parametrize("model", ["DynamicUnicycle2D", "SingleIntegrator2D", "DoubleIntegrator2D"])
def test_mpc_superellipsoid_rows(model):
    """Obstacle lists mixing circles and superellipsoids (flag 1): flagged agents go through the general-row launch,
    the others through the fast path, and both must be KKT points of the oracle's NLP (which restates the reference's
    if_else(obs[6] < 0.5, circle, superellipsoid) barrier).  Without mpc_superellipsoid the flagged agents are refused."""
    from safe_control_b200 import BatchedMPCCBF, scenes
    N, H, M = 96, 8, 8
    base = scenes.make_scene(model, N, M, seed=99)
    sc = scenes.with_superellipsoids(base, every=3)
    sc["OBS"][::4] = base["OBS"][::4]                                   # every 4th agent keeps circles only
    flagged = (sc["OBS"][:, :, 6] >= 0.5).any(axis=1) & (sc["nobs"] > 0)
    assert flagged.any() and (~flagged).any()
    ctrl = BatchedMPCCBF(sc["spec"], num_obs=M, horizon=H)
    assert ctrl.params.mpc_superellipsoid == 1
    out = solve(ctrl, sc, sc["goal"])
    assert (out["status"] == 0).mean() > 0.8, np.bincount(out["status"])
    sample = np.concatenate([np.nonzero(flagged)[0][:8], np.nonzero(~flagged)[0][:4]])
    spec = {k: v for k, v in ctrl.robot_spec.items() if k != "mpc_superellipsoid"}
    stats = check_mpc(spec, M, H, sc["X"], sc["goal"], sc["u_prev"], sc["OBS"], sc["nobs"], out, sample=sample, min_agree=0.75)
    print(model, stats)
    # circles-only agents: identical to a plain run of the fast path
    plain = BatchedMPCCBF(base["spec"], num_obs=M, horizon=H)
    ref = solve(plain, sc, sc["goal"])
    np.testing.assert_array_equal(out["U"][~flagged], ref["U"][~flagged])
    assert (ref["status"][flagged] == 3).all()                           # refused loudly without the flag


@pytest.mark.parametrize("model,sum_rterms", [("KinematicBicycle2D", False), ("DynamicUnicycle2D", False), ("Quad2D", True)])
def test_optimal_decay_mpc_vs_oracle(model, sum_rterms):
    """optimal_decay_mpc_cbf: omega1 / omega2 as extra stage inputs, bilinear CBF rows, 5 obstacle slots,
    horizon 10; [u, omega1, omega2] columns in u_prev / U / pred_u.  KKT of the oracle's NLP, u0 vs SLSQP, active masks."""
    from safe_control_b200 import BatchedOptimalDecayMPCCBF, scenes
    N = 96
    sc = scenes.make_scene(model, N, 5, seed=77, dense=True)
    ctrl = BatchedOptimalDecayMPCCBF(dict(sc["spec"], od_sum_rterms=sum_rterms), num_obs=5)
    assert ctrl.horizon == 10 and ctrl.nu == ctrl.nu_model + 2 and ctrl.params.od_mpc == 1
    up = np.zeros((N, ctrl.nu))
    out = ctrl.solve(dev(sc["X"]), dev(sc["goal"]), dev(up), dev(sc["OBS"]), dev(sc["nobs"]), want_pred=True, want_active=True)
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in out.items()}
    assert out["U"].shape == (N, ctrl.nu) and out["pred_u"].shape == (N, 10, ctrl.nu)
    assert (out["status"] == 0).mean() > 0.6, np.bincount(out["status"])
    stats = check_mpc_parallel(ctrl.robot_spec, 5, 10, sc["X"], sc["goal"], up, sc["OBS"], sc["nobs"], out, sample=range(0, N, 4),
                               min_agree=0.75, optimal_decay=True, sum_rterms=sum_rterms)
    print(model, sum_rterms, stats)
    assert stats["optimal"] >= 10


def test_vtol2d_reference_horizon():
    """VTOL2D at the reference's horizon 30 (mpc_cbf.py:41; 96 KB of workspace per agent): statuses are definite, inputs in
    the box, predictions follow the kernel's own Euler map (x_{k+1}[0:3] = x_k[0:3] + dt x_k[3:6]), state bounds hold."""
    from safe_control_b200 import BatchedMPCCBF, scenes
    N, M = 300, 8
    sc = scenes.make_scene("VTOL2D", N, M, seed=11)
    ctrl = BatchedMPCCBF(sc["spec"], num_obs=M)
    assert ctrl.horizon == 30 and ctrl.active_words == (30 * 8 + 2 * 30 * 4 + 5 * 30 + 63) // 64
    out = solve(ctrl, sc, sc["goal"], want_active=True)
    ok = out["status"] == 0
    # (random scenes: 8 obstacles in an 11 m arena at 6-12 m/s over a 1.5 s horizon with pitch / descent limits leave about
    #  half of the agents without a feasible plan; those must be REPORTED infeasible / at the iteration limit, not optimal)
    assert ok.mean() > 0.35 and set(np.unique(out["status"])) <= {0, 1, 2}, np.bincount(out["status"])
    lb = np.array(list(ctrl.params.u_lb)); ub = np.array(list(ctrl.params.u_ub))
    assert ((out["U"] >= lb - 1e-12) & (out["U"] <= ub + 1e-12)).all()
    px = out["pred_x"][ok]
    assert np.abs(px[:, 1:, 0:3] - (px[:, :-1, 0:3] + 0.05 * px[:, :-1, 3:6])).max() < 1e-11
    pitch = ctrl.params.pitch_max * 3.14159 / 180
    assert (np.abs(px[:, 1:, 3]) <= ctrl.params.v_max + 1e-7).all() and (px[:, 1:, 4] >= -ctrl.params.descent_speed_max - 1e-7).all()
    assert (np.abs(px[:, 1:, 2]) <= pitch + 1e-7).all()


def test_config5_share_full_size_properties():
    """One GPU's share of BASELINE config 5 (8192 mixed DynamicUnicycle2D / KinematicBicycle2D / Quad3D agents,
    mpc_cbf horizon 10, 64 obstacle slots) through MixedMPCCBF (three concurrent launches): for every agent the
    predictions follow the Euler model from the given state, the first predicted input is the returned one, inputs
    are inside the box, |v| <= v_max on every optimal trajectory, and every row of the discrete CBF holds at stage 0."""
    from safe_control_b200 import scenes
    from safe_control_b200.mixed import MixedMPCCBF, split_counts
    models = ["DynamicUnicycle2D", "KinematicBicycle2D", "Quad3D"]
    counts = split_counts(8192, 3)
    H, M = 10, 64
    scs = [scenes.make_scene(m, n, M, seed=1234 + i) for i, (m, n) in enumerate(zip(models, counts))]
    mixed = MixedMPCCBF([s["spec"] for s in scs], num_obs=M, horizon=H)
    ins = [dict(X=dev(s["X"]), goal=dev(s["goal"]), u_prev=dev(s["u_prev"]), OBS=dev(s["OBS"]), nobs=dev(s["nobs"])) for s in scs]
    outs = []
    for g, a in zip(mixed.groups, ins):                       # (want_pred through the groups directly)
        outs.append({k: v.cpu().numpy() for k, v in g.solve(a["X"], a["goal"], a["u_prev"], a["OBS"], a["nobs"], want_pred=True).items()})
    torch.cuda.synchronize()
    both = [{k: v.cpu().numpy() for k, v in o.items()} for o in mixed.solve(ins)]
    for m, sc, g, o, o2 in zip(models, scs, mixed.groups, outs, both):
        np.testing.assert_array_equal(o["U"], o2["U"])        # concurrent launches == one after the other
        ok = o["status"] == 0
        assert ok.mean() > 0.9, (m, np.bincount(o["status"]))
        nu = g.nu
        lb = np.array(list(g.params.u_lb)[:nu]); ub = np.array(list(g.params.u_ub)[:nu])
        assert ((o["U"] >= lb - 1e-12) & (o["U"] <= ub + 1e-12)).all()
        np.testing.assert_array_equal(o["pred_x"][:, 0], sc["X"])
        assert np.abs(o["pred_u"][ok][:, 0] - o["U"][ok]).max() < 1e-12
        px, pu = o["pred_x"][ok], o["pred_u"][ok]
        if m != "Quad3D":
            assert (np.abs(px[:, :, 3]) <= g.params.v_max + 1e-7).all()
            th, v = px[:, :-1, 2], px[:, :-1, 3]
            if m == "DynamicUnicycle2D":
                nxt = np.stack([px[:, :-1, 0] + 0.05 * v * np.cos(th), px[:, :-1, 1] + 0.05 * v * np.sin(th),
                                th + 0.05 * pu[:, :, 1], v + 0.05 * pu[:, :, 0]], -1)
            else:
                b = pu[:, :, 1]; lr = g.params.rear_ax_dist
                nxt = np.stack([px[:, :-1, 0] + 0.05 * (v * np.cos(th) - v * np.sin(th) * b),
                                px[:, :-1, 1] + 0.05 * (v * np.sin(th) + v * np.cos(th) * b),
                                th + 0.05 * v / lr * b, v + 0.05 * pu[:, :, 0]], -1)
            assert np.abs(nxt - px[:, 1:]).max() < 1e-11, m
        else:
            assert np.abs(px[:, 1:, 0:6] - (px[:, :-1, 0:6] + 0.05 * px[:, :-1, 6:12])).max() < 1e-11


def test_mpc_schedule_does_not_change_results():
    """The hardest-first schedule (scb_mpccbf_solve_ws + workspace) only reorders when agents START: every agent's
    output must be bit-identical to the index-order launch (scb_mpccbf_solve), and the launch count says which ran."""
    from safe_control_b200 import BatchedMPCCBF, scenes
    N, H, M = 3000, 8, 16                                             # more agents than one persistent wave
    sc = scenes.make_scene("DynamicUnicycle2D", N, M, seed=77)
    ctrl = BatchedMPCCBF(sc["spec"], num_obs=M, horizon=H)
    a = solve(ctrl, sc, sc["goal"], want_active=True); n_sched = ctrl.launches
    ctrl.schedule = False
    b = solve(ctrl, sc, sc["goal"], want_active=True); n_plain = ctrl.launches - n_sched
    assert (n_sched, n_plain) == (3, 1)
    for k in ("U", "status", "iters", "pred_u", "pred_x", "kkt", "active"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


def test_mpc_track_mask_and_host_path():
    from safe_control_b200 import BatchedMPCCBF, HostContext, scenes
    N, H, M = 96, 8, 16
    sc = scenes.make_scene("DynamicUnicycle2D", N, M, seed=3)
    ctrl = BatchedMPCCBF(sc["spec"], num_obs=M, horizon=H)
    track = np.ones(N, np.int32); track[::3] = 0
    out = ctrl.solve(dev(sc["X"]), dev(sc["goal"]), dev(sc["u_prev"]), dev(sc["OBS"]), dev(sc["nobs"]),
                     U_ref=dev(sc["U_ref"]), track=dev(track))
    U = out["U"].cpu().numpy()
    np.testing.assert_array_equal(U[::3], sc["U_ref"][::3])          # not 'track' -> u_ref untouched (mpc_cbf.py:379-381)
    ref = solve(ctrl, sc, sc["goal"])
    np.testing.assert_allclose(U[track == 1], ref["U"][track == 1], atol=0)
    ctx = HostContext(0)
    h = ctx.mpccbf_solve(ctrl.params, M, H, sc["X"], sc["goal"], sc["u_prev"], sc["OBS"], sc["nobs"], want_pred=True,
                         want_active=True)
    np.testing.assert_array_equal(h["U"], ref["U"]); np.testing.assert_array_equal(h["status"], ref["status"])
    np.testing.assert_array_equal(h["pred_u"], ref["pred_u"])
    refa = solve(ctrl, sc, sc["goal"], want_active=True)
    np.testing.assert_array_equal(h["active"].view(np.int64), refa["active"])
    # track < 0: the agent is skipped, its outputs are left untouched (closed loop: frozen agents)
    track2 = np.ones(N, np.int32); track2[1::4] = -1
    keep = ctrl.solve(dev(sc["X"]), dev(sc["goal"]), dev(sc["u_prev"]), dev(sc["OBS"]), dev(sc["nobs"]),
                      U_ref=dev(sc["U_ref"]), track=dev(track2))
    assert keep["U"].shape == (N, 2)
    with pytest.raises(TypeError):                                   # raw-pointer arguments are validated in Python
        ctrl.solve(dev(sc["X"]), dev(sc["goal"]), dev(sc["u_prev"]), dev(sc["OBS"]), dev(sc["nobs"]),
                   U_ref=dev(sc["U_ref"]), track=torch.ones(N, dtype=torch.int64, device="cuda"))
    with pytest.raises(ValueError):
        ctx.mpccbf_solve(ctrl.params, M, H, sc["X"][:, :3].copy(), sc["goal"], sc["u_prev"], sc["OBS"], sc["nobs"])
    ctx.close()


def test_mpc_unsupported_models_fail_loudly():
    from safe_control_b200 import BatchedMPCCBF
    from safe_control_b200._lib import ScbError
    from safe_control_b200.params import NotCompatibleError
    with pytest.raises(NotCompatibleError):                     # no agent_barrier_dt in the reference -> no MPC
        BatchedMPCCBF({"model": "Manipulator2D"}, num_obs=4, horizon=4)
    ctrl = BatchedMPCCBF({"model": "DynamicUnicycle2D"}, num_obs=4, horizon=40)      # beyond the compiled horizon
    z = lambda *s: torch.zeros(s, dtype=torch.float64, device="cuda")
    with pytest.raises(ScbError):
        ctrl.solve(z(2, 4), z(2, 2), z(2, 2), z(2, 4, 7))
