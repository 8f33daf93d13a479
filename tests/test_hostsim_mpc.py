"""CPU check of the MPC-CBF kernel body (g++ build of scb_mpc.cuh at LANES = 1) against the oracle NLP."""
import numpy as np
import pytest

from hostsim_util import hostsim, hs_mpccbf_solve
from parity_util import check_mpc
from safe_control_b200.params import resolve_params
from safe_control_b200 import scenes


def near_goal(sc, dist=3.0):
    X = sc["X"]
    th = X[:, 2] if X.shape[1] == 4 else np.zeros(len(X))
    g = X[:, :2] + dist * np.stack([np.cos(th), np.sin(th)], 1)
    if X.shape[1] == 12:                       # Quad3D goals are (x, y, z)
        g = np.concatenate([g, X[:, 2:3]], axis=1)
    return g


@pytest.mark.parametrize("model,N,H,M,near", [("DynamicUnicycle2D", 10, 8, 16, False),
                                               ("DynamicUnicycle2D", 8, 8, 16, True),
                                               ("KinematicBicycle2D", 6, 6, 8, False),
                                               ("SingleIntegrator2D", 8, 10, 8, False),
                                               ("Quad3D", 6, 6, 8, False),
                                               ("Quad3D", 5, 5, 6, True),
                                               ("DoubleIntegrator2D", 8, 8, 8, False),
                                               ("Quad2D", 6, 6, 8, False),
                                               ("Unicycle2D", 8, 10, 8, False),
                                               ("KinematicBicycle2D_C3BF", 10, 6, 8, False),
                                               ("KinematicBicycle2D_DPCBF", 10, 6, 8, False),
                                               ("VTOL2D", 6, 8, 6, False)])
def test_mpc_vs_oracle(model, N, H, M, near):
    sc = scenes.make_scene(model, N, M, seed=4321 if model != "VTOL2D" else 11, dense=(model == "Quad3D" and near))
    goal = near_goal(sc) if near else sc["goal"]
    p, spec = resolve_params(sc["spec"], "mpc_cbf", lib=hostsim())
    out = hs_mpccbf_solve(p, H, sc["X"], goal, sc["u_prev"], sc["OBS"], sc["nobs"], want_active=True)
    assert (out["status"] == 0).mean() >= 0.8, out["status"]
    stats = check_mpc(spec, M, H, sc["X"], goal, sc["u_prev"], sc["OBS"], sc["nobs"], out, min_agree=0.9,
                      second_solver=1 if model in ("SingleIntegrator2D", "DynamicUnicycle2D") and not near else 0)
    print(model, stats, "iters", out["iters"])
    assert stats["masks_compared"] >= 1 or stats["agree"] == 0, stats


@pytest.mark.parametrize("model", ["DynamicUnicycle2D", "SingleIntegrator2D", "DoubleIntegrator2D"])
def test_mpc_superellipsoid_rows_vs_oracle(model):
    """if_else(obs[6] < 0.5, circle, superellipsoid) rows (e.g. dynamic_unicycle2D.py:204-228) through the general-row path."""
    N, H, M = 10, 6, 8
    sc = scenes.with_superellipsoids(scenes.make_scene(model, N, M, seed=4321))
    spec = {k: v for k, v in sc["spec"].items() if k != "mpc_superellipsoid"}
    p, spec = resolve_params(spec, "mpc_cbf", lib=hostsim())
    out = hs_mpccbf_solve(p, H, sc["X"], sc["goal"], sc["u_prev"], sc["OBS"], sc["nobs"])
    assert (out["status"] == 0).mean() >= 0.8, out["status"]
    stats = check_mpc(spec, M, H, sc["X"], sc["goal"], sc["u_prev"], sc["OBS"], sc["nobs"], out, min_agree=0.8)
    print(model, stats)


def test_vtol2d_full_horizon():
    """VTOL2D at the reference's own horizon of 30 (mpc_cbf.py:41): the answers must be KKT points of the oracle's NLP
    (SLSQP itself rarely converges at 300 variables, so u0 agreement is covered by the H = 8 case above)."""
    sc = scenes.make_scene("VTOL2D", 3, 6, seed=11)
    p, spec = resolve_params(sc["spec"], "mpc_cbf", lib=hostsim())
    assert spec["mpc_horizon"] == 30
    out = hs_mpccbf_solve(p, 30, sc["X"], sc["goal"], sc["u_prev"], sc["OBS"], sc["nobs"], want_active=True)
    assert (out["status"] == 0).sum() >= 2, out["status"]
    from oracle.mpc_cbf import OracleMPCCBF
    o = OracleMPCCBF(spec, num_obs=6, horizon=30)
    for i in np.nonzero(out["status"] == 0)[0]:
        k = int(sc["nobs"][i])
        kk, gmin, comp = o.kkt_error(sc["X"][i], sc["goal"][i], sc["u_prev"][i], sc["OBS"][i][:k], out["pred_u"][i])
        assert gmin >= -1e-7 and kk <= 2e-4 and comp <= 1e-5, (i, kk, gmin, comp)


@pytest.mark.parametrize("model", ["KinematicBicycle2D_C3BF", "KinematicBicycle2D_DPCBF", "Unicycle2D", "DoubleIntegrator2D",
                                   "Quad2D", "DynamicUnicycle2D", "Quad3D"])
def test_mpc_edge_shapes(model):
    """Smallest and largest compiled shapes, agents without any obstacle (all dummy rows, mpc_cbf.py:346-364) and with
    one: finite inputs inside the box, a definite status, and the no-obstacle agent always solves."""
    for M, H in ((1, 1), (4, 3), (16, 16 if model != "Quad3D" else 12)):
        sc = scenes.make_scene(model, 5, M, seed=5)
        p, spec = resolve_params(sc["spec"], "mpc_cbf", lib=hostsim())
        nobs = sc["nobs"].copy(); nobs[0] = 0; nobs[1] = min(1, M)
        out = hs_mpccbf_solve(p, H, sc["X"], sc["goal"], sc["u_prev"], sc["OBS"], nobs)
        nu = p.nu
        lb = np.array(list(p.u_lb)[:nu]); ub = np.array(list(p.u_ub)[:nu])
        assert np.isfinite(out["U"]).all() and ((out["U"] >= lb - 1e-12) & (out["U"] <= ub + 1e-12)).all(), (model, M, H)
        assert set(np.unique(out["status"])) <= {0, 1, 2}, (model, M, H, out["status"])
        assert out["status"][0] == 0, (model, M, H)


def test_mpc_no_obstacles_is_box_clipped_tracking():
    """Without obstacles the solution must equal the unconstrained-by-CBF MPC; with the goal far
    ahead and straight, full acceleration saturates: u0 = [a_max, ~0]."""
    p, spec = resolve_params({"model": "DynamicUnicycle2D"}, "mpc_cbf", lib=hostsim())
    X = np.array([[0.0, 0.0, 0.0, 0.2]]); goal = np.array([[10.0, 0.0]]); up = np.zeros((1, 2))
    OBS = np.zeros((1, 4, 7)); nobs = np.array([0], np.int32)
    out = hs_mpccbf_solve(p, 8, X, goal, up, OBS, nobs)
    assert out["status"][0] == 0
    assert abs(out["U"][0, 0] - 0.5) < 1e-6 and abs(out["U"][0, 1]) < 1e-6
    # Euler prediction is consistent with the inputs
    x = X[0].copy()
    for k in range(8):
        u = out["pred_u"][0, k]
        x = x + 0.05 * np.array([x[3] * np.cos(x[2]), x[3] * np.sin(x[2]), u[1], u[0]])
        np.testing.assert_allclose(out["pred_x"][0, k + 1], x, atol=1e-12)


@pytest.mark.parametrize("fma", [False, True], ids=["nofma", "fma"])
def test_kernel_statement_matches_reference(fma):
    """(Both CPU builds of the kernel source: without FMA contraction and with it, as nvcc contracts.)
    The MPC kernel's own problem statement (Euler map, stage cost, CBF constraint of every obstacle slot incl.
    the model's own step, dummy-obstacle padding) vs what the REFERENCE'S mpc_cbf.py hands to do-mpc at seeded probe
    points (tests/golden/ref_mpc_statement.npz, generated through oracle/refshim's probing do_mpc stand-in)."""
    import ctypes as C
    from hostsim_util import ptr
    from test_oracle_pinned import _load, _spec_from_tag
    import hostsim_util
    hostsim_util.use_fma(fma)
    try:
        _statement_check(_load, _spec_from_tag, C, ptr)
    finally:
        hostsim_util.use_fma(False)


def _statement_check(_load, _spec_from_tag, C, ptr):
    lib = hostsim()
    seen = n_se = 0
    for tag, d in _load("ref_mpc_statement.npz").items():
        spec = _spec_from_tag(tag)
        if spec["model"] not in ("SingleIntegrator2D", "DynamicUnicycle2D", "KinematicBicycle2D", "Quad3D", "DoubleIntegrator2D",
                                 "Quad2D", "Unicycle2D", "KinematicBicycle2D_C3BF", "KinematicBicycle2D_DPCBF", "VTOL2D"):
            continue
        spec.pop("mpc_horizon", None)
        p, _ = resolve_params(spec, "mpc_cbf", lib=lib)
        M = d["cbf"].shape[1]
        for i in range(len(d["X"])):
            k = int(d["NOBS"][i])
            obs = np.nan_to_num(d["OBS"][i][:M].copy(), nan=0.0)
            n_se += int((obs[:k, 6] != 0).any())            # superellipsoid rows -> the general-row variant of the model
            x, u, goal = np.ascontiguousarray(d["X"][i]), np.ascontiguousarray(d["U"][i]), np.ascontiguousarray(d["GOAL"][i])
            xn = np.zeros(p.nx); cost = C.c_double(); cbf = np.zeros(M)
            rc = lib.hostsim_mpc_statement(C.byref(p), M, k, ptr(x), ptr(u), ptr(goal), ptr(np.ascontiguousarray(obs)),
                                           ptr(xn), C.byref(cost), ptr(cbf))
            assert rc == 0
            np.testing.assert_allclose(xn, d["x_next"][i], rtol=1e-13, atol=1e-13, err_msg=tag)
            np.testing.assert_allclose(cost.value, d["cost"][i], rtol=1e-12, err_msg=tag)
            np.testing.assert_allclose(cbf, d["cbf"][i], rtol=1e-9, atol=1e-8, err_msg=f"{tag} probe {i}")
            seen += 1
    assert seen > 275 and n_se >= 15
