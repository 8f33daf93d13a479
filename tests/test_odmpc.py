"""optimal_decay_mpc_cbf: statement pinned against the reference's own file, solutions against the oracle.

The reference class (position_control/optimal_decay_mpc_cbf.py) is constructed UNMODIFIED by
tests/golden/gen_odmpc_from_reference.py through oracle/refshim; the fixtures hold x_next, the state cost, the five CBF
constraint values with the bilinear omega terms, the values of its two expression rterms, bounds, horizon, tvp padding and
alphas at seeded probe points.  Checked against them: the oracle's NLP ingredients (oracle/mpc_cbf.py OracleODMPCCBF) and
the CUDA kernel's own statement (CPU build of the kernel source).  Then the kernel's solutions must be KKT points of the
oracle's NLP and agree with its SLSQP solve.  do-mpc's handling of the two set_rterm calls is UNPINNED (both readings are
implemented and tested: scb_params.od_sum_rterms)."""
import ctypes as C

import numpy as np
import pytest

from hostsim_util import hostsim, hs_mpccbf_solve, ptr
from parity_util import check_mpc
from safe_control_b200.params import resolve_params, NotCompatibleError
from safe_control_b200 import scenes
from test_oracle_pinned import _load, _spec_from_tag


def test_reference_itself_raises_for_dynamic_unicycle():
    """DynamicUnicycle2D's agent_barrier_dt reads obs[6] of the 5-column obstacle row this controller defines
    (optimal_decay_mpc_cbf.py:125; dynamic_unicycle2D.py:224): the reference cannot construct the controller for its own
    first-listed model.  Recorded by the fixture generator; we implement the circle branch that row can only mean."""
    d = _load("ref_odmpc_statement.npz")
    assert "IndexError" in str(d["DynamicUnicycle2D"]["raises"])
    assert set(d) == {"DynamicUnicycle2D", "KinematicBicycle2D", "Quad2D", "VTOL2D", "KinematicBicycle2D+a_max=2.0,v_max=2.0"}


def _cases():
    for tag, d in _load("ref_odmpc_statement.npz").items():
        if "raises" not in d:
            yield tag, d


def test_oracle_statement_matches_reference():
    import torch
    from oracle.mpc_cbf import OracleODMPCCBF
    seen = 0
    for tag, d in _cases():
        spec = _spec_from_tag(tag)
        o = OracleODMPCCBF(spec)
        nm = o.nu_model
        assert o.H == int(d["horizon"][0]) and float(d["t_step"][0]) == o.dt and float(d["lterm_is_mterm"][0]) == 1.0
        np.testing.assert_array_equal(np.asarray(o.par["R"], float), d["R"][0])
        assert (o.par["alpha1"], o.par["alpha2"]) == tuple(d["alphas"][0]) and o.p_sb == tuple(d["p_sb"][0]) and o.omega0 == tuple(d["omega0"][0])
        np.testing.assert_array_equal(o.u_lb[:nm], d["lb_u"][0]); np.testing.assert_array_equal(o.u_ub[:nm], d["ub_u"][0])
        assert float(d["omega_bounded"][0]) == 0.0 and np.isinf(o.u_lb[nm:]).all()          # the omegas are free variables
        lbx = np.full(o.nx, -np.inf); ubx = np.full(o.nx, np.inf)
        for i_, sgn, off in o.state_bounds:
            if sgn < 0:
                ubx[i_] = off
            else:
                lbx[i_] = -off
        np.testing.assert_array_equal(d["ub_x"][0], ubx); np.testing.assert_array_equal(d["lb_x"][0], lbx)
        assert (d["cons_ub"] == 0).all()
        for i in range(len(d["X"])):
            x = torch.tensor(d["X"][i])[None]
            u = torch.tensor(np.concatenate([d["U"][i], d["OMEGA"][i]]))[None]
            k = int(d["NOBS"][i])
            obs = d["OBS"][i][:k] if k else None
            np.testing.assert_allclose(o.tm.euler(x, u[:, :nm])[0].numpy(), d["x_next"][i], rtol=1e-13, atol=1e-13, err_msg=tag)
            ob = o.pad_obs(obs)
            np.testing.assert_array_equal(ob.numpy()[:, :5], d["tvp_obs"][i])                # 5 x 5 tvp, dummy rows [1000, 1000, 0, 0, 0]
            g = np.zeros(o.nx); g[:2] = d["GOAL"][i]
            np.testing.assert_array_equal(g, d["tvp_goal"][i])
            e = d["X"][i] - g
            np.testing.assert_allclose(float((e * e * o.Q.numpy()).sum()), d["cost"][i], rtol=1e-12, err_msg=tag)
            # the two expression rterms, in call order (:178-185): sum R u^2, then the omega penalty
            ru = float((np.asarray(o.par["R"]) * d["U"][i] ** 2).sum())
            ro = float(o.p_sb[0] * (d["OMEGA"][i][0] - 1) ** 2 + o.p_sb[1] * (d["OMEGA"][i][1] - 1) ** 2)
            np.testing.assert_allclose(d["rterm_calls"][i], [ru, ro], rtol=1e-12)
            w = torch.cat([x.reshape(-1), torch.zeros(o.nx), u.reshape(-1)])                # H = 1 slice: [x_0, x_1 | u_0]
            o1 = OracleODMPCCBF(spec, horizon=1)
            np.testing.assert_allclose(o1.cbf(w, ob).numpy(), d["cbf"][i], rtol=1e-9, atol=1e-8, err_msg=f"{tag} probe {i}")
            seen += 1
    assert seen == 80


@pytest.mark.parametrize("sum_rterms", [False, True])
def test_kernel_statement_matches_reference(sum_rterms):
    """The CUDA kernel's own stage map / cost / rows for the optimal-decay variants (CPU build of scb_mpc.cuh)."""
    lib = hostsim()
    seen = 0
    for tag, d in _cases():
        spec = dict(_spec_from_tag(tag), od_sum_rterms=sum_rterms)
        p, _ = resolve_params(spec, "optimal_decay_mpc_cbf", lib=lib)
        assert p.od_mpc == 1 and p.od_sum_rterms == int(sum_rterms)
        for i in range(len(d["X"])):
            k = min(int(d["NOBS"][i]), 5)                                                   # more than 5: the first 5 (:278-280)
            obs = np.zeros((5, 7)); obs[:k, :5] = d["OBS"][i][:k]
            x = np.ascontiguousarray(d["X"][i]); u = np.ascontiguousarray(np.concatenate([d["U"][i], d["OMEGA"][i]]))
            goal = np.ascontiguousarray(d["GOAL"][i])
            xn = np.zeros(p.nx); cost = C.c_double(); cbf = np.zeros(5)
            rc = lib.hostsim_mpc_statement(C.byref(p), 5, k, ptr(x), ptr(u), ptr(goal), ptr(np.ascontiguousarray(obs)), ptr(xn),
                                           C.byref(cost), ptr(cbf))
            assert rc == 0
            np.testing.assert_allclose(xn, d["x_next"][i], rtol=1e-13, atol=1e-13, err_msg=tag)
            want = d["cost"][i] + d["rterm_calls"][i][1] + (d["rterm_calls"][i][0] if sum_rterms else 0.0)
            np.testing.assert_allclose(cost.value, want, rtol=1e-12, err_msg=tag)
            np.testing.assert_allclose(cbf, d["cbf"][i], rtol=1e-9, atol=1e-8, err_msg=f"{tag} probe {i}")
            seen += 1
    assert seen == 80


@pytest.mark.parametrize("model,N,sum_rterms", [("KinematicBicycle2D", 8, False), ("Quad2D", 6, False), ("DynamicUnicycle2D", 8, False),
                                                ("DynamicUnicycle2D", 6, True), ("KinematicBicycle2D", 6, True)])
def test_solutions_vs_oracle(model, N, sum_rterms):
    sc = scenes.make_scene(model, N, 5, seed=77, dense=True)
    p, spec = resolve_params(dict(sc["spec"], od_sum_rterms=sum_rterms), "optimal_decay_mpc_cbf", lib=hostsim())
    H = int(spec["mpc_horizon"])
    assert H == 10
    up = np.zeros((N, p.nu + 2))                      # do-mpc's u0 at the first call: zeros, omegas included
    out = hs_mpccbf_solve(p, H, sc["X"], sc["goal"], up, sc["OBS"], sc["nobs"], want_active=True)
    assert out["U"].shape == (N, p.nu + 2) and (out["status"] == 0).mean() >= 0.6, out["status"]
    stats = check_mpc(spec, 5, H, sc["X"], sc["goal"], up, sc["OBS"], sc["nobs"], out, min_agree=0.75, optimal_decay=True,
                      sum_rterms=sum_rterms)
    print(model, sum_rterms, stats, "iters", out["iters"], "omega", out["U"][:, -2:].round(3).tolist())
    assert stats["optimal"] >= 1


def test_unsupported_models_are_refused():
    for m in ("SingleIntegrator2D", "Quad3D", "KinematicBicycle2D_C3BF", "DoubleIntegrator2D"):
        with pytest.raises(NotCompatibleError):
            resolve_params({"model": m}, "optimal_decay_mpc_cbf", lib=hostsim())
