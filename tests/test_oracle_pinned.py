"""The oracle restatement vs fixtures produced by the REFERENCE'S OWN CODE
(tests/golden/gen_from_reference.py, run in the build container through
oracle/refshim).  CPU only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.models import make_model, MODELS, MODELS_QP_EXTRA, MODELS_EXTRA3
from oracle.controllers import OracleCBFQP, OracleOptimalDecayCBFQP

RTOL, ATOL = 1e-12, 1e-12


def _load(fname):
    """All fixture sets: <name>.npz (first models), <name>2.npz (DI / Quad2D / DPCBF), <name>3.npz (Unicycle2D)."""
    out = {}
    for f in (fname, fname.replace(".npz", "2.npz"), fname.replace(".npz", "3.npz"), fname.replace(".npz", "4.npz")):
        path = os.path.join(GOLDEN, f)
        if not os.path.exists(path):
            continue
        z = np.load(path)
        for k in z.files:
            tag, key = k.rsplit("/", 1)
            out.setdefault(tag, {})[key] = z[k]
    return out


@pytest.mark.parametrize("name", MODELS + MODELS_QP_EXTRA + MODELS_EXTRA3)
def test_models_match_reference(name):
    d = _load("ref_models.npz")[name]
    m = make_model({"model": name})
    for i in range(len(d["X"])):
        x, u, goal = d["X"][i], d["U"][i], d["GOAL"][i]
        np.testing.assert_allclose(m.f(x), d["F"][i], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(m.g(x), d["G"][i], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(m.step(x.copy(), u), d["STEP"][i], rtol=RTOL, atol=ATOL)
        if name == "SingleIntegrator2D":
            nom = m.nominal_input(x, goal[:2])
        elif name == "Quad3D":
            nom = m.nominal_input(x, goal)
        elif name.startswith("KinematicBicycle2D"):
            nom = m.nominal_input(x, goal[:2], 0.05, 2.0, 1.0, 1.0)
        elif name == "Quad2D":
            nom = d["NOM"][i]                 # cascaded PD law off the solve path: not restated
        elif name == "Unicycle2D":
            nom = m.nominal_input(x, goal[:2], 0.05, 2.0, 1.0)
        else:
            nom = m.nominal_input(x, goal[:2])
        np.testing.assert_allclose(nom, d["NOM"][i], rtol=1e-11, atol=1e-12)
        for j, o in enumerate(d["OBS"][i]):
            if name != "Quad3D":
                parts = m.agent_barrier(x, o)
                got = np.concatenate([np.asarray(p, float).reshape(-1) for p in parts])
                np.testing.assert_allclose(got, d["CT"][i][j], rtol=1e-11, atol=1e-11)
            if name == "Manipulator2D":
                continue                          # no agent_barrier_dt in the reference
            got = np.array(m.barrier_dt(x.copy(), u, o), float)
            np.testing.assert_allclose(got, d["DT"][i][j], rtol=1e-10, atol=1e-11)


def test_quad3d_agent_barrier_raises_like_reference():
    m = make_model({"model": "Quad3D"})
    with pytest.raises(NotImplementedError):
        m.agent_barrier(np.zeros(12), np.zeros(7))


def _spec_from_tag(tag):
    name, _, rest = tag.partition("+")
    spec = {"model": name}
    if rest:
        for kv in rest.split(","):
            k, v = kv.split("=")
            spec[k] = v if k == "cbf_mode" else float(v)
    return spec


def test_cbfqp_matches_reference_end_to_end():
    data = _load("ref_cbfqp.npz")
    assert len(data) == 15
    for tag, d in data.items():
        spec = _spec_from_tag(tag)
        num_obs = d["A"].shape[1]
        ctrl = OracleCBFQP(spec, num_obs=num_obs)
        for i in range(len(d["X"])):
            k = int(d["NOBS"][i])
            obs = d["OBS"][i][:k] if k else None
            A, b = ctrl.rows(d["X"][i], obs)
            np.testing.assert_allclose(A, d["A"][i], rtol=1e-11, atol=1e-11, err_msg=tag)
            np.testing.assert_allclose(b, d["B"][i], rtol=1e-11, atol=1e-11, err_msg=tag)
            u, info = ctrl.solve(d["X"][i], d["UREF"][i], obs)
            assert info["status"] == d["STATUS"][i], (tag, i)
            if d["STATUS"][i] == 0:
                np.testing.assert_allclose(u, d["U"][i], rtol=1e-9, atol=1e-10, err_msg=tag)


def test_optimal_decay_matches_reference_end_to_end():
    data = _load("ref_odcbf.npz")
    for name, d in data.items():
        ctrl = OracleOptimalDecayCBFQP({"model": name})
        for i in range(len(d["X"])):
            obs = d["OBS"][i] if d["HAS"][i] else None
            u, om, info = ctrl.solve(d["X"][i], d["UREF"][i], obs)
            assert info["status"] == d["STATUS"][i]
            if d["STATUS"][i] == 0:
                np.testing.assert_allclose(u, d["U"][i], rtol=1e-9, atol=1e-10, err_msg=name)
                np.testing.assert_allclose(om, d["OMEGA"][i], rtol=1e-9, atol=1e-10, err_msg=name)


def test_qp_fixtures_against_an_independent_solver():
    """The fixture's u comes from oracle/qp_exact.py (through the cvxpy stand-in): cross-check the same stated QPs --
    (P, q, G, h) as the reference's own problem statement produces them -- with scipy's SLSQP, which shares no code with
    the enumeration solver (an independent third solver)."""
    from scipy.optimize import minimize
    checked = 0
    for fname, make in (("ref_cbfqp.npz", lambda tag: OracleCBFQP(_spec_from_tag(tag), num_obs=None)),
                        ("ref_odcbf.npz", lambda tag: OracleOptimalDecayCBFQP({"model": tag}))):
        for tag, d in _load(fname).items():
            ctrl = make(tag)
            if isinstance(ctrl, OracleCBFQP):
                ctrl.num_obs = d["A"].shape[1]
            for i in range(0, len(d["X"]), 3):
                if d["STATUS"][i] != 0:
                    continue
                if isinstance(ctrl, OracleCBFQP):
                    k = int(d["NOBS"][i])
                    if k == 0:
                        continue
                    P, q, G, h = ctrl.qp(d["X"][i], d["UREF"][i], d["OBS"][i][:k])
                    want = d["U"][i]
                else:
                    P, q, G, h = ctrl.qp(d["X"][i], d["UREF"][i], d["OBS"][i] if d["HAS"][i] else None)
                    want = np.concatenate([d["U"][i], d["OMEGA"][i]])
                sc = 1.0 / max(1.0, float(np.abs(np.diag(P)).max()))           # (the slack penalty is 1e4: scale the cost)
                res = minimize(lambda z: sc * (0.5 * z @ P @ z + q @ z), np.zeros(q.size), jac=lambda z: sc * (P @ z + q), method="SLSQP",
                               constraints=[{"type": "ineq", "fun": lambda z: h - G @ z, "jac": lambda z: -G}],
                               options={"ftol": 1e-14, "maxiter": 400})
                if not res.success:
                    continue
                J = lambda z: 0.5 * z @ P @ z + q @ z
                assert np.all(G @ res.x - h <= 1e-7)
                assert J(want) <= J(res.x) + 1e-7 * max(1.0, abs(J(want))), (fname, tag, i)      # the fixture's point is at least as good
                np.testing.assert_allclose(res.x, want, rtol=2e-5, atol=2e-5, err_msg=f"{fname} {tag} {i}")
                checked += 1
    assert checked >= 60


# ---- MPC-CBF problem statement: oracle/mpc_cbf.py vs what the REFERENCE'S OWN mpc_cbf.py hands to do-mpc ----------
# (tests/golden/gen_mpc_from_reference.py; probing stand-in for do_mpc, nothing solved).  Pins the Euler rhs, the
# stage cost (Q, goal padding), every CBF constraint value incl. the model's own step and the alphas, the dummy
# obstacle padding, input / state bounds, the rterm weights and the horizon.  do-mpc's transcription of these pieces
# into the NLP (sum over stages + terminal cost, rterm on input increments) stays as documented in SURVEY.md 8a.
MPC_ORACLE_MODELS = ("SingleIntegrator2D", "DynamicUnicycle2D", "KinematicBicycle2D", "KinematicBicycle2D_C3BF", "Quad3D",
                     "DoubleIntegrator2D", "Quad2D", "Unicycle2D", "KinematicBicycle2D_DPCBF", "VTOL2D")


def test_mpc_statement_matches_reference():
    import torch
    from oracle.mpc_cbf import OracleMPCCBF
    data = _load("ref_mpc_statement.npz")
    seen = 0
    for tag, d in data.items():
        spec = _spec_from_tag(tag)
        if "mpc_horizon" in spec:
            spec["mpc_horizon"] = int(spec["mpc_horizon"])
        if spec["model"] not in MPC_ORACLE_MODELS:
            continue                                   # DPCBF: recorded, MPC not built yet
        seen += 1
        M = d["cbf"].shape[1]
        o = OracleMPCCBF(spec, num_obs=M)
        assert o.H == int(d["horizon"][0]) and int(d["n_robust"][0]) == 0 and float(d["t_step"][0]) == o.dt
        assert float(d["lterm_is_mterm"][0]) == 1.0                       # terminal cost == stage cost expression
        np.testing.assert_array_equal(o.Rw.numpy(), d["R"][0])
        np.testing.assert_array_equal(o.u_lb, d["lb_u"][0]); np.testing.assert_array_equal(o.u_ub, d["ub_u"][0])
        if spec["model"] == "VTOL2D":                                       # mpc_cbf.py:227-232
            lbx = np.full(6, -np.inf); ubx = np.full(6, np.inf)
            for i_, sgn, off in o.state_bounds:
                if sgn < 0:
                    ubx[i_] = off
                else:
                    lbx[i_] = -off
            np.testing.assert_array_equal(d["ub_x"][0], ubx); np.testing.assert_array_equal(d["lb_x"][0], lbx)
            assert len(o.state_bounds) == 5
        elif o.has_vbound:                                                  # |x[3]| <= v_max, nothing else bounded
            v = o.spec["v_max"]
            np.testing.assert_array_equal(d["ub_x"][0], [np.inf, np.inf, np.inf, v])
            np.testing.assert_array_equal(d["lb_x"][0], [-np.inf, -np.inf, -np.inf, -v])
        else:
            assert np.isinf(d["ub_x"][0]).all() and np.isinf(d["lb_x"][0]).all()
        assert (d["cons_ub"] == 0).all()
        al = d["alphas"][0]
        if "alpha" in o.par:
            assert o.par["alpha"] == al[0]
        else:
            assert (o.par["alpha1"], o.par["alpha2"]) == (al[1], al[2])
        for i in range(len(d["X"])):
            x, u = torch.tensor(d["X"][i])[None], torch.tensor(d["U"][i])[None]
            k = int(d["NOBS"][i])
            obs = d["OBS"][i][:k] if k else None
            np.testing.assert_allclose(o.tm.euler(x, u)[0].numpy(), d["x_next"][i], rtol=1e-13, atol=1e-13, err_msg=tag)
            ob = o.pad_obs(obs)
            np.testing.assert_array_equal(ob.numpy(), d["tvp_obs"][i])       # dummy rows [1000, 1000, 0, ...]
            g = np.zeros(o.nx); gl = d["GOAL"][i][: (3 if spec["model"] == "Quad3D" else 2)]; g[: gl.size] = gl
            np.testing.assert_array_equal(g, d["tvp_goal"][i])               # goal padded with zeros
            e = d["X"][i] - g
            np.testing.assert_allclose(float((e * e * o.Q.numpy()).sum()), d["cost"][i], rtol=1e-12, err_msg=tag)
            # CBF constraints of one stage: w = [x_0, x_1 | u_0] with H = 1 semantics -> call the stage pieces directly
            tm, p = o.tm, o.par
            x1 = tm.own_step(x, u); h0, h1 = tm.h(x, ob), tm.h(x1, ob)
            if "alpha" in p:
                c = (h1 - h0) + p["alpha"] * h0
            else:
                x2 = tm.own_step(x1, u); h2 = tm.h(x2, ob)
                c = (h2 - 2 * h1 + h0) + (p["alpha1"] + p["alpha2"]) * (h1 - h0) + p["alpha1"] * p["alpha2"] * h0
            np.testing.assert_allclose(c[0].numpy(), d["cbf"][i], rtol=1e-9, atol=1e-9, err_msg=f"{tag} probe {i}")
    assert seen == 13
