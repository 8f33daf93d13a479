"""The oracle restatement vs fixtures produced by the REFERENCE'S OWN CODE
(tests/golden/gen_from_reference.py, run in the build container through
oracle/refshim).  CPU only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.models import make_model, MODELS, MODELS_QP_EXTRA
from oracle.controllers import OracleCBFQP, OracleOptimalDecayCBFQP

RTOL, ATOL = 1e-12, 1e-12


def _load(fname):
    """Both fixture sets: <name>.npz (round-1 models) and <name>2.npz (models added for SURVEY 8f-2)."""
    out = {}
    for f in (fname, fname.replace(".npz", "2.npz")):
        path = os.path.join(GOLDEN, f)
        if not os.path.exists(path):
            continue
        z = np.load(path)
        for k in z.files:
            tag, key = k.rsplit("/", 1)
            out.setdefault(tag, {})[key] = z[k]
    return out


@pytest.mark.parametrize("name", MODELS + MODELS_QP_EXTRA)
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
        else:
            nom = m.nominal_input(x, goal[:2])
        np.testing.assert_allclose(nom, d["NOM"][i], rtol=1e-11, atol=1e-12)
        for j, o in enumerate(d["OBS"][i]):
            if name != "Quad3D":
                parts = m.agent_barrier(x, o)
                got = np.concatenate([np.asarray(p, float).reshape(-1) for p in parts])
                np.testing.assert_allclose(got, d["CT"][i][j], rtol=1e-11, atol=1e-11)
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
    assert len(data) == 11
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
