"""CPU checks of the CUDA source's per-agent math (compiled with g++ at LANES=1, see
tests/_hostsim/hostsim.cpp) against the reference-generated fixtures and the oracle, and of the
host-side scene preparation against the oracle.  The same checks run on the real kernels in
test_gpu_qp.py; this file exists so kernel-math regressions are caught on the CPU-only box."""
import numpy as np
import pytest

import hostsim_util
from hostsim_util import hostsim, hs_cbfqp_rows, hs_cbfqp_solve, hs_odcbf_solve
from parity_util import check_cbfqp, check_odcbf
from safe_control_b200.params import resolve_params
from safe_control_b200 import scenes
from test_oracle_pinned import _load, _spec_from_tag
from oracle.models import make_model
from oracle.controllers import nearest_unpassed_obs as oracle_select


@pytest.fixture(autouse=True, params=[False, True], ids=["nofma", "fma"])
def _fp_variant(request):
    """Run every check on both CPU builds: without FMA contraction and with it (nvcc contracts,
    and a rounding-path bug once hid behind that difference)."""
    hostsim_util.use_fma(request.param)
    yield
    hostsim_util.use_fma(False)


def test_reference_fixtures_cbfqp():
    for tag, d in _load("ref_cbfqp.npz").items():
        spec = _spec_from_tag(tag)
        p, _ = resolve_params(spec, "cbf_qp", lib=hostsim())
        M = d["A"].shape[1]
        obs = np.nan_to_num(d["OBS"][:, :M].copy(), nan=0.0)
        if obs.shape[1] < M:                          # Manipulator2D: M is the ROW budget (25 rows per obstacle)
            obs = np.concatenate([obs, np.zeros((obs.shape[0], M - obs.shape[1], 7))], axis=1)
        nobs = np.minimum(d["NOBS"], M).astype(np.int32)
        A, b = hs_cbfqp_rows(p, d["X"], obs, nobs)
        np.testing.assert_allclose(A, d["A"], rtol=1e-11, atol=1e-11, err_msg=tag)
        np.testing.assert_allclose(b, d["B"], rtol=1e-11, atol=1e-11, err_msg=tag)
        nobs_none = np.where(d["NOBS"] == 0, -1, nobs).astype(np.int32)     # generator passed obs=None when k == 0
        U, st, _ = hs_cbfqp_solve(p, d["X"], d["UREF"], obs, nobs_none)
        assert np.array_equal(st, d["STATUS"]), tag
        ok = d["STATUS"] == 0
        np.testing.assert_allclose(U[ok], d["U"][ok], rtol=1e-8, atol=1e-9, err_msg=tag)


def test_reference_fixtures_odcbf():
    for name, d in _load("ref_odcbf.npz").items():
        p, _ = resolve_params({"model": name}, "optimal_decay_cbf_qp", lib=hostsim())
        obs = d["OBS"][:, None, :].copy()
        nobs = np.where(d["HAS"], 1, 0).astype(np.int32)
        U, om, sel, st, act = hs_odcbf_solve(p, d["X"], d["UREF"], obs, nobs)
        assert np.array_equal(st, d["STATUS"])
        np.testing.assert_allclose(U, d["U"], rtol=1e-8, atol=1e-9, err_msg=name)
        nw = d["OMEGA"].shape[1]
        np.testing.assert_allclose(om[:, :nw], d["OMEGA"], rtol=1e-8, atol=1e-9, err_msg=name)


@pytest.mark.parametrize("model,dense", [("DynamicUnicycle2D", False), ("DynamicUnicycle2D", True),
                                         ("KinematicBicycle2D", True), ("KinematicBicycle2D_C3BF", True),
                                         ("SingleIntegrator2D", True), ("DoubleIntegrator2D", True),
                                         ("Quad2D", True), ("KinematicBicycle2D_DPCBF", True),
                                         ("Unicycle2D", True), ("Unicycle2D", False),
                                         ("Manipulator2D", True), ("Manipulator2D", False)])
def test_scene_cbfqp_vs_oracle(model, dense):
    M, N = (60, 20) if model == "Manipulator2D" else (16, 160)      # (exact enumeration: C(66, 3) vertices per arm)
    sc = scenes.make_scene(model, N, M, seed=1234, dense=dense)
    p, spec = resolve_params(sc["spec"], "cbf_qp", lib=hostsim())
    U, st, act = hs_cbfqp_solve(p, sc["X"], sc["U_ref"], sc["OBS"], sc["nobs"])
    stats = check_cbfqp(spec, M, sc["X"], sc["U_ref"], sc["OBS"], sc["nobs"], U, st, act)
    print(model, dense, stats)
    assert stats["masks_compared"] + stats["infeasible"] > N // 2 or model == "Manipulator2D"


@pytest.mark.parametrize("model", ["KinematicBicycle2D_C3BF", "DynamicUnicycle2D", "KinematicBicycle2D", "Quad2D"])
def test_scene_odcbf_vs_oracle(model):
    M, N = 32, 96
    sc = scenes.make_scene(model, N, M, seed=11, dense=True, dynamic=True, optimal_decay=True)
    p, spec = resolve_params(sc["spec"], "optimal_decay_cbf_qp", lib=hostsim())
    U, om, sel, st, act = hs_odcbf_solve(p, sc["X"], sc["U_ref"], sc["OBS"], sc["nobs"])
    check_odcbf(spec, M, sc["X"], sc["U_ref"], sc["OBS"], sc["nobs"], U, om, sel, st, act)


def test_known_answers():
    """Analytic cases: far obstacle -> clip(u_ref); one active half-plane; infeasible pair."""
    p, spec = resolve_params({"model": "SingleIntegrator2D"}, "cbf_qp", lib=hostsim())
    X = np.array([[0.0, 0.0]])
    far = np.array([[[100.0, 100.0, 0.5, 0, 0, 0, 0]]])
    U, st, act = hs_cbfqp_solve(p, X, np.array([[3.0, -0.2]]), far)
    assert st[0] == 0 and np.allclose(U[0], [1.0, -0.2]) and int(act[0, 0]) == 1 << 1   # u_0 at upper bound: bit M+0
    # head-on: obstacle at (1,0) r=0.25 -> d=0.5, h = 1 - 1.01*0.25; A = 2[dx,dy] = [-2, 0], b = alpha h
    obs = np.array([[[1.0, 0.0, 0.25, 0, 0, 0, 0]]])
    h = 1.0 - 1.01 * 0.25
    U, st, act = hs_cbfqp_solve(p, X, np.array([[0.9, 0.3]]), obs)
    assert st[0] == 0 and np.allclose(U[0], [h / 2.0, 0.3], atol=1e-14) and int(act[0, 0]) == 1
    # two opposing half-planes that cannot both hold inside the box -> infeasible
    obs2 = np.array([[[0.45, 0.0, 0.25, 0, 0, 0, 0], [-0.45, 0.0, 0.25, 0, 0, 0, 0]]])
    p2, _ = resolve_params({"model": "SingleIntegrator2D", "cbf_alpha": 50.0}, "cbf_qp", lib=hostsim())
    U, st, act = hs_cbfqp_solve(p2, X, np.array([[0.0, 0.0]]), obs2)
    assert st[0] == 1 and np.all(np.abs(U[0]) <= 1.0)


def test_vacuous_and_none():
    p, spec = resolve_params({"model": "DynamicUnicycle2D"}, "cbf_qp", lib=hostsim())
    X = np.array([[0.0, 0.0, 0.3, 0.5]] * 3)
    Ur = np.array([[0.9, 0.2], [0.9, 0.2], [0.1, 0.2]])
    OBS = np.tile(np.array([0.4, 0.0, 0.2, 0, 0, 0, 0.0]), (3, 4, 1))     # would be violated if used
    nobs = np.array([-1, 0, 0], np.int32)
    U, st, act = hs_cbfqp_solve(p, X, Ur, OBS, nobs)
    assert np.array_equal(st, [0, 0, 0])
    assert np.allclose(U[0], [0.9, 0.2])          # obs None -> u_ref UNCLIPPED (cbf_qp.py:113-118)
    assert np.allclose(U[1], [0.5, 0.2])          # empty list -> box only
    assert np.allclose(U[2], [0.1, 0.2])
    # flag neither 0 nor 1 -> vacuous row for SI/DU (quirk 5)
    OBS[..., 6] = 2.0
    U, st, act = hs_cbfqp_solve(p, X[:1], Ur[2:3], OBS[:1], np.array([4], np.int32))
    assert st[0] == 0 and np.allclose(U[0], [0.1, 0.2])


def test_scene_inputs_match_oracle():
    """Vectorised nominal_input / obstacle selection of scenes.py vs the per-agent oracle restatement."""
    for model in ("SingleIntegrator2D", "DynamicUnicycle2D", "KinematicBicycle2D_C3BF", "Quad3D"):
        sc = scenes.make_scene(model, 40, 8, seed=3)
        m = make_model(sc["spec"])
        for i in range(40):
            x, g = sc["X"][i], sc["goal"][i]
            if model.startswith("KinematicBicycle2D"):
                want = m.nominal_input(x, g, 0.05, 2.0, 1.0, 1.0)
            else:
                want = m.nominal_input(x, g)
            np.testing.assert_allclose(sc["U_ref"][i], want, rtol=1e-12, atol=1e-12)
            yaw = x[2] if m.nx == 4 else (x[5] if m.nx == 12 else 0.0)
            rows, idx = oracle_select(model, x[:2], yaw, sc["scene_obs"], 8)
            k = int(sc["nobs"][i])
            assert k == len(rows)
            np.testing.assert_array_equal(sc["OBS"][i][:k], rows)
            assert np.all(sc["OBS"][i][k:] == scenes.DUMMY_OBS)


def test_active_set_fuzz_near_degenerate_rows():
    """Fuzz of the closed-form 2-variable active-set solver (scb_gi.cuh gi_solve2) on hand-made row geometries:
    SingleIntegrator2D rows are A = 2 (p - o), b = alpha h, so obstacle placement controls the half-planes directly --
    clusters of nearly parallel rows, rows that just touch the input box, wedges whose apex is the optimum (drop / re-add
    sequences, where a re-added row keeps the dual steps it already took).  Everything must equal the exact enumeration
    oracle: status, u to 1e-8, and the active mask wherever strict complementarity holds."""
    rng = np.random.default_rng(2026)
    N, M = 4000, 12
    X = rng.uniform(-1, 1, (N, 2))
    OBS = np.zeros((N, M, 7))
    kind = rng.integers(0, 4, N)
    for i in range(N):
        base = rng.uniform(-np.pi, np.pi)
        if kind[i] == 0:      # a fan of nearly parallel rows (angles within 1e-3 .. 1e-1 rad)
            ang = base + rng.normal(0, 10.0 ** rng.uniform(-3, -1), M)
        elif kind[i] == 1:    # a wedge: two clusters 60..170 degrees apart
            half = rng.uniform(0.5, 1.5)
            ang = base + np.where(rng.random(M) < 0.5, -half, half) + rng.normal(0, 0.02, M)
        elif kind[i] == 2:    # surrounded (often infeasible)
            ang = base + np.linspace(0, 2 * np.pi, M, endpoint=False) + rng.normal(0, 0.05, M)
        else:
            ang = rng.uniform(-np.pi, np.pi, M)
        r = rng.uniform(0.2, 0.6, M)
        gapd = 10.0 ** rng.uniform(-2.5, 0.3, M)                       # distance of the boundary beyond the barrier radius
        if kind[i] == 2:                                               # already inside some barriers: b < 0, opposing rows
            gapd = np.where(rng.random(M) < 0.4, -gapd * 0.2, gapd)
        d = np.sqrt(1.01) * (r + 0.25) + gapd
        OBS[i, :, 0] = X[i, 0] + d * np.cos(ang); OBS[i, :, 1] = X[i, 1] + d * np.sin(ang); OBS[i, :, 2] = r
    nobs = rng.integers(1, M + 1, N).astype(np.int32)
    head = rng.uniform(-np.pi, np.pi, N)
    Uref = np.stack([np.cos(head), np.sin(head)], 1) * rng.uniform(0.2, 1.6, (N, 1))   # some beyond the box (v_max = 1)
    p, spec = resolve_params({"model": "SingleIntegrator2D"}, "cbf_qp", lib=hostsim())
    U, st, act = hs_cbfqp_solve(p, X, Uref, OBS, nobs)
    stats = check_cbfqp(spec, M, X, Uref, OBS, nobs, U, st, act)
    assert stats["infeasible"] > 50 and stats["cbf_active"] > 1500 and stats["masks_compared"] > 2000, stats
