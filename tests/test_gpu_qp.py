"""Parity tests proper for the QP paths: the CUDA kernels, called through the C ABI
(include/scb.h), against the oracle and the reference-generated fixtures."""
import numpy as np
import pytest
import torch

from parity_util import check_cbfqp, check_odcbf
from test_oracle_pinned import _load, _spec_from_tag

pytestmark = pytest.mark.gpu


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def run_cbfqp(ctrl, sc_or_arrays):
    X, Ur, OBS, nobs = sc_or_arrays
    U, st, act = ctrl.solve(dev(X), dev(Ur), dev(OBS), None if nobs is None else dev(nobs))
    torch.cuda.synchronize()
    return U.cpu().numpy(), st.cpu().numpy(), act.cpu().numpy().view(np.uint64)


def test_library_loads_and_reports_device():
    from safe_control_b200._lib import lib
    assert lib().scb_device_count() >= 1
    assert lib().scb_version() == 210


def test_reference_fixtures_cbfqp_rows_and_solve():
    from safe_control_b200 import BatchedCBFQP
    for tag, d in _load("ref_cbfqp.npz").items():
        spec = _spec_from_tag(tag)
        M = d["A"].shape[1]
        ctrl = BatchedCBFQP(spec, num_obs=M)
        obs = np.nan_to_num(d["OBS"][:, :M].copy(), nan=0.0)
        if obs.shape[1] < M:                          # Manipulator2D: M is the ROW budget (25 rows per obstacle)
            obs = np.concatenate([obs, np.zeros((obs.shape[0], M - obs.shape[1], 7))], axis=1)
        nobs = np.minimum(d["NOBS"], M).astype(np.int32)
        A, b = ctrl.rows(dev(d["X"]), dev(obs), dev(nobs))
        np.testing.assert_allclose(A.cpu().numpy(), d["A"], rtol=1e-11, atol=1e-11, err_msg=tag)
        np.testing.assert_allclose(b.cpu().numpy(), d["B"], rtol=1e-11, atol=1e-11, err_msg=tag)
        nobs_none = np.where(d["NOBS"] == 0, -1, nobs).astype(np.int32)
        U, st, _ = run_cbfqp(ctrl, (d["X"], d["UREF"], obs, nobs_none))
        assert np.array_equal(st, d["STATUS"]), tag
        ok = d["STATUS"] == 0
        np.testing.assert_allclose(U[ok], d["U"][ok], rtol=1e-8, atol=1e-9, err_msg=tag)


def test_reference_fixtures_odcbf():
    from safe_control_b200 import BatchedOptimalDecayCBFQP
    for name, d in _load("ref_odcbf.npz").items():
        ctrl = BatchedOptimalDecayCBFQP({"model": name}, num_obs=1)
        obs = d["OBS"][:, None, :].copy()
        nobs = np.where(d["HAS"], 1, 0).astype(np.int32)
        U, om, sel, st, act = ctrl.solve(dev(d["X"]), dev(d["UREF"]), dev(obs), dev(nobs))
        assert np.array_equal(st.cpu().numpy(), d["STATUS"])
        np.testing.assert_allclose(U.cpu().numpy(), d["U"], rtol=1e-8, atol=1e-9, err_msg=name)
        nw = d["OMEGA"].shape[1]
        np.testing.assert_allclose(om.cpu().numpy()[:, :nw], d["OMEGA"], rtol=1e-8, atol=1e-9, err_msg=name)


@pytest.mark.parametrize("model,M,dense", [
    ("DynamicUnicycle2D", 16, False),      # BASELINE config 2
    ("DynamicUnicycle2D", 16, True),
    ("KinematicBicycle2D", 16, True),
    ("KinematicBicycle2D_C3BF", 32, False),
    ("SingleIntegrator2D", 2, True),       # config-1 sized rows
    ("DynamicUnicycle2D", 40, True),       # RPL=2 path
    ("DynamicUnicycle2D", 100, False),     # RPL=4 path
    ("DoubleIntegrator2D", 16, True),      # SURVEY 8f-2 barrier families
    ("Quad2D", 16, True),
    ("KinematicBicycle2D_DPCBF", 16, True),
    ("KinematicBicycle2D_DPCBF", 8, False),
    ("Unicycle2D", 16, True),              # third fixture set (sigma-shaped barrier, rel. degree 1)
    ("Manipulator2D", 20, True),           # 3-input QP, 25 link-circle rows per obstacle: RPL = 1 / 2 / 4 paths
    ("Manipulator2D", 50, True),
    ("Manipulator2D", 100, False),
])
def test_scene_cbfqp_vs_oracle(model, M, dense):
    from safe_control_b200 import BatchedCBFQP, scenes
    N = 1024 if (model, M, dense) == ("DynamicUnicycle2D", 16, False) else 160
    sc = scenes.make_scene(model, N, M, seed=1234, dense=dense)
    ctrl = BatchedCBFQP(sc["spec"], num_obs=M)
    U, st, act = run_cbfqp(ctrl, (sc["X"], sc["U_ref"], sc["OBS"], sc["nobs"]))
    sample = np.arange(0, N, 10) if model == "Manipulator2D" else None      # (oracle: C(M + 6, 3) vertices per arm)
    stats = check_cbfqp(ctrl.robot_spec, M, sc["X"], sc["U_ref"], sc["OBS"], sc["nobs"], U, st, act, sample=sample)
    print(model, M, dense, stats)
    if model == "Manipulator2D":                                            # whole batch: inside the box, rows satisfied
        lib = ctrl.params
        assert (np.abs(U[st == 0]) <= lib.u_ub[0] + 1e-9).all()


@pytest.mark.parametrize("model", ["KinematicBicycle2D_C3BF", "DynamicUnicycle2D", "KinematicBicycle2D", "Quad2D"])
def test_scene_odcbf_vs_oracle(model):
    from safe_control_b200 import BatchedOptimalDecayCBFQP, scenes
    M, N = 32, 256
    sc = scenes.make_scene(model, N, M, seed=1234, dense=False, dynamic=True, optimal_decay=True)
    ctrl = BatchedOptimalDecayCBFQP(sc["spec"], num_obs=M)
    out = ctrl.solve(dev(sc["X"]), dev(sc["U_ref"]), dev(sc["OBS"]), dev(sc["nobs"]))
    U, om, sel, st, act = [t.cpu().numpy() for t in out]
    stats = check_odcbf(ctrl.robot_spec, M, sc["X"], sc["U_ref"], sc["OBS"], sc["nobs"], U, om, sel, st,
                        act.view(np.uint64))
    print(model, stats)


def test_edge_cases():
    from safe_control_b200 import BatchedCBFQP
    ctrl = BatchedCBFQP({"model": "DynamicUnicycle2D"}, num_obs=4)
    # empty batch
    U, st, act = ctrl.solve(torch.empty((0, 4), dtype=torch.float64, device="cuda"),
                            torch.empty((0, 2), dtype=torch.float64, device="cuda"),
                            torch.empty((0, 4, 7), dtype=torch.float64, device="cuda"))
    assert U.shape == (0, 2)
    X = np.array([[0.0, 0.0, 0.3, 0.5]] * 3)
    Ur = np.array([[0.9, 0.2], [0.9, 0.2], [0.1, 0.2]])
    OBS = np.tile(np.array([0.4, 0.0, 0.2, 0, 0, 0, 0.0]), (3, 4, 1))
    nobs = np.array([-1, 0, 0], np.int32)
    U, st, act = run_cbfqp(ctrl, (X, Ur, OBS, nobs))
    assert np.array_equal(st, [0, 0, 0])
    assert np.allclose(U[0], [0.9, 0.2]) and np.allclose(U[1], [0.5, 0.2]) and np.allclose(U[2], [0.1, 0.2])
    assert int(act[1, 0]) == 1 << 4      # u_0 upper bound: bit M + 0
    # shared obstacle list [M, 7] == per-agent copies
    sh = np.array([[3.0, 0.5, 0.3, 0, 0, 0, 0], [1.5, -2.0, 0.4, 0, 0, 0, 0], [-2, 0, 0.2, 0, 0, 0, 0], [1000, 1000, 0, 0, 0, 0, 0.0]])
    Xs = np.array([[0.0, 0.0, 0.1 * i, 0.8] for i in range(64)])
    Urs = np.tile([0.3, 0.1], (64, 1))
    U1, s1, a1 = run_cbfqp(ctrl, (Xs, Urs, sh, None))
    U2, s2, a2 = run_cbfqp(ctrl, (Xs, Urs, np.tile(sh, (64, 1, 1)), None))
    assert np.array_equal(U1, U2) and np.array_equal(s1, s2) and np.array_equal(a1, a2)
    # too many obstacle slots -> error code, not a crash
    from safe_control_b200._lib import ScbError
    big = BatchedCBFQP({"model": "DynamicUnicycle2D"}, num_obs=500)
    with pytest.raises(ScbError):
        big.solve(dev(Xs), dev(Urs), dev(np.zeros((64, 500, 7))))
    # Quad3D has no continuous-time barrier in the reference either
    from safe_control_b200 import NotCompatibleError
    with pytest.raises(NotCompatibleError):
        BatchedCBFQP({"model": "Quad3D"}, num_obs=4)


def test_full_size_properties():
    """Size-independent properties at large N: returned inputs are feasible for the assembled
    rows, inside the box, the projection is idempotent, and the large-batch launch geometry
    (8 lanes per QP) agrees with the small-batch one (32 lanes per QP) to rounding (the two template
    instantiations may contract FMAs differently, so not bit for bit)."""
    from safe_control_b200 import BatchedCBFQP, scenes
    M, N = 16, 1 << 18
    base = scenes.make_scene("DynamicUnicycle2D", 4096, M, seed=99, dense=True)
    rep = N // 4096
    ctrl = BatchedCBFQP(base["spec"], num_obs=M)
    X = dev(np.tile(base["X"], (rep, 1))); Ur = dev(np.tile(base["U_ref"], (rep, 1)))
    OBS = dev(np.tile(base["OBS"], (rep, 1, 1))); nobs = dev(np.tile(base["nobs"], rep))
    U, st, act = ctrl.solve(X, Ur, OBS, nobs)
    Us, sts, acts = ctrl.solve(X[:4096], Ur[:4096], OBS[:4096], nobs[:4096])
    assert (U.view(rep, 4096, 2) - Us[None]).abs().max().item() < 1e-11
    assert torch.equal(st.view(rep, 4096), sts.expand(rep, -1))
    assert (act.view(rep, 4096, -1) != acts[None]).float().mean().item() < 1e-4   # ties only
    assert torch.equal(U.view(rep, 4096, 2)[0], U.view(rep, 4096, 2)[-1])         # same geometry: deterministic
    ok = st == 0
    assert ok.float().mean() > 0.3
    A, b = ctrl.rows(X, OBS, nobs)
    slack = (A * U[:, None, :]).sum(-1) + b
    scale = 1.0 + b.abs() + A.abs().sum(-1)
    assert (slack[ok] >= -1e-9 * scale[ok]).all()
    lb = torch.tensor(list(ctrl.params.u_lb)[:2], device="cuda", dtype=torch.float64)
    ub = torch.tensor(list(ctrl.params.u_ub)[:2], device="cuda", dtype=torch.float64)
    assert ((U >= lb - 1e-12) & (U <= ub + 1e-12)).all()
    U2, st2, _ = ctrl.solve(X, U, OBS, nobs)          # idempotence of the projection
    assert torch.equal(st2[ok], st[ok])
    assert (U2[ok] - U[ok]).abs().max().item() < 1e-9


def test_config4_full_size_properties():
    """BASELINE config 4 at its full size (8192 KinematicBicycle2D_C3BF agents, optimal_decay_cbf_qp, 32 moving
    obstacles): the selected row is the nearest obstacle, u is inside the box, the decay variable is feasible
    (the relaxed row holds at the returned (u, omega)), and without obstacles the filter is the identity
    (omega = 1, u = clip(u_ref)); a sample is checked against the oracle in test_scene_odcbf_vs_oracle."""
    from safe_control_b200 import BatchedOptimalDecayCBFQP, BatchedCBFQP, scenes
    M, N = 32, 8192
    sc = scenes.make_scene("KinematicBicycle2D_C3BF", N, M, seed=1234, dynamic=True, optimal_decay=True)
    ctrl = BatchedOptimalDecayCBFQP(sc["spec"], num_obs=M)
    X, Ur, OBS, nobs = dev(sc["X"]), dev(sc["U_ref"]), dev(sc["OBS"]), dev(sc["nobs"])
    U, om, sel, st, act = ctrl.solve(X, Ur, OBS, nobs)
    ok = st == 0
    assert ok.float().mean() > 0.95
    d = ((OBS[:, :, :2] - X[:, None, :2]) ** 2).sum(-1)
    d = torch.where(torch.arange(M, device="cuda")[None] < nobs[:, None], d, torch.full_like(d, float("inf")))
    has = nobs > 0
    assert torch.equal(sel[has].long(), d.argmin(1)[has])
    lb = torch.tensor(list(ctrl.params.u_lb)[:2], device="cuda", dtype=torch.float64)
    ub = torch.tensor(list(ctrl.params.u_ub)[:2], device="cuda", dtype=torch.float64)
    assert ((U >= lb - 1e-9) & (U <= ub + 1e-9)).all()       # (4-variable active-set solve: bounds hold to rounding)
    # the relaxed row  A u + dh.f + alpha_od h omega >= 0  (optimal_decay_cbf_qp.py:98-103) at the returned (u, omega):
    # A, dh.f and h are recovered from the cbf_qp rows of the same model at two gains (b = dh.f + alpha h, cbf_qp.py:165)
    qa = BatchedCBFQP(dict(sc["spec"], cbf_alpha=1.5), num_obs=M)
    qb = BatchedCBFQP(dict(sc["spec"], cbf_alpha=0.5), num_obs=M)
    A, b15 = qa.rows(X, OBS, nobs); _, b05 = qb.rows(X, OBS, nobs)
    ar = torch.arange(N, device="cuda"); idx = sel.clamp(min=0).long()
    a_sel, h_sel = A[ar, idx], (b15 - b05)[ar, idx]
    lf_sel = b05[ar, idx] - 0.5 * h_sel
    row = (a_sel * U).sum(-1) + lf_sel + ctrl.params.alpha * h_sel * om[:, 0]
    scale = 1.0 + a_sel.abs().sum(-1) + lf_sel.abs() + h_sel.abs()
    m = has & ok
    assert (row[m] >= -1e-7 * scale[m]).all(), float((row[m] / scale[m]).min())
    # without any obstacle the filter is the identity: omega = 1, u = clip(u_ref)
    U2, om2, sel2, st2, _ = ctrl.solve(X, Ur, OBS, torch.zeros_like(nobs))
    assert (st2 == 0).all() and (sel2 == -1).all() and (om2[:, 0] - 1.0).abs().max().item() < 1e-9
    assert (U2 - torch.minimum(torch.maximum(Ur, lb), ub)).abs().max().item() < 1e-9


def test_host_pointer_path_matches_device_path():
    from safe_control_b200 import BatchedCBFQP, HostContext, scenes
    M, N = 16, 1024
    sc = scenes.make_scene("DynamicUnicycle2D", N, M, seed=5)
    ctrl = BatchedCBFQP(sc["spec"], num_obs=M)
    Ud, std, actd = run_cbfqp(ctrl, (sc["X"], sc["U_ref"], sc["OBS"], sc["nobs"]))
    ctx = HostContext(0)
    U, st, act = ctx.cbfqp_solve(ctrl.params, M, sc["X"], sc["U_ref"], sc["OBS"], sc["nobs"])
    assert np.array_equal(U, Ud) and np.array_equal(st, std) and np.array_equal(act, actd)
    assert ctx.launches == 1
    ctx.close()


def test_every_launch_geometry_is_instantiated():
    """Batch sizes on both sides of every lanes-per-QP threshold, for both QP controllers (a missing template
    instantiation shows up as SCB_ERR_TOO_LARGE)."""
    from safe_control_b200 import BatchedCBFQP, BatchedOptimalDecayCBFQP, scenes
    for M in (16, 32, 40, 100):
        for N in (64, 4096, 40000):
            sc = scenes.make_scene("DynamicUnicycle2D", 64, M, seed=2)
            rep = -(-N // 64)
            tile = lambda a: dev(np.tile(a, (rep,) + (1,) * (a.ndim - 1))[:N])
            X, Ur, OBS, nobs = tile(sc["X"]), tile(sc["U_ref"]), tile(sc["OBS"]), tile(sc["nobs"])
            U, st, _ = BatchedCBFQP(sc["spec"], num_obs=M).solve(X, Ur, OBS, nobs)
            U2, om, sel, st2, _ = BatchedOptimalDecayCBFQP(sc["spec"], num_obs=M).solve(X, Ur, OBS, nobs)
            torch.cuda.synchronize()
            assert torch.equal(st[:64], st[-64:] if N % 64 == 0 else st[:64])
