"""Shared parity checks: any backend (CUDA C-ABI or CPU host-sim) vs the oracle."""
import numpy as np

from oracle.controllers import OracleCBFQP, OracleOptimalDecayCBFQP

U_TOL = 1e-8          # ||u - u*||_inf <= U_TOL * max(1, ||u*||_inf)     (SURVEY 8c "stated tolerances")
GAP_MIN = 1e-7        # active masks compared bit-exactly when strict complementarity gap > GAP_MIN


def mask_from_bool(active, words):
    out = np.zeros(words, dtype=np.uint64)
    for j in np.nonzero(active)[0]:
        out[j >> 6] |= np.uint64(1) << np.uint64(j & 63)
    return out


def check_cbfqp(spec, M, X, Uref, OBS, nobs, U, status, active, sample=None):
    """-> stats dict; asserts parity."""
    ctrl = OracleCBFQP(spec, num_obs=M)
    N = X.shape[0]
    idx = range(N) if sample is None else sample
    words = (M + 2 * ctrl.model.nu + 63) // 64
    n_inf = n_act = n_box = n_cmp = 0
    for i in idx:
        k = M if nobs is None else int(nobs[i])
        obs = None if k < 0 else OBS[i][:k]
        u, info = ctrl.solve(X[i], Uref[i], obs)
        assert info["status"] == status[i], f"agent {i}: status {status[i]} vs oracle {info['status']}"
        if info["status"] != 0:
            n_inf += 1
            continue
        err = np.max(np.abs(u - U[i]))
        assert err <= U_TOL * max(1.0, np.max(np.abs(u))), f"agent {i}: |du|={err:.3e} u*={u} got={U[i]}"
        if k < 0:
            continue
        act = info["active"]
        n_act += bool(act[:M].any()); n_box += bool(act[M:].any())
        if active is not None and info["gap"] > GAP_MIN:
            n_cmp += 1
            want = mask_from_bool(act, words)
            got = np.asarray(active[i]).view(np.uint64).reshape(-1)
            assert np.array_equal(want, got), f"agent {i}: active {got} vs oracle {want} (gap {info['gap']:.2e})"
    return dict(n=len(list(idx)), infeasible=n_inf, cbf_active=n_act, box_active=n_box, masks_compared=n_cmp)


def check_odcbf(spec, M, X, Uref, OBS, nobs, U, omega, sel, status, active, sample=None):
    ctrl = OracleOptimalDecayCBFQP(spec)
    N = X.shape[0]
    idx = range(N) if sample is None else sample
    n_cmp = n_act = 0
    for i in idx:
        k = M if nobs is None else max(int(nobs[i]), 0)
        if k > 0:
            d = np.linalg.norm(OBS[i][:k, :2] - X[i][:2], axis=1)
            j = int(np.argmin(d))
            assert sel[i] == j, f"agent {i}: sel {sel[i]} vs nearest {j}"
            obs = OBS[i][j]
        else:
            assert sel[i] == -1
            obs = None
        u, om, info = ctrl.solve(X[i], Uref[i], obs)
        assert info["status"] == status[i]
        if info["status"] != 0:
            continue
        assert np.max(np.abs(u - U[i])) <= U_TOL * max(1.0, np.max(np.abs(u))), (i, u, U[i])
        nw = om.size
        assert np.max(np.abs(om - omega[i][:nw])) <= U_TOL * max(1.0, np.max(np.abs(om))), (i, om, omega[i])
        n_act += bool(info["active"][0])
        if active is not None and info["gap"] > GAP_MIN:
            n_cmp += 1
            want = mask_from_bool(info["active"], 1)[0]
            assert np.uint64(active[i]) == want, f"agent {i}: active {active[i]} vs {want}"
    return dict(n=len(list(idx)), cbf_active=n_act, masks_compared=n_cmp)


def check_mpc(spec, M, H, X, goal, u_prev, OBS, nobs, out, sample=None, u0_tol=1e-4, min_agree=0.9, second_solver=0,
              optimal_decay=False, sum_rterms=False):
    """MPC parity: (a) every 'optimal' answer must be a KKT point of the ORACLE's restated NLP
    (independent derivatives: torch.autograd on oracle/mpc_cbf.py: stationarity <= 1e-4 with non-negative least-squares
    multipliers, complementarity max lam_i g_i <= 1e-5), feasible to 1e-7; (b) u0 must agree
    with the oracle's own SLSQP solve within u0_tol (box-normalised).  The NLP is non-convex, so a miss is classified by
    the COST of the two points (both feasible KKT points of the same NLP from the same cold start):
      same_cost   |J - J_slsqp| <= 1e-6 max(1, |J_slsqp|): same optimum value, flat direction / SLSQP stopped early
      better      the kernel's point is cheaper: SLSQP sits in a worse basin (or stopped early)
      worse       the kernel's point is more expensive: a different, worse local optimum
    and the assertion is that agree + same_cost + better >= min_agree of the compared cases (the kernel's optimum is at
    least as good as the oracle solver's), with every count reported.
    (c) when out['active'] is present: the kernel's active mask must equal, bit for bit, the oracle's active set at the
    oracle's own solution (OracleMPCCBF.active_set) on every agreeing case whose strict-complementarity gap is >= 1.
    second_solver > 0: on that many agreeing cases also solve with scipy trust-constr warm-started at the cold start
    ('two solvers agree to 1e-6') and require the same u0."""
    import warnings
    import torch
    from oracle.mpc_cbf import OracleMPCCBF
    warnings.filterwarnings("ignore")
    if optimal_decay:                                        # optimal_decay_mpc_cbf: [u, omega1, omega2] per stage, 5 obstacle slots
        from oracle.mpc_cbf import OracleODMPCCBF
        o = OracleODMPCCBF(spec, sum_rterms=sum_rterms, horizon=H)
        assert M == 5
    else:
        o = OracleMPCCBF(spec, num_obs=M, horizon=H)
    N = X.shape[0]
    idx = list(range(N) if sample is None else sample)
    rng = getattr(o, "u_range", None)
    if rng is None:
        rng = o.u_ub - o.u_lb
    n_ok = n_cmp = n_agree = n_same = n_better = n_worse = n_mask = n_second = n_ofail = n_second_agree = 0
    worst_kkt = worst_du = 0.0
    for i in idx:
        k = M if nobs is None else max(int(nobs[i]), 0)
        obs = OBS[i][:k]
        if out["status"][i] != 0:
            continue
        n_ok += 1
        kk, gmin, comp = o.kkt_error(X[i], goal[i], u_prev[i], obs, out["pred_u"][i])
        scale = 1.0 + float(np.abs(out["pred_u"][i]).max())
        assert gmin >= -1e-7, f"agent {i}: infeasible point reported optimal (min g = {gmin:.2e})"
        assert kk <= 1e-4 * scale, f"agent {i}: not a KKT point of the oracle NLP (residual {kk:.2e})"
        assert comp <= 1e-5 * scale, f"agent {i}: complementarity violated (max lam g = {comp:.2e})"
        worst_kkt = max(worst_kkt, kk)
        u, info = o.solve(X[i], goal[i], u_prev[i], obs)
        if not info["success"] or info["cbf_min"] < -1e-6:
            n_ofail += 1                                  # the oracle's SLSQP did not converge: nothing to compare u0 with
            continue
        n_cmp += 1
        du = float(np.max(np.abs(out["U"][i] - u) / rng))
        Jo = info["fun"]
        Jm = float(o.condensed(X[i], goal[i], u_prev[i], obs, torch.tensor(out["pred_u"][i].reshape(-1)))[0])
        if du <= u0_tol:
            n_agree += 1; worst_du = max(worst_du, du)
            if "active" in out and out["active"] is not None:
                act, gap = o.active_set(X[i], goal[i], u_prev[i], obs, info["u_pred"])
                # (the mask covers the whole horizon: compare it only where the two input TRAJECTORIES coincide -- with
                #  flat directions, e.g. optimal decay without an input cost, u0 can agree while later stages differ)
                dtraj = float(np.max(np.abs(np.asarray(out["pred_u"][i]) - info["u_pred"]) / rng))
                if gap >= 1.0 and dtraj <= 10 * u0_tol:
                    n_mask += 1
                    words = np.asarray(out["active"][i]).view(np.uint64).reshape(-1)
                    want = np.zeros(words.size, dtype=np.uint64)
                    for r in np.nonzero(act)[0]:
                        b = o.kernel_bit_of_row(int(r))
                        want[b >> 6] |= np.uint64(1) << np.uint64(b & 63)
                    assert np.array_equal(want, words), f"agent {i}: MPC active mask {words} vs oracle {want} (gap {gap:.2f})"
            if n_second < second_solver:
                n_second += 1
                u2, info2 = o.solve(X[i], goal[i], u_prev[i], obs, method="trust-constr")
                du2 = float(np.max(np.abs(u2 - u) / rng))
                # trust-constr stops on gtol / xtol, a few 1e-4 from the optimum in a flat direction is what it delivers:
                # the two solvers must land in the same basin (2e-3 box-normalised, or a different cost = a different basin)
                assert du2 <= 2e-3 or abs(info2["fun"] - Jo) > 1e-6 * max(1.0, abs(Jo)), \
                    f"agent {i}: trust-constr and SLSQP disagree at equal cost (du {du2:.2e})"
                if du2 <= 2e-3:
                    n_second_agree += 1
                    assert float(np.max(np.abs(out["U"][i] - u2) / rng)) <= 2.5e-3, f"agent {i}: kernel vs trust-constr"
        elif abs(Jm - Jo) <= 1e-6 * max(1.0, abs(Jo)):
            n_same += 1
        elif Jm < Jo:
            n_better += 1
        else:
            n_worse += 1
    stats = dict(n=len(idx), optimal=n_ok, compared=n_cmp, agree=n_agree, same_cost=n_same, better=n_better, worse=n_worse,
                 other_local=n_same + n_better + n_worse, masks_compared=n_mask, second_solver=n_second, second_solver_agree=n_second_agree,
                 oracle_failed=n_ofail, worst_kkt=worst_kkt, worst_du=worst_du)
    assert n_cmp == 0 or (n_agree + n_same + n_better) >= min_agree * n_cmp, stats
    return stats


def check_mpc_parallel(spec, M, H, X, goal, u_prev, OBS, nobs, out, sample, min_agree=0.9, procs=None, timeout_s=900, **kw):
    """check_mpc over `sample` on all host cores (the oracle costs 2-6 s per agent at config-5 shapes): every per-agent
    assertion of check_mpc still fires (a worker's AssertionError fails the test with its message); the agreement fraction
    is asserted on the merged counts.  The workers are separate python processes (tests/mpc_check_worker.py) fed through
    an .npz file -- never a fork of this process: the parent has initialised CUDA and torch's thread pools, and a forked
    child's autograd engine hangs (observed: a whole GPU lease lost)."""
    import json
    import os
    import subprocess
    import sys
    import tempfile
    sample = [int(i) for i in sample]
    procs = min(procs or os.cpu_count() or 1, max(1, len(sample)))
    if procs <= 1:
        return check_mpc(spec, M, H, X, goal, u_prev, OBS, nobs, out, sample=sample, min_agree=min_agree, **kw)
    idx = np.array(sample)
    sub = lambda a: np.ascontiguousarray(np.asarray(a)[idx])          # ship only the sampled rows
    here = os.path.dirname(os.path.abspath(__file__))
    with tempfile.TemporaryDirectory() as tmp:
        arrays = dict(X=sub(X), goal=sub(goal), u_prev=sub(u_prev), OBS=sub(OBS),
                      nobs=sub(nobs) if nobs is not None else np.full(len(sample), M, np.int32))
        for k in ("U", "status", "pred_u", "active"):
            if k in out and out[k] is not None:
                arrays["out_" + k] = sub(out[k])
        np.savez(os.path.join(tmp, "in.npz"), **arrays)
        with open(os.path.join(tmp, "meta.json"), "w") as f:
            json.dump(dict(spec={k: v for k, v in spec.items() if isinstance(v, (int, float, str, bool))}, M=M, H=H, kw=kw), f)
        ps = []
        for r in range(procs):
            env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
            ps.append(subprocess.Popen([sys.executable, os.path.join(here, "mpc_check_worker.py"), tmp, str(r), str(procs)],
                                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
        parts = []
        for r, pr in enumerate(ps):
            try:
                so, se = pr.communicate(timeout=timeout_s)
            except subprocess.TimeoutExpired:
                for q in ps:
                    q.kill()
                raise AssertionError(f"oracle worker {r} exceeded {timeout_s} s")
            assert pr.returncode == 0, f"oracle worker {r} failed:\n{se[-2000:]}"
            parts.append(json.loads(so.strip().splitlines()[-1]))
    stats = {k: (max(p[k] for p in parts) if k.startswith("worst") else sum(p[k] for p in parts)) for k in parts[0]}
    n_cmp = stats["compared"]
    assert n_cmp == 0 or (stats["agree"] + stats["same_cost"] + stats["better"]) >= min_agree * n_cmp, stats
    return stats
