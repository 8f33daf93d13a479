"""Exact solver for tiny strictly convex QPs (TEST INFRASTRUCTURE -- oracle).

    min 1/2 x'Px + q'x   s.t.  G x <= h          (P symmetric positive definite)

The reference hands these problems to cvxpy -> GUROBI
(position_control/cbf_qp.py:190, optimal_decay_cbf_qp.py:156); neither is
installed here.  Because P > 0 the optimum is unique,
so any exact method yields *the* answer GUROBI converges to (within its 1e-6
tolerances).  We enumerate working sets W of size 0..n, solve the
equality-constrained KKT system on each, and keep the unique candidate that is
primal feasible and dual feasible.  Cost is O(C(m, <=n)) -- fine for n <= 4,
m <= ~70, which covers every QP on the hot path (a1, a2).

Returns the solution, its multipliers and the active working set, which is
what "active-constraint indices" means in this repo (SURVEY.md section 8a,
quirk 9: the reference exposes none; the oracle defines them).
"""
from itertools import combinations

import numpy as np

OPTIMAL, INFEASIBLE = 0, 1


def solve_qp_exact(P, q, G, h, tol=1e-9):
    """-> dict(x, status, lam, active, obj, gap)

    status   0 optimal / 1 infeasible (x is None)
    lam      (m,) multipliers of G x <= h (>= 0), zero off the working set
    active   (m,) bool, the optimal working set (lam > 0 rows; for degenerate
             problems the smallest such set found)
    gap      strict-complementarity margin: min(min active lam, min inactive
             slack); tests only compare active sets bit-exactly when gap is
             comfortably positive.
    """
    P = np.asarray(P, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64).reshape(-1)
    G = np.asarray(G, dtype=np.float64).reshape(-1, q.size)
    h = np.asarray(h, dtype=np.float64).reshape(-1)
    n, m = q.size, h.size
    scale = max(1.0, float(np.max(np.abs(h))) if m else 1.0)
    rown = np.linalg.norm(G, axis=1)
    # rows with a zero normal are constants: 0 <= h_i
    const_rows = rown == 0.0
    if np.any(h[const_rows] < -tol * scale):
        return dict(x=None, status=INFEASIBLE, lam=None, active=None, obj=None, gap=0.0)
    rows = [i for i in range(m) if not const_rows[i]]
    Pinv = np.linalg.inv(P)
    best = None
    for k in range(0, n + 1):
        for W in combinations(rows, k):
            W = list(W)
            if k == 0:
                x = -Pinv @ q
                lam_w = np.zeros(0)
            else:
                Gw = G[W]
                S = Gw @ Pinv @ Gw.T
                if np.linalg.cond(S) > 1e13:
                    continue
                # stationarity: P x + q + Gw' lam = 0, Gw x = h_w
                rhs = -(h[W] + Gw @ Pinv @ q)
                lam_w = np.linalg.solve(S, rhs)
                x = -Pinv @ (q + Gw.T @ lam_w)
            if k and np.any(lam_w < -tol * scale):
                continue
            viol = G @ x - h
            feas_tol = tol * np.maximum(1.0, rown * max(1.0, float(np.max(np.abs(x)))) + np.abs(h))
            if np.any(viol > feas_tol):
                continue
            obj = 0.5 * x @ P @ x + q @ x
            if best is None or obj < best[0] - 1e-14 * max(1.0, abs(obj)):
                best = (obj, x, W, lam_w)
        if best is not None and k >= len(best[2]):
            # a KKT point of a strictly convex QP is THE optimum; stop early
            break
    if best is None:
        return dict(x=None, status=INFEASIBLE, lam=None, active=None, obj=None, gap=0.0)
    obj, x, W, lam_w = best
    lam = np.zeros(m)
    active = np.zeros(m, dtype=bool)
    for i, l in zip(W, lam_w):
        lam[i] = l
        active[i] = l > 0.0
    slack = h - G @ x
    inactive = ~active & ~const_rows
    lam_min = float(np.min(lam[active] / np.maximum(rown[active], 1e-300))) if active.any() else np.inf
    sl_min = float(np.min(slack[inactive] / np.maximum(rown[inactive], 1e-300))) if inactive.any() else np.inf
    return dict(x=x, status=OPTIMAL, lam=lam, active=active, obj=float(obj),
                gap=min(lam_min, sl_min))


def kkt_residual(P, q, G, h, x, lam):
    """Max-norm KKT residuals (stationarity, primal, dual, complementarity)."""
    P = np.asarray(P, float); q = np.asarray(q, float).reshape(-1)
    G = np.asarray(G, float).reshape(-1, q.size); h = np.asarray(h, float).reshape(-1)
    stat = P @ x + q + G.T @ lam
    prim = np.maximum(G @ x - h, 0.0)
    dual = np.maximum(-lam, 0.0)
    comp = lam * (h - G @ x)
    f = lambda v: float(np.max(np.abs(v))) if v.size else 0.0
    return dict(stationarity=f(stat), primal=f(prim), dual=f(dual), complementarity=f(comp))
