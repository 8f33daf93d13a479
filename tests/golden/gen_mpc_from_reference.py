#!/usr/bin/env python
"""MPC-CBF problem-statement fixtures from the REFERENCE'S OWN position_control/mpc_cbf.py.

    python tests/golden/gen_mpc_from_reference.py        # writes tests/golden/ref_mpc_statement.npz

The reference's MPCCBF is imported UNMODIFIED and constructed at seeded probe points through
oracle/refshim (casadi -> numeric stand-in, do_mpc -> probing stand-in that records every call).
Per (model, probe) it stores what the reference hands to do-mpc: x_next of set_rhs, the 'cost'
expression, the CBF constraint value of every obstacle slot, input / state bounds, the rterm
weights, horizon / t_step / n_robust and the tvp values its tvp_fun produces (goal padded with
zeros, dummy obstacle rows, alphas).  tests/test_oracle_pinned.py compares oracle/mpc_cbf.py's NLP
ingredients with them.  Nothing is solved here: do-mpc's own assembly and IPOPT stay unpinned.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import refshim  # noqa: E402

refshim.install()
from oracle.refshim import fake_do_mpc as fdm  # noqa: E402

from gen_from_reference import Facade, rand_state, rand_input, rand_circle, rand_superellipsoid  # noqa: E402
from safe_control.position_control.mpc_cbf import MPCCBF  # noqa: E402

CASES = [("SingleIntegrator2D", {}), ("DynamicUnicycle2D", {}), ("KinematicBicycle2D", {}), ("Quad3D", {}),
         ("KinematicBicycle2D_C3BF", {}), ("DynamicUnicycle2D", {"mpc_horizon": 8, "mpc_cbf_alpha1": 0.3, "a_max": 1.0}),
         ("DoubleIntegrator2D", {}), ("Quad2D", {}), ("KinematicBicycle2D_DPCBF", {})]


def probe(name, spec, x, u, goal, obs, M):
    """-> dict of what the reference's MPCCBF states at (x, u; goal, obs)."""
    nx = x.size
    fdm.PROBE["_x"] = {"x": x}; fdm.PROBE["_u"] = {"u": u}; fdm.PROBE["_tvp"] = {}
    fac = Facade(name, x, spec)
    ctrl = MPCCBF(fac, fac.robot_spec, num_obs=M)
    ctrl.update_tvp(goal, obs)                                   # what solve_control_problem does first (:366-371)
    tvp = ctrl.mpc.tvp_fun(0.0)
    fdm.PROBE["_tvp"] = dict(tvp)                                # second pass: the same problem at the actual tvp values
    fac = Facade(name, x, spec)
    ctrl = MPCCBF(fac, fac.robot_spec, num_obs=M)
    mdl, mpc = fdm.LAST["model"], fdm.LAST["mpc"]
    nu = u.size
    lb_u = mpc.bounds.get(("lower", "_u", "u"), np.full(nu, -np.inf)); ub_u = mpc.bounds.get(("upper", "_u", "u"), np.full(nu, np.inf))
    lb_x = np.full(nx, -np.inf); ub_x = np.full(nx, np.inf)
    for key, v in mpc.bounds.items():
        if key[1] == "_x" and len(key) == 4:
            (lb_x if key[0] == "lower" else ub_x)[key[3]] = float(v)
    alphas = np.array([float(tvp.get("alpha", np.array(np.nan)).reshape(-1)[0]),
                       float(tvp.get("alpha1", np.array(np.nan)).reshape(-1)[0]),
                       float(tvp.get("alpha2", np.array(np.nan)).reshape(-1)[0])])
    return dict(x_next=mdl.rhs["x"].reshape(-1), cost=float(mdl.aux["cost"].reshape(-1)[0]),
                cbf=np.array([-mpc.nl_cons[f"cbf_{i}"][0] for i in range(M)]),
                cons_ub=np.array([mpc.nl_cons[f"cbf_{i}"][1] for i in range(M)], float),
                lb_u=lb_u, ub_u=ub_u, lb_x=lb_x, ub_x=ub_x, R=np.asarray(mpc.rterm["u"], float).reshape(-1),
                horizon=int(mpc.params["n_horizon"]), t_step=float(mpc.params["t_step"]), n_robust=int(mpc.params["n_robust"]),
                tvp_goal=tvp["goal"].reshape(-1), tvp_obs=tvp["obs"].reshape(M, 7), alphas=alphas,
                lterm_is_mterm=float(np.array_equal(mpc.objective["lterm"], mpc.objective["mterm"])))


CASES2 = [("Unicycle2D", {}), ("Unicycle2D", {"mpc_horizon": 6, "mpc_cbf_alpha": 0.2, "w_max": 1.0})]   # second file
CASES3 = [("VTOL2D", {}), ("VTOL2D", {"mpc_cbf_alpha1": 0.35, "mpc_cbf_alpha2": 0.35, "v_max": 20.0, "pitch_max": 20.0})]   # third file


def main(cases=CASES, seed=20261018, fname="ref_mpc_statement.npz"):
    rng = np.random.default_rng(seed)
    M, n = 5, 24
    flat = {}
    for name, spec in cases:
        tag = name + ("" if not spec else "+" + ",".join(f"{k}={v}" for k, v in spec.items()))
        rows = {}
        for i in range(n):
            x = rand_state(rng, name); u = rand_input(rng, spec, name)
            k = int(rng.integers(0, M + 2))
            dyn = name.endswith("C3BF") or name.endswith("DPCBF")
            obs = [rand_circle(rng, x, dyn) for _ in range(k)]
            if k and name in ("SingleIntegrator2D", "DynamicUnicycle2D", "DoubleIntegrator2D") and i % 3 == 0:
                obs[0] = rand_superellipsoid(rng, x)
            goal = rng.uniform(0, 10, 3 if name == "Quad3D" else 2)
            rec = probe(name, spec, x, u, goal, np.array(obs) if k else None, M)
            rec.update(X=x, U=u, GOAL=np.pad(goal, (0, 3 - goal.size)), NOBS=k,
                       OBS=np.vstack([np.array(obs).reshape(-1, 7), np.full((M + 1 - k, 7), np.nan)]))
            for kk, v in rec.items():
                rows.setdefault(kk, []).append(v)
        for kk, v in rows.items():
            flat[f"{tag}/{kk}"] = np.asarray(v)
        print(tag, "horizon", rows["horizon"][0], "R", rows["R"][0], "alphas", rows["alphas"][0], "min cbf", float(np.min(rows["cbf"])))
    np.savez_compressed(os.path.join(HERE, fname), **flat)


if __name__ == "__main__":
    if "--third" in sys.argv:
        main(CASES3, 20261020, "ref_mpc_statement3.npz")
    elif "--second" in sys.argv:
        main(CASES2, 20261019, "ref_mpc_statement2.npz")
    else:
        main()
