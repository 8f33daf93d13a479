#!/usr/bin/env python
"""Optimal-decay MPC-CBF problem-statement fixtures from the REFERENCE'S OWN position_control/optimal_decay_mpc_cbf.py.

    python tests/golden/gen_odmpc_from_reference.py        # writes tests/golden/ref_odmpc_statement.npz

Same method as gen_mpc_from_reference.py: the unmodified OptimalDecayMPCCBF is constructed at seeded probe points
(x, u, omega1, omega2) through oracle/refshim (numeric casadi, probing do_mpc) and what it hands to do-mpc is recorded:
x_next of set_rhs, the 'cost' expression, the 5 CBF constraint values (with the bilinear omega terms, :296-300), the
VALUES of the two expression rterms it passes to set_rterm (:178-185, in call order), bounds, horizon, tvp goal / obstacle
padding (5 x 5) and alphas.  Nothing is solved; how do-mpc combines the two set_rterm calls stays unpinned.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import refshim  # noqa: E402

refshim.install()
from oracle.refshim import fake_do_mpc as fdm  # noqa: E402

from gen_from_reference import Facade, rand_state, rand_input, rand_circle  # noqa: E402
from safe_control.position_control.optimal_decay_mpc_cbf import OptimalDecayMPCCBF  # noqa: E402

CASES = [("DynamicUnicycle2D", {}), ("KinematicBicycle2D", {}), ("Quad2D", {}), ("VTOL2D", {}),
         ("KinematicBicycle2D", {"a_max": 2.0, "v_max": 2.0})]


def probe(name, spec, x, u, om, goal, obs):
    nx = x.size
    fdm.PROBE["_x"] = {"x": x}; fdm.PROBE["_u"] = {"u": u, "omega1": om[0:1], "omega2": om[1:2]}; fdm.PROBE["_tvp"] = {}
    fac = Facade(name, x, spec)
    ctrl = OptimalDecayMPCCBF(fac, fac.robot_spec)
    ctrl.update_tvp(goal, obs)
    tvp = ctrl.mpc.tvp_fun(0.0)
    fdm.PROBE["_tvp"] = dict(tvp)
    fac = Facade(name, x, spec)
    ctrl = OptimalDecayMPCCBF(fac, fac.robot_spec)
    mdl, mpc = fdm.LAST["model"], fdm.LAST["mpc"]
    nu = u.size
    lb_u = mpc.bounds.get(("lower", "_u", "u"), np.full(nu, -np.inf)); ub_u = mpc.bounds.get(("upper", "_u", "u"), np.full(nu, np.inf))
    lb_x = np.full(nx, -np.inf); ub_x = np.full(nx, np.inf)
    for key, v in mpc.bounds.items():
        if key[1] == "_x" and len(key) == 4:
            (lb_x if key[0] == "lower" else ub_x)[key[3]] = float(v)
    omega_bounded = float(any(k[1] == "_u" and k[2] in ("omega1", "omega2") for k in mpc.bounds))
    return dict(x_next=mdl.rhs["x"].reshape(-1), cost=float(mdl.aux["cost"].reshape(-1)[0]),
                cbf=np.array([-mpc.nl_cons[f"cbf_{i}"][0] for i in range(5)]),
                cons_ub=np.array([mpc.nl_cons[f"cbf_{i}"][1] for i in range(5)], float),
                rterm_calls=np.array(mpc.rterm_calls, float), lb_u=lb_u, ub_u=ub_u, lb_x=lb_x, ub_x=ub_x,
                omega_bounded=omega_bounded, horizon=int(mpc.params["n_horizon"]), t_step=float(mpc.params["t_step"]),
                tvp_goal=tvp["goal"].reshape(-1), tvp_obs=tvp["obs"].reshape(5, 5),
                alphas=np.array([float(tvp["alpha1"].reshape(-1)[0]), float(tvp["alpha2"].reshape(-1)[0])]),
                R=np.asarray(ctrl.R, float), p_sb=np.array([ctrl.cbf_param["p_sb1"], ctrl.cbf_param["p_sb2"]], float),
                omega0=np.array([ctrl.cbf_param["omega1"], ctrl.cbf_param["omega2"]], float),
                lterm_is_mterm=float(np.array_equal(mpc.objective["lterm"], mpc.objective["mterm"])))


def main(seed=20261021, fname="ref_odmpc_statement.npz"):
    rng = np.random.default_rng(seed)
    n = 20
    flat = {}
    for name, spec in CASES:
        tag = name + ("" if not spec else "+" + ",".join(f"{k}={v}" for k, v in spec.items()))
        rows = {}
        try:                                   # does the reference construct at all for this model?
            x = rand_state(np.random.default_rng(0), name)
            probe(name, spec, x, rand_input(np.random.default_rng(0), spec, name), np.ones(2), np.zeros(2), None)
        except Exception as e:                 # DynamicUnicycle2D: its agent_barrier_dt reads obs[6] of the 5-column tvp row
            flat[f"{tag}/raises"] = np.array(f"{type(e).__name__}: {e}")
            print(tag, "REFERENCE RAISES:", type(e).__name__, e)
            continue
        for i in range(n):
            x = rand_state(rng, name); u = rand_input(rng, spec, name)
            om = rng.uniform(0.0, 2.0, 2)
            k = int(rng.integers(0, 7))
            obs = [rand_circle(rng, x, True)[:5] for _ in range(k)]                  # rows [x, y, r, vx, vy] (:336-345)
            goal = rng.uniform(0, 10, 2)
            rec = probe(name, spec, x, u, om, goal, np.array(obs) if k else None)
            rec.update(X=x, U=u, OMEGA=om, GOAL=goal, NOBS=k,
                       OBS=np.vstack([np.array(obs).reshape(-1, 5), np.full((7 - k, 5), np.nan)]))
            for kk, v in rec.items():
                rows.setdefault(kk, []).append(v)
        for kk, v in rows.items():
            flat[f"{tag}/{kk}"] = np.asarray(v)
        print(tag, "horizon", rows["horizon"][0], "R", rows["R"][0], "alphas", rows["alphas"][0], "rterm calls", rows["rterm_calls"][0],
              "omega bounded", rows["omega_bounded"][0])
    np.savez_compressed(os.path.join(HERE, fname), **flat)


if __name__ == "__main__":
    main()
