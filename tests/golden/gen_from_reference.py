#!/usr/bin/env python
"""Generate golden fixtures by running the REFERENCE'S OWN CODE in the build container.

    python tests/golden/gen_from_reference.py        # writes tests/golden/ref_*.npz

Runs only where /root/reference exists (never on the GPU box, never from the
tests themselves).  The reference's third-party solver stack is absent
(SURVEY.md 8c), so it is imported through oracle/refshim:

  * robots/*.py, dynamic_env/kinematic_bicycle2D_c3bf.py  -- imported unmodified;
    f, g, step, nominal_input, agent_barrier, agent_barrier_dt are evaluated by
    the reference's own formulas (casadi -> numeric stand-in).
  * position_control/cbf_qp.py, optimal_decay_cbf_qp.py -- imported unmodified;
    CBFQP / OptimalDecayCBFQP run end to end with cvxpy replaced by an
    affine-expression layer feeding oracle/qp_exact.py (strictly convex QP =>
    unique optimum, so the substitution cannot change the answer beyond GUROBI's
    own 1e-6 tolerances).

What this pins: constants, row assembly, problem statement, bounds, status
semantics.  What it cannot pin: GUROBI's / IPOPT's floating-point output.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import refshim  # noqa: E402

refshim.install()

from safe_control.robots.single_integrator2D import SingleIntegrator2D  # noqa: E402
from safe_control.robots.dynamic_unicycle2D import DynamicUnicycle2D  # noqa: E402
from safe_control.robots.kinematic_bicycle2D import KinematicBicycle2D  # noqa: E402
from safe_control.dynamic_env.kinematic_bicycle2D_c3bf import KinematicBicycle2D_C3BF  # noqa: E402
from safe_control.robots.quad3D import Quad3D  # noqa: E402
from safe_control.robots.double_integrator2D import DoubleIntegrator2D  # noqa: E402
from safe_control.robots.quad2D import Quad2D  # noqa: E402
from safe_control.robots.unicycle2D import Unicycle2D  # noqa: E402
from safe_control.robots.manipulator2D import Manipulator2D  # noqa: E402
from safe_control.robots.vtol2D import VTOL2D  # noqa: E402
from safe_control.dynamic_env.kinematic_bicycle2D_dpcbf import KinematicBicycle2D_DPCBF  # noqa: E402
from safe_control.position_control.cbf_qp import CBFQP  # noqa: E402
from safe_control.position_control.optimal_decay_cbf_qp import OptimalDecayCBFQP  # noqa: E402

MODEL_CLS = {
    "SingleIntegrator2D": SingleIntegrator2D,
    "DynamicUnicycle2D": DynamicUnicycle2D,
    "KinematicBicycle2D": KinematicBicycle2D,
    "KinematicBicycle2D_C3BF": KinematicBicycle2D_C3BF,
    "Quad3D": Quad3D,
    "DoubleIntegrator2D": DoubleIntegrator2D,
    "Quad2D": Quad2D,
    "KinematicBicycle2D_DPCBF": KinematicBicycle2D_DPCBF,
    "Unicycle2D": Unicycle2D,
    "Manipulator2D": Manipulator2D,
    "VTOL2D": VTOL2D,
}
BASE_MODELS = ("SingleIntegrator2D", "DynamicUnicycle2D", "KinematicBicycle2D", "KinematicBicycle2D_C3BF", "Quad3D")
EXTRA_MODELS = ("DoubleIntegrator2D", "Quad2D", "KinematicBicycle2D_DPCBF")      # SURVEY 8f-2, second fixture set
EXTRA3_MODELS = ("Unicycle2D", "Manipulator2D")                                                   # third fixture set
DT = 0.05


class Facade:
    """The forwarding half of robots/robot.py:BaseRobot (31-50, 389-399, 435-439)
    without its shapely/matplotlib half."""

    def __init__(self, name, X, spec=None):
        self.robot_spec = dict(spec or {}, model=name)
        self.robot_spec.setdefault("radius", 0.25)          # robots/robot.py:49 (before the model ctor)
        self.robot_radius = self.robot_spec["radius"]
        self.dt = DT
        self.robot = MODEL_CLS[name](DT, self.robot_spec)
        self.X = np.array(X, float).reshape(-1, 1)

    def f(self): return self.robot.f(self.X)
    def g(self): return self.robot.g(self.X)
    def f_casadi(self, X): return self.robot.f(X, casadi=True)          # robots/robot.py:395-399
    def g_casadi(self, X): return self.robot.g(X, casadi=True)
    def agent_barrier(self, obs): return self.robot.agent_barrier(self.X, obs, self.robot_radius)
    def agent_barrier_dt(self, x, u, obs): return self.robot.agent_barrier_dt(x, u, obs, self.robot_radius)


def rand_state(rng, name):
    if name == "SingleIntegrator2D":
        return rng.uniform(0, 10, 2)
    if name == "DynamicUnicycle2D":
        return np.array([*rng.uniform(0, 10, 2), rng.uniform(-np.pi, np.pi), rng.uniform(0, 1.0)])
    if name == "Unicycle2D":
        return np.array([*rng.uniform(0, 10, 2), rng.uniform(-np.pi, np.pi)])
    if name == "Manipulator2D":
        return rng.uniform(-np.pi, np.pi, 3)
    if name.startswith("KinematicBicycle2D"):
        return np.array([*rng.uniform(0, 10, 2), rng.uniform(-np.pi, np.pi), rng.uniform(0.2, 3.5)])
    if name == "DoubleIntegrator2D":
        return np.array([*rng.uniform(0, 10, 2), *rng.uniform(-0.7, 0.7, 2)])
    if name == "Quad2D":
        return np.array([*rng.uniform(0, 10, 2), rng.uniform(-0.5, 0.5), *rng.uniform(-1, 1, 2), rng.uniform(-0.3, 0.3)])
    if name == "VTOL2D":          # cruise-like states: forward speed 3..12 m/s, small pitch / pitch rate
        return np.array([rng.uniform(0, 10), rng.uniform(2, 10), rng.uniform(-0.25, 0.25), rng.uniform(3, 12), rng.uniform(-1.5, 1.5),
                         rng.uniform(-0.2, 0.2)])
    x = np.zeros(12)
    x[0:2] = rng.uniform(0, 10, 2); x[2] = rng.uniform(1, 3)
    x[3:6] = rng.normal(0, 0.05, 3); x[6:9] = rng.normal(0, 0.5, 3); x[9:12] = rng.normal(0, 0.05, 3)
    return x


def rand_input(rng, spec, name):
    if name == "SingleIntegrator2D":
        return rng.uniform(-1, 1, 2)
    if name == "DynamicUnicycle2D":
        return rng.uniform(-0.5, 0.5, 2)
    if name == "Unicycle2D":
        return np.array([rng.uniform(-1, 1), rng.uniform(-0.5, 0.5)])
    if name == "Manipulator2D":
        return rng.uniform(-2, 2, 3)
    if name.startswith("KinematicBicycle2D"):
        return np.array([rng.uniform(-5, 5), rng.uniform(-0.3, 0.3)])
    if name == "DoubleIntegrator2D":
        return rng.uniform(-1, 1, 2)
    if name == "Quad2D":
        return rng.uniform(1, 10, 2)
    if name == "VTOL2D":
        return np.array([*rng.uniform(0, 1, 3), rng.uniform(-0.5, 0.5)])
    return rng.uniform(-10, 10, 4)


def rand_circle(rng, near, dynamic=False):
    ang = rng.uniform(-np.pi, np.pi); dist = rng.uniform(1.2, 5.0)
    o = np.zeros(7)
    o[0] = near[0] + dist * np.cos(ang); o[1] = near[1] + dist * np.sin(ang); o[2] = rng.uniform(0.2, 0.6)
    if dynamic:
        o[3:5] = rng.uniform(-0.5, 0.5, 2)
    return o


def rand_superellipsoid(rng, near):
    # reference formula uses real powers of the (signed) rotated offsets
    # (single_integrator2D.py:137): keep e an even integer so it is defined everywhere.
    ang = rng.uniform(-np.pi, np.pi); dist = rng.uniform(2.0, 5.0)
    return np.array([near[0] + dist * np.cos(ang), near[1] + dist * np.sin(ang),
                     rng.uniform(0.3, 1.0), rng.uniform(0.3, 1.0), float(rng.choice([2, 4, 6])),
                     rng.uniform(-np.pi, np.pi), 1.0])


def gen_models(rng, n=48, names=BASE_MODELS, fname="ref_models.npz"):
    out = {}
    for name in names:
        X, U, OBS, G = [], [], [], []
        F, Gm, STEP, NOM = [], [], [], []
        CT, DTB = [], []
        for _ in range(n):
            x = rand_state(rng, name); fac = Facade(name, x); m = fac.robot
            u = rand_input(rng, fac.robot_spec, name)
            dyn = name.endswith("C3BF") or name.endswith("DPCBF")
            obs = [rand_circle(rng, x, dyn) for _ in range(3)]
            if name in ("SingleIntegrator2D", "DynamicUnicycle2D", "DoubleIntegrator2D"):
                obs.append(rand_superellipsoid(rng, x))
            else:
                obs.append(rand_circle(rng, x, dyn))
            goal = rng.uniform(0, 10, 3)
            X.append(x); U.append(u); OBS.append(np.stack(obs)); G.append(goal)
            F.append(np.asarray(fac.f(), float).reshape(-1))
            Gm.append(np.asarray(fac.g(), float))
            xc, uc = x.reshape(-1, 1).copy(), u.reshape(-1, 1)
            if name in ("SingleIntegrator2D", "DynamicUnicycle2D", "DoubleIntegrator2D", "Quad2D", "Unicycle2D", "Manipulator2D"):
                STEP.append(np.asarray(m.step(xc, uc), float).reshape(-1))
            else:
                STEP.append(np.asarray(m.step(xc, uc, casadi=False), float).reshape(-1))
            if name == "SingleIntegrator2D":
                NOM.append(np.asarray(m.nominal_input(fac.X, goal[:2]), float).reshape(-1))
            elif name == "DoubleIntegrator2D":   # facade: (X, goal, d_min, k_v, k_a) (robots/robot.py:408-409)
                NOM.append(np.asarray(m.nominal_input(fac.X, goal[:2], 0.05, 1.0, 1.0), float).reshape(-1))
            elif name == "Unicycle2D":           # facade: (X, goal, d_min, k_omega, k_v) (robots/robot.py:404-405)
                NOM.append(np.asarray(m.nominal_input(fac.X, goal[:2], 0.05, 2.0, 1.0), float).reshape(-1))
            elif name == "Manipulator2D":        # facade: (X, goal) (robots/robot.py:414-415)
                NOM.append(np.asarray(m.nominal_input(fac.X, goal[:2].reshape(-1, 1)), float).reshape(-1))
            elif name == "Quad2D":               # cascaded PD law, not on the solve path: not restated
                NOM.append(np.full(2, np.nan))
            elif name == "Quad3D":
                NOM.append(np.asarray(m.nominal_input(fac.X, goal), float).reshape(-1))
            else:   # facade passes (X, goal, d_min, k_omega, k_a, k_v) positionally (robots/robot.py:406-407)
                NOM.append(np.asarray(m.nominal_input(fac.X, goal[:2], 0.05, 2.0, 1.0, 1.0), float).reshape(-1))
            ct_rows, dt_rows = [], []
            for o in obs:
                if name != "Quad3D":
                    # Unicycle2D.agent_barrier indexes obs[2][0] (unicycle2D.py:109): it only accepts COLUMN obstacles
                    parts = fac.agent_barrier(o.reshape(-1, 1) if name == "Unicycle2D" else o)
                    ct_rows.append(np.concatenate([np.asarray(p, float).reshape(-1) for p in parts]))
                if name == "Manipulator2D":                  # no agent_barrier_dt (no MPC for the arm)
                    dt_rows.append(np.zeros(2)); continue
                parts = fac.agent_barrier_dt(x.reshape(-1, 1).copy(), u.reshape(-1, 1), o)
                dt_rows.append(np.array([float(np.asarray(p).reshape(-1)[0]) for p in parts]))
            CT.append(np.stack(ct_rows) if ct_rows else np.zeros((4, 0)))
            DTB.append(np.stack(dt_rows))
        out[name] = dict(X=np.stack(X), U=np.stack(U), OBS=np.stack(OBS), GOAL=np.stack(G), F=np.stack(F),
                         G=np.stack(Gm), STEP=np.stack(STEP), NOM=np.stack(NOM), CT=np.stack(CT), DT=np.stack(DTB))
    flat = {f"{k}/{kk}": v for k, d in out.items() for kk, v in d.items()}
    np.savez_compressed(os.path.join(HERE, fname), **flat)
    print(fname, {k: v["X"].shape for k, v in out.items()})


BASE_CBFQP = [("SingleIntegrator2D", {}), ("DynamicUnicycle2D", {}), ("KinematicBicycle2D", {}),
              ("KinematicBicycle2D_C3BF", {}), ("DynamicUnicycle2D", {"cbf_mode": "hard"}),
              ("DynamicUnicycle2D", {"cbf_alpha1": 0.7, "cbf_alpha2": 2.5, "a_max": 1.0, "radius": 0.3})]
EXTRA_CBFQP = [("DoubleIntegrator2D", {}), ("Quad2D", {}), ("KinematicBicycle2D_DPCBF", {}),
               ("KinematicBicycle2D_DPCBF", {"cbf_mode": "hard"}), ("DoubleIntegrator2D", {"cbf_mode": "hard", "a_max": 2.0})]


def gen_cbfqp(rng, n=64, num_obs=6, cases=BASE_CBFQP, fname="ref_cbfqp.npz", max_k=None):
    """Reference CBFQP end to end (row assembly + problem + status)."""
    flat = {}
    for name, spec in cases:
        tag = name + ("" if not spec else "+" + ",".join(f"{k}={v}" for k, v in spec.items()))
        X, UR, OBS, NOBS, A, B, Uo, ST = [], [], [], [], [], [], [], []
        for i in range(n):
            x = rand_state(rng, name); fac = Facade(name, x, spec)
            ctrl = CBFQP(fac, fac.robot_spec, num_obs=num_obs)
            k = int(rng.integers(0, (max_k if max_k is not None else num_obs + 2) + 1))   # also more obstacles than rows
            dyn = name.endswith("C3BF") or name.endswith("DPCBF")
            obs = np.stack([rand_circle(rng, x, dyn) for _ in range(k)]) if k else None
            if k and name in ("SingleIntegrator2D", "DynamicUnicycle2D", "DoubleIntegrator2D") and i % 4 == 0:
                obs[0] = rand_superellipsoid(rng, x)
            u_ref = rand_input(rng, fac.robot_spec, name) * 1.3   # sometimes outside the box
            # Unicycle2D: rows of a 2-D obstacle array crash its agent_barrier (obs[2][0] on a scalar), i.e. the
            # reference's own tracking.py:611-616 call is dead for this model; feed the column form it expects
            obs_arg = [o.reshape(-1, 1) for o in obs] if (name == "Unicycle2D" and obs is not None) else obs
            u = ctrl.solve_control_problem(fac.X, {"u_ref": u_ref.reshape(-1, 1)}, obs_arg)
            pad = np.full(((max_k if max_k is not None else num_obs + 2), 7), np.nan)
            if k:
                pad[:k] = obs
            X.append(x); UR.append(u_ref); OBS.append(pad); NOBS.append(k)
            A.append(ctrl.A1.value.copy()); B.append(ctrl.b1.value.reshape(-1).copy())
            ST.append(0 if ctrl.status == "optimal" else 1)
            Uo.append(np.full(u_ref.size, np.nan) if u is None else np.asarray(u, float).reshape(-1))
        for k_, v in dict(X=X, UREF=UR, OBS=OBS, NOBS=NOBS, A=A, B=B, U=Uo, STATUS=ST).items():
            flat[f"{tag}/{k_}"] = np.asarray(v)
        print(tag, "infeasible:", int(np.sum(ST)), "of", n)
    np.savez_compressed(os.path.join(HERE, fname), **flat)


def gen_odcbf(rng, n=64, names=("KinematicBicycle2D_C3BF", "DynamicUnicycle2D", "KinematicBicycle2D"), fname="ref_odcbf.npz"):
    flat = {}
    for name in names:
        X, UR, OBS, HAS, Uo, OM, ST = [], [], [], [], [], [], []
        for i in range(n):
            x = rand_state(rng, name); fac = Facade(name, x)
            ctrl = OptimalDecayCBFQP(fac, fac.robot_spec)
            has = i % 8 != 7
            o = rand_circle(rng, x, name.endswith("C3BF"))
            u_ref = rand_input(rng, fac.robot_spec, name) * 1.3
            # intended argument = the single nearest obstacle (tracking.py:585-586)
            u = ctrl.solve_control_problem(fac.X, {"u_ref": u_ref.reshape(-1, 1)}, o if has else None)
            X.append(x); UR.append(u_ref); OBS.append(o); HAS.append(has)
            ST.append(0 if ctrl.status == "optimal" else 1)
            Uo.append(np.full(2, np.nan) if u is None else np.asarray(u, float).reshape(-1))
            om = [ctrl.omega1.value]
            if name != "KinematicBicycle2D_C3BF":
                om.append(ctrl.omega2.value)
            OM.append([np.nan if v is None else float(np.asarray(v).reshape(-1)[0]) for v in om])
        for k_, v in dict(X=X, UREF=UR, OBS=OBS, HAS=HAS, U=Uo, OMEGA=OM, STATUS=ST).items():
            flat[f"{name}/{k_}"] = np.asarray(v)
        print("od", name, "infeasible:", int(np.sum(ST)), "of", n)
    np.savez_compressed(os.path.join(HERE, fname), **flat)


if __name__ == "__main__" and "--third" not in sys.argv:
    rng = np.random.default_rng(20260925)
    gen_models(rng)
    gen_cbfqp(rng)
    gen_odcbf(rng)
    rng2 = np.random.default_rng(20261017)                   # second fixture set: the models added for SURVEY 8f-2
    gen_models(rng2, names=EXTRA_MODELS, fname="ref_models2.npz")
    gen_cbfqp(rng2, cases=EXTRA_CBFQP, fname="ref_cbfqp2.npz")
    gen_odcbf(rng2, names=("Quad2D",), fname="ref_odcbf2.npz")


def third_set():
    rng3 = np.random.default_rng(20261018)                   # third fixture set: Unicycle2D
    gen_models(rng3, names=EXTRA3_MODELS, fname="ref_models3.npz")
    gen_cbfqp(rng3, cases=[("Unicycle2D", {}), ("Unicycle2D", {"cbf_mode": "hard", "v_max": 2.0})], fname="ref_cbfqp3.npz")
    # Manipulator2D: 25 rows per obstacle, so give CBFQP a row budget that cuts inside the 2nd / 3rd obstacle
    gen_cbfqp(rng3, num_obs=60, cases=[("Manipulator2D", {}), ("Manipulator2D", {"cbf_mode": "hard", "w_max": 1.0})],
              fname="ref_cbfqp4.npz", max_k=4)


if __name__ == "__main__" and "--third" in sys.argv:
    third_set()
