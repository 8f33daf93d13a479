"""Oracle restatement of MPCCBF (TEST INFRASTRUCTURE).

The reference builds this NLP with do-mpc/CasADi and solves it with IPOPT+MUMPS
(position_control/mpc_cbf.py:108-160 model, 162-259 MPC, 295-325 CBF rows, 366-402 solve);
do-mpc >= 5.1.1 and casadi are NOT in /root/reference and not installed.
We restate the NLP exactly as do-mpc's MPC._prepare_nlp formulates it (multiple shooting,
n_robust = 0, discrete model) and solve it with scipy (SLSQP, cross-checkable with
trust-constr), derivatives by torch.autograd in float64:

  vars   x_0..x_H, u_0..u_{H-1}
  min    sum_{k<H} [(x_k-g)'Q(x_k-g) + sum_i R_i (u_k,i - u_{k-1,i})^2] + (x_H-g)'Q(x_H-g)
         u_{-1} = u_prev  (set_rterm penalises the input RATE, mpc_cbf.py:180)
         g = [goal, 0...]  (mpc_cbf.py:267)
  s.t.   x_0 = x_init;  x_{k+1} = x_k + (f(x_k) + g(x_k) u_k) dt        (mpc_cbf.py:135-141)
         cbf_j(x_k, u_k) >= 0   k < H, j < num_obs                       (mpc_cbf.py:301-325)
             built from the model's agent_barrier_dt, i.e. the model's OWN step (clip / RK4)
         u_lb <= u_k <= u_ub;  DU/KB: |x_k[3]| <= v_max                   (mpc_cbf.py:183-221)
         missing obstacle slots = [1000, 1000, 0, 0, 0, 0, 0]             (mpc_cbf.py:346-364)
  start  x_k = x_init, u_k = u_prev for all k (set_initial_guess, mpc_cbf.py:368-369)

PARITY: the problem STATEMENT is pinned against the reference's own position_control/mpc_cbf.py --
tests/golden/gen_mpc_from_reference.py constructs the unmodified MPCCBF through oracle/refshim
(numeric casadi, probing do_mpc stand-in) and records what it hands to do-mpc at seeded probe points
(x_next of set_rhs, the cost expression, every CBF constraint value, bounds, rterm weights, horizon,
tvp goal / dummy-obstacle padding, alphas); tests/test_oracle_pinned.py::test_mpc_statement_matches_reference
checks euler(), the stage cost, cbf pieces, pad_obs, bounds and constants of this file against them.
PARITY UNPINNED for what do-mpc itself does with those pieces (lterm summed over k < H plus mterm, rterm
as a penalty on u_k - u_{k-1}, nl_cons at every stage but the terminal node) and for IPOPT's output:
neither do-mpc nor casadi can be installed here.
"""
import math

import numpy as np
import torch
from scipy.optimize import minimize, NonlinearConstraint, Bounds

from .models import resolve_spec

torch.set_default_dtype(torch.float64)

MPC_PARAM = {   # mpc_cbf.py:19-39 (Q diag, R), 49-82 (alpha)
    "SingleIntegrator2D": dict(Q=[50, 50], R=[5, 5], alpha=0.05),
    "Unicycle2D": dict(Q=[50, 50, 0.01], R=[0.5, 0.5], alpha=0.05),
    "DynamicUnicycle2D": dict(Q=[50, 50, 0.01, 30], R=[0.5, 0.5], alpha1=0.15, alpha2=0.15),
    "KinematicBicycle2D": dict(Q=[50, 50, 1, 1], R=[0.5, 5000.0], alpha1=0.1, alpha2=0.1),
    "KinematicBicycle2D_C3BF": dict(Q=[50, 50, 1, 1], R=[0.5, 5000.0], alpha=0.15),
    "KinematicBicycle2D_DPCBF": dict(Q=[50, 50, 1, 1], R=[0.5, 5000.0], alpha=0.15),
    "Quad3D": dict(Q=[30, 30, 5, 20, 20, 1, 10, 10, 10, 20, 20, 1], R=[1, 1, 1, 1], alpha=0.15),
    "DoubleIntegrator2D": dict(Q=[50, 50, 20, 20], R=[0.5, 0.5], alpha1=0.2, alpha2=0.2),
    "Quad2D": dict(Q=[25, 25, 50, 10, 10, 50], R=[0.5, 0.5], alpha1=0.15, alpha2=0.15),
    "VTOL2D": dict(Q=[10, 10, 250, 10, 10, 50], R=[0.5, 0.5, 0.5, 50000.0], alpha1=0.05, alpha2=0.05),   # mpc_cbf.py:40-43, 83-87
}
DUMMY = [1000.0, 1000.0, 0.0, 0.0, 0.0, 0.0, 0.0]


class TorchModel:
    """f, g, the model's own step and the discrete barrier h, in torch (batched over stages)."""

    def __init__(self, spec, dt):
        self.s, self.dt, self.name = spec, dt, spec["model"]
        self.R = float(spec["radius"])
        n = self.name
        self.nx, self.nu = {"SingleIntegrator2D": (2, 2), "Quad3D": (12, 4), "Quad2D": (6, 2), "Unicycle2D": (3, 2),
                            "VTOL2D": (6, 4)}.get(n, (4, 2))
        if n == "Quad3D":
            L, nu, gr = spec["L"], spec["nu"], 9.8
            B2 = torch.tensor([[1, 1, 1, 1], [0, L, 0, -L], [L, 0, -L, 0], [nu, -nu, nu, -nu]], dtype=torch.float64)
            A = torch.zeros(12, 12)
            for i in range(6):
                A[i, 6 + i] = 1
            A[6, 3] = gr; A[7, 4] = -gr
            B1 = torch.zeros(12, 4)
            B1[8, 0] = 1 / spec["mass"]; B1[9, 1] = 1 / spec["Iy"]; B1[10, 2] = 1 / spec["Ix"]; B1[11, 3] = 1 / spec["Iz"]
            self.A, self.B = A, B1 @ B2

    def rhs(self, x, u):
        """x_dot = f(x) + g(x) u, rows = stages."""
        n = self.name
        if n == "SingleIntegrator2D":
            return u
        if n == "Quad3D":
            return x @ self.A.T + u @ self.B.T
        if n == "Unicycle2D":                                   # unicycle2D.py:43-63
            return torch.stack([u[:, 0] * torch.cos(x[:, 2]), u[:, 0] * torch.sin(x[:, 2]), u[:, 1]], dim=1)
        if n == "DoubleIntegrator2D":                           # double_integrator2D.py:46-78
            return torch.stack([x[:, 2], x[:, 3], u[:, 0], u[:, 1]], dim=1)
        if n == "VTOL2D":                                       # vtol2D.py:115-300 (f: baseline aero + gravity; g: rotors, aero at delta_e = 1)
            sp = self.s
            th, xd, zd, thd = x[:, 2], x[:, 3], x[:, 4], x[:, 5]
            c, sn = torch.cos(th), torch.sin(th)
            ub, wb = c * xd + sn * zd, -sn * xd + c * zd
            V = torch.sqrt(ub * ub + wb * wb)
            al = torch.atan2(-wb, ub)
            CLlin = sp["C_L0"] + sp["C_Lalpha"] * al
            CLnl = 2 * torch.sin(al) * torch.cos(al)
            t1 = torch.exp(-sp["M"] * (al - sp["alpha_0"])); t2 = torch.exp(sp["M"] * (al + sp["alpha_0"]))
            sig = (1 + t1 + t2) / ((1 + t1) * (1 + t2))
            CLa = (1 - sig) * CLlin + sig * CLnl
            qS = 0.5 * sp["rho"] * V ** 2 * sp["S_wing"]

            def ldm(de):
                return (qS * (CLa + sp["C_Ldelta_e"] * de), qS * (sp["C_D0"] + sp["C_Dalpha"] * al ** 2 + sp["C_Ddelta_e"] * de),
                        qS * (sp["C_m0"] + sp["C_malpha"] * al + sp["C_mdelta_e"] * de) * sp["chord"])
            ch, sh = torch.cos(th + al), torch.sin(th + al)
            L0, D0, M0 = ldm(0.0); L1, D1, M1 = ldm(1.0)
            w2i = lambda D, L: (ch * (-D) - sh * L, sh * (-D) + ch * L)
            fx0, fz0 = w2i(D0, L0); fx1, fz1 = w2i(D1, L1)
            m, I = sp["mass"], sp["inertia"]
            kf, kr, kp = sp["k_front"], sp["k_rear"], sp["k_pusher"]
            xdd = fx0 / m + (-sn * kf / m) * u[:, 0] + (-sn * kr / m) * u[:, 1] + (c * kp / m) * u[:, 2] + (fx1 / m) * u[:, 3]
            zdd = (fz0 - m * 9.81) / m + (c * kf / m) * u[:, 0] + (c * kr / m) * u[:, 1] + (sn * kp / m) * u[:, 2] + (fz1 / m) * u[:, 3]
            tdd = M0 / I + (sp["ell_f"] * kf / I) * u[:, 0] + (-sp["ell_r"] * kr / I) * u[:, 1] + (M1 / I) * u[:, 3]
            return torch.stack([xd, zd, thd, xdd, zdd, tdd], dim=1)
        if n == "Quad2D":                                       # quad2D.py:46-85
            m, I, r = self.s["mass"], self.s["inertia"], self.R
            th, us = x[:, 2], u[:, 0] + u[:, 1]
            return torch.stack([x[:, 3], x[:, 4], x[:, 5], -torch.sin(th) / m * us, -9.81 + torch.cos(th) / m * us,
                                r / I * (u[:, 0] - u[:, 1])], dim=1)
        th, v = x[:, 2], x[:, 3]
        c, s = torch.cos(th), torch.sin(th)
        if n == "DynamicUnicycle2D":
            return torch.stack([v * c, v * s, u[:, 1], u[:, 0]], dim=1)
        Lr = self.s["rear_ax_dist"]
        b = u[:, 1]
        return torch.stack([v * c - v * s * b, v * s + v * c * b, v / Lr * b, u[:, 0]], dim=1)

    def euler(self, x, u):
        return x + self.rhs(x, u) * self.dt

    def own_step(self, x, u):
        """The model's step() as used inside agent_barrier_dt (wrap omitted: the barriers only
        see the heading through cos/sin, or not at all)."""
        n = self.name
        if n == "Quad3D":
            dt = self.dt
            k1 = self.rhs(x, u); k2 = self.rhs(x + dt / 2 * k1, u)
            k3 = self.rhs(x + dt / 2 * k2, u); k4 = self.rhs(x + dt * k3, u)
            return x + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        xn = self.euler(x, u)
        if n == "DoubleIntegrator2D":                           # velocity rescaled to |v| <= v_max (:80-108)
            vmag = torch.sqrt(xn[:, 2] ** 2 + xn[:, 3] ** 2)
            sc = torch.where(vmag > self.s["v_max"], self.s["v_max"] / vmag, torch.ones_like(vmag))
            xn = torch.cat([xn[:, :2], xn[:, 2:4] * sc[:, None]], dim=1)
        if n.startswith("KinematicBicycle2D"):
            v = torch.clamp(xn[:, 3], self.s["v_min"], self.s["v_max"])
            xn = torch.cat([xn[:, :3], v[:, None]], dim=1)
        return xn

    def h(self, x, obs):
        """x [K, nx], obs [M, 7] -> [K, M] barrier values."""
        n = self.name
        if n == "KinematicBicycle2D_C3BF":                      # kinematic_bicycle2D_c3bf.py:82-108
            th, v = x[:, 2:3], x[:, 3:4]
            ego = (obs[None, :, 2] + self.R) * 1.01
            px, py = obs[None, :, 0] - x[:, 0:1], obs[None, :, 1] - x[:, 1:2]
            vx, vy = obs[None, :, 3] - v * torch.cos(th), obs[None, :, 4] - v * torch.sin(th)
            pm = torch.sqrt(px * px + py * py); vm = torch.sqrt(vx * vx + vy * vy)
            return (px * vx + py * vy) + pm * vm * torch.sqrt(torch.clamp(pm ** 2 - ego ** 2, min=0.0)) / pm
        if n == "KinematicBicycle2D_DPCBF":                     # kinematic_bicycle2D_dpcbf.py:86-136
            th, v = x[:, 2:3], x[:, 3:4]
            sm = 1.05
            ego = (obs[None, :, 2] + self.R) * sm
            px, py = obs[None, :, 0] - x[:, 0:1], obs[None, :, 1] - x[:, 1:2]
            vx, vy = obs[None, :, 3] - v * torch.cos(th), obs[None, :, 4] - v * torch.sin(th)
            pm = torch.sqrt(px * px + py * py); vm = torch.sqrt(vx * vx + vy * vy)
            rot = torch.atan2(py, px)
            vnx = torch.cos(rot) * vx + torch.sin(rot) * vy
            vny = -torch.sin(rot) * vx + torch.cos(rot) * vy
            d_safe = torch.clamp(pm ** 2 - ego ** 2, min=1e-6)
            kl, km = 0.1 * math.sqrt(sm ** 2 - 1) / ego, 0.5 * math.sqrt(sm ** 2 - 1) / ego
            return vnx + kl * torch.sqrt(d_safe) / vm * vny ** 2 + km * torch.sqrt(d_safe)
        beta = 1.1 if n == "KinematicBicycle2D" else 1.01
        dx, dy = x[:, 0:1] - obs[None, :, 0], x[:, 1:2] - obs[None, :, 1]
        circ = dx * dx + dy * dy - beta * (obs[None, :, 2] + self.R) ** 2
        if n in ("SingleIntegrator2D", "DynamicUnicycle2D", "DoubleIntegrator2D") and bool((obs[:, 6] >= 0.5).any()):
            a = torch.clamp(obs[:, 2].abs(), min=1e-3); b = torch.clamp(obs[:, 3].abs(), min=1e-3)
            e = torch.clamp(obs[:, 4].abs(), min=2.0)
            ct, st = torch.cos(obs[:, 5]), torch.sin(obs[:, 5])
            xp = ct[None] * dx + st[None] * dy; yp = -st[None] * dx + ct[None] * dy
            sup = (xp.abs() / (a + self.R)[None]) ** e[None] + (yp.abs() / (b + self.R)[None]) ** e[None] - 1
            return torch.where(obs[None, :, 6] < 0.5, circ, sup)
        return circ


class OracleMPCCBF:
    def __init__(self, robot_spec, num_obs=5, horizon=None, dt=0.05):
        self.spec = resolve_spec(robot_spec)
        self.name = self.spec["model"]
        self.dt, self.M = dt, num_obs
        self.H = int(horizon if horizon is not None else self.spec.get("mpc_horizon", 10))
        if self.name == "VTOL2D" and horizon is None:            # mpc_cbf.py:40-41 overrides the horizon for this model
            self.H = 30
        par = dict(MPC_PARAM[self.name])
        for k in ("alpha", "alpha1", "alpha2"):                  # mpc_cbf.py:90-95
            if "mpc_cbf_" + k in self.spec:
                par[k] = float(self.spec["mpc_cbf_" + k])
        self.par = par
        self.tm = TorchModel(self.spec, dt)
        self.nx, self.nu = self.tm.nx, self.tm.nu
        self.Q = torch.tensor(par["Q"], dtype=torch.float64)
        self.Rw = torch.tensor(par["R"], dtype=torch.float64)
        s = self.spec
        if self.name == "SingleIntegrator2D":
            lb, ub = [-s["v_max"]] * 2, [s["v_max"]] * 2
        elif self.name == "Unicycle2D":                          # mpc_cbf.py:188-192
            lb, ub = [-s["v_max"], -s["w_max"]], [s["v_max"], s["w_max"]]
        elif self.name == "DynamicUnicycle2D":
            lb, ub = [-s["a_max"], -s["w_max"]], [s["a_max"], s["w_max"]]
        elif self.name.startswith("KinematicBicycle2D"):
            lb, ub = [-s["a_max"], -s["beta_max"]], [s["a_max"], s["beta_max"]]
        elif self.name == "DoubleIntegrator2D":                  # mpc_cbf.py:200-204
            lb, ub = [-s["ax_max"], -s["ay_max"]], [s["ax_max"], s["ay_max"]]
        elif self.name == "Quad2D":                              # mpc_cbf.py:212-216
            lb, ub = [s["f_min"]] * 2, [s["f_max"]] * 2
        elif self.name == "VTOL2D":                              # mpc_cbf.py:222-226
            lb, ub = [s["throttle_min"]] * 3 + [s["elevator_min"]], [s["throttle_max"]] * 3 + [s["elevator_max"]]
        else:
            lb, ub = [s["u_min"]] * 4, [s["u_max"]] * 4
        self.u_lb, self.u_ub = np.array(lb, float), np.array(ub, float)
        self.has_vbound = self.name == "DynamicUnicycle2D" or self.name.startswith("KinematicBicycle2D")
        # state bounds at the nodes 1..H, in the order of the CUDA kernel's active mask (include/scb.h): (state index, sign, offset)
        # meaning sign * x[i] + offset >= 0
        self.state_bounds = []
        if self.has_vbound:                                      # mpc_cbf.py:193-199, 205-211
            self.state_bounds = [(3, -1.0, s["v_max"]), (3, 1.0, s["v_max"])]
        if self.name == "VTOL2D":                                # mpc_cbf.py:227-232
            pitch = s["pitch_max"] * 3.14159 / 180
            self.state_bounds = [(3, -1.0, s["v_max"]), (3, 1.0, s["v_max"]), (4, 1.0, s["descent_speed_max"]),
                                 (2, -1.0, pitch), (2, 1.0, pitch)]
        self.status = "optimal"

    # ---- packing: w = [x_0..x_H | u_0..u_{H-1}] ----
    def split(self, w):
        H, nx, nu = self.H, self.nx, self.nu
        return w[: (H + 1) * nx].reshape(H + 1, nx), w[(H + 1) * nx:].reshape(H, nu)

    def pad_obs(self, obs):
        rows = [] if obs is None else [list(np.asarray(o, float).reshape(-1)) for o in obs][: self.M]
        rows = [r + [0.0] * (7 - len(r)) for r in rows]
        rows += [DUMMY] * (self.M - len(rows))
        return torch.tensor(rows, dtype=torch.float64).reshape(self.M, 7)

    def cost(self, w, goal_full, u_prev):
        x, u = self.split(w)
        e = x - goal_full[None]
        du = u - torch.cat([u_prev[None], u[:-1]], dim=0)
        return (e * e * self.Q[None]).sum() + (du * du * self.Rw[None]).sum()

    def eq(self, w, x0):
        x, u = self.split(w)
        return torch.cat([(x[0] - x0), (x[1:] - self.tm.euler(x[:-1], u)).reshape(-1)])

    def cbf(self, w, obs):
        """[H*M] values of the CBF constraints (>= 0)."""
        x, u = self.split(w)
        xk = x[:-1]
        tm, p = self.tm, self.par
        x1 = tm.own_step(xk, u)
        h0, h1 = tm.h(xk, obs), tm.h(x1, obs)
        if "alpha" in p:                                          # mpc_cbf.py:312-315
            return ((h1 - h0) + p["alpha"] * h0).reshape(-1)
        x2 = tm.own_step(x1, u)
        h2 = tm.h(x2, obs)
        return ((h2 - 2 * h1 + h0) + (p["alpha1"] + p["alpha2"]) * (h1 - h0) + p["alpha1"] * p["alpha2"] * h0).reshape(-1)

    def solve(self, x_init, goal, u_prev, obs, method="SLSQP", w0=None, tol=1e-10, maxiter=400):
        """-> u0 (nu,), info dict(w, fun, success, x_pred, u_pred, cbf_min, eq_res)."""
        H, nx, nu = self.H, self.nx, self.nu
        x0 = torch.tensor(np.asarray(x_init, float).reshape(-1))
        up = torch.tensor(np.asarray(u_prev, float).reshape(-1))
        g = np.zeros(nx); gl = np.asarray(goal, float).reshape(-1); g[: gl.size] = gl
        gf = torch.tensor(g)
        ob = self.pad_obs(obs)
        if w0 is None:
            w0 = np.concatenate([np.tile(x0.numpy(), H + 1), np.tile(up.numpy(), H)])     # cold start

        def tn(v):
            return torch.tensor(v, dtype=torch.float64, requires_grad=True)

        def f(v):
            t = tn(v); c = self.cost(t, gf, up); c.backward()
            return float(c.detach()), t.grad.numpy().copy()

        def jac(fn):
            def j(v):
                return torch.autograd.functional.jacobian(fn, torch.tensor(v, dtype=torch.float64), vectorize=True).numpy()
            return j

        eqf = lambda t: self.eq(t, x0)
        cbff = lambda t: self.cbf(t, ob)
        lo = np.full(w0.size, -np.inf); hi = np.full(w0.size, np.inf)
        xl, ul = self.split(lo); xh, uh = self.split(hi)
        ul[:] = self.u_lb; uh[:] = self.u_ub
        for i, sgn, off in self.state_bounds:
            if sgn < 0:
                xh[:, i] = off
            else:
                xl[:, i] = -off
        if method == "SLSQP":
            cons = [dict(type="eq", fun=lambda v: eqf(torch.tensor(v)).numpy(), jac=jac(eqf)),
                    dict(type="ineq", fun=lambda v: cbff(torch.tensor(v)).numpy(), jac=jac(cbff))]
            res = minimize(f, w0, jac=True, method="SLSQP", bounds=list(zip(lo, hi)), constraints=cons,
                           options=dict(ftol=tol, maxiter=maxiter))
        else:
            def lag_hess(v, lam_eq, lam_in):
                return None
            cons = [NonlinearConstraint(lambda v: eqf(torch.tensor(v)).numpy(), 0.0, 0.0, jac=jac(eqf)),
                    NonlinearConstraint(lambda v: cbff(torch.tensor(v)).numpy(), 0.0, np.inf, jac=jac(cbff))]
            res = minimize(f, w0, jac=True, method="trust-constr", bounds=Bounds(lo, hi), constraints=cons,
                           options=dict(gtol=1e-9, xtol=1e-12, maxiter=3000))
        w = torch.tensor(res.x)
        xs, us = self.split(w)
        info = dict(w=res.x, fun=float(res.fun), success=bool(res.success), nit=int(res.nit),
                    x_pred=xs.numpy().copy(), u_pred=us.numpy().copy(),
                    cbf_min=float(self.cbf(w, ob).min()), eq_res=float(self.eq(w, x0).abs().max()))
        return us[0].numpy().copy(), info

    # ---- condensed (single-shooting) view used to verify KKT points independently of the solver ----
    def rollout(self, x_init, u_seq):
        x = [torch.as_tensor(x_init, dtype=torch.float64).reshape(1, -1)]
        for k in range(self.H):
            x.append(self.tm.euler(x[-1], u_seq[k:k + 1]))
        return torch.cat(x, dim=0)

    def condensed(self, x_init, goal, u_prev, obs, z):
        """z = flattened u sequence (torch, requires_grad ok) -> (J, g_ineq>=0 incl. bounds)."""
        H, nu = self.H, self.nu
        u = z.reshape(H, nu)
        x = self.rollout(x_init, u)
        g = torch.zeros(self.nx); gl = torch.as_tensor(np.asarray(goal, float).reshape(-1)); g[: gl.numel()] = gl
        w = torch.cat([x.reshape(-1), u.reshape(-1)])
        J = self.cost(w, g, torch.as_tensor(np.asarray(u_prev, float).reshape(-1)))
        parts = [self.cbf(w, self.pad_obs(obs)),
                 (torch.as_tensor(self.u_ub)[None] - u).reshape(-1), (u - torch.as_tensor(self.u_lb)[None]).reshape(-1)]
        parts += [sgn * x[1:, i] + off for i, sgn, off in self.state_bounds]
        return J, torch.cat(parts)

    def active_set(self, x_init, goal, u_prev, obs, z, g_tol=1e-6, lam_tol=1e-6):
        """Active set of the NLP at the point z, defined independently of any solver: rows within g_tol of their bound
        whose non-negative least-squares multiplier (stationarity grad J = sum lam_i grad g_i over those rows) exceeds
        lam_tol.  -> (active [n_rows] bool in condensed() order, gap): gap = strict-complementarity margin
        min_i max(g_i / 1e-4, lam_i / 1e-5) over all rows (>= 1: every row is clearly active or clearly inactive)."""
        from scipy.optimize import nnls
        zt = torch.tensor(np.asarray(z, float).reshape(-1), requires_grad=True)
        J, g = self.condensed(x_init, goal, u_prev, obs, zt)
        gradJ = torch.autograd.grad(J, zt, retain_graph=True)[0].numpy()
        gv = g.detach().numpy()
        near = np.nonzero(gv < g_tol)[0]
        lam = np.zeros(gv.size)
        if near.size:
            A = np.stack([torch.autograd.grad(g[i], zt, retain_graph=True)[0].numpy() for i in near], axis=1)
            lam[near], _ = nnls(A, gradJ)
        active = lam > lam_tol
        gap = float(np.min(np.maximum(np.maximum(gv, 0.0) / 1e-4, lam / 1e-5))) if gv.size else np.inf
        return active, gap

    def kernel_bit_of_row(self, r):
        """condensed() row index -> bit index of the CUDA kernel's active mask (include/scb.h, scb_mpccbf_solve)."""
        H, M, nu = self.H, self.M, self.nu
        if r < H * M:
            return r                                             # CBF row (stage k, slot j): k*M + j in both
        r -= H * M
        if r < H * nu:                                           # u_ub - u >= 0, row k*nu + i -> upper bound bit
            k, i = divmod(r, nu)
            return H * M + k * 2 * nu + 2 * i
        r -= H * nu
        if r < H * nu:                                           # u - u_lb >= 0 -> lower bound bit
            k, i = divmod(r, nu)
            return H * M + k * 2 * nu + 2 * i + 1
        r -= H * nu                                              # state bounds: condensed() groups them by kind, the kernel by node
        kind, k = divmod(r, H)
        return H * M + 2 * H * nu + k * len(self.state_bounds) + kind

    def kkt_error(self, x_init, goal, u_prev, obs, z, res_tol=1e-4):
        """KKT check of a point -> (stationarity residual, min g, complementarity max_i lam_i g_i).
        Multipliers: non-negative least squares over the constraints within tau of their bound, for the SMALLEST
        tau in 1e-7 .. 1e-2 that brings the residual under res_tol (the last one otherwise).  A fixed small tau is
        not enough: a badly scaled row (an e = 6 superellipsoid has gradients ~1e5) can sit 1e-4 away from its
        bound with a multiplier of 1e-4 and still carry an O(1) share of stationarity -- an interior-point
        solution at mu = 1e-9 has exactly such rows -- which is why complementarity is reported separately."""
        from scipy.optimize import nnls
        zt = torch.tensor(np.asarray(z, float).reshape(-1), requires_grad=True)
        J, g = self.condensed(x_init, goal, u_prev, obs, zt)
        gradJ = torch.autograd.grad(J, zt, retain_graph=True)[0].numpy()
        gv = g.detach().numpy()
        scale = 1.0 + float(np.abs(np.asarray(z, float)).max())
        res, comp, grads = float(np.abs(gradJ).max()), 0.0, {}
        for tau in (1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2):
            act = np.nonzero(gv < tau)[0]
            if act.size == 0:
                continue
            for i in act:
                if i not in grads:
                    grads[i] = torch.autograd.grad(g[i], zt, retain_graph=True)[0].numpy()
            A = np.stack([grads[i] for i in act], axis=1)
            # multipliers: non-negative least squares on stationarity, with complementarity as a soft equation
            # (10 g_i lam_i = 0: the same 10:1 weighting as the two acceptance thresholds).  Without it, nearly collinear rows
            # (the same obstacle at consecutive stages) let NNLS shift multiplier mass onto a row that is 1e-3 away from
            # its bound -- a KKT point then "violates" complementarity only because of how the checker chose lam.
            gpos = np.maximum(gv[act], 0.0)
            lam, _ = nnls(np.vstack([A, np.diag(10.0 * gpos)]), np.concatenate([gradJ, np.zeros(act.size)]))
            res = float(np.abs(gradJ - A @ lam).max())
            comp = float(np.max(lam * gpos))
            if res <= res_tol * scale:
                break
        return res, float(gv.min()), comp


# ---------------------------------------------------------------------------------------------------------------------
OD_MPC_PARAM = {   # optimal_decay_mpc_cbf.py:28-50 (Q diag, R), 60-85 (alphas); omega0 = 1, p_sb = 10 (:87-90)
    "DynamicUnicycle2D": dict(Q=[50, 50, 0.01, 30], R=[0.5, 0.5], alpha1=0.01, alpha2=0.01),
    "KinematicBicycle2D": dict(Q=[50, 50, 1, 1], R=[0.5, 50.0], alpha1=0.05, alpha2=0.05),
    "Quad2D": dict(Q=[25, 25, 50, 10, 10, 50], R=[0.5, 0.5], alpha1=0.15, alpha2=0.15),
    "VTOL2D": dict(Q=[10, 10, 250, 10, 10, 50], R=[0.5, 0.5, 0.5, 50000.0], alpha1=0.35, alpha2=0.35),
}


class OracleODMPCCBF(OracleMPCCBF):
    """Oracle restatement of OptimalDecayMPCCBF (TEST INFRASTRUCTURE): the MPC-CBF NLP above with two more inputs per
    stage, omega1 and omega2 (optimal_decay_mpc_cbf.py:122-124), the bilinear row
        dd_h + (alpha1 omega1 + alpha2 omega2) d_h + alpha1 alpha2 omega1 omega2 h_k >= 0            (:296-300)
    for each of 5 obstacle slots (:271-280), horizon 10 (30 for VTOL2D, :24, 47) and an input cost evaluated at u_k (no
    rate penalty): the reference hands do-mpc two expression rterms (:178-185), sum_i R_i u_i^2 and
    p_sb1 (omega1 - 1)^2 + p_sb2 (omega2 - 1)^2.  `sum_rterms` False = do-mpc's assignment semantics (the second call
    replaces the first), True = both.  The per-stage pieces are pinned by tests/golden/ref_odmpc_statement.npz
    (gen_odmpc_from_reference.py); do-mpc's use of them is UNPINNED like the rest of its transcription.
    (DynamicUnicycle2D: the reference itself raises IndexError at construction -- its agent_barrier_dt reads obs[6] of
    the 5-column obstacle row; restated here with the circle branch that row can only mean.)"""

    def __init__(self, robot_spec, sum_rterms=False, dt=0.05, horizon=None):
        name = robot_spec["model"]
        if name not in OD_MPC_PARAM:
            raise ValueError(f"optimal_decay_mpc_cbf: {name} not restated")
        MPC_PARAM_SAVE = MPC_PARAM.get(name)
        MPC_PARAM[name] = dict(OD_MPC_PARAM[name])               # same ctor, this controller's weights / gains
        try:
            super().__init__(robot_spec, num_obs=5, horizon=horizon if horizon is not None else (30 if name == "VTOL2D" else 10), dt=dt)
        finally:
            MPC_PARAM[name] = MPC_PARAM_SAVE
        self.nu_model = self.nu
        self.nu = self.nu_model + 2
        self.p_sb = (10.0, 10.0); self.omega0 = (1.0, 1.0)
        self.sum_rterms = bool(sum_rterms)
        self.u_range = np.concatenate([self.u_ub - self.u_lb, [1.0, 1.0]])
        self.u_lb = np.concatenate([self.u_lb, [-np.inf, -np.inf]]); self.u_ub = np.concatenate([self.u_ub, [np.inf, np.inf]])
        ra = list(self.par["R"]) if self.sum_rterms else [0.0] * self.nu_model
        self.Ra = torch.tensor(ra + list(self.p_sb), dtype=torch.float64)
        self.ut = torch.tensor([0.0] * self.nu_model + list(self.omega0), dtype=torch.float64)

    def cost(self, w, goal_full, u_prev):
        x, u = self.split(w)
        e = x - goal_full[None]
        d = u - self.ut[None]
        return (e * e * self.Q[None]).sum() + (d * d * self.Ra[None]).sum()

    def eq(self, w, x0):
        x, u = self.split(w)
        return torch.cat([(x[0] - x0), (x[1:] - self.tm.euler(x[:-1], u[:, : self.nu_model])).reshape(-1)])

    def cbf(self, w, obs):
        x, u = self.split(w)
        xk, ur, o1, o2 = x[:-1], u[:, : self.nu_model], u[:, self.nu_model:self.nu_model + 1], u[:, self.nu_model + 1:]
        tm, p = self.tm, self.par
        x1 = tm.own_step(xk, ur); x2 = tm.own_step(x1, ur)
        h0, h1, h2 = tm.h(xk, obs), tm.h(x1, obs), tm.h(x2, obs)
        a1, a2 = p["alpha1"], p["alpha2"]
        return ((h2 - 2 * h1 + h0) + (a1 * o1 + a2 * o2) * (h1 - h0) + a1 * a2 * h0 * o1 * o2).reshape(-1)

    def rollout(self, x_init, u_seq):
        return super().rollout(x_init, u_seq[:, : self.nu_model])

    def condensed(self, x_init, goal, u_prev, obs, z):
        H, nu, nm = self.H, self.nu, self.nu_model
        u = z.reshape(H, nu)
        x = self.rollout(x_init, u)
        g = torch.zeros(self.nx); gl = torch.as_tensor(np.asarray(goal, float).reshape(-1)); g[: gl.numel()] = gl
        w = torch.cat([x.reshape(-1), u.reshape(-1)])
        J = self.cost(w, g, None)
        parts = [self.cbf(w, self.pad_obs(obs)),
                 (torch.as_tensor(self.u_ub[:nm])[None] - u[:, :nm]).reshape(-1), (u[:, :nm] - torch.as_tensor(self.u_lb[:nm])[None]).reshape(-1)]
        parts += [sgn * x[1:, i] + off for i, sgn, off in self.state_bounds]
        return J, torch.cat(parts)

    def kernel_bit_of_row(self, r):
        H, M, nm = self.H, self.M, self.nu_model
        if r < H * M:
            return r
        r -= H * M
        if r < H * nm:
            k, i = divmod(r, nm)
            return H * M + k * 2 * nm + 2 * i
        r -= H * nm
        if r < H * nm:
            k, i = divmod(r, nm)
            return H * M + k * 2 * nm + 2 * i + 1
        r -= H * nm
        kind, k = divmod(r, H)
        return H * M + 2 * H * nm + k * len(self.state_bounds) + kind
