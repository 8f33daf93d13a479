"""Oracle restatement of the reference's QP position controllers and of the
caller-side obstacle selection (TEST INFRASTRUCTURE).

  position_control/cbf_qp.py:5-45 (alpha defaults), 47-106 (problem), 108-199 (rows + solve)
  position_control/optimal_decay_cbf_qp.py:14-54, 56-130, 132-159
  tracking.py:345-403 (get_nearest_unpassed_obs)

The QPs are solved exactly (oracle/qp_exact.py).  Unlike the reference, X is
an explicit argument (the reference reads robot.X through the facade,
SURVEY 8a quirk 1).
"""
import numpy as np

from .models import make_model, angle_normalize
from .qp_exact import solve_qp_exact, OPTIMAL

REL1 = ("SingleIntegrator2D", "Unicycle2D", "KinematicBicycle2D_C3BF", "KinematicBicycle2D_DPCBF", "Quad3D")
REL2 = ("DynamicUnicycle2D", "KinematicBicycle2D", "DoubleIntegrator2D", "Quad2D")

CBFQP_ALPHA = {                       # cbf_qp.py:12-35
    "SingleIntegrator2D": dict(alpha=1.0),
    "Unicycle2D": dict(alpha=1.0),
    "Manipulator2D": dict(alpha=1.0),
    "DynamicUnicycle2D": dict(alpha1=1.5, alpha2=1.5),
    "KinematicBicycle2D": dict(alpha1=1.5, alpha2=1.5),
    "KinematicBicycle2D_C3BF": dict(alpha=1.5),
    "Quad3D": dict(alpha=1.5),
    "DoubleIntegrator2D": dict(alpha1=1.5, alpha2=1.5),
    "Quad2D": dict(alpha1=1.5, alpha2=1.5),
    "KinematicBicycle2D_DPCBF": dict(alpha=1.5),
}


class OracleCBFQP:
    def __init__(self, robot_spec, num_obs=1, dt=0.05):
        self.model = make_model(robot_spec, dt)
        self.spec = self.model.spec
        self.name = self.spec["model"]
        self.num_obs = num_obs
        self.dt = dt
        self.cbf_param = dict(CBFQP_ALPHA[self.name])
        for k in ("alpha", "alpha1", "alpha2"):          # cbf_qp.py:38-43
            if "cbf_" + k in self.spec:
                self.cbf_param[k] = float(self.spec["cbf_" + k])
        self.mode = self.spec.get("cbf_mode", "cbf")
        self.status = None

    def rows(self, X, obs_list):
        """A1 (num_obs, nu), b1 (num_obs,): constraint A1 u + b1 >= 0 (cbf_qp.py:122-185)."""
        m = self.model
        A = np.zeros((self.num_obs, m.nu)); b = np.zeros(self.num_obs)
        if obs_list is None:
            return A, b
        r, dt = 0, self.dt
        for obs in obs_list:
            if obs is None:
                continue
            if r >= self.num_obs:
                break
            obs = np.asarray(obs, float)
            if self.name == "Manipulator2D":                  # several rows per obstacle (cbf_qp.py:131-149)
                h_list, dh_list = m.agent_barrier(X, obs)
                for h, dh in zip(h_list, dh_list):
                    if r >= self.num_obs:
                        break
                    if self.mode == "hard":
                        A[r] = dh @ m.g(X); b[r] = h / dt + dh @ m.f(X)
                    else:
                        A[r] = dh; b[r] = self.cbf_param["alpha"] * h
                    r += 1
                continue
            if self.name in REL1:
                h, dh = m.agent_barrier(X, obs)
                A[r] = dh @ m.g(X)
                if self.mode == "hard":
                    b[r] = h / dt + dh @ m.f(X)
                else:
                    b[r] = dh @ m.f(X) + self.cbf_param["alpha"] * h
            else:
                h, hd, dhd = m.agent_barrier(X, obs)
                A[r] = dhd @ m.g(X)
                if self.mode == "hard":
                    b[r] = h / dt ** 2 + 2 * hd / dt + dhd @ m.f(X)
                else:
                    g1 = self.cbf_param["alpha1"] + self.cbf_param["alpha2"]
                    g2 = self.cbf_param["alpha1"] * self.cbf_param["alpha2"]
                    b[r] = dhd @ m.f(X) + g1 * hd + g2 * h
            r += 1
        return A, b

    def qp(self, X, u_ref, obs_list):
        """(P, q, G, h) of min ||u-u_ref||^2 s.t. A u + b >= 0, box (cbf_qp.py:47-106).
        Row order of G: CBF rows 0..num_obs-1, then for each input i: u_i<=ub_i, -u_i<=-lb_i."""
        nu = self.model.nu
        A, b = self.rows(X, obs_list)
        lb, ub = self.model.u_bounds()
        G = [-A]; h = [b]
        for i in range(nu):
            e = np.zeros(nu); e[i] = 1
            G += [e[None], -e[None]]; h += [[ub[i]], [-lb[i]]]
        return 2 * np.eye(nu), -2 * np.asarray(u_ref, float).reshape(-1), np.vstack(G), np.concatenate(h)

    def solve(self, X, u_ref, obs_list):
        """-> (u or None, info).  obs_list None -> u_ref unclipped (cbf_qp.py:113-118)."""
        if obs_list is None:
            self.status = "optimal"
            return np.asarray(u_ref, float).reshape(-1).copy(), dict(status=OPTIMAL, active=None, gap=np.inf)
        P, q, G, h = self.qp(X, u_ref, obs_list)
        res = solve_qp_exact(P, q, G, h)
        self.status = "optimal" if res["status"] == OPTIMAL else "infeasible"
        return res["x"], res


OD_PARAM = {                          # optimal_decay_cbf_qp.py:17-50
    "DynamicUnicycle2D": dict(alpha1=0.5, alpha2=0.5, omega1=1.0, p_sb1=1e4, omega2=1.0, p_sb2=1e4),
    "KinematicBicycle2D": dict(alpha1=0.5, alpha2=0.5, omega1=1.0, p_sb1=1e4, omega2=1.0, p_sb2=1e4),
    "KinematicBicycle2D_C3BF": dict(alpha=0.5, omega1=1.0, p_sb1=1e4),
    "Quad2D": dict(alpha1=0.5, alpha2=0.5, omega1=1.0, p_sb1=1e4, omega2=1.0, p_sb2=1e4),
}


class OracleOptimalDecayCBFQP:
    """vars z = [u (nu), omega1 (, omega2)]; ONE CBF row (A1 is 1 x nu, :61)."""

    def __init__(self, robot_spec, dt=0.05):
        self.model = make_model(robot_spec, dt)
        self.spec = self.model.spec
        self.name = self.spec["model"]
        self.cbf_param = dict(OD_PARAM[self.name])
        self.two = "alpha1" in self.cbf_param
        self.status = None

    def qp(self, X, u_ref, nearest_obs):
        m, p = self.model, self.cbf_param
        nu = m.nu
        A = np.zeros(nu); b = 0.0; h = 0.0; hd = 0.0
        if nearest_obs is not None:
            obs = np.asarray(nearest_obs, float).reshape(-1)
            if self.name == "KinematicBicycle2D_C3BF":            # :139-143
                h, dh = m.agent_barrier(X, obs)
                A = dh @ m.g(X); b = dh @ m.f(X)
            elif self.name in ("DynamicUnicycle2D", "Quad2D"):    # :144-149
                h, hd, dhd = m.agent_barrier(X, obs)
                A = dhd @ m.g(X); b = dhd @ m.f(X)
            # plain KinematicBicycle2D: no branch -> row stays zero
        n = nu + (2 if self.two else 1)
        P = np.zeros((n, n)); q = np.zeros(n)
        P[:nu, :nu] = 2 * np.eye(nu); q[:nu] = -2 * np.asarray(u_ref, float).reshape(-1)
        P[nu, nu] = 2 * p["p_sb1"]; q[nu] = -2 * p["p_sb1"] * p["omega1"]
        row = np.zeros(n); row[:nu] = A
        if self.two:
            P[nu + 1, nu + 1] = 2 * p["p_sb2"]; q[nu + 1] = -2 * p["p_sb2"] * p["omega2"]
            row[nu] = (p["alpha1"] + p["alpha2"]) * hd
            row[nu + 1] = p["alpha1"] * p["alpha2"] * h
        else:
            row[nu] = p["alpha"] * h
        lb, ub = m.u_bounds()
        G = [-row[None]]; hh = [[b]]
        for i in range(nu):
            e = np.zeros(n); e[i] = 1
            G += [e[None], -e[None]]; hh += [[ub[i]], [-lb[i]]]
        return P, q, np.vstack(G), np.concatenate(hh)

    def solve(self, X, u_ref, nearest_obs):
        P, q, G, h = self.qp(X, u_ref, nearest_obs)
        res = solve_qp_exact(P, q, G, h)
        self.status = "optimal" if res["status"] == OPTIMAL else "infeasible"
        if res["x"] is None:
            return None, None, res
        nu = self.model.nu
        return res["x"][:nu], res["x"][nu:], res


ANGLE_UNPASSED = {                    # tracking.py:352-357
    "SingleIntegrator2D": 2.0 * np.pi, "Quad3D": 2.0 * np.pi,
    "DynamicUnicycle2D": 1.2 * np.pi,
    "KinematicBicycle2D": 2.0 * np.pi, "KinematicBicycle2D_C3BF": 2.0 * np.pi,
    "KinematicBicycle2D_DPCBF": 2.0 * np.pi, "DoubleIntegrator2D": 2.0 * np.pi, "Quad2D": 2.0 * np.pi,
    "Unicycle2D": 1.2 * np.pi,
}


def nearest_unpassed_obs(model_name, pos, yaw, all_obs, obs_num):
    """tracking.py:345-403 -> (selected rows (k,7), their indices into all_obs)."""
    all_obs = np.asarray(all_obs, float)
    if all_obs.size == 0:
        return None, None
    if all_obs.ndim == 1:
        all_obs = all_obs.reshape(1, -1)
    half = ANGLE_UNPASSED[model_name] / 2
    keep = []
    for i, o in enumerate(all_obs):
        ang = np.arctan2(o[1] - pos[1], o[0] - pos[0])
        if abs(angle_normalize(ang - yaw)) <= half:
            keep.append(i)
    idx = np.array(keep, dtype=int) if keep else np.arange(len(all_obs))
    d = np.linalg.norm(all_obs[idx, :2] - np.asarray(pos, float), axis=1)
    order = np.argsort(d, kind="stable")[:obs_num]
    return all_obs[idx[order]], idx[order]
