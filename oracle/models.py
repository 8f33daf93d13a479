"""Oracle restatement of the reference's robot models (TEST INFRASTRUCTURE).

numpy float64, one agent at a time, written from the formulas in

  robots/single_integrator2D.py    (f 44, g 55, step 64, nominal_input 72, agent_barrier 114-146, _dt 148-195)
  robots/dynamic_unicycle2D.py     (f 42, g 64, step 75, nominal_input 80, agent_barrier 121-186, _dt 188-238)
  robots/kinematic_bicycle2D.py    (f 75, g 93, step 112-123, nominal_input 125, agent_barrier 160-173, _dt 175-199)
  dynamic_env/kinematic_bicycle2D_c3bf.py (agent_barrier 15-75, _dt 77-115)
  robots/quad3D.py                 (A,B 81-97, step 121-158, nominal_input 160-206, _dt 275-297)
  robots/robot.py                  (radius default 49-50; facade forwarding 389-439)

States/inputs are 1-D arrays (the reference uses (n,1) columns).  Obstacles are
1-D rows of width 7: circle [x, y, r, vx, vy, ., flag=0], superellipsoid
[x, y, a, b, e, theta, flag=1] (README.md:133-138).

Pinned against the reference's own code by tests/golden/gen_from_reference.py
(run in the build container through oracle/refshim; fixtures committed).
"""
import math

import numpy as np

MODELS = ("SingleIntegrator2D", "DynamicUnicycle2D", "KinematicBicycle2D",
          "KinematicBicycle2D_C3BF", "Quad3D")
MODELS_QP_EXTRA = ("DoubleIntegrator2D", "Quad2D", "KinematicBicycle2D_DPCBF")     # SURVEY 8f-2, second fixture set
MODELS_EXTRA3 = ("Unicycle2D", "Manipulator2D")                                     # third fixture set


def angle_normalize(x):
    """numpy branch: floored modulo (e.g. robots/dynamic_unicycle2D.py:14-16)."""
    return ((x + np.pi) % (2 * np.pi)) - np.pi


def resolve_spec(robot_spec):
    """Apply the reference's `setdefault` cascades -> a plain dict of floats.

    robots/robot.py:49 sets radius=0.25 BEFORE the model ctor runs, so
    KinematicBicycle2D's own 0.3 default never applies.
    """
    s = dict(robot_spec)
    model = s["model"]
    s.setdefault("radius", 0.25)
    if model == "SingleIntegrator2D":
        s.setdefault("v_max", 1.0); s.setdefault("w_max", 0.5)
    elif model == "DynamicUnicycle2D":
        s.setdefault("a_max", 0.5); s.setdefault("w_max", 0.5); s.setdefault("v_max", 1.0)
    elif model == "Unicycle2D":                               # unicycle2D.py:40-41
        s.setdefault("v_max", 1.0); s.setdefault("w_max", 0.5)
    elif model == "Manipulator2D":                            # manipulator2D.py:19-20
        s.setdefault("w_max", 2.0); s.setdefault("Kp", 3.0)
    elif model == "DoubleIntegrator2D":                       # double_integrator2D.py:40-44
        s.setdefault("a_max", 1.0); s.setdefault("v_max", 1.0)
        s.setdefault("ax_max", s["a_max"]); s.setdefault("ay_max", s["a_max"]); s.setdefault("w_max", 0.5)
    elif model == "Quad2D":                                   # quad2D.py:40-46
        s.setdefault("mass", 1.0); s.setdefault("inertia", 0.01); s.setdefault("f_min", 1.0); s.setdefault("f_max", 10.0)
    elif model in ("KinematicBicycle2D", "KinematicBicycle2D_C3BF", "KinematicBicycle2D_DPCBF"):
        s.setdefault("wheel_base", 0.4); s.setdefault("front_ax_dist", 0.2)
        s.setdefault("rear_ax_dist", 0.2); s.setdefault("v_max", 3.5)
        s.setdefault("a_max", 5.0); s.setdefault("delta_max", np.deg2rad(32))
        s.setdefault("beta_max", math.atan(s["rear_ax_dist"] / s["wheel_base"] * math.tan(s["delta_max"])))
        s.setdefault("v_min", 0.2)
    elif model == "Quad3D":
        s.setdefault("mass", 3.0); s.setdefault("Ix", 0.5); s.setdefault("Iy", 0.5)
        s.setdefault("Iz", 0.5); s.setdefault("L", 0.3); s.setdefault("nu", 0.1)
        s.setdefault("u_max", 10.0); s.setdefault("u_min", -10.0)
    elif model == "VTOL2D":                                   # vtol2D.py:57-110
        for k, v in (("mass", 11.0), ("inertia", 1.135), ("S_wing", 0.55), ("rho", 1.2682), ("C_L0", 0.23), ("C_Lalpha", 5.61),
                     ("M", 50.0), ("alpha_0", np.deg2rad(15)), ("C_Ldelta_e", 0.13), ("C_D0", 0.043), ("C_Dalpha", 0.03),
                     ("C_Ddelta_e", 0.0), ("C_m0", 0.0135), ("C_malpha", -2.74), ("C_mdelta_e", -0.99), ("chord", 0.18994),
                     ("k_front", 70.0), ("k_rear", 70.0), ("k_pusher", 60.0), ("ell_f", 0.5), ("ell_r", 0.5),
                     ("throttle_min", 0.0), ("throttle_max", 1.0), ("elevator_min", -0.5), ("elevator_max", 0.5),
                     ("v_max", 15.0), ("pitch_max", 15.0), ("descent_speed_max", 5.0)):
            s.setdefault(k, v)
    else:
        raise ValueError(f"oracle does not restate model {model!r}")
    return s


class Model:
    """Common facade: f, g, step, agent_barrier (continuous), barrier_dt (discrete)."""
    rel_degree = 1
    nx = nu = 0

    def __init__(self, robot_spec, dt=0.05):
        self.spec = resolve_spec(robot_spec)
        self.dt = dt
        self.radius = float(self.spec["radius"])


# ---------------------------------------------------------------------------
def _circle_h(px, py, obs, radius, beta):
    d_min = obs[2] + radius
    return (px - obs[0]) ** 2 + (py - obs[1]) ** 2 - beta * d_min ** 2


def _superellipsoid(px, py, obs, radius):
    """Continuous-time superellipsoid pieces (single_integrator2D.py:128-143)."""
    ox, oy, a, b, e, th = obs[0], obs[1], obs[2], obs[3], obs[4], obs[5]
    c, s = math.cos(th), math.sin(th)
    pxp = c * (px - ox) + s * (py - oy)
    pyp = -s * (px - ox) + c * (py - oy)
    h = (pxp / (a + radius)) ** e + (pyp / (b + radius)) ** e - 1
    dh = np.array([
        e * pxp ** (e - 1) * (c / (a + radius) ** e) + e * pyp ** (e - 1) * (-s / (b + radius) ** e),
        e * pxp ** (e - 1) * (s / (a + radius) ** e) + e * pyp ** (e - 1) * (c / (b + radius) ** e)])
    return h, dh, (pxp, pyp, c, s)


def _superellipsoid_dt(px, py, obs, radius):
    """Discrete-time (guarded) superellipsoid h (dynamic_unicycle2D.py:204-220)."""
    ox, oy = obs[0], obs[1]
    a = max(abs(obs[2]), 1e-3); b = max(abs(obs[3]), 1e-3); e = max(abs(obs[4]), 2.0)
    c, s = math.cos(obs[5]), math.sin(obs[5])
    pxp = c * (px - ox) + s * (py - oy)
    pyp = -s * (px - ox) + c * (py - oy)
    return (abs(pxp) / (a + radius)) ** e + (abs(pyp) / (b + radius)) ** e - 1


def _h_dt_flag(px, py, obs, radius, beta):
    """`if_else(obs[6] < 0.5, circle, superellipsoid)` (dynamic_unicycle2D.py:222-228)."""
    if obs[6] < 0.5:
        return _circle_h(px, py, obs, radius, beta)
    return _superellipsoid_dt(px, py, obs, radius)


# ---------------------------------------------------------------------------
class SingleIntegrator2D(Model):
    nx, nu, rel_degree = 2, 2, 1
    beta = 1.01

    def f(self, X): return np.zeros(2)
    def g(self, X): return np.eye(2)
    def step(self, X, U): return X + U * self.dt
    def u_bounds(self):
        v = self.spec["v_max"]; return np.array([-v, -v]), np.array([v, v])

    def nominal_input(self, X, G, d_min=0.05, k_v=1.0):
        v_max = self.spec["v_max"]
        err = np.asarray(G[0:2], float) - X[0:2]
        err = np.sign(err) * np.maximum(np.abs(err) - d_min, 0.0)
        v_des = k_v * err
        mag = np.linalg.norm(v_des)
        if mag > v_max:
            v_des = v_des * v_max / mag
        return v_des

    def agent_barrier(self, X, obs):
        """-> h, dh_dx(2,)   (single_integrator2D.py:114-146)"""
        h, dh = 0.0, np.zeros(2)
        if obs[-1] == 0:
            h = _circle_h(X[0], X[1], obs, self.radius, self.beta)
            dh = 2 * (X[0:2] - obs[0:2])
        elif obs[-1] == 1:
            h, dh, _ = _superellipsoid(X[0], X[1], obs, self.radius)
        return h, dh

    def barrier_dt(self, x, u, obs):
        """-> h_k, d_h   (single_integrator2D.py:148-195)"""
        x1 = self.step(x, u)
        hk = _h_dt_flag(x[0], x[1], obs, self.radius, self.beta)
        h1 = _h_dt_flag(x1[0], x1[1], obs, self.radius, self.beta)
        return hk, h1 - hk


class Unicycle2D(Model):
    """robots/unicycle2D.py: X = [x, y, theta], U = [v, omega]; relative degree 1 through the sigma(s) term."""
    nx, nu, rel_degree = 3, 2, 1
    beta = 1.01
    k1, k2 = 0.5, 1.8                                           # :36-37

    def f(self, X): return np.zeros(3)                          # :43-51
    def g(self, X):                                             # :53-63
        return np.array([[math.cos(X[2]), 0.0], [math.sin(X[2]), 0.0], [0.0, 1.0]])

    def step(self, X, U, wrap=True):                            # :65-68
        Xn = X + (self.f(X) + self.g(X) @ U) * self.dt
        if wrap:
            Xn[2] = angle_normalize(Xn[2])
        return Xn

    def u_bounds(self):                                         # cbf_qp.py:58-61, mpc_cbf.py:188-192
        v, w = self.spec["v_max"], self.spec["w_max"]
        return np.array([-v, -w]), np.array([v, w])

    def nominal_input(self, X, G, d_min=0.05, k_omega=2.0, k_v=1.0):     # :70-86
        distance = max(np.linalg.norm(X[0:2] - np.asarray(G[0:2], float)) - d_min, 0.05)
        theta_d = math.atan2(G[1] - X[1], G[0] - X[0])
        err = angle_normalize(theta_d - X[2])
        omega = k_omega * err
        v = 0.0 if abs(err) > np.deg2rad(90) else k_v * distance * math.cos(err)
        return np.array([v, omega])

    def sigma(self, s):                                         # :100-102
        return self.k2 * (math.exp(self.k1 - s) - 1) / (math.exp(self.k1 - s) + 1)

    def sigma_der(self, s):                                     # :104-105
        return -self.k2 * math.exp(self.k1 - s) / (1 + math.exp(self.k1 - s)) * (1 - self.sigma(s) / self.k2)

    def agent_barrier(self, X, obs):
        """-> h, dh_dx(3,)   (unicycle2D.py:107-125; circle only, the flag column is never read)"""
        d = X[0:2] - obs[0:2]
        c, s_ = math.cos(X[2]), math.sin(X[2])
        h = d @ d - self.beta * (obs[2] + self.radius) ** 2
        s = d[0] * c + d[1] * s_
        h = h - self.sigma(s)
        ds = self.sigma_der(s)
        dh = np.array([2 * d[0] - ds * c, 2 * d[1] - ds * s_, -ds * (-s_ * d[0] + c * d[1])])
        return h, dh

    def barrier_dt(self, x, u, obs):
        """-> h_k, d_h   (unicycle2D.py:127-145: plain circle h, no sigma term)"""
        x1 = self.step(x, u)
        hk = _circle_h(x[0], x[1], obs, self.radius, self.beta)
        h1 = _circle_h(x1[0], x1[1], obs, self.radius, self.beta)
        return hk, h1 - hk


class Manipulator2D(Model):
    """robots/manipulator2D.py: 3-link planar arm, X = joint angles, U = joint velocities (q_dot = u).
    agent_barrier returns one row PER LINK CIRCLE (25 per obstacle), cbf_qp.py:131-149 stacks them until its
    num_obs rows are full."""
    nx, nu, rel_degree = 3, 3, 1
    beta = 1.3                                                   # agent_barrier's default (:189)
    link_lengths = np.array([80, 70, 50]) / 60.0                 # :15-16
    step_len = 10.0 / 60.0                                       # :126

    def f(self, X): return np.zeros(3)                           # :24-27
    def g(self, X): return np.eye(3)                             # :29-32
    def step(self, X, U): return X + U * self.dt                 # :34-35

    def u_bounds(self):                                          # cbf_qp.py:96-105: |u| <= w_max
        w = self.spec["w_max"]; return np.full(3, -w), np.full(3, w)

    def joints(self, X):                                         # get_joint_positions :46-53
        P = [np.zeros(2)]; ang = 0.0
        for i in range(3):
            ang += X[i]
            P.append(P[-1] + self.link_lengths[i] * np.array([math.cos(ang), math.sin(ang)]))
        return P

    def jacobian(self, X):                                       # get_jacobian :55-92
        J = np.zeros((2, 3))
        for i in range(3):
            jx = jy = 0.0
            ang = sum(X[k] for k in range(i))
            for k in range(i, 3):
                ang += X[k]
                jx -= self.link_lengths[k] * math.sin(ang)
                jy += self.link_lengths[k] * math.cos(ang)
            J[0, i] = jx; J[1, i] = jy
        return J

    def nominal_input(self, X, G, d_min=0.05):                   # :94-108
        ee = self.joints(X)[-1]
        v = self.spec["Kp"] * (np.asarray(G, float).reshape(-1)[0:2] - ee)
        om = self.jacobian(X).T @ v
        return np.clip(om, -self.spec["w_max"], self.spec["w_max"])

    def link_circles(self, X):                                   # get_link_circles :110-138 -> [(x, y, link_idx)]
        out = []; p0 = np.zeros(2); ang = 0.0
        for i in range(3):
            ang += X[i]
            d = self.link_lengths[i] * np.array([math.cos(ang), math.sin(ang)])
            n = int(np.ceil(self.link_lengths[i] / self.step_len))       # 8, 8, 6 (7.000000000000001 -> 8)
            for j in range(n + 1):
                pos = p0 + (j / n) * d
                out.append((pos[0], pos[1], i))
            p0 = p0 + d
        return out

    def point_jacobian(self, X, pt, link_idx):                   # get_points_jacobian :140-160
        J = np.zeros((2, 3)); P = [np.zeros(2)]; ang = 0.0
        for i in range(link_idx + 1):
            if i > 0:
                ang += X[i - 1]
                P.append(P[-1] + self.link_lengths[i - 1] * np.array([math.cos(ang), math.sin(ang)]))
        for k in range(link_idx + 1):
            J[0, k] = -(pt[1] - P[k][1]); J[1, k] = pt[0] - P[k][0]
        return J

    def agent_barrier(self, X, obs):
        """-> h_list, dh_dx_list (25 each)   (manipulator2D.py:163-198)"""
        hs, dhs = [], []
        for cx, cy, li in self.link_circles(X):
            d_min = self.radius + obs[2]
            dx, dy = cx - obs[0], cy - obs[1]
            hs.append(dx * dx + dy * dy - self.beta * d_min ** 2)
            dhs.append(2 * np.array([dx, dy]) @ self.point_jacobian(X, (cx, cy), li))
        return hs, dhs

    def barrier_dt(self, x, u, obs):
        raise NotImplementedError("Manipulator2D has no agent_barrier_dt (no MPC in the reference)")


class DynamicUnicycle2D(Model):
    nx, nu, rel_degree = 4, 2, 2
    beta = 1.01

    def f(self, X):
        return np.array([X[3] * math.cos(X[2]), X[3] * math.sin(X[2]), 0.0, 0.0])

    def df_dx(self, X):
        c, s, v = math.cos(X[2]), math.sin(X[2]), X[3]
        return np.array([[0, 0, -v * s, c], [0, 0, v * c, s], [0, 0, 0, 0], [0, 0, 0, 0.0]])

    def g(self, X):
        return np.array([[0, 0], [0, 0], [0, 1], [1, 0.0]])

    def step(self, X, U, wrap=True):
        Xn = X + (self.f(X) + self.g(X) @ U) * self.dt
        if wrap:
            Xn[2] = angle_normalize(Xn[2])
        return Xn

    def u_bounds(self):
        a, w = self.spec["a_max"], self.spec["w_max"]
        return np.array([-a, -w]), np.array([a, w])

    def nominal_input(self, X, G, d_min=0.05, k_omega=2.0, k_a=1.0, k_v=1.0):
        k_omega = self.spec.get("nominal_k_omega", k_omega)
        k_a = self.spec.get("nominal_k_a", k_a)
        k_v = self.spec.get("nominal_k_v", k_v)
        v_max = self.spec["v_max"]
        distance = max(np.linalg.norm(X[0:2] - np.asarray(G[0:2], float)) - d_min, 0.0)
        theta_d = math.atan2(G[1] - X[1], G[0] - X[0])
        err = angle_normalize(theta_d - X[2])
        omega = k_omega * err
        if abs(err) > np.deg2rad(90):
            v = 0.0
        else:
            v = min(k_v * distance * math.cos(err), v_max)
        return np.array([k_a * (v - X[3]), omega])

    def agent_barrier(self, X, obs):
        """-> h, h_dot, dh_dot_dx(4,)   (dynamic_unicycle2D.py:121-186)"""
        h, h_dot, dhd = 0.0, 0.0, np.zeros(4)
        fx = self.f(X)
        if obs[-1] == 0:
            h = np.linalg.norm(X[0:2] - obs[0:2]) ** 2 - self.beta * (obs[2] + self.radius) ** 2
            h_dot = 2 * (X[0:2] - obs[0:2]) @ fx[0:2]
            dhd = np.append(2 * fx[0:2], [0.0, 0.0]) + 2 * (X[0:2] - obs[0:2]) @ self.df_dx(X)[0:2, :]
        elif obs[-1] == 1:
            R = self.radius
            a, b, e = obs[2], obs[3], obs[4]
            _, dh2, (pxp, pyp, c, s) = _superellipsoid(X[0], X[1], obs, R)
            h = (pxp / (a + R)) ** e + (pyp / (b + R)) ** e - 1
            dh = np.array([dh2[0], dh2[1], 0.0, 0.0])
            h_dot = dh @ fx
            ka = e * (e - 1) / (a + R) ** e * pxp ** (e - 2)
            kb = e * (e - 1) / (b + R) ** e * pyp ** (e - 2)
            ga = e / (a + R) ** e * pxp ** (e - 1)
            gb = e / (b + R) ** e * pyp ** (e - 1)
            ct, st, v = math.cos(X[2]), math.sin(X[2]), X[3]
            hxx = ka * c * c + kb * s * s
            hxy = (ka - kb) * c * s
            hyy = ka * s * s + kb * c * c
            gx = ga * c - gb * s
            gy = ga * s + gb * c
            dhd = np.array([hxx * v * ct + hxy * v * st,
                            hxy * v * ct + hyy * v * st,
                            gx * (-v * st) + gy * (v * ct),
                            gx * ct + gy * st])
        return h, h_dot, dhd

    def barrier_dt(self, x, u, obs):
        """-> h_k, d_h, dd_h   (dynamic_unicycle2D.py:188-238)"""
        x1 = self.step(x, u); x2 = self.step(x1, u)
        hk = _h_dt_flag(x[0], x[1], obs, self.radius, self.beta)
        h1 = _h_dt_flag(x1[0], x1[1], obs, self.radius, self.beta)
        h2 = _h_dt_flag(x2[0], x2[1], obs, self.radius, self.beta)
        return hk, h1 - hk, h2 - 2 * h1 + hk


class KinematicBicycle2D(DynamicUnicycle2D):
    nx, nu, rel_degree = 4, 2, 2
    beta = 1.1          # kinematic_bicycle2D.py:160

    def g(self, X):
        th, v, Lr = X[2], X[3], self.spec["rear_ax_dist"]
        return np.array([[0, -v * math.sin(th)], [0, v * math.cos(th)], [0, v / Lr], [1, 0.0]])

    def step(self, X, U, wrap=True):
        Xn = X + (self.f(X) + self.g(X) @ U) * self.dt
        if wrap:
            Xn[2] = angle_normalize(Xn[2])
        Xn[3] = min(max(Xn[3], self.spec["v_min"]), self.spec["v_max"])   # :116-121
        return Xn

    def u_bounds(self):
        a, b = self.spec["a_max"], self.spec["beta_max"]
        return np.array([-a, -b]), np.array([a, b])

    def beta_of_delta(self, delta):
        return math.atan(self.spec["rear_ax_dist"] / self.spec["wheel_base"] * math.tan(delta))

    def nominal_input(self, X, G, d_min=0.05, k_theta=0.5, k_a=1.5, k_v=0.5):
        """NB the facade (robots/robot.py:406-407) passes (d_min, k_omega, k_a, k_v)
        positionally, so k_theta receives the facade's k_omega (default 2.0)."""
        v_max, v_min, dmax = self.spec["v_max"], self.spec["v_min"], self.spec["delta_max"]
        distance = max(np.linalg.norm(X[0:2] - np.asarray(G[0:2], float)) - d_min, 0.05)
        theta_d = math.atan2(G[1] - X[1], G[0] - X[0])
        err = angle_normalize(theta_d - X[2])
        delta = min(max(k_theta * err, -dmax), dmax)
        beta = self.beta_of_delta(delta)
        v = min(max(k_v * distance * max(0.0, math.cos(err)), v_min), v_max)
        return np.array([k_a * (v - X[3]), beta])

    def agent_barrier(self, X, obs):
        """circle only, flag ignored   (kinematic_bicycle2D.py:160-173)"""
        fx = self.f(X)
        h = np.linalg.norm(X[0:2] - obs[0:2]) ** 2 - self.beta * (obs[2] + self.radius) ** 2
        h_dot = 2 * (X[0:2] - obs[0:2]) @ fx[0:2]
        dhd = np.append(2 * fx[0:2], [0.0, 0.0]) + 2 * (X[0:2] - obs[0:2]) @ self.df_dx(X)[0:2, :]
        return h, h_dot, dhd

    def barrier_dt(self, x, u, obs):
        """(kinematic_bicycle2D.py:175-199) circle h, model's own clipped step."""
        x1 = self.step(x, u); x2 = self.step(x1, u)
        hk = _circle_h(x[0], x[1], obs, self.radius, self.beta)
        h1 = _circle_h(x1[0], x1[1], obs, self.radius, self.beta)
        h2 = _circle_h(x2[0], x2[1], obs, self.radius, self.beta)
        return hk, h1 - hk, h2 - 2 * h1 + hk


class KinematicBicycle2D_C3BF(KinematicBicycle2D):
    rel_degree = 1

    def agent_barrier(self, X, obs):
        """-> h, dh_dx(4,)  collision-cone CBF (kinematic_bicycle2D_c3bf.py:15-75),
        hand-written gradient with its +eps terms transcribed literally."""
        th, v = X[2], X[3]
        ovx, ovy = (obs[3], obs[4]) if len(obs) > 3 else (0.0, 0.0)
        ego = (obs[2] + self.radius) * 1.0
        px, py = obs[0] - X[0], obs[1] - X[1]
        vx, vy = ovx - v * math.cos(th), ovy - v * math.sin(th)
        pm = math.sqrt(px * px + py * py)
        vm = math.sqrt(vx * vx + vy * vy)
        eps = 1e-6
        sq = math.sqrt(max(pm ** 2 - ego ** 2, eps))
        cos_phi = sq / (pm + eps)
        h = (px * vx + py * vy) + pm * vm * cos_phi
        dh = np.array([
            -vx - vm * px / (sq + eps),
            -vy - vm * py / (sq + eps),
            v * math.sin(th) * px - v * math.cos(th) * py + (sq + eps) / vm * (v * (ovx * math.sin(th) - ovy * math.cos(th))),
            -math.cos(th) * px - math.sin(th) * py + (sq + eps) / vm * (v - (ovx * math.cos(th) + ovy * math.sin(th)))])
        return h, dh

    def _h_dt(self, x, obs):
        """(kinematic_bicycle2D_c3bf.py:82-108) NB beta=1.01 here, no eps."""
        th, v = x[2], x[3]
        ovx, ovy = (obs[3], obs[4]) if len(obs) > 3 else (0.0, 0.0)
        ego = (obs[2] + self.radius) * 1.01
        px, py = obs[0] - x[0], obs[1] - x[1]
        vx, vy = ovx - v * math.cos(th), ovy - v * math.sin(th)
        pm = math.sqrt(px * px + py * py); vm = math.sqrt(vx * vx + vy * vy)
        return (px * vx + py * vy) + pm * vm * math.sqrt(max(pm ** 2 - ego ** 2, 0.0)) / pm

    def barrier_dt(self, x, u, obs):
        x1 = self.step(x, u)
        hk = self._h_dt(x, obs)
        return hk, self._h_dt(x1, obs) - hk


class KinematicBicycle2D_DPCBF(KinematicBicycle2D):
    """Dynamic-parabolic CBF (dynamic_env/kinematic_bicycle2D_dpcbf.py).  The hand-written dh/dx is not the
    gradient of h: both are transcribed literally."""
    rel_degree = 1
    k_lambda, k_mu, s_margin = 0.1, 0.5, 1.05

    def _parts(self, x, obs):
        th, v = x[2], x[3]
        ovx, ovy = (obs[3], obs[4]) if len(obs) > 3 else (0.0, 0.0)
        ego = (obs[2] + self.radius) * self.s_margin
        px, py = obs[0] - x[0], obs[1] - x[1]
        vx, vy = ovx - v * math.cos(th), ovy - v * math.sin(th)
        pm = math.sqrt(px * px + py * py); vm = math.sqrt(vx * vx + vy * vy)
        rot = math.atan2(py, px)
        vnx = math.cos(rot) * vx + math.sin(rot) * vy
        vny = -math.sin(rot) * vx + math.cos(rot) * vy
        d_safe = max(pm ** 2 - ego ** 2, 1e-6)
        return th, v, ovx, ovy, ego, px, py, pm, vm, rot, vnx, vny, d_safe

    def agent_barrier(self, X, obs):
        """-> h, dh_dx(4,)   (:16-84)"""
        th, v, ovx, ovy, ego, px, py, pm, vm, rot, vnx, vny, d_safe = self._parts(X, obs)
        kl, km, s = self.k_lambda, self.k_mu, self.s_margin
        sd = math.sqrt(d_safe)
        lam = kl * sd / vm * math.sqrt(s ** 2 - 1) / ego
        mu = km * sd * math.sqrt(s ** 2 - 1) / ego
        h = vnx + lam * vny ** 2 + mu
        dh = np.array([
            py * vny / pm ** 2 - kl * px * vny ** 2 / vm / sd - 2 * kl * sd / vm * vny * py / pm ** 2 * vnx - km * px / sd,
            -px * vny / pm ** 2 - kl * py * vny ** 2 / vm / sd + 2 * kl * sd / vm * vny * px / pm ** 2 * vnx - km * py / sd,
            -v * math.sin(rot - th) - kl * sd * v * (ovx * math.sin(th) - ovy * math.cos(th)) * vny ** 2 / vm ** 3
            - 2 * kl * sd * vny * v * math.cos(rot - th) / vm,
            -math.cos(rot - th) - kl * sd / vm ** 3 * (v - ovx * math.cos(th) - ovy * math.sin(th)) * vny ** 2
            - 2 * kl * sd * vny * math.sin(rot - th) / vm])
        return h, dh

    def _h_dt(self, x, obs):
        """(:86-136)"""
        th, v, ovx, ovy, ego, px, py, pm, vm, rot, vnx, vny, d_safe = self._parts(x, obs)
        s = self.s_margin
        k_l, k_m = 0.1 * math.sqrt(s ** 2 - 1) / ego, 0.5 * math.sqrt(s ** 2 - 1) / ego
        return vnx + k_l * math.sqrt(d_safe) / vm * vny ** 2 + k_m * math.sqrt(d_safe)

    def barrier_dt(self, x, u, obs):
        x1 = self.step(x, u)
        hk = self._h_dt(x, obs)
        return hk, self._h_dt(x1, obs) - hk


class DoubleIntegrator2D(Model):
    """X = [x, y, vx, vy] (yaw kept outside the state), U = [ax, ay]   (robots/double_integrator2D.py)."""
    nx, nu, rel_degree = 4, 2, 2
    beta = 1.01

    def f(self, X): return np.array([X[2], X[3], 0.0, 0.0])
    def g(self, X): return np.array([[0.0, 0], [0, 0], [1, 0], [0, 1]])

    def step(self, X, U):                                      # :82-110 (velocity magnitude clipped to v_max)
        Xn = X + (self.f(X) + self.g(X) @ U) * self.dt
        v_max = self.spec.get("v_max")
        if v_max is not None:
            vm = math.sqrt(Xn[2] ** 2 + Xn[3] ** 2)
            if vm > v_max:
                Xn[2] *= v_max / vm; Xn[3] *= v_max / vm
        return Xn

    def u_bounds(self):
        a = self.spec["a_max"]; return np.array([-a, -a]), np.array([a, a])

    def nominal_input(self, X, G, d_min=0.05, k_v=1.0, k_a=1.0):     # :116-143
        k_v = self.spec.get("nominal_k_v", k_v); k_a = self.spec.get("nominal_k_a", k_a)
        v_max, a_max = self.spec["v_max"], self.spec["a_max"]
        err = np.asarray(G[0:2], float) - X[0:2]
        err = np.sign(err) * np.maximum(np.abs(err) - d_min, 0.0)
        v_des = k_v * err
        mag = np.linalg.norm(v_des)
        if mag > v_max:
            v_des = v_des * v_max / mag
        a = k_a * (v_des - X[2:4])
        am = np.linalg.norm(a)
        if am > a_max:
            a = a * a_max / am
        return a

    def agent_barrier(self, X, obs):
        """-> h, h_dot, dh_dot_dx(4,)   (:167-222)"""
        h, h_dot, dhd = 0.0, 0.0, np.zeros(4)
        if obs[-1] == 0:
            h = np.linalg.norm(X[0:2] - obs[0:2]) ** 2 - self.beta * (obs[2] + self.radius) ** 2
            h_dot = 2 * (X[0:2] - obs[0:2]) @ X[2:4]
            dhd = np.append(2 * X[2:4], 2 * (X[0:2] - obs[0:2]))
        elif obs[-1] == 1:
            R = self.radius
            a, b, e = obs[2], obs[3], obs[4]
            h, dh2, (pxp, pyp, c, s) = _superellipsoid(X[0], X[1], obs, R)
            h_dot = dh2 @ X[2:4]
            ka = e * (e - 1) / (a + R) ** e * pxp ** (e - 2)
            kb = e * (e - 1) / (b + R) ** e * pyp ** (e - 2)
            dhd = np.array([(ka * c * c + kb * s * s) * X[2] + ((ka - kb) * c * s) * X[3],
                            ((ka - kb) * c * s) * X[2] + (ka * s * s + kb * c * c) * X[3], dh2[0], dh2[1]])
        return h, h_dot, dhd

    def barrier_dt(self, x, u, obs):
        x1 = self.step(x, u); x2 = self.step(x1, u)
        hk = _h_dt_flag(x[0], x[1], obs, self.radius, self.beta)
        h1 = _h_dt_flag(x1[0], x1[1], obs, self.radius, self.beta)
        h2 = _h_dt_flag(x2[0], x2[1], obs, self.radius, self.beta)
        return hk, h1 - hk, h2 - 2 * h1 + hk


class Quad2D(Model):
    """X = [x, z, theta, vx, vz, theta_dot], U = [f_right, f_left]   (robots/quad2D.py)."""
    nx, nu, rel_degree = 6, 2, 2
    beta = 1.01
    gravity = 9.81

    def f(self, X): return np.array([X[3], X[4], X[5], 0.0, -self.gravity, 0.0])

    def g(self, X):
        m, I, r = self.spec["mass"], self.spec["inertia"], self.radius
        th = X[2]
        return np.array([[0, 0, 0, -math.sin(th) / m, math.cos(th) / m, r / I],
                         [0, 0, 0, -math.sin(th) / m, math.cos(th) / m, -r / I]]).T

    def step(self, X, U):
        Xn = X + (self.f(X) + self.g(X) @ U) * self.dt
        Xn[2] = angle_normalize(Xn[2])
        return Xn

    def u_bounds(self):
        return np.full(2, self.spec["f_min"]), np.full(2, self.spec["f_max"])

    def nominal_input(self, X, G, k_px=3.0, k_dx=0.5, k_pz=0.1, k_dz=0.5, k_p_theta=0.05, k_d_theta=0.05):
        """cascaded PD law (quad2D.py:87-150)"""
        m, g = self.spec["mass"], 9.81
        f_min, f_max = self.spec["f_min"], self.spec["f_max"]
        r = self.radius
        x, z, theta, x_dot, z_dot, theta_dot = X
        a_d_x = k_px * (G[0] - x) + k_dx * (-x_dot)
        a_d_z = k_pz * (G[1] - z) + k_dz * (-z_dot) + g
        T = m * math.sqrt(a_d_x ** 2 + a_d_z ** 2)
        theta_d = -math.atan2(a_d_x, a_d_z)
        e = theta_d - theta
        e = math.atan2(math.sin(e), math.cos(e))
        tau = float(np.clip(k_p_theta * e + k_d_theta * (-theta_dot), -1, 1))
        return np.array([np.clip((T + tau / r) / 2.0, f_min, f_max), np.clip((T - tau / r) / 2.0, f_min, f_max)])

    def agent_barrier(self, X, obs):
        """-> h, h_dot, dh_dot_dx(6,)   (:166-177; circle only, flag ignored)"""
        d = X[0:2] - obs[0:2]
        h = np.linalg.norm(d) ** 2 - self.beta * (obs[2] + self.radius) ** 2
        h_dot = 2 * d @ X[3:5]
        dhd = np.array([2 * X[3], 2 * X[4], 0.0, 2 * d[0], 2 * d[1], 0.0])
        return h, h_dot, dhd

    def barrier_dt(self, x, u, obs):
        x1 = self.step(x, u); x2 = self.step(x1, u)
        hk = _circle_h(x[0], x[1], obs, self.radius, self.beta)
        h1 = _circle_h(x1[0], x1[1], obs, self.radius, self.beta)
        h2 = _circle_h(x2[0], x2[1], obs, self.radius, self.beta)
        return hk, h1 - hk, h2 - 2 * h1 + hk


class Quad3D(Model):
    nx, nu, rel_degree = 12, 4, 1
    beta = 1.01

    def __init__(self, robot_spec, dt=0.05):
        super().__init__(robot_spec, dt)
        s = self.spec
        self.gravity = 9.8
        L, nu = s["L"], s["nu"]
        self.B2 = np.array([[1, 1, 1, 1], [0, L, 0, -L], [L, 0, -L, 0], [nu, -nu, nu, -nu]], float)
        A = np.zeros((12, 12))
        for i in range(6):
            A[i, 6 + i] = 1
        A[6, 3] = self.gravity; A[7, 4] = -self.gravity
        B1 = np.zeros((12, 4))
        B1[8, 0] = 1 / s["mass"]; B1[9, 1] = 1 / s["Iy"]; B1[10, 2] = 1 / s["Ix"]; B1[11, 3] = 1 / s["Iz"]
        self.A, self.B = A, B1 @ self.B2

    def f(self, X): return self.A @ X
    def g(self, X): return self.B
    def u_bounds(self):
        return np.full(4, self.spec["u_min"]), np.full(4, self.spec["u_max"])

    def step(self, X, U, wrap=True):
        """RK4 + wrap of the three angles (quad3D.py:121-158)."""
        A, B, dt = self.A, self.B, self.dt
        k1 = A @ X + B @ U
        k2 = A @ (X + dt / 2 * k1) + B @ U
        k3 = A @ (X + dt / 2 * k2) + B @ U
        k4 = A @ (X + dt * k3) + B @ U
        Xn = X + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        if wrap:
            Xn[3:6] = angle_normalize(Xn[3:6])
        return Xn

    def nominal_input(self, X, goal, k_p=1.0, k_d=2.0, k_ang=5.0):
        s = self.spec
        pos_err = np.asarray(goal[0:3], float) - X[0:3]
        vel_err = -X[6:9]
        acc = k_p * pos_err + k_d * vel_err
        theta_des, phi_des, F_des = acc[0] / self.gravity, -acc[1] / self.gravity, s["mass"] * acc[2]
        tau_y = s["Iy"] * (k_ang * (theta_des - X[3]) + k_d * (-X[9]))
        tau_x = s["Ix"] * (k_ang * (phi_des - X[4]) + k_d * (-X[10]))
        tau_z = s["Iz"] * (k_ang * (0 - X[5]) + k_d * (-X[11]))
        u = np.linalg.pinv(self.B2) @ np.array([F_des, tau_y, tau_x, tau_z])
        return np.clip(u, s["u_min"], s["u_max"])

    def agent_barrier(self, X, obs):
        raise NotImplementedError("Cannot implement with nominal distance based CBF")  # quad3D.py:269-273

    def barrier_dt(self, x, u, obs):
        x1 = self.step(x, u)
        hk = _circle_h(x[0], x[1], obs, self.radius, self.beta)
        return hk, _circle_h(x1[0], x1[1], obs, self.radius, self.beta) - hk


_REGISTRY = {
    "SingleIntegrator2D": SingleIntegrator2D,
    "DynamicUnicycle2D": DynamicUnicycle2D,
    "KinematicBicycle2D": KinematicBicycle2D,
    "KinematicBicycle2D_C3BF": KinematicBicycle2D_C3BF,
    "Quad3D": Quad3D,
    "Unicycle2D": Unicycle2D,
    "Manipulator2D": Manipulator2D,
    "DoubleIntegrator2D": DoubleIntegrator2D,
    "Quad2D": Quad2D,
    "KinematicBicycle2D_DPCBF": KinematicBicycle2D_DPCBF,
}


def make_model(robot_spec, dt=0.05) -> Model:
    return _REGISTRY[robot_spec["model"]](robot_spec, dt)
