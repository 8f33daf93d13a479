"""Synthetic waypoint/obstacle scenes for tests and bench.py.

Vectorised numpy, float64, seeded.  Produces exactly what the reference's caller hands the
controller each step:
  X        robot.X                                                  (robots/robot.py:38)
  U_ref    robot.nominal_input(goal)                                (tracking.py:589-604)
  OBS/nobs get_nearest_unpassed_obs(..., obs_num=num_constraints)   (tracking.py:345-403), padded
           to M rows with the reference's own dummy row [1000, 1000, 0, 0, 0, 0, 0] (mpc_cbf.py:346)
  goal     current waypoint

Host-side input preparation only -- nothing here is on the timed hot path.
"""
import math

import numpy as np

DUMMY_OBS = np.array([1000.0, 1000.0, 0.0, 0.0, 0.0, 0.0, 0.0])
ANGLE_UNPASSED = {   # tracking.py:352-357
    "SingleIntegrator2D": 2.0 * np.pi, "Quad3D": 2.0 * np.pi, "DynamicUnicycle2D": 1.2 * np.pi,
    "KinematicBicycle2D": 2.0 * np.pi, "KinematicBicycle2D_C3BF": 2.0 * np.pi,
    "KinematicBicycle2D_DPCBF": 2.0 * np.pi, "DoubleIntegrator2D": 2.0 * np.pi, "Quad2D": 2.0 * np.pi,
    "Unicycle2D": 1.2 * np.pi, "VTOL2D": 1.2 * np.pi,
}
BARRIER_BETA = {"SingleIntegrator2D": 1.01, "DynamicUnicycle2D": 1.01, "KinematicBicycle2D": 1.1,
                "KinematicBicycle2D_C3BF": 1.1, "Quad3D": 1.01, "KinematicBicycle2D_DPCBF": 1.1,
                "DoubleIntegrator2D": 1.01, "Quad2D": 1.01, "Unicycle2D": 1.01, "VTOL2D": 1.01}


def angle_normalize(x):
    return ((x + np.pi) % (2 * np.pi)) - np.pi


# ------------------------------------------------------------------ nominal inputs (vectorised)
def nominal_input(model, spec, X, goal, optimal_decay=False):
    """Vectorised robots/<model>.py:nominal_input as called through the facade
    (robots/robot.py:401-415; optimal-decay gains from tracking.py:601-602)."""
    X = np.asarray(X, float); goal = np.asarray(goal, float)
    if model == "SingleIntegrator2D":                          # single_integrator2D.py:72-90
        err = goal[:, 0:2] - X[:, 0:2]
        err = np.sign(err) * np.maximum(np.abs(err) - 0.05, 0.0)
        mag = np.linalg.norm(err, axis=1, keepdims=True)
        vmax = spec["v_max"]
        return np.where(mag > vmax, err * vmax / np.maximum(mag, 1e-300), err)
    if model == "Unicycle2D":                                  # unicycle2D.py:70-86
        dist = np.maximum(np.linalg.norm(X[:, 0:2] - goal[:, 0:2], axis=1) - 0.05, 0.05)
        err = angle_normalize(np.arctan2(goal[:, 1] - X[:, 1], goal[:, 0] - X[:, 0]) - X[:, 2])
        v = np.where(np.abs(err) > np.deg2rad(90), 0.0, 1.0 * dist * np.cos(err))
        return np.stack([v, 2.0 * err], axis=1)
    if model == "DynamicUnicycle2D":                           # dynamic_unicycle2D.py:80-104
        k_omega, k_a, k_v = (3.0, 0.5, 0.5) if optimal_decay else (2.0, 1.0, 1.0)
        k_omega = spec.get("nominal_k_omega", k_omega); k_a = spec.get("nominal_k_a", k_a)
        k_v = spec.get("nominal_k_v", k_v)
        dist = np.maximum(np.linalg.norm(X[:, 0:2] - goal[:, 0:2], axis=1) - 0.05, 0.0)
        err = angle_normalize(np.arctan2(goal[:, 1] - X[:, 1], goal[:, 0] - X[:, 0]) - X[:, 2])
        v = np.where(np.abs(err) > np.deg2rad(90), 0.0, np.minimum(k_v * dist * np.cos(err), spec["v_max"]))
        return np.stack([k_a * (v - X[:, 3]), k_omega * err], axis=1)
    if model.startswith("KinematicBicycle2D"):                 # kinematic_bicycle2D.py:125-147
        # facade passes (d_min, k_omega, k_a, k_v) positionally -> k_theta = k_omega
        k_theta, k_a, k_v = (3.0, 0.5, 0.5) if optimal_decay else (2.0, 1.0, 1.0)
        dist = np.maximum(np.linalg.norm(X[:, 0:2] - goal[:, 0:2], axis=1) - 0.05, 0.05)
        err = angle_normalize(np.arctan2(goal[:, 1] - X[:, 1], goal[:, 0] - X[:, 0]) - X[:, 2])
        delta = np.clip(k_theta * err, -spec["delta_max"], spec["delta_max"])
        beta = np.arctan(spec["rear_ax_dist"] / spec["wheel_base"] * np.tan(delta))
        v = np.clip(k_v * dist * np.maximum(0.0, np.cos(err)), spec["v_min"], spec["v_max"])
        return np.stack([k_a * (v - X[:, 3]), beta], axis=1)
    if model == "DoubleIntegrator2D":                          # double_integrator2D.py:116-143
        err = goal[:, 0:2] - X[:, 0:2]
        err = np.sign(err) * np.maximum(np.abs(err) - 0.05, 0.0)
        mag = np.linalg.norm(err, axis=1, keepdims=True)
        v_des = np.where(mag > spec["v_max"], err * spec["v_max"] / np.maximum(mag, 1e-300), err)
        a = v_des - X[:, 2:4]
        am = np.linalg.norm(a, axis=1, keepdims=True)
        return np.where(am > spec["a_max"], a * spec["a_max"] / np.maximum(am, 1e-300), a)
    if model == "Quad2D":      # the reference's cascaded PD law is off the solve path: hover thrust + a seeded perturbation
        rng = np.random.default_rng(X.shape[0])
        hover = spec["mass"] * 9.81 / 2.0
        return hover + rng.uniform(-2.0, 2.0, (X.shape[0], 2))
    if model == "VTOL2D":                                      # vtol2D.py:446-448: "not implemented" -> zeros
        return np.zeros((X.shape[0], 4))
    if model == "Quad3D":                                      # quad3D.py:160-206
        g, m = 9.8, spec["mass"]
        k_p, k_d, k_ang = 1.0, 2.0, 5.0
        acc = k_p * (goal[:, 0:3] - X[:, 0:3]) + k_d * (-X[:, 6:9])
        th_des, ph_des, F = acc[:, 0] / g, -acc[:, 1] / g, m * acc[:, 2]
        tau_y = spec["Iy"] * (k_ang * (th_des - X[:, 3]) + k_d * (-X[:, 9]))
        tau_x = spec["Ix"] * (k_ang * (ph_des - X[:, 4]) + k_d * (-X[:, 10]))
        tau_z = spec["Iz"] * (k_ang * (0 - X[:, 5]) + k_d * (-X[:, 11]))
        L, nu = spec["L"], spec["nu"]
        B2 = np.array([[1, 1, 1, 1], [0, L, 0, -L], [L, 0, -L, 0], [nu, -nu, nu, -nu]], float)
        u = np.stack([F, tau_y, tau_x, tau_z], axis=1) @ np.linalg.pinv(B2).T
        return np.clip(u, spec["u_min"], spec["u_max"])
    raise ValueError(model)


# ------------------------------------------------------------------ obstacle selection (vectorised)
def nearest_unpassed_obs(model, pos, yaw, scene_obs, obs_num):
    """Vectorised tracking.py:345-403 for N agents over one shared scene.
    -> OBS [N, obs_num, 7] (padded with DUMMY_OBS), nobs [N] int32, idx [N, obs_num] (-1 = pad)."""
    N, K = pos.shape[0], scene_obs.shape[0]
    d = scene_obs[None, :, 0:2] - pos[:, None, 0:2]
    ang = np.arctan2(d[..., 1], d[..., 0])
    keep = np.abs(angle_normalize(ang - yaw[:, None])) <= ANGLE_UNPASSED[model] / 2
    none = ~keep.any(axis=1)
    keep[none] = True                                          # fallback: nearest of all (tracking.py:393-397)
    dist = np.linalg.norm(d, axis=2)
    dist = np.where(keep, dist, np.inf)
    order = np.argsort(dist, axis=1, kind="stable")[:, :obs_num]
    cnt = np.minimum(keep.sum(axis=1), obs_num).astype(np.int32)
    take = min(obs_num, K)
    valid = np.arange(take)[None, :] < cnt[:, None]
    OBS = np.tile(DUMMY_OBS, (N, obs_num, 1))
    sel = scene_obs[order]                                     # [N, take, 7]
    OBS[:, :take][valid] = sel[valid]
    idx = np.full((N, obs_num), -1, dtype=np.int64)
    idx[:, :take][valid] = order[valid]
    return OBS, cnt, idx


# ------------------------------------------------------------------ scenes
def default_spec(model):
    s = {"model": model, "radius": 0.25}
    if model == "SingleIntegrator2D":
        s.update(v_max=1.0, w_max=0.5)
    elif model == "DynamicUnicycle2D":
        s.update(a_max=0.5, w_max=0.5, v_max=1.0)
    elif model == "Unicycle2D":
        s.update(v_max=1.0, w_max=0.5)
    elif model.startswith("KinematicBicycle2D"):
        dmax = math.radians(32)
        s.update(wheel_base=0.4, front_ax_dist=0.2, rear_ax_dist=0.2, v_max=3.5, a_max=5.0, delta_max=dmax,
                 beta_max=math.atan(0.5 * math.tan(dmax)), v_min=0.2)
    elif model == "Quad3D":
        s.update(mass=3.0, Ix=0.5, Iy=0.5, Iz=0.5, L=0.3, nu=0.1, u_max=10.0, u_min=-10.0)
    elif model == "DoubleIntegrator2D":
        s.update(a_max=1.0, v_max=1.0, w_max=0.5)
    elif model == "Quad2D":
        s.update(mass=1.0, inertia=0.01, f_min=1.0, f_max=10.0)
    elif model == "VTOL2D":
        s.update(v_max=15.0, pitch_max=15.0, descent_speed_max=5.0, throttle_min=0.0, throttle_max=1.0, elevator_min=-0.5,
                 elevator_max=0.5)
    return s


def with_superellipsoids(sc, every=3, seed=7):
    """Turn every `every`-th obstacle row of each agent's list into a superellipsoid [x, y, a, b, e, theta, 1]
    (README.md:133-138) of about the same size: exercises the if_else(obs[6] < 0.5, ...) branch of the discrete
    barriers (single_integrator2D.py / dynamic_unicycle2D.py / double_integrator2D.py agent_barrier_dt)."""
    rng = np.random.default_rng(seed)
    OBS = sc["OBS"].copy()
    N, M, _ = OBS.shape
    sel = (np.arange(M)[None, :] % every == 0) & (np.arange(M)[None, :] < sc["nobs"][:, None])
    r = OBS[..., 2]
    a = r * rng.uniform(0.7, 1.1, (N, M)); b = r * rng.uniform(0.7, 1.1, (N, M))
    e = rng.choice([2.0, 4.0, 6.0], (N, M)); th = rng.uniform(-np.pi, np.pi, (N, M))
    for col, v in ((2, a), (3, b), (4, e), (5, th), (6, np.ones((N, M)))):
        OBS[..., col] = np.where(sel, v, OBS[..., col])
    out = dict(sc); out["OBS"] = np.ascontiguousarray(OBS)
    out["spec"] = dict(sc["spec"], mpc_superellipsoid=True)
    return out


def make_manipulator_scene(N, M, seed=1234, dense=False, spec=None):
    """Manipulator2D (robots/manipulator2D.py): N arms at the origin, joint angles uniform, M = CBFQP's row budget
    (25 link-circle rows per obstacle, cbf_qp.py:131-149) and also the slot count of OBS; every arm gets its own
    ceil(M / 25) (+1: one more than fits) circular obstacles placed around its reach, u_ref = the reference's
    Jacobian-transpose law towards a random goal, scaled so that the box is sometimes active."""
    rng = np.random.default_rng(seed)
    spec = dict({"model": "Manipulator2D", "radius": 0.25, "w_max": 2.0, "Kp": 3.0}, **(spec or {}))
    L = np.array([80, 70, 50]) / 60.0
    X = rng.uniform(-np.pi, np.pi, (N, 3))
    ang = np.cumsum(X, axis=1)
    J = np.concatenate([np.zeros((N, 1, 2)), np.cumsum(L[None, :, None] * np.stack([np.cos(ang), np.sin(ang)], -1), axis=1)], axis=1)
    ee = J[:, 3]
    k = min(M, (M + 24) // 25 + 1)
    OBS = np.zeros((N, M, 7))
    r_lo, r_hi = (0.8, 2.6) if dense else (1.5, 4.0)
    a = rng.uniform(-np.pi, np.pi, (N, k)); rad = rng.uniform(r_lo, r_hi, (N, k))
    OBS[:, :k, 0] = rad * np.cos(a); OBS[:, :k, 1] = rad * np.sin(a); OBS[:, :k, 2] = rng.uniform(0.1, 0.4, (N, k))
    nobs = rng.integers(0, k + 1, N).astype(np.int32)
    goal = rng.uniform(-3, 3, (N, 2))
    v = spec["Kp"] * (goal - ee)
    U_ref = np.zeros((N, 3))
    for i in range(3):                                   # J^T v with the geometric Jacobian of the end effector (:55-108)
        U_ref[:, i] = -(ee[:, 1] - J[:, i, 1]) * v[:, 0] + (ee[:, 0] - J[:, i, 0]) * v[:, 1]
    U_ref = np.clip(U_ref, -1.3 * spec["w_max"], 1.3 * spec["w_max"])
    return dict(model="Manipulator2D", spec=spec, X=np.ascontiguousarray(X), U_ref=np.ascontiguousarray(U_ref), goal=goal,
                OBS=np.ascontiguousarray(OBS), nobs=nobs, obs_idx=None, scene_obs=None, u_prev=np.zeros((N, 3)), L=4.0)


def make_scene(model, N, M, seed=1234, dense=False, dynamic=None, optimal_decay=False, spec=None):
    """-> dict(X, U_ref, goal, OBS [N,M,7], nobs [N] i32, scene_obs [M,7], u_prev, spec)"""
    if model == "Manipulator2D":
        return make_manipulator_scene(N, M, seed, dense, spec)
    rng = np.random.default_rng(seed)
    spec = dict(default_spec(model), **(spec or {}))
    if dynamic is None:
        dynamic = model.endswith("C3BF") or model.endswith("DPCBF")
    L = (2.5 if dense else 4.0) * math.sqrt(M)
    scene = np.zeros((M, 7))
    scene[:, 0:2] = rng.uniform(0, L, (M, 2))
    scene[:, 2] = rng.uniform(0.2, 0.6, M)
    if dynamic:
        scene[:, 3:5] = rng.uniform(-0.5, 0.5, (M, 2))
    R, beta = spec["radius"], BARRIER_BETA[model]
    # agents: rejection-sample positions so every h > 0.05 (start safe)
    pos = np.empty((N, 2)); todo = np.arange(N)
    while todo.size:
        cand = rng.uniform(0, L, (todo.size, 2))
        d2 = ((cand[:, None, :] - scene[None, :, 0:2]) ** 2).sum(-1)
        h = d2 - beta * (scene[None, :, 2] + R) ** 2
        ok = (h > 0.05).all(axis=1)
        pos[todo[ok]] = cand[ok]
        todo = todo[~ok]
    goal2 = rng.uniform(0, L, (N, 2))
    nx = {"SingleIntegrator2D": 2, "Quad3D": 12, "Quad2D": 6, "Unicycle2D": 3, "VTOL2D": 6}.get(model, 4)
    X = np.zeros((N, nx)); X[:, 0:2] = pos
    yaw = np.zeros(N)
    if model == "DoubleIntegrator2D":
        sp = rng.uniform(0, spec["v_max"], N); hd = rng.uniform(-np.pi, np.pi, N)
        if dense:   # head at the nearest obstacle
            d = scene[None, :, 0:2] - pos[:, None, :]
            j = np.argmin((d ** 2).sum(-1), axis=1)
            hd = np.arctan2(d[np.arange(N), j, 1], d[np.arange(N), j, 0]) + rng.normal(0, 0.2, N)
            sp = rng.uniform(0.5, 1.0, N) * spec["v_max"]
        X[:, 2] = sp * np.cos(hd); X[:, 3] = sp * np.sin(hd)
        yaw = rng.uniform(-np.pi, np.pi, N)
        goal = goal2
    elif model == "VTOL2D":       # cruise: forward speed 6..12 m/s, small pitch; the goal lies 20..40 m ahead at a similar height
        X[:, 2] = rng.uniform(-0.1, 0.1, N); X[:, 3] = rng.uniform(6.0, 12.0, N); X[:, 4] = rng.uniform(-0.5, 0.5, N)
        X[:, 5] = rng.normal(0, 0.05, N)
        yaw = X[:, 2].copy()
        goal = np.stack([pos[:, 0] + rng.uniform(20, 40, N), np.maximum(pos[:, 1] + rng.uniform(-2, 2, N), 2.0)], axis=1)
    elif model == "Quad2D":
        X[:, 2] = rng.uniform(-0.4, 0.4, N); X[:, 3:5] = rng.uniform(-1.0, 1.0, (N, 2)); X[:, 5] = rng.normal(0, 0.2, N)
        if dense:
            d = scene[None, :, 0:2] - pos[:, None, :]
            j = np.argmin((d ** 2).sum(-1), axis=1)
            dirn = d[np.arange(N), j]; dirn /= np.linalg.norm(dirn, axis=1, keepdims=True)
            X[:, 3:5] = dirn * rng.uniform(0.5, 1.5, (N, 1))
        yaw = X[:, 2].copy()
        goal = goal2
    elif model == "Unicycle2D":
        theta = rng.uniform(-np.pi, np.pi, N)
        if dense:   # head at the nearest obstacle
            d = scene[None, :, 0:2] - pos[:, None, :]
            j = np.argmin((d ** 2).sum(-1), axis=1)
            theta = angle_normalize(np.arctan2(d[np.arange(N), j, 1], d[np.arange(N), j, 0]) + rng.normal(0, 0.2, N))
        X[:, 2] = theta; yaw = theta
        goal = goal2
    elif nx == 4:
        theta = rng.uniform(-np.pi, np.pi, N)
        if model == "DynamicUnicycle2D":
            v = rng.uniform(0, spec["v_max"], N)
        else:
            v = rng.uniform(spec["v_min"], spec["v_max"], N)
        if dense:   # head at the nearest obstacle, fast
            d = scene[None, :, 0:2] - pos[:, None, :]
            j = np.argmin((d ** 2).sum(-1), axis=1)
            theta = np.arctan2(d[np.arange(N), j, 1], d[np.arange(N), j, 0]) + rng.normal(0, 0.2, N)
            theta = angle_normalize(theta)
            v = rng.uniform(0.5, 1.0, N) * spec["v_max"]
        X[:, 2] = theta; X[:, 3] = v; yaw = theta
        goal = goal2
    elif nx == 12:
        X[:, 2] = rng.uniform(1, 3, N)
        X[:, 3:6] = rng.normal(0, 0.05, (N, 3)); X[:, 6:9] = rng.normal(0, 0.5, (N, 3))
        X[:, 9:12] = rng.normal(0, 0.05, (N, 3)); yaw = X[:, 5].copy()
        goal = np.concatenate([goal2, rng.uniform(1, 3, (N, 1))], axis=1)
    else:
        goal = goal2
    U_ref = nominal_input(model, spec, X, goal, optimal_decay=optimal_decay)
    OBS, nobs, idx = nearest_unpassed_obs(model, pos, yaw, scene, M)
    nu = 4 if model in ("Quad3D", "VTOL2D") else 2
    return dict(model=model, spec=spec, X=np.ascontiguousarray(X), U_ref=np.ascontiguousarray(U_ref),
                goal=np.ascontiguousarray(goal), OBS=np.ascontiguousarray(OBS), nobs=nobs, obs_idx=idx,
                scene_obs=scene, u_prev=np.zeros((N, nu)), L=L)


def make_evade_batch(n, seed=1234, k_mov=1, hallway_length=60.0, pocket=(25.0, 35.0, 2.0, 6.0), a_max=2.0, v_max=1.5):
    """Seeded batch for the Backup-CBF QP path (examples/evade/test_evade.py's scene): states spread over the hallway (70 %)
    and the pocket / its mouth (30 %), speeds up to v_max, nominal inputs up to 1.25 a_max, one bullet per agent somewhere
    between its spawn point and the end of the hallway (10 % inactive) [+ optional slow discs].
    -> X [n, 4], U_ref [n, 2], MOV [n, k_mov, 8] (safe_control_b200.backup row layout)."""
    from .backup import bullet_row, KIND_CIRCLE, KIND_NONE
    rng = np.random.default_rng(seed)
    X = np.zeros((n, 4)); MOV = np.zeros((n, max(k_mov, 1), 8))
    X[:, 0] = rng.uniform(0.3, hallway_length - 0.3, n)
    in_pocket = rng.random(n) < 0.3
    X[:, 0] = np.where(in_pocket, rng.uniform(pocket[0] + 0.2, pocket[1] - 0.2, n), X[:, 0])
    X[:, 1] = np.where(in_pocket, rng.uniform(0.0, pocket[3] - 0.6, n), rng.uniform(-1.6, 1.6, n))
    X[:, 2:] = rng.uniform(-1.0, 1.0, (n, 2)) * rng.uniform(0.0, v_max, (n, 1))
    U_ref = rng.uniform(-1.25 * a_max, 1.25 * a_max, (n, 2))
    bx = rng.uniform(-10.0, hallway_length + 2.0, n); act = rng.random(n) < 0.9
    base = bullet_row(0.0)
    MOV[:, 0, :] = base[None, :]
    MOV[:, 0, 0] = bx + base[0]
    MOV[:, 0, 7] = np.where(act, base[7], float(KIND_NONE))
    for k in range(1, k_mov):
        MOV[:, k, 0] = rng.uniform(0, hallway_length, n); MOV[:, k, 1] = rng.uniform(-1.5, 1.5, n)
        MOV[:, k, 2] = rng.uniform(-0.5, 0.5, n); MOV[:, k, 3] = rng.uniform(-0.2, 0.2, n)
        MOV[:, k, 6] = rng.uniform(0.2, 0.8, n)
        MOV[:, k, 7] = np.where(rng.random(n) < 0.5, float(KIND_CIRCLE), float(KIND_NONE))
    return np.ascontiguousarray(X), np.ascontiguousarray(U_ref), np.ascontiguousarray(MOV)


def make_evade_plans(X, T=100, dt=0.1, a_max=2.0, v_max=1.5):
    """Nominal plans of the evade example for a batch (examples/evade/test_evade.py:141-168, 387-408: PD law towards the
    hallway axis at v_max, rolled out with DoubleIntegrator2D.step), vectorised over agents.
    -> NOMX [n, T + 1, 4], NOMU [n, T, 2]"""
    n = X.shape[0]
    NOMX = np.zeros((n, T + 1, 4)); NOMU = np.zeros((n, T, 2))
    s = np.array(X, dtype=np.float64); NOMX[:, 0] = s
    for k in range(T):
        ax = 2.0 * (v_max - s[:, 2]); ay = 2.0 * (0.0 - s[:, 1]) + 2.0 * (0.0 - s[:, 3])
        am = np.sqrt(ax ** 2 + ay ** 2); f = np.where(am > a_max, a_max / np.maximum(am, 1e-300), 1.0)
        ax, ay = ax * f, ay * f
        nxt = np.stack([s[:, 0] + s[:, 2] * dt, s[:, 1] + s[:, 3] * dt, s[:, 2] + ax * dt, s[:, 3] + ay * dt], axis=1)
        vm = np.sqrt(nxt[:, 2] ** 2 + nxt[:, 3] ** 2); g = np.where(vm > v_max, v_max / np.maximum(vm, 1e-300), 1.0)
        nxt[:, 2] *= g; nxt[:, 3] *= g
        NOMU[:, k, 0] = ax; NOMU[:, k, 1] = ay; NOMX[:, k + 1] = nxt; s = nxt
    return NOMX, NOMU


def evade_static_rects(MOV, bullet_length=3.0, bullet_width=4.0):
    """the bullet's current hitbox per agent (envs/evade_env.py:469-473) from the bullet rows of make_evade_batch"""
    bx = MOV[:, 0, 0] - bullet_length / 6
    S = np.zeros((MOV.shape[0], 5))
    S[:, 0] = bx - bullet_length / 2; S[:, 1] = bx + bullet_length / 2 + bullet_length / 3
    S[:, 2] = MOV[:, 0, 1] - bullet_width / 2; S[:, 3] = MOV[:, 0, 1] + bullet_width / 2; S[:, 4] = (MOV[:, 0, 7] != 0)
    return S
