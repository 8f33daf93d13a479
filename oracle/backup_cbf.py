"""Oracle restatement of the reference's Backup-CBF QP for the double integrator in the evade scene
(TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this).

  position_control/backup_cbf_qp.py:41-126   parameters (N = int(backup_horizon / dt), alpha = 1, alpha_terminal = 2, Q_u)
  position_control/backup_cbf_qp.py:236-318  _integrate_backup_trajectory: robot.step rollout under the backup policy,
                                             S_{k+1} = A_k S_k with A_k by forward differences (eps = 1e-5) of one CLOSED-LOOP step
  position_control/backup_cbf_qp.py:341-444  _h_safety: evade walls (hallway + pocket) and moving obstacles (rectangle / circle)
  position_control/backup_cbf_qp.py:446-458  _grad_h_safety: forward differences, eps = 1e-5
  position_control/backup_cbf_qp.py:460-539  _h_terminal: pocket box with margin radius + 0.2, v_max - |v|, h_safety at t = backup_horizon
  position_control/backup_cbf_qp.py:563-794  solve_control_problem: rows, variable scaling, QP, fall-backs
  position_control/backup_controller.py:420-575  EvadeBackupController.compute_control / _clamp_control
  robots/double_integrator2D.py:46-107       f, g, step (Euler + speed clamp)
  envs/evade_env.py:30-85, 386-406           geometry, bullet state (collision centre x + L/6, length 4L/3)

Everything that is finite-differenced is evaluated in the reference's own operation order (each product and sum rounded
separately, as numpy does): an error of one ulp in h is an error of 1e-11 .. 1e-10 in a gradient, so the CUDA kernel
(safe_control_b200/csrc/scb_backup.cuh) follows the same order with non-contracted arithmetic and the two agree to ~1e-12.

Moving obstacles: the reference takes a callable t -> dict (examples/evade/test_evade.py:373-385: x0 + vx * t, the bullet
flies at constant speed).  Here they are rows  [x, y, vx, vy, length, width, radius, kind]  (kind 0 = absent / inactive,
1 = rectangle, 2 = circle) advanced as  x + vx * t,  y + vy * t.

The QP (2 scaled variables, <= N rows + 4 box rows) is solved exactly by enumerating working sets (qp2_exact below).
The reference solves it with OSQP (approximate); the fixture generator runs the reference's own code with cvxpy replaced
by oracle/refshim/fake_cvxpy.py, i.e. with the exact optimum of the reference's own problem statement.
"""
import numpy as np

EPS_FD = 1e-5        # backup_cbf_qp.py:285, 448, 543


class EvadeScene:
    """The numbers the path reads from EvadeEnv, EvadeBackupController and robot_spec."""

    def __init__(self, hallway_length=60.0, hallway_width=4.0, pocket_x=25.0, pocket_length=10.0, pocket_width=4.0,
                 goal_length=5.0, radius=0.5, a_max=2.0, v_max=1.5, safety_margin=0.5, use_goal=True,
                 dt=0.1, backup_horizon=12.0):
        self.hallway_length = float(hallway_length)
        self.half_width = hallway_width / 2                                   # evade_env.py:62
        self.pocket_x_min = float(pocket_x)                                   # :65-68
        self.pocket_x_max = pocket_x + pocket_length
        self.pocket_y_min = self.half_width
        self.pocket_y_max = self.half_width + pocket_width
        self.center = ((self.pocket_x_min + self.pocket_x_max) / 2, (self.pocket_y_min + self.pocket_y_max) / 2)   # :69-72
        self.goal = None
        if use_goal:                                                          # test_evade.py:308-313
            self.goal = (hallway_length - goal_length, float(hallway_length), -self.half_width, self.half_width)
        self.radius, self.a_max, self.v_max, self.safety_margin = float(radius), float(a_max), float(v_max), float(safety_margin)
        self.Kp, self.Kd = 2.0, 2.0                                           # backup_controller.py:449-450
        self.dt, self.backup_horizon = float(dt), float(backup_horizon)
        self.N = int(backup_horizon / dt)                                     # backup_cbf_qp.py:56
        self.alpha, self.alpha_terminal = 1.0, 2.0                            # :91-92
        self.Q_u = (1.0, 1.0)                                                 # :110

    def as_vector(self):
        """the flat parameter block of the C ABI (include/scb.h scb_backup_params, same order)"""
        g = self.goal if self.goal is not None else (0.0, 0.0, 0.0, 0.0)
        return dict(hallway_length=self.hallway_length, half_width=self.half_width, pocket_x_min=self.pocket_x_min,
                    pocket_x_max=self.pocket_x_max, pocket_y_min=self.pocket_y_min, pocket_y_max=self.pocket_y_max,
                    center_x=self.center[0], center_y=self.center[1], goal_x_min=g[0], goal_x_max=g[1], goal_y_min=g[2],
                    goal_y_max=g[3], use_goal=int(self.goal is not None), radius=self.radius, a_max=self.a_max,
                    v_max=self.v_max, safety_margin=self.safety_margin, Kp=self.Kp, Kd=self.Kd, dt=self.dt,
                    backup_horizon=self.backup_horizon, N=self.N, alpha=self.alpha, alpha_terminal=self.alpha_terminal,
                    q0=self.Q_u[0], q1=self.Q_u[1])


def bullet_row(bullet_x, bullet_length=3.0, bullet_width=4.0, bullet_speed=3.0, bullet_y=0.0, active=True):
    """EvadeEnv.get_bullet_state (evade_env.py:386-406) as a mover row."""
    return np.array([bullet_x + (bullet_length / 6), bullet_y, bullet_speed, 0.0, bullet_length * (1 + 1 / 3), bullet_width,
                     0.0, 1.0 if active else 0.0])


def _clamp(sc, ax, ay):                                   # backup_controller.py:568-575
    a_mag = np.sqrt(ax ** 2 + ay ** 2)
    if a_mag > sc.a_max:
        ax = ax * sc.a_max / a_mag
        ay = ay * sc.a_max / a_mag
    return ax, ay


def backup_control(sc, s):
    """EvadeBackupController.compute_control (backup_controller.py:456-566)."""
    x, y, vx, vy = float(s[0]), float(s[1]), float(s[2]), float(s[3])
    if sc.goal is not None:                                                                # :476-484
        gx0, gx1, gy0, gy1 = sc.goal
        if gx0 <= x <= gx1 and gy0 <= y <= gy1:
            return _clamp(sc, -sc.Kd * vx, -sc.Kd * vy)
    x_min, x_max, y_min, y_max = sc.pocket_x_min, sc.pocket_x_max, sc.pocket_y_min, sc.pocket_y_max
    cx, cy = sc.center
    margin = sc.radius + 0.1                                                               # :496
    dist_to_center = np.sqrt((x - cx) ** 2 + (y - cy) ** 2)
    if x_min + margin <= x <= x_max - margin and y_min + margin <= y <= y_max - margin and dist_to_center < 1.0:   # :502-508
        ax = -sc.Kd * vx
        ay = -sc.Kd * vy
        return _clamp(sc, ax, ay)
    elif x_min - 2.0 <= x <= x_max + 2.0:                                                  # :513-542
        if x_min + margin <= x <= x_max - margin:
            error_x = cx - x
            error_y = cy - y
            ax = sc.Kp * error_x - sc.Kd * vx
            ay = sc.Kp * error_y - sc.Kd * vy
        else:
            error_x = cx - x
            target_y = max(y, 3.0) if y > y_min else 0.0
            error_y = target_y - y
            ax = sc.Kp * error_x - sc.Kd * vx
            ay = sc.Kp * error_y - sc.Kd * vy
    else:                                                                                  # :546-563
        target_y = max(y, 3.0) if (y > y_min and x > x_max) else 0.0
        error_x = cx - x
        error_y = target_y - y
        ax = sc.Kp * np.sign(error_x) * min(abs(error_x), 3.0) - sc.Kd * vx
        ay = sc.Kp * error_y - sc.Kd * vy
    return _clamp(sc, ax, ay)


def di_step(sc, s, u):
    """DoubleIntegrator2D.step (double_integrator2D.py:79-107): X + (f + g U) dt, then the speed clamp."""
    dt = sc.dt
    n = [s[0] + (s[2] + 0.0) * dt, s[1] + (s[3] + 0.0) * dt, s[2] + (0.0 + u[0]) * dt, s[3] + (0.0 + u[1]) * dt]
    v_mag = np.sqrt(n[2] ** 2 + n[3] ** 2)
    if v_mag > sc.v_max:
        scale = sc.v_max / v_mag
        n[2] *= scale
        n[3] *= scale
    return np.array(n, dtype=np.float64)


def closed_loop_step(sc, s):
    return di_step(sc, s, backup_control(sc, s))


def integrate(sc, x0):
    """_integrate_backup_trajectory (backup_cbf_qp.py:236-318) -> phi [N, 4], S [N, 4, 4]."""
    N = sc.N
    phi = np.zeros((N, 4)); S = np.zeros((N, 4, 4))
    x = np.array(x0, dtype=np.float64).reshape(-1).copy()
    S_curr = np.eye(4)
    phi[0] = x; S[0] = S_curr
    for i in range(1, N):
        x_next = closed_loop_step(sc, x)
        A = np.zeros((4, 4))
        for j in range(4):
            xp = x.copy(); xp[j] += EPS_FD
            A[:, j] = (closed_loop_step(sc, xp) - x_next) / EPS_FD
        S_curr = A @ S_curr
        x = x_next
        phi[i] = x; S[i] = S_curr
    return phi, S


def h_safety(sc, s, t, movers):
    """_h_safety (backup_cbf_qp.py:341-444), evade branch: walls of hallway + pocket, then the moving obstacles at time t."""
    px, py = float(s[0]), float(s[1])
    r = sc.radius
    h = float("inf")
    h = min(h, py + sc.half_width - r)                            # bottom :362
    h = min(h, px - r)                                            # left :366
    h = min(h, sc.hallway_length - px - r)                        # right :370
    if sc.pocket_x_min <= px <= sc.pocket_x_max:                  # :376-389
        h = min(h, sc.pocket_y_max - py - r)
        if py > sc.half_width:
            h = min(h, px - sc.pocket_x_min - r, sc.pocket_x_max - px - r)
    else:
        h = min(h, sc.half_width - py - r)
    if movers is not None:
        for o in np.asarray(movers, dtype=np.float64).reshape(-1, 8):                        # :419-442
            kind = int(o[7])
            if kind == 0:
                continue
            ox = o[0] + o[2] * t
            oy = o[1] + o[3] * t
            if kind == 1:
                dx = max(abs(px - ox) - o[4] / 2, 0)
                dy = max(abs(py - oy) - o[5] / 2, 0)
                dist = np.sqrt(dx ** 2 + dy ** 2)
                h = min(h, dist - r - sc.safety_margin)
            else:
                dist = np.sqrt((px - ox) ** 2 + (py - oy) ** 2)
                h = min(h, dist - r - o[6] - sc.safety_margin)
    return h if h != float("inf") else 1.0


def h_terminal(sc, s, movers):
    """_h_terminal (backup_cbf_qp.py:460-539), evade branch."""
    px, py = float(s[0]), float(s[1])
    margin = sc.radius + 0.2                                      # :480
    h = min(px - sc.pocket_x_min - margin, sc.pocket_x_max - px - margin,
            py - sc.pocket_y_min - margin, sc.pocket_y_max - py - margin)
    velocity = np.sqrt(s[2] ** 2 + s[3] ** 2)                     # :522
    h = min(h, sc.v_max - velocity)
    h = min(h, h_safety(sc, s, sc.backup_horizon, movers))        # :533-536
    return h


def _fd_grad(fun, s):
    s = np.array(s, dtype=np.float64).reshape(-1)
    h0 = fun(s)
    g = np.zeros(s.size)
    for i in range(s.size):
        sp = s.copy(); sp[i] += EPS_FD
        g[i] = (fun(sp) - h0) / EPS_FD
    return g


def rows(sc, x0, movers, phi=None, S=None):
    """The CBF rows of solve_control_problem (backup_cbf_qp.py:613-673) -> G [N, 2], h [N], keep [N] bool, h_min.

    Row i - 1 (i = 1 .. N-1) is the safety row of backup step i, row N - 1 the terminal row; `keep` is the reference's
    ||lhs|| > 1e-6 filter (rows it never hands to the QP).  G u >= h."""
    if phi is None:
        phi, S = integrate(sc, x0)
    N, dt = sc.N, sc.dt
    x0 = np.array(x0, dtype=np.float64).reshape(-1)
    f0 = np.array([x0[2], x0[3], 0.0, 0.0])
    g0 = np.array([[0, 0], [0, 0], [1, 0], [0, 1]], dtype=np.float64)
    G = np.zeros((N, 2)); hh = np.zeros(N); keep = np.zeros(N, dtype=bool)
    for i in range(1, N):
        t_i = i * dt
        h_val = h_safety(sc, phi[i], t_i, movers)
        grad = _fd_grad(lambda s: h_safety(sc, s, t_i, movers), phi[i])
        dh_dt = (h_safety(sc, phi[i], t_i + dt, movers) - h_val) / dt                      # :628-632 (a callable is always set)
        f_pi = (phi[i + 1] - phi[i]) / dt if i < N - 1 else (phi[i] - phi[i - 1]) / dt     # :636-639
        lhs = grad @ S[i] @ g0
        rhs = -(grad @ S[i] @ f0) + (grad @ f_pi) - dh_dt - sc.alpha * h_val
        G[i - 1] = lhs; hh[i - 1] = rhs; keep[i - 1] = np.linalg.norm(lhs) > 1e-6
    h_T = h_terminal(sc, phi[-1], movers)
    gT = _fd_grad(lambda s: h_terminal(sc, s, movers), phi[-1])
    lhs = gT @ S[-1] @ g0
    rhs = -(gT @ S[-1] @ f0 + sc.alpha_terminal * h_T)
    G[N - 1] = lhs; hh[N - 1] = rhs; keep[N - 1] = np.linalg.norm(lhs) > 1e-6
    h_vals = [h_safety(sc, phi[i], i * dt, movers) for i in range(N)]                      # :587-590
    h_min = min(float(np.min(h_vals)), h_T)
    return G, hh, keep, h_min


def qp2_exact(Qd, c, A, b, tol=1e-9):
    """min ||diag(Qd) (z - c)||^2  s.t.  A z >= b  (2 variables), by enumerating working sets of size 0, 1, 2.
    -> (z or None, working set as a tuple of row indices, strict-complementarity gap)."""
    A = np.asarray(A, float).reshape(-1, 2); b = np.asarray(b, float).reshape(-1)
    Qd = np.asarray(Qd, float); c = np.asarray(c, float)
    m = b.size
    scale = max(1.0, float(np.max(np.abs(b))) if m else 1.0)
    An = np.linalg.norm(A, axis=1)

    def feas(z):
        return np.all(A @ z - b >= -tol * scale * np.maximum(An, 1.0))

    Hinv = 1.0 / (2.0 * Qd ** 2)
    best = None

    def offer(z, W, lam):
        nonlocal best
        if not feas(z) or np.any(lam < -tol):
            return
        J = float(np.sum((Qd * (z - c)) ** 2))
        slack = A @ z - b
        inact = np.ones(m, bool); inact[list(W)] = False
        gap = min([float(np.min(lam))] if len(W) else [np.inf]) if len(W) else np.inf
        if np.any(inact):
            gap = min(gap, float(np.min(slack[inact] / np.maximum(An[inact], 1e-300))))
        if best is None or J < best[0] - 1e-14 or (abs(J - best[0]) <= 1e-14 and len(W) < len(best[2])):
            best = (J, z, tuple(W), gap)

    offer(c.copy(), (), np.zeros(0))
    for i in range(m):                                       # one active row: z = c + Hinv a lam, a.z = b
        a = A[i]
        d = float(a @ (Hinv * a))
        if d <= 0:
            continue
        lam = (b[i] - a @ c) / d
        offer(c + Hinv * a * lam, (i,), np.array([lam]))
    if m >= 2:                                               # two active rows: the vertex, multipliers from stationarity
        I, Jx = np.triu_indices(m, 1)
        det = A[I, 0] * A[Jx, 1] - A[I, 1] * A[Jx, 0]
        ok = np.abs(det) > 1e-12 * An[I] * An[Jx]
        I, Jx, det = I[ok], Jx[ok], det[ok]
        z0 = (b[I] * A[Jx, 1] - A[I, 1] * b[Jx]) / det
        z1 = (A[I, 0] * b[Jx] - b[I] * A[Jx, 0]) / det
        Z = np.stack([z0, z1], axis=1)
        viol = (Z @ A.T - b[None, :]) < -tol * scale * np.maximum(An, 1.0)[None, :]
        cand = np.where(~viol.any(axis=1))[0]
        for k in cand:
            z = Z[k]
            gr = 2.0 * Qd ** 2 * (z - c)                     # = A_W^T lam
            M = np.array([[A[I[k], 0], A[Jx[k], 0]], [A[I[k], 1], A[Jx[k], 1]]])
            lam = np.linalg.solve(M, gr)
            offer(z, (int(I[k]), int(Jx[k])), lam)
    if best is None:
        return None, (), 0.0
    return best[1], best[2], best[3]


def solve(sc, x0, u_ref, movers, backup_u=None):
    """solve_control_problem (backup_cbf_qp.py:563-794) for one agent.
    -> dict(u [2], status 0 optimal / 1 QP infeasible (fall-back used), intervene bool, h_min, G, h, keep, phi, active rows)."""
    x0 = np.array(x0, dtype=np.float64).reshape(-1)
    phi, S = integrate(sc, x0)
    G, hh, keep, h_min = rows(sc, x0, movers, phi, S)
    u_ref = np.array(u_ref, dtype=np.float64).reshape(-1)
    out = dict(G=G, h=hh, keep=keep, h_min=h_min, phi=phi, S=S, active=())
    if not keep.any():                                                         # :785-789: u_ref as it is (not clipped)
        out.update(u=u_ref.copy(), status=0, intervene=False)
        return out
    u_scale = np.array([sc.a_max, sc.a_max])                                   # :689-691
    u_ref = np.clip(u_ref, -u_scale, u_scale)                                  # :702
    urs = (1.0 / u_scale) * u_ref                                              # :709-710
    Q = np.array(sc.Q_u)
    idx = np.where(keep)[0]
    A = np.vstack([G[idx] * u_scale[None, :], np.eye(2), -np.eye(2)])          # (G S) z >= h, z >= -1, -z >= -1   :722-729
    b = np.concatenate([hh[idx], -np.ones(2), -np.ones(2)])
    z, W, gap = qp2_exact(Q, urs, A, b)
    if z is not None:
        out["u"] = u_scale * z                                                 # :752
        out["status"] = 0
        out["intervene"] = bool(np.linalg.norm(Q * (z - urs)) > 0.1)           # :757-766
        # active rows in the kernel's numbering: safety / terminal rows 0 .. N-1, then N + (0: z0 >= -1, 1: z1 >= -1, 2: z0 <= 1, 3: z1 <= 1)
        out["active"] = tuple(sorted(int(idx[w]) if w < idx.size else sc.N + (w - idx.size) for w in W))
        out["gap"] = gap
    else:                                                                      # :768-783
        out["status"] = 1
        if h_min > 0.01:
            out["u"] = u_ref
            out["intervene"] = False
        else:
            out["u"] = np.array(backup_control(sc, x0), dtype=np.float64)
            out["intervene"] = True
    return out


def nominal_control(sc, s):
    """EvadeNominalController.compute_control (examples/evade/test_evade.py:141-168): the u_ref of the evade scenario."""
    x, y, vx, vy = float(s[0]), float(s[1]), float(s[2]), float(s[3])
    error_y = 0.0 - y
    error_vx = sc.v_max - vx
    error_vy = 0.0 - vy
    ax = 2.0 * error_vx
    ay = 2.0 * error_y + 2.0 * error_vy
    a_mag = np.sqrt(ax ** 2 + ay ** 2)
    if a_mag > sc.a_max:
        ax = ax * sc.a_max / a_mag
        ay = ay * sc.a_max / a_mag
    return np.array([ax, ay])
