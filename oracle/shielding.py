"""Oracle restatement of the reference's trajectory-rollout shields for the double integrator in the evade scene
(TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this).

  shielding/gatekeeper.py:43-133   parameters and state (committed trajectory, event timing)
  shielding/gatekeeper.py:271-307  _forward_simulate_backup (robot.step under the backup policy; excludes the initial state)
  shielding/gatekeeper.py:309-367  _generate_candidate_trajectory (external nominal trajectory prefix + backup rollout)
  shielding/gatekeeper.py:380-471  _is_collision / _check_moving_obstacle_collision
  shielding/gatekeeper.py:499-527  _is_candidate_valid (time-synchronised obstacle states)
  shielding/gatekeeper.py:553-672  Gatekeeper.solve_control_problem (backward search over the nominal horizon)
  shielding/mps.py:59-160          MPS.solve_control_problem (one nominal step, re-evaluated every call)
  envs/evade_env.py:408-485        check_collision (walls), check_obstacle_collision (the bullet where it is NOW)

The nominal trajectory is external (set_nominal_trajectory, the way examples/evade/test_evade.py:424-426 drives both
classes); the policy / dynamics / scene are those of oracle/backup_cbf.py.  Moving obstacles are constant-velocity rows
[x, y, vx, vy, length, width, radius, kind] evaluated at t = k dt (test_evade.py:373-385); `static_rect` is the bullet's
hitbox at the current time (x_min, x_max, y_min, y_max, active), which the reference checks against EVERY state of a
candidate with the bare robot radius (gatekeeper.py:407-413 -> evade_env.py:454-485).
"""
import numpy as np

from . import backup_cbf as B


def wall_collision(sc, px, py, r):
    """EvadeEnv.check_collision (evade_env.py:408-452)."""
    if py - r < -sc.half_width:
        return True
    if py + r > sc.half_width:
        if sc.pocket_x_min <= px <= sc.pocket_x_max:
            if py + r > sc.pocket_y_max:
                return True
            if px - r < sc.pocket_x_min:
                if py > sc.half_width:
                    return True
            if px + r > sc.pocket_x_max:
                if py > sc.half_width:
                    return True
        else:
            return True
    if px - r < 0:
        return True
    if px + r > sc.hallway_length:
        return True
    return False


def rect_hit(px, py, x_min, x_max, y_min, y_max, r):
    """circle vs axis-aligned rectangle (evade_env.py:476-483, gatekeeper.py:452-462): dist(closest point) < r"""
    cx = min(max(px, x_min), x_max)
    cy = min(max(py, y_min), y_max)
    return np.sqrt((px - cx) ** 2 + (py - cy) ** 2) < r


def is_collision(sc, s, t, movers, static_rect):
    """Gatekeeper._is_collision (gatekeeper.py:380-424) with safety_margin = sc.safety_margin."""
    px, py = float(s[0]), float(s[1])
    if wall_collision(sc, px, py, sc.radius):
        return True
    if static_rect is not None and static_rect[4] != 0 and rect_hit(px, py, static_rect[0], static_rect[1], static_rect[2],
                                                                   static_rect[3], sc.radius):
        return True
    rr = sc.radius + sc.safety_margin
    if movers is not None:
        for o in np.asarray(movers, dtype=np.float64).reshape(-1, 8):
            kind = int(o[7])
            if kind == 0:
                continue
            ox = o[0] + o[2] * t
            oy = o[1] + o[3] * t
            if kind == 1:
                if rect_hit(px, py, ox - o[4] / 2, ox + o[4] / 2, oy - o[5] / 2, oy + o[5] / 2, rr):
                    return True
            else:
                if np.sqrt((px - ox) ** 2 + (py - oy) ** 2) < (rr + o[6]):
                    return True
    return False


def bullet_static_rect(bullet_x, bullet_length=3.0, bullet_width=4.0, bullet_y=0.0, active=True):
    """the hitbox of EvadeEnv.check_obstacle_collision (evade_env.py:469-473)"""
    return np.array([bullet_x - bullet_length / 2, bullet_x + bullet_length / 2 + bullet_length / 3,
                     bullet_y - bullet_width / 2, bullet_y + bullet_width / 2, 1.0 if active else 0.0])


class OracleShield:
    """Gatekeeper (mode 'gatekeeper') or MPS (mode 'mps') for ONE agent; state lives in the object like in the reference."""

    def __init__(self, sc, mode="gatekeeper", event_offset=0.05, horizon_discount=None):
        self.sc, self.mode = sc, mode
        self.event_offset = event_offset
        self.horizon_discount = horizon_discount if horizon_discount is not None else 5 * sc.dt      # gatekeeper.py:68
        self.Nb = int(sc.backup_horizon / sc.dt)
        self.committed_u = None
        self.committed_x = None
        self.current_time_idx = self.Nb                                                              # :104
        self.next_event_time = 0.0
        self.committed_horizon = 0.0
        self.actual_nominal_steps = 0
        self.matches_nominal = False

    def backup_rollout(self, s):
        xs, us = np.zeros((self.Nb, 4)), np.zeros((self.Nb, 2))
        s = np.array(s, dtype=np.float64).reshape(-1)
        for i in range(self.Nb):
            u = np.array(B.backup_control(self.sc, s), dtype=np.float64)
            us[i] = u
            s = B.di_step(self.sc, s, u)
            xs[i] = s
        return xs, us

    def candidate(self, x, nom_x, nom_u, steps):
        """gatekeeper.py:309-367 with an external nominal trajectory -> (x_traj, u_traj, actual nominal steps)"""
        n_use = min(steps + 1, len(nom_x))
        actual = max(0, n_use - 1)
        if n_use > 0:
            nx, nu = nom_x[:n_use], (nom_u[:actual] if actual > 0 else np.empty((0, 2)))
        else:
            nx, nu, actual = np.array(x, dtype=np.float64).reshape(1, -1), np.empty((0, 2)), 0
        bx, bu = self.backup_rollout(nx[-1])
        return np.vstack([nx, bx]), np.vstack([nu, bu]), actual

    def valid(self, cx, movers, static_rect):
        for k, s in enumerate(cx):
            if is_collision(self.sc, s, k * self.sc.dt, movers, static_rect):
                return False
        return True

    def commit(self, cx, cu, actual):                                                                # :529-551
        self.committed_x, self.committed_u = cx.copy(), cu.copy()
        self.next_event_time = self.event_offset
        self.current_time_idx = 0
        self.actual_nominal_steps = actual
        self.committed_horizon = actual * self.sc.dt

    def solve(self, x, nom_x, nom_u, movers, static_rect):
        sc, dt = self.sc, self.sc.dt
        x = np.array(x, dtype=np.float64).reshape(-1)
        if self.committed_u is None:                                                                 # :571-583
            bx, bu = self.backup_rollout(x)
            self.committed_x, self.committed_u = np.vstack([x.reshape(1, -1), bx]), bu
            self.committed_horizon, self.actual_nominal_steps = 0.0, 0
            self.current_time_idx, self.next_event_time = 0, 0.0
        if self.mode == "mps":                                                                       # mps.py:86-127
            if nom_x is not None and len(nom_x) > 1:
                cx, cu, actual = self.candidate(x, nom_x, nom_u, 1)
                if self.valid(cx, movers, static_rect):
                    self.commit(cx, cu, actual)
                else:
                    self.next_event_time = self.current_time_idx * dt + self.event_offset
        elif self.current_time_idx >= self.next_event_time / dt:                                     # gatekeeper.py:590-654
            max_steps = len(nom_x) - 1 if nom_x is not None else 0
            disc = max(1, int(self.horizon_discount / dt))
            found = False
            for i in range(max_steps // disc + 2):
                steps = max(max_steps - i * disc, 0)
                cx, cu, actual = self.candidate(x, nom_x, nom_u, steps)
                if self.valid(cx, movers, static_rect):
                    self.commit(cx, cu, actual)
                    found = True
                    break
            if not found:
                self.next_event_time = self.current_time_idx * dt + self.event_offset
        if self.current_time_idx < len(self.committed_u):                                            # :656-667
            u = self.committed_u[self.current_time_idx].copy()
        else:
            u = np.array(B.backup_control(sc, x), dtype=np.float64)
        if self.mode == "mps":                                                                       # mps.py:142-153
            self.matches_nominal = bool(nom_u is not None and len(nom_u) > 0 and np.linalg.norm(u - nom_u[0]) < 1e-2)
        self.current_time_idx += 1
        return u

    def is_using_backup(self):
        if self.mode == "mps":
            return not self.matches_nominal                                                          # mps.py:55-57
        return self.current_time_idx >= int(self.committed_horizon / self.sc.dt)                     # gatekeeper.py:741-744


def nominal_rollout(sc, x, horizon_time=10.0):
    """rollout_nominal of examples/evade/test_evade.py:387-408 -> (x_traj [T+1, 4], u_traj [T, 2])"""
    steps = int(horizon_time / sc.dt)
    xs, us = [np.array(x, dtype=np.float64).reshape(-1)], []
    s = xs[0]
    for _ in range(steps):
        u = B.nominal_control(sc, s)
        s = B.di_step(sc, s, u)
        xs.append(s); us.append(u)
    return np.array(xs), np.array(us)
