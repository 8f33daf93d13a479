#!/usr/bin/env python
"""Backup-CBF QP fixtures from the REFERENCE'S OWN code (run here once; /root/reference does not travel to the GPU box).

    python tests/golden/gen_backupcbf_from_reference.py        # writes tests/golden/ref_backupcbf.npz

The unmodified  position_control/backup_cbf_qp.py::BackupCBF,  position_control/backup_controller.py::EvadeBackupController,
robots/double_integrator2D.py::DoubleIntegrator2D  and  envs/evade_env.py::EvadeEnv  are imported through oracle/refshim
(cvxpy -> the affine-probing stand-in that hands the reference's own QP statement to an exact solver; matplotlib -> inert
mocks) and driven the way examples/evade/test_evade.py:265-470 drives them:

  loop    the evade scenario itself (dt 0.1, backup horizon 12 s -> 120 backup steps), closed loop from x = 20 with the
          bullet chasing the robot, every step recorded until the robot has left the pocket again or 400 steps;
  probe   seeded random states / bullet positions / nominal inputs in and around the hallway and the pocket, including
          states outside the safe set (QP infeasible -> both fall-backs of backup_cbf_qp.py:768-783).

Recorded per call: state, u_ref, bullet x, the returned input, is_using_backup(), _last_h_min, the backup trajectory
(latest_backup_trajectory) and the QP the reference stated (rows of (G S) z >= h in its own order, box rows last).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from unittest import mock  # noqa: E402
from oracle import refshim  # noqa: E402

sys.modules.setdefault("matplotlib.collections", mock.MagicMock(name="matplotlib.collections"))
refshim.install()
import cvxpy as fake_cp  # noqa: E402  (the refshim stand-in)

from safe_control.envs.evade_env import EvadeEnv  # noqa: E402
from safe_control.robots.double_integrator2D import DoubleIntegrator2D  # noqa: E402
from safe_control.position_control.backup_controller import EvadeBackupController  # noqa: E402
from safe_control.position_control.backup_cbf_qp import BackupCBF  # noqa: E402

LAST = {}
_Problem = fake_cp.Problem


class RecordingProblem(_Problem):
    def solve(self, *a, **kw):
        v = super().solve(*a, **kw)
        LAST["qp"] = self.qp
        return v


fake_cp.Problem = RecordingProblem

ROWS_MAX = 124          # 119 safety rows + terminal + 4 box rows


def make(dt=0.1, horizon=12.0, safety_margin=0.5, use_goal=True):
    """test_evade.py:272-371 (default configuration, :60-100)"""
    env = EvadeEnv(hallway_length=60.0, hallway_width=4.0, pocket_x=25.0, pocket_length=10.0, pocket_width=4.0,
                   goal_length=5.0, bullet_speed=3.0, bullet_length=3.0, bullet_start_x=-10.0)
    env._draw_bullet_bill = lambda: None
    spec = {"radius": 0.5, "a_max": 2.0, "v_max": 1.5, "model": "DoubleIntegrator2D", "safety_margin": safety_margin}
    goal_bounds = {"x_min": env.goal_x_min, "x_max": env.goal_x_max, "y_min": -env.half_width, "y_max": env.half_width}
    backup = EvadeBackupController(spec, dt, env.get_pocket_center(), env.get_pocket_bounds(), goal_bounds if use_goal else None)
    dyn = DoubleIntegrator2D(dt, spec)
    sh = BackupCBF(robot=dyn, robot_spec=spec, dt=dt, backup_horizon=horizon, ax=None)
    sh.set_backup_controller(backup)
    sh.set_environment(env)

    def get_obstacles(t=0.0):                    # test_evade.py:373-385
        st = env.get_bullet_state()
        if not st["active"]:
            return None
        fut = st.copy()
        fut["x"] = st["x"] + st["vx"] * t
        return fut

    sh.set_moving_obstacles(get_obstacles)
    return env, spec, dyn, sh


def nominal(spec, state):
    """EvadeNominalController.compute_control (test_evade.py:141-168), restated: it lives in an example script that pulls in
    the animation stack; only u_ref = its output enters the path and u_ref is an INPUT of the fixture."""
    x, y, vx, vy = state.flatten()
    ax = 2.0 * (spec["v_max"] - vx)
    ay = 2.0 * (0.0 - y) + 2.0 * (0.0 - vy)
    a = np.sqrt(ax ** 2 + ay ** 2)
    if a > spec["a_max"]:
        ax, ay = ax * spec["a_max"] / a, ay * spec["a_max"] / a
    return np.array([[ax], [ay]])


def record(rec, sh, env, state, u_ref_rows):
    LAST.pop("qp", None)
    sh.set_nominal_trajectory(np.tile(state.reshape(1, -1), (u_ref_rows.shape[0] + 1, 1)), u_ref_rows)
    u = sh.solve_control_problem(state.reshape(-1, 1))
    qp = LAST.get("qp")
    A = np.zeros((ROWS_MAX, 2)); b = np.zeros(ROWS_MAX); m = 0; st = -1
    if qp is not None:
        m = qp["h"].size
        A[:m] = -qp["G"]; b[:m] = -qp["h"]           # fake_cvxpy states G z <= h; the reference wrote (G S) z >= h
        st = int(qp["res"]["status"])
    rec["state"].append(state.flatten().copy()); rec["u_ref"].append(u_ref_rows[0].copy())
    rec["bullet_x"].append(env.bullet_x); rec["bullet_active"].append(float(env.bullet_active))
    rec["u"].append(np.asarray(u, float).flatten()); rec["using_backup"].append(float(sh.is_using_backup()))
    rec["h_min"].append(float(sh._last_h_min)); rec["phi"].append(sh.latest_backup_trajectory.copy())
    rec["qp_A"].append(A); rec["qp_b"].append(b); rec["qp_m"].append(m); rec["qp_status"].append(st)
    return np.asarray(u, float).reshape(-1, 1)


def main():
    out = {}
    # ---- loop: the scenario itself --------------------------------------------------------------------------------------
    env, spec, dyn, sh = make()
    rec = {k: [] for k in ("state", "u_ref", "bullet_x", "bullet_active", "u", "using_backup", "h_min", "phi", "qp_A", "qp_b",
                           "qp_m", "qp_status")}
    state = np.array([20.0, 0.0, 0.0, 0.0]).reshape(-1, 1)
    hidden = False
    for step in range(400):
        uref = nominal(spec, state).reshape(1, 2)
        u = record(rec, sh, env, state.copy(), np.tile(uref, (3, 1)))
        state = dyn.step(state, u)
        vx, vy = state[2, 0], state[3, 0]                      # test_evade.py:452-456
        vm = np.sqrt(vx ** 2 + vy ** 2)
        if vm > spec["v_max"]:
            state[2, 0] = vx * spec["v_max"] / vm
            state[3, 0] = vy * spec["v_max"] / vm
        env.step_bullet(0.1)
        hidden = hidden or state[1, 0] > 2.5
        if env.check_goal_reached(state[:2, 0]):
            break
    for k, v in rec.items():
        out["loop_" + k] = np.array(v)
    out["loop_final_state"] = state.flatten()
    out["loop_reached_goal"] = np.array(float(env.check_goal_reached(state[:2, 0])))
    out["loop_hid_in_pocket"] = np.array(float(hidden))
    print("loop: steps", len(rec["u"]), "goal", bool(out["loop_reached_goal"]), "hid", hidden, "backup steps", int(np.sum(rec["using_backup"])),
          "h_min range", np.min(rec["h_min"]), np.max(rec["h_min"]), "qp statuses", np.bincount(np.array(rec["qp_status"]) + 1))

    # ---- probe: seeded random calls ------------------------------------------------------------------------------------------
    rng = np.random.default_rng(20260117)
    for tag, (dt, hor, n) in {"probe": (0.1, 12.0, 96), "short": (0.05, 2.0, 64)}.items():
        env, spec, dyn, sh = make(dt=dt, horizon=hor, use_goal=(tag == "probe"))
        rec = {k: [] for k in rec}
        for q in range(n):
            kind = q % 4
            if kind == 0:      # hallway
                p = [rng.uniform(1.0, 59.0), rng.uniform(-1.4, 1.4)]
            elif kind == 1:    # pocket and its mouth
                p = [rng.uniform(25.2, 34.8), rng.uniform(0.0, 5.4)]
            elif kind == 2:    # near walls / corners (some outside the safe set)
                p = [rng.uniform(22.0, 38.0), rng.uniform(1.0, 2.2)]
            else:
                p = [rng.uniform(0.2, 59.8), rng.uniform(-1.8, 1.8)]
            v = rng.uniform(-1.0, 1.0, 2) * rng.uniform(0.0, 1.5)
            state = np.array([p[0], p[1], v[0], v[1]])
            env.bullet_x = rng.uniform(-10.0, 62.0) if q % 5 else state[0] - rng.uniform(2.0, 12.0)
            env.bullet_active = bool(q % 11)
            uref = rng.uniform(-2.5, 2.5, (1, 2))
            record(rec, sh, env, state, np.tile(uref, (3, 1)))
        for k, v in rec.items():
            out[tag + "_" + k] = np.array(v)
        out[tag + "_cfg"] = np.array([dt, hor, float(tag == "probe")])
        print(tag, "calls", n, "qp statuses (none/optimal/infeasible)", np.bincount(np.array(rec["qp_status"]) + 1, minlength=3),
              "backup", int(np.sum(rec["using_backup"])))
    path = os.path.join(HERE, "ref_backupcbf.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
