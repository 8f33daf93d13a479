#!/usr/bin/env python
"""Gatekeeper / MPS fixtures from the REFERENCE'S OWN code (run here once; /root/reference does not travel to the GPU box).

    python tests/golden/gen_shield_from_reference.py        # writes tests/golden/ref_shield.npz

The unmodified  shielding/gatekeeper.py::Gatekeeper,  shielding/mps.py::MPS,  position_control/backup_controller.py::
EvadeBackupController,  robots/double_integrator2D.py::DoubleIntegrator2D  and  envs/evade_env.py::EvadeEnv  are imported
through oracle/refshim (matplotlib -> inert mocks) and driven exactly like examples/evade/test_evade.py:265-470 (default
configuration: dt 0.1, backup horizon 12 s, nominal horizon 10 s, event offset 0.05 s, safety margin 0.5): closed loop
from x = 20 with the bullet chasing the robot, until the goal is reached (or tf = 60 s).

Recorded per step: state, bullet x / active flag, the returned input, is_using_backup(), current_time_idx,
committed_horizon, len(committed_u_traj), next_event_time, a checksum of the nominal trajectory handed over (the tests
rebuild it with oracle/shielding.py::nominal_rollout, a restatement of the example's rollout_nominal), and the whole
committed input trajectory every 25 steps.  A second, shorter run per class starts inside the pocket mouth with the bullet
close (every candidate invalid at first -> the 'keep the committed backup' branch, gatekeeper.py:648-654).
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

from safe_control.envs.evade_env import EvadeEnv  # noqa: E402
from safe_control.robots.double_integrator2D import DoubleIntegrator2D  # noqa: E402
from safe_control.position_control.backup_controller import EvadeBackupController  # noqa: E402
from safe_control.shielding.gatekeeper import Gatekeeper  # noqa: E402
from safe_control.shielding.mps import MPS  # noqa: E402

DT, TB, TN, OFFSET, MARGIN = 0.1, 12.0, 10.0, 0.05, 0.5


def nominal(spec, state):
    """EvadeNominalController.compute_control (test_evade.py:141-168), restated (the example script pulls in the animation stack)"""
    x, y, vx, vy = np.asarray(state, float).flatten()
    ax = 2.0 * (spec["v_max"] - vx)
    ay = 2.0 * (0.0 - y) + 2.0 * (0.0 - vy)
    a = np.sqrt(ax ** 2 + ay ** 2)
    if a > spec["a_max"]:
        ax, ay = ax * spec["a_max"] / a, ay * spec["a_max"] / a
    return np.array([[ax], [ay]])


def run(algo, x0, bullet_x0, max_steps):
    env = EvadeEnv(hallway_length=60.0, hallway_width=4.0, pocket_x=25.0, pocket_length=10.0, pocket_width=4.0,
                   goal_length=5.0, bullet_speed=3.0, bullet_length=3.0, bullet_start_x=-10.0)
    env._draw_bullet_bill = lambda: None
    env.bullet_x = bullet_x0
    spec = {"radius": 0.5, "a_max": 2.0, "v_max": 1.5, "model": "DoubleIntegrator2D", "safety_margin": MARGIN}
    goal_bounds = {"x_min": env.goal_x_min, "x_max": env.goal_x_max, "y_min": -env.half_width, "y_max": env.half_width}
    backup = EvadeBackupController(spec, DT, env.get_pocket_center(), env.get_pocket_bounds(), goal_bounds)
    dyn = DoubleIntegrator2D(DT, spec)
    if algo == "mps":
        sh = MPS(robot=dyn, robot_spec=spec, dt=DT, backup_horizon=TB, event_offset=OFFSET, ax=None, safety_margin=MARGIN)
    else:
        sh = Gatekeeper(robot=dyn, robot_spec=spec, dt=DT, backup_horizon=TB, nominal_horizon=TN, event_offset=OFFSET, ax=None,
                        safety_margin=MARGIN)
    sh.visualize_backup = False
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

    def rollout_nominal(start):                  # test_evade.py:387-408
        xs, us = [start.flatten()], []
        cur = start.reshape(-1, 1)
        for _ in range(int(TN / DT)):
            u = nominal(spec, cur)
            nxt = dyn.step(cur, u)
            xs.append(nxt.flatten()); us.append(u.flatten())
            cur = nxt
        return np.array(xs), np.array(us)

    rec = {k: [] for k in ("state", "bullet_x", "bullet_active", "u", "using_backup", "idx", "horizon", "clen", "next_event", "nom_sum")}
    snaps, snap_at = [], []
    state = np.array(x0, dtype=float).reshape(-1, 1)
    reached = False
    for step in range(max_steps):
        nom_x, nom_u = rollout_nominal(state)
        sh.set_nominal_trajectory(nom_x, nom_u)
        rec["state"].append(state.flatten().copy()); rec["bullet_x"].append(env.bullet_x); rec["bullet_active"].append(float(env.bullet_active))
        rec["nom_sum"].append(float(nom_x.sum() + nom_u.sum()))
        u = sh.solve_control_problem(state)
        rec["u"].append(np.asarray(u, float).flatten()); rec["using_backup"].append(float(sh.is_using_backup()))
        rec["idx"].append(sh.current_time_idx); rec["horizon"].append(sh.committed_horizon)
        rec["clen"].append(len(sh.committed_u_traj)); rec["next_event"].append(sh.next_event_time)
        if step % 25 == 0:
            cu = np.zeros((220, 2)); cu[: len(sh.committed_u_traj)] = sh.committed_u_traj
            snaps.append(cu); snap_at.append(step)
        state = dyn.step(state, u)
        vx, vy = state[2, 0], state[3, 0]        # test_evade.py:452-456
        vm = np.sqrt(vx ** 2 + vy ** 2)
        if vm > spec["v_max"]:
            state[2, 0] = vx * spec["v_max"] / vm
            state[3, 0] = vy * spec["v_max"] / vm
        env.step_bullet(DT)
        if env.check_goal_reached(state[:2, 0]):
            reached = True
            break
    out = {k: np.array(v) for k, v in rec.items()}
    out["snap_cu"] = np.array(snaps); out["snap_at"] = np.array(snap_at)
    out["reached_goal"] = np.array(float(reached)); out["final_state"] = state.flatten()
    print(algo, "x0", list(x0), "steps", len(rec["u"]), "goal", reached, "backup steps", int(np.sum(rec["using_backup"])),
          "max committed horizon", float(np.max(rec["horizon"])), "y max", float(np.max(np.array(rec["state"])[:, 1])))
    return out


def main():
    out = {}
    for algo in ("gatekeeper", "mps"):
        for tag, (x0, bx, n) in {"scenario": ([20.0, 0.0, 0.0, 0.0], -10.0, 600), "cornered": ([23.0, 0.3, 0.6, 0.0], 14.0, 120)}.items():
            for k, v in run(algo, x0, bx, n).items():
                out[f"{algo}_{tag}_{k}"] = v
    out["cfg"] = np.array([DT, TB, TN, OFFSET, MARGIN])
    path = os.path.join(HERE, "ref_shield.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
