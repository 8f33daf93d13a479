#!/usr/bin/env python
"""Closed-loop golden fixtures from the REFERENCE'S OWN LocalTrackingController.

    python tests/golden/gen_tracking_from_reference.py        # writes tests/golden/ref_tracking.npz

Runs only in the build container (needs /root/reference).  The reference's tracking.py,
robots/robot.py, robots/<model>.py, position_control/cbf_qp.py, utils/env.py and
dynamic_env/main.py are imported UNMODIFIED through oracle/refshim (casadi -> numeric
stand-in, cvxpy -> affine layer over the exact QP solver, matplotlib / shapely -> inert
mocks: with show_animation=False and no 'sensor' key no geometry result is ever consumed).

Per scenario it records, for every control step k, the tracker state BEFORE the step
(X, yaw, state machine, waypoint index, goal, u_att, obstacle array) and what the step
produced (u_ref and obstacle rows handed to the controller, u, status, return code), plus
the state after the last step.  tests/test_oracle_pinned.py replays them through
oracle/tracking.py; the GPU / host-sim tests replay them through scb_control_step, both
teacher-forced (every recorded step as one agent of a batch) and free-running.
"""
import contextlib
import io
import json
import math
import os
import sys
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import refshim  # noqa: E402

plt = mock.MagicMock(name="matplotlib.pyplot")
plt.colormaps.get_cmap.return_value.colors = [(0.5, 0.5, 0.5)] * 9      # robots/robot.py:45-47 indexes the palette
sys.modules["matplotlib.pyplot"] = plt
mpl = mock.MagicMock(name="matplotlib"); mpl.pyplot = plt
sys.modules["matplotlib"] = mpl
for n in ("shapely", "shapely.geometry", "shapely.ops", "shapely.validation"):
    sys.modules[n] = mock.MagicMock(name=n)
refshim.install()
sys.path.insert(0, refshim.REFERENCE_ROOT)                               # dynamic_env/main.py:36 imports `position_control`

from safe_control.tracking import LocalTrackingController  # noqa: E402
from safe_control.dynamic_env.main import LocalTrackingControllerDyn  # noqa: E402
from safe_control.utils import env as envmod  # noqa: E402
from safe_control.utils.headless_plot import NullArtist, NullAxes, NullFigure  # noqa: E402

SM = {"idle": 0, "track": 1, "stop": 2, "rotate": 3}


class Ax(NullAxes):
    patches = []

    def __getattr__(self, _n):
        return lambda *a, **k: NullArtist()


def pad7(obs):
    obs = np.asarray(obs, float)
    if obs.shape[1] < 7:
        obs = np.hstack((obs, np.zeros((obs.shape[0], 7 - obs.shape[1]))))
    return obs


def run(cls, x0, spec, waypoints, known_obs, steps, enable_rotation=True, M=None):
    spec = dict(spec)
    if M is not None:
        spec["num_constraints"] = M
    with contextlib.redirect_stdout(io.StringIO()):
        tc = cls(np.asarray(x0, float), spec, controller_type={"pos": "cbf_qp"}, dt=0.05, show_animation=False,
                 enable_rotation=enable_rotation, env=envmod.Env(), ax=Ax(NullFigure()), fig=NullFigure())
        tc.obs = pad7(known_obs).copy()
        tc.set_waypoints(np.asarray(waypoints, float))
    M = tc.pos_controller.num_obs if hasattr(tc.pos_controller, "num_obs") else tc.num_constraints
    nu = 2
    rec = {k: [] for k in ("X", "yaw", "sm", "wp_idx", "has_goal", "goal", "u_att", "scene", "u_ref", "sel", "nsel",
                           "u", "status", "ret")}
    cap = {}
    orig = tc.pos_controller.solve_control_problem

    def spy(robot_state, control_ref, obs):
        cap["u_ref"] = np.asarray(control_ref["u_ref"], float).reshape(-1).copy()
        cap["obs"] = None if obs is None else np.asarray(obs, float).copy()
        u = orig(robot_state, control_ref, obs)
        cap["u"] = np.full(nu, np.nan) if u is None else np.asarray(u, float).reshape(-1).copy()
        return u

    tc.pos_controller.solve_control_problem = spy

    def snap():
        rec["X"].append(tc.robot.X.reshape(-1).copy()); rec["yaw"].append(float(tc.robot.yaw))
        rec["sm"].append(SM[tc.state_machine]); rec["wp_idx"].append(tc.current_goal_index)
        g = np.full(3, np.nan)
        if tc.goal is not None:
            g[: len(tc.goal)] = tc.goal
        rec["goal"].append(g); rec["has_goal"].append(0 if tc.goal is None else 1)
        rec["u_att"].append(np.nan if tc.u_att is None else float(np.asarray(tc.u_att).reshape(-1)[0]))
        rec["scene"].append(np.asarray(tc.obs, float).copy())

    for _ in range(steps):
        snap()
        with contextlib.redirect_stdout(io.StringIO()):
            ret = tc.control_step()
        rec["ret"].append(ret)
        rec["u_ref"].append(cap["u_ref"])
        sel = np.full((M, 7), np.nan); k = 0
        if cap["obs"] is not None:
            k = min(len(cap["obs"]), M); sel[:k] = cap["obs"][:k]
        rec["sel"].append(sel); rec["nsel"].append(-1 if cap["obs"] is None else len(cap["obs"]))
        ok = tc.pos_controller.status == "optimal"
        rec["status"].append(0 if ok else 1)
        rec["u"].append(cap["u"])
        if ret in (-1, -2):
            break
    snap()
    out = {k: np.asarray(v) for k, v in rec.items()}
    out["waypoints"] = np.asarray(tc.waypoints, float)
    out["M"] = np.array(M)
    out["enable_rotation"] = np.array(int(enable_rotation))
    return out, spec


TEST_TRACKING_OBS = np.array([[2.2, 5.0, 0.2], [3.0, 5.0, 0.2], [4.0, 9.0, 0.3], [1.5, 10.0, 0.5], [9.0, 11.0, 1.0],
                              [7.0, 7.0, 3.0], [4.0, 3.5, 1.5], [10.0, 7.3, 0.4], [6.0, 13.0, 0.7], [5.0, 10.0, 0.6],
                              [11.0, 5.0, 0.8], [13.5, 11.0, 0.6], [2.0, 7.0, 0.7], [2.0, 8.0, 0.5]])   # examples/test_tracking.py:52-54
TEST_TRACKING_WP = np.array([[2, 2, math.pi / 2], [2, 12, 0], [12, 12, 0], [12, 2, 0]], float)          # :44-49


def scenarios():
    wp = TEST_TRACKING_WP
    yield ("du_test_tracking", LocalTrackingController, np.append(wp[0], 1.0),
           {"model": "DynamicUnicycle2D", "w_max": 0.5, "a_max": 1.0, "radius": 0.25}, wp, TEST_TRACKING_OBS, 300, True, None)
    yield ("si_test_tracking", LocalTrackingController, wp[0],
           {"model": "SingleIntegrator2D", "v_max": 1.0, "radius": 0.25}, wp, TEST_TRACKING_OBS, 900, True, None)
    yield ("kb_test_tracking", LocalTrackingController, np.append(wp[0], 1.0),
           {"model": "KinematicBicycle2D", "a_max": 0.5, "radius": 0.5}, wp, TEST_TRACKING_OBS, 300, True, None)
    # goal behind the robot -> 'stop' -> 'rotate' -> 'track' (tracking.py:224-235, 569-578)
    yield ("du_stop_rotate", LocalTrackingController, np.array([2.0, 2.0, -math.pi / 2, 0.8]),
           {"model": "DynamicUnicycle2D", "w_max": 0.5, "a_max": 0.5, "radius": 0.25},
           np.array([[2, 2, 0], [2.5, 6.5, 0], [6, 4.2, 0]], float), TEST_TRACKING_OBS[:6], 400, True, 16)
    yield ("du_no_rotation", LocalTrackingController, np.array([1.0, 1.0, 0.3, 0.0]),
           {"model": "DynamicUnicycle2D", "w_max": 0.5, "a_max": 0.5, "radius": 0.25},
           np.array([[1, 1, 0], [3.0, 2.0, 0], [5.5, 2.2, 0]], float), TEST_TRACKING_OBS[[6, 0, 1]], 300, False, 4)
    # dynamic_env/main.py:243-268: 8 moving circles, C3BF
    known = np.array([[8.0, 9.0, 0.5], [10.0, 4.0, 0.5], [12.0, 5.0, 0.5], [14.0, 9.0, 0.5], [16.0, 6.0, 0.5],
                      [18.0, 14.0, 0.5], [20.0, 4.0, 0.5], [22.0, 12.0, 0.5]])
    dyn = [[o[0], o[1], o[2], -0.5, 0.5 if i % 2 == 0 else -0.5, 0.0, 15.0] for i, o in enumerate(known)]
    wpd = np.array([[1, 7.5, 0], [20, 7.5, 0]], float)
    yield ("c3bf_dynamic_env", LocalTrackingControllerDyn, np.append(wpd[0], 1.0),
           {"model": "KinematicBicycle2D_C3BF", "a_max": 5.0, "radius": 0.3}, wpd, np.array(dyn), 400, True, None)
    # no obstacles at all -> obs None -> u_ref unclipped (cbf_qp.py:113-118)
    yield ("du_no_obstacles", LocalTrackingController, np.array([0.0, 0.0, 0.1, 0.2]),
           {"model": "DynamicUnicycle2D", "radius": 0.25}, np.array([[0, 0, 0], [2.0, 0.5, 0]], float),
           np.zeros((0, 7)), 200, True, None)


def scenarios2():
    """Second file (ref_tracking2.npz): DoubleIntegrator2D -- yaw outside the state, VelocityTrackingYaw on the state
    velocity, own step rescales the velocity (examples/test_tracking.py --model di)."""
    wp = TEST_TRACKING_WP
    yield ("di_test_tracking", LocalTrackingController, wp[0],
           {"model": "DoubleIntegrator2D", "v_max": 1.0, "a_max": 1.0, "ax_max": 1.0, "ay_max": 1.0, "radius": 0.25},
           wp, TEST_TRACKING_OBS, 700, True, None)
    yield ("di_stop_rotate", LocalTrackingController, np.array([2.0, 2.0, 0.6, -0.3, -math.pi / 2]),
           {"model": "DoubleIntegrator2D", "v_max": 1.0, "a_max": 1.0, "radius": 0.25},
           np.array([[2, 2, 0], [2.5, 6.5, 0], [6, 4.2, 0]], float), TEST_TRACKING_OBS[:6], 400, True, 16)
    # Quad2D (x-z plane, two rotor forces; cascaded PD nominal law, quad2D.py:92-150)
    yield ("quad2d_tracking", LocalTrackingController, np.array([2.0, 2.0, 0.0]),
           {"model": "Quad2D", "f_min": 3.0, "f_max": 10.0, "sensor": None, "radius": 0.25},
           np.array([[2, 2, 0], [2, 12, 0], [12, 12, 0]], float), TEST_TRACKING_OBS, 500, True, None)


def main(gen=scenarios, fname="ref_tracking.npz"):
    flat = {}
    for name, cls, x0, spec, wp, obs, steps, rot, M in gen():
        out, spec = run(cls, x0, spec, wp, obs, steps, rot, M)
        for k, v in out.items():
            flat[f"{name}/{k}"] = v
        clean = {k: (float(v) if isinstance(v, (float, np.floating)) else v) for k, v in spec.items()
                 if isinstance(v, (int, float, str, bool, np.floating))}
        flat[f"{name}/spec"] = np.array(json.dumps(clean))
        flat[f"{name}/dynamic"] = np.array(int(cls is LocalTrackingControllerDyn))
        r = out["ret"]
        print(f"{name}: {len(r)} steps, last ret {r[-1]}, sm counts {np.bincount(out['sm'], minlength=4)}, "
              f"infeasible {int(out['status'].sum())}, final X {np.round(out['X'][-1], 3)}")
    np.savez_compressed(os.path.join(HERE, fname), **flat)


if __name__ == "__main__":
    if "--second" in sys.argv:
        main(scenarios2, "ref_tracking2.npz")
    else:
        main()
