"""Shared closed-loop parity helpers: any backend (CUDA scb_control_step or the CPU host-sim) vs
the reference's golden LocalTrackingController runs and vs oracle/tracking.py."""
import json
import os

import numpy as np

from conftest import GOLDEN

X_TOL = 1e-9          # closed-loop state / input agreement with the reference run (free-running, hundreds of steps)
STEP_TOL = 1e-10      # one teacher-forced step

TEST_TRACKING_WP = np.array([[2, 2, np.pi / 2], [2, 12, 0], [12, 12, 0], [12, 2, 0]], float)


def load_tracking_golden():
    out = {}
    for f in ("ref_tracking.npz", "ref_tracking2.npz"):              # 2: DoubleIntegrator2D runs
        z = np.load(os.path.join(GOLDEN, f))
        for k in z.files:
            tag, key = k.rsplit("/", 1)
            out.setdefault(tag, {})[key] = z[k]
    for d in out.values():
        d["spec"] = json.loads(str(d["spec"]))
        for junk in ("robot_id", "exploration", "unknown_obs_detection"):
            d["spec"].pop(junk, None)
        d["dynamic"] = bool(int(d["dynamic"]))
        d["enable_rotation"] = bool(int(d["enable_rotation"]))
        d["M"] = int(d["M"])
        d["spec"]["num_constraints"] = d["M"]
    return out


def golden_initial_state(d):
    """X0 as LocalTrackingController received it (yaw appended for SingleIntegrator2D)."""
    if d["spec"]["model"] in ("SingleIntegrator2D", "DoubleIntegrator2D"):
        return np.append(d["X"][0], d["yaw"][0])[None]
    return d["X"][0][None]


def forced_arrays(d):
    """Every recorded step of a golden run as one agent of a batch: state BEFORE step k for k < T."""
    T = len(d["ret"])
    npos = d["goal"].shape[1] if d["spec"]["model"] == "Quad3D" else 2
    goal = np.nan_to_num(d["goal"][:T, :npos], nan=0.0)
    return dict(X=d["X"][:T], yaw=d["yaw"][:T], sm=d["sm"][:T].astype(np.int32), wp_idx=d["wp_idx"][:T].astype(np.int32),
                goal=goal, has_goal=d["has_goal"][:T].astype(np.int32), u_att=d["u_att"][:T])


def check_forced(d, out, nu=2):
    """`out`: dict of arrays after ONE control step of the forced batch -> asserts vs the golden step results."""
    T = len(d["ret"])
    assert np.array_equal(out["ret"], d["ret"].astype(np.int32)), np.nonzero(out["ret"] != d["ret"])[0][:10]
    np.testing.assert_allclose(out["Uref"], d["u_ref"], rtol=0, atol=STEP_TOL)
    assert np.array_equal(out["nobs"], np.minimum(d["nsel"], d["M"]).astype(np.int32))
    for k in range(T):
        n = max(int(out["nobs"][k]), 0)
        np.testing.assert_allclose(out["OBS"][k, :n], d["sel"][k, :n], rtol=0, atol=1e-12)
    assert np.array_equal(out["status"] != 0, d["status"] != 0)
    ok = d["status"] == 0
    np.testing.assert_allclose(out["U"][ok], d["u"][ok], rtol=0, atol=1e-8)
    # state after the step (= recorded state before step k+1)
    np.testing.assert_allclose(out["X"], d["X"][1:T + 1], rtol=0, atol=STEP_TOL)
    np.testing.assert_allclose(out["yaw"], d["yaw"][1:T + 1], rtol=0, atol=STEP_TOL)
    # the state machine / goal are compared where the golden run carries on (after the next update_goal they are
    # whatever that step left): sm, wp_idx, has_goal recorded before step k+1 are the values after step k
    assert np.array_equal(out["sm"], d["sm"][1:T + 1])
    assert np.array_equal(out["wp_idx"], d["wp_idx"][1:T + 1])
    assert np.array_equal(out["has_goal"], d["has_goal"][1:T + 1])
    ua, ub = out["u_att"], d["u_att"][1:T + 1]
    assert np.array_equal(np.isnan(ua), np.isnan(ub))
    np.testing.assert_allclose(ua[~np.isnan(ua)], ub[~np.isnan(ub)], rtol=0, atol=STEP_TOL)


def random_closed_loop_case(model, N, K, seed, dynamic=False):
    """Seeded closed-loop scene: N agents, K circles, 3 waypoints each -> (X0, scene, waypoints list)."""
    rng = np.random.default_rng(seed)
    L = 3.0 * np.sqrt(K)
    scene = np.zeros((K, 7))
    scene[:, 0:2] = rng.uniform(0, L, (K, 2)); scene[:, 2] = rng.uniform(0.2, 0.5, K)
    if dynamic:
        scene[:, 3:5] = rng.uniform(-0.3, 0.3, (K, 2))
    pos = np.empty((N, 2)); todo = np.arange(N)
    while todo.size:
        cand = rng.uniform(0, L, (todo.size, 2))
        d = np.sqrt(((cand[:, None] - scene[None, :, :2]) ** 2).sum(-1)) - scene[None, :, 2] - 0.7
        ok = (d > 0).all(1)
        pos[todo[ok]] = cand[ok]; todo = todo[~ok]
    if model == "SingleIntegrator2D":
        X0 = np.hstack([pos, rng.uniform(-np.pi, np.pi, (N, 1))])
    elif model == "DoubleIntegrator2D":
        X0 = np.hstack([pos, rng.uniform(-0.6, 0.6, (N, 2)), rng.uniform(-np.pi, np.pi, (N, 1))])
    elif model == "Unicycle2D":
        X0 = np.hstack([pos, rng.uniform(-np.pi, np.pi, (N, 1))])
    elif model == "Quad2D":
        X0 = np.hstack([pos, rng.uniform(-0.2, 0.2, (N, 1)), rng.uniform(-0.5, 0.5, (N, 2)), rng.uniform(-0.1, 0.1, (N, 1))])
    elif model == "Quad3D":
        X0 = np.zeros((N, 12)); X0[:, :2] = pos; X0[:, 2] = rng.uniform(1, 2, N)
    else:
        v = rng.uniform(0.3, 0.9, N)
        X0 = np.hstack([pos, rng.uniform(-np.pi, np.pi, (N, 1)), v[:, None]])
    wps = []
    for i in range(N):
        w = np.zeros((3, 3)); w[:, :2] = pos[i] + np.cumsum(rng.uniform(-2.5, 2.5, (3, 2)), axis=0)
        w[:, 2] = rng.uniform(1, 2, 3) if model == "Quad3D" else 0.0
        wps.append(w)
    return X0, scene, wps
