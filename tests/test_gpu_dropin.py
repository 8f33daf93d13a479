"""The reference-surface classes (safe_control_b200.position_control) driven the way tracking.py drives
the reference's: one agent, numpy (n,1) columns, obstacle rows, `.status` strings."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class Robot:
    """The part of robots/robot.py:BaseRobot the controllers touch."""
    def __init__(self, X, dt=0.05):
        self.X = np.asarray(X, float).reshape(-1, 1)
        self.dt = dt


def test_config1_closed_loop_single_integrator():
    """BASELINE config 1: examples/test_tracking.py --model si --algo cbf_qp with the first two obstacles
    of its list (test_tracking.py:52), dt = 0.05: 300 closed-loop steps, every input equal to the oracle's."""
    from safe_control_b200.position_control import CBFQP
    from oracle.controllers import OracleCBFQP
    from oracle.models import make_model
    spec = {"model": "SingleIntegrator2D", "v_max": 1.0, "radius": 0.25}
    obs = np.array([[2.2, 5.0, 0.2, 0, 0, 0, 0], [3.0, 5.0, 0.2, 0, 0, 0, 0.0]])
    robot = Robot([2.0, 2.0])
    ctrl = CBFQP(robot, dict(spec), num_obs=2)
    oracle = OracleCBFQP(dict(spec), num_obs=2)
    model = make_model(dict(spec))
    goal = np.array([2.0, 12.0])
    n_active = 0
    for step in range(300):
        x = robot.X[:, 0].copy()
        u_ref = model.nominal_input(x, goal)
        u = ctrl.solve_control_problem(robot.X, {"u_ref": u_ref.reshape(-1, 1), "state_machine": "track", "goal": goal}, obs)
        u_o, info = oracle.solve(x, u_ref, obs)
        assert ctrl.status == "optimal" and info["status"] == 0
        assert u.shape == (2, 1)
        np.testing.assert_allclose(u[:, 0], u_o, rtol=0, atol=1e-9)
        n_active += int(info["active"][:2].any())
        robot.X = (x + u[:, 0] * 0.05).reshape(-1, 1)
    assert n_active > 20                      # the filter actually deflected the robot around the obstacles
    assert robot.X[1, 0] > 10.0               # and it got past them


def test_shims_status_and_none():
    from safe_control_b200.position_control import CBFQP, OptimalDecayCBFQP, MPCCBF
    r = Robot([0.0, 0.0, 0.0, 0.8])
    c = CBFQP(r, {"model": "DynamicUnicycle2D"}, num_obs=4)
    assert c.cbf_param == {"alpha1": 1.5, "alpha2": 1.5}
    u_ref = np.array([[0.9], [0.1]])
    u = c.solve_control_problem(r.X, {"u_ref": u_ref}, None)
    assert np.array_equal(u, u_ref) and c.status == "optimal"          # unclipped, cbf_qp.py:113-118
    u = c.solve_control_problem(r.X, {"u_ref": u_ref}, np.array([[0.45, 0.0, 0.2]]))   # 3-column row, too close
    assert c.status == "infeasible" and np.all(np.abs(u) <= 0.5 + 1e-12)
    od = OptimalDecayCBFQP(Robot([0.0, 0.0, 0.0, 1.0]), {"model": "KinematicBicycle2D_C3BF"})
    u = od.solve_control_problem(None, {"u_ref": np.array([[0.5], [0.0]])}, np.array([3.0, 0.2, 0.4, -0.3, 0.0, 0, 0]))
    assert od.status == "optimal" and u.shape == (2, 1) and od.omega[0] <= 1.0 + 1e-9
    m = MPCCBF(r, {"model": "DynamicUnicycle2D", "mpc_horizon": 8}, num_obs=4)
    ref = {"u_ref": u_ref, "state_machine": "stop", "goal": np.array([5.0, 0.0])}
    assert m.solve_control_problem(r.X, ref, None) is u_ref             # not tracking -> u_ref (mpc_cbf.py:379-381)
    ref["state_machine"] = "track"
    u = m.solve_control_problem(r.X, ref, np.array([[2.0, 0.1, 0.3, 0, 0, 0, 0]]))
    assert m.status == "optimal" and m.solver_status == "optimal" and u.shape == (2, 1)
    assert m.pred_x.shape == (9, 4) and np.allclose(m.u_prev[0], u[:, 0])
