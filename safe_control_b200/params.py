"""robot_spec dict -> scb_params POD.

Mirrors the reference's `setdefault` cascades and gain overrides:
  robots/robot.py:49, robots/<model>.py ctor defaults      (model parameters)
  position_control/cbf_qp.py:12-43                         ('cbf_alpha*', 'cbf_mode')
  position_control/optimal_decay_cbf_qp.py:17-50
  position_control/mpc_cbf.py:15,19-95                     ('mpc_cbf_alpha*', 'mpc_horizon')
The library supplies the defaults (scb_params_default); this module only applies
what the caller put in robot_spec, using the reference's own key names.
"""
import math

from ._abi import MODEL_IDS, ScbParams, ERR_UNSUPPORTED

CONTROLLERS = ("cbf_qp", "optimal_decay_cbf_qp", "mpc_cbf", "optimal_decay_mpc_cbf")


class NotCompatibleError(Exception):
    """Same name as position_control/optimal_decay_cbf_qp.py:4-12."""

    def __init__(self, message="Currently not compatible with the robot model."):
        self.message = message
        super().__init__(self.message)


def resolve_params(robot_spec, controller, dt=0.05, lib=None):
    """-> (ScbParams, resolved_spec dict).  `lib` defaults to the CUDA library."""
    if lib is None:
        from ._lib import lib as _get
        lib = _get()
    if controller not in CONTROLLERS:
        raise ValueError(f"Unknown controller type: {controller}")
    model = robot_spec.get("model")
    if model not in MODEL_IDS:
        raise ValueError(f"Invalid robot model: {model!r} (supported: {sorted(MODEL_IDS)})")
    p = ScbParams()
    rc = lib.scb_params_default(p, MODEL_IDS[model], controller.encode())
    if rc == ERR_UNSUPPORTED:
        raise NotCompatibleError(f"{controller} is not compatible with {model}")
    if rc != 0:
        raise RuntimeError(lib.scb_strerror(rc).decode())
    s = dict(robot_spec)
    p.dt = float(dt)
    p.radius = float(s.setdefault("radius", 0.25))
    p.cbf_mode = 1 if s.get("cbf_mode", "cbf") == "hard" else 0

    if model == "SingleIntegrator2D":
        v = float(s.setdefault("v_max", 1.0))
        s.setdefault("w_max", 0.5)
        p.u_lb[0] = p.u_lb[1] = -v
        p.u_ub[0] = p.u_ub[1] = v
    elif model == "Unicycle2D":                               # unicycle2D.py:40-41
        v = float(s.setdefault("v_max", 1.0)); w = float(s.setdefault("w_max", 0.5))
        p.u_lb[0], p.u_ub[0], p.u_lb[1], p.u_ub[1] = -v, v, -w, w
    elif model == "Manipulator2D":                            # manipulator2D.py:19-20
        w = float(s.setdefault("w_max", 2.0)); s.setdefault("Kp", 3.0)
        for i in range(3):
            p.u_lb[i], p.u_ub[i] = -w, w
    elif model == "DynamicUnicycle2D":
        a = float(s.setdefault("a_max", 0.5)); w = float(s.setdefault("w_max", 0.5))
        v = float(s.setdefault("v_max", 1.0))
        p.u_lb[0], p.u_ub[0], p.u_lb[1], p.u_ub[1] = -a, a, -w, w
        p.v_max = v; p.v_min = -v
    elif model.startswith("KinematicBicycle2D"):
        s.setdefault("wheel_base", 0.4); s.setdefault("front_ax_dist", 0.2)
        lr = float(s.setdefault("rear_ax_dist", 0.2))
        v = float(s.setdefault("v_max", 3.5)); a = float(s.setdefault("a_max", 5.0))
        dmax = float(s.setdefault("delta_max", math.radians(32)))
        b = float(s.setdefault("beta_max", math.atan(lr / s["wheel_base"] * math.tan(dmax))))
        vmin = float(s.setdefault("v_min", 0.2))
        p.rear_ax_dist = lr
        p.u_lb[0], p.u_ub[0], p.u_lb[1], p.u_ub[1] = -a, a, -b, b
        p.v_min, p.v_max = vmin, v
    elif model == "DoubleIntegrator2D":                       # double_integrator2D.py:40-44
        a = float(s.setdefault("a_max", 1.0)); v = float(s.setdefault("v_max", 1.0))
        s.setdefault("ax_max", a); s.setdefault("ay_max", a); s.setdefault("w_max", 0.5)
        p.u_lb[0] = p.u_lb[1] = -a                              # cbf_qp.py:66-69 bounds both inputs by a_max
        p.u_ub[0] = p.u_ub[1] = a
        if controller in ("mpc_cbf", "optimal_decay_mpc_cbf"):   # mpc_cbf.py:200-204 uses ax_max / ay_max
            ax, ay = float(s["ax_max"]), float(s["ay_max"])
            p.u_lb[0], p.u_ub[0], p.u_lb[1], p.u_ub[1] = -ax, ax, -ay, ay
        p.v_max = v; p.v_min = -v
    elif model == "Quad2D":                                   # quad2D.py:40-46
        p.mass = float(s.setdefault("mass", 1.0)); p.Iy = float(s.setdefault("inertia", 0.01))
        fmin = float(s.setdefault("f_min", 1.0)); fmax = float(s.setdefault("f_max", 10.0))
        p.u_lb[0] = p.u_lb[1] = fmin
        p.u_ub[0] = p.u_ub[1] = fmax
    elif model == "Quad3D":
        p.mass = float(s.setdefault("mass", 3.0))
        p.Ix = float(s.setdefault("Ix", 0.5)); p.Iy = float(s.setdefault("Iy", 0.5)); p.Iz = float(s.setdefault("Iz", 0.5))
        p.arm_L = float(s.setdefault("L", 0.3)); p.nu_coef = float(s.setdefault("nu", 0.1))
        umax = float(s.setdefault("u_max", 10.0)); umin = float(s.setdefault("u_min", -10.0))
        for i in range(4):
            p.u_lb[i], p.u_ub[i] = umin, umax

    elif model == "VTOL2D":                                   # vtol2D.py:57-110 (same key names)
        p.mass = float(s.setdefault("mass", 11.0)); p.Iy = float(s.setdefault("inertia", 1.135))
        for key, field, dflt in (("S_wing", "S_wing", 0.55), ("rho", "rho", 1.2682), ("C_L0", "C_L0", 0.23), ("C_Lalpha", "C_Lalpha", 5.61),
                                 ("M", "blend_M", 50.0), ("alpha_0", "alpha_0", math.radians(15)), ("C_Ldelta_e", "C_Ldelta_e", 0.13),
                                 ("C_D0", "C_D0", 0.043), ("C_Dalpha", "C_Dalpha", 0.03), ("C_Ddelta_e", "C_Ddelta_e", 0.0),
                                 ("C_m0", "C_m0", 0.0135), ("C_malpha", "C_malpha", -2.74), ("C_mdelta_e", "C_mdelta_e", -0.99),
                                 ("chord", "chord", 0.18994), ("k_front", "k_front", 70.0), ("k_rear", "k_rear", 70.0),
                                 ("k_pusher", "k_pusher", 60.0), ("ell_f", "ell_f", 0.5), ("ell_r", "ell_r", 0.5),
                                 ("pitch_max", "pitch_max", 15.0), ("descent_speed_max", "descent_speed_max", 5.0)):
            setattr(p, field, float(s.setdefault(key, dflt)))
        v = float(s.setdefault("v_max", 15.0)); p.v_max = v; p.v_min = -v
        tmin = float(s.setdefault("throttle_min", 0.0)); tmax = float(s.setdefault("throttle_max", 1.0))
        emin = float(s.setdefault("elevator_min", -0.5)); emax = float(s.setdefault("elevator_max", 0.5))
        for i in range(3):
            p.u_lb[i], p.u_ub[i] = tmin, tmax
        p.u_lb[3], p.u_ub[3] = emin, emax
        s.setdefault("mpc_horizon", 30)                         # mpc_cbf.py:41

    prefix = {"cbf_qp": "cbf_", "mpc_cbf": "mpc_cbf_"}.get(controller)
    if prefix:
        for k in ("alpha", "alpha1", "alpha2"):
            if prefix + k in s:
                setattr(p, k, float(s[prefix + k]))
    if controller == "mpc_cbf" and s.get("mpc_superellipsoid"):
        if model not in ("SingleIntegrator2D", "DynamicUnicycle2D", "DoubleIntegrator2D"):
            raise ValueError(f"{model} has no superellipsoid branch in its agent_barrier_dt")
        p.mpc_superellipsoid = 1
    if controller == "optimal_decay_mpc_cbf":                  # ours: which of the two do-mpc rterm readings (include/scb.h)
        p.od_sum_rterms = int(bool(s.get("od_sum_rterms", False)))
        if model == "VTOL2D":
            s["mpc_horizon"] = 30                               # optimal_decay_mpc_cbf.py:47 (hard-wired, like 10 for the others)
        else:
            s["mpc_horizon"] = int(s.get("od_mpc_horizon", 10))  # :24 ignores robot_spec['mpc_horizon']
    if "mpc_max_iter" in s:
        p.mpc_max_iter = int(s["mpc_max_iter"])
    if "mpc_tol" in s:
        p.mpc_tol = float(s["mpc_tol"])
    return p, s


def cbf_param_dict(p, controller, model):
    """The `.cbf_param` dict the reference exposes (tracking.py:738 reads alpha1/alpha2)."""
    rel2 = model in ("DynamicUnicycle2D", "KinematicBicycle2D", "DoubleIntegrator2D", "Quad2D", "VTOL2D")
    d = {"alpha1": p.alpha1, "alpha2": p.alpha2} if rel2 else {"alpha": p.alpha}
    if controller == "optimal_decay_cbf_qp":
        d.update(omega1=p.omega1_0, p_sb1=p.p_sb1)
        if rel2:
            d.update(omega2=p.omega2_0, p_sb2=p.p_sb2)
    if controller == "optimal_decay_mpc_cbf":                  # optimal_decay_mpc_cbf.py:87-90
        d.update(omega1=p.omega1_0, p_sb1=p.p_sb1, omega2=p.omega2_0, p_sb2=p.p_sb2)
    return d
