"""ctypes mirror of include/scb.h (struct scb_params + prototypes)."""
import ctypes as C

MODEL_IDS = {
    "SingleIntegrator2D": 0,
    "DynamicUnicycle2D": 1,
    "KinematicBicycle2D": 2,
    "KinematicBicycle2D_C3BF": 3,
    "Quad3D": 4,
    "DoubleIntegrator2D": 5,
    "Quad2D": 6,
    "Unicycle2D": 8,
    "Manipulator2D": 9,
    "KinematicBicycle2D_DPCBF": 7,
    "VTOL2D": 10,
}
MODEL_NAMES = {v: k for k, v in MODEL_IDS.items()}
MODEL_DIMS = {0: (2, 2), 1: (4, 2), 2: (4, 2), 3: (4, 2), 4: (12, 4), 5: (4, 2), 6: (6, 2), 7: (4, 2), 8: (3, 2), 9: (3, 3), 10: (6, 4)}

OPTIMAL, INFEASIBLE, MAXITER, NUMERICAL = 0, 1, 2, 3
STATUS_STR = {0: "optimal", 1: "infeasible", 2: "user_limit", 3: "solver_error"}   # cvxpy's vocabulary

ERR_UNSUPPORTED = -2


class ScbParams(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("cbf_mode", C.c_int32), ("nx", C.c_int32), ("nu", C.c_int32),
        ("dt", C.c_double), ("radius", C.c_double),
        ("alpha", C.c_double), ("alpha1", C.c_double), ("alpha2", C.c_double),
        ("u_lb", C.c_double * 4), ("u_ub", C.c_double * 4),
        ("v_min", C.c_double), ("v_max", C.c_double), ("rear_ax_dist", C.c_double),
        ("omega1_0", C.c_double), ("omega2_0", C.c_double), ("p_sb1", C.c_double), ("p_sb2", C.c_double),
        ("Q", C.c_double * 12), ("R", C.c_double * 4),
        ("mass", C.c_double), ("Ix", C.c_double), ("Iy", C.c_double), ("Iz", C.c_double),
        ("arm_L", C.c_double), ("nu_coef", C.c_double), ("gravity", C.c_double),
        ("mpc_max_iter", C.c_int32), ("mpc_superellipsoid", C.c_int32), ("mpc_tol", C.c_double),
        # VTOL2D (robots/vtol2D.py:57-110)
        ("S_wing", C.c_double), ("rho", C.c_double), ("C_L0", C.c_double), ("C_Lalpha", C.c_double), ("blend_M", C.c_double),
        ("alpha_0", C.c_double), ("C_Ldelta_e", C.c_double), ("C_D0", C.c_double), ("C_Dalpha", C.c_double),
        ("C_Ddelta_e", C.c_double), ("C_m0", C.c_double), ("C_malpha", C.c_double), ("C_mdelta_e", C.c_double),
        ("chord", C.c_double), ("k_front", C.c_double), ("k_rear", C.c_double), ("k_pusher", C.c_double),
        ("ell_f", C.c_double), ("ell_r", C.c_double), ("pitch_max", C.c_double), ("descent_speed_max", C.c_double),
        ("od_mpc", C.c_int32), ("od_sum_rterms", C.c_int32),
    ]


class ScbBackupParams(C.Structure):
    """Mirror of `struct scb_backup_params` (include/scb.h): evade scene + Backup-CBF parameters."""
    _fields_ = [(k, C.c_double) for k in (
        "hallway_length", "half_width", "pocket_x_min", "pocket_x_max", "pocket_y_min", "pocket_y_max", "center_x", "center_y",
        "goal_x_min", "goal_x_max", "goal_y_min", "goal_y_max", "radius", "a_max", "v_max", "safety_margin", "Kp", "Kd",
        "dt", "backup_horizon", "alpha", "alpha_terminal", "q0", "q1")] + [("use_goal", C.c_int32), ("n_backup", C.c_int32)]


class ScbShieldParams(C.Structure):
    """Mirror of `struct scb_shield_params` (include/scb.h)."""
    _fields_ = [("scene", ScbBackupParams), ("event_offset", C.c_double), ("mode", C.c_int32), ("discount_steps", C.c_int32),
                ("nom_cap", C.c_int32), ("reserved", C.c_int32)]


class ScbShieldState(C.Structure):
    """Mirror of `struct scb_shield_state`: device pointers of the per-agent shield state."""
    _fields_ = [("CU", C.c_void_p), ("CX", C.c_void_p), ("clen", C.c_void_p), ("cidx", C.c_void_p), ("nsteps", C.c_void_p),
                ("next_event", C.c_void_p), ("cbuf", C.c_void_p), ("work", C.c_void_p)]


_P = C.POINTER(ScbParams)
_B = C.POINTER(ScbBackupParams)
_vp = C.c_void_p

SM_IDLE, SM_TRACK, SM_STOP, SM_ROTATE = 0, 1, 2, 3
SM_NAMES = {0: "idle", 1: "track", 2: "stop", 3: "rotate"}       # tracking.py:49
SM_IDS = {v: k for k, v in SM_NAMES.items()}
CONTROLLER_IDS = {"cbf_qp": 0, "optimal_decay_cbf_qp": 1, "mpc_cbf": 2}


class ScbTrack(C.Structure):
    """Mirror of `struct scb_track` (include/scb.h): closed-loop configuration + caller-owned arrays."""
    _fields_ = [
        ("controller", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("M", C.c_int32), ("W", C.c_int32),
        ("H", C.c_int32), ("enable_rotation", C.c_int32), ("dynamic_obs", C.c_int32),
        ("att_velocity_tracking", C.c_int32), ("mpc_strict", C.c_int32),
        ("reached_threshold", C.c_double), ("rotation_threshold", C.c_double),
        ("k_omega", C.c_double), ("k_a", C.c_double), ("k_v", C.c_double), ("k_a_stop", C.c_double),
        ("w_max", C.c_double), ("att_kp", C.c_double), ("wheel_base", C.c_double), ("delta_max", C.c_double),
        ("X", _vp), ("yaw", _vp), ("sm", _vp), ("wp_idx", _vp), ("WP", _vp), ("nwp", _vp), ("goal", _vp),
        ("has_goal", _vp), ("u_att", _vp), ("u_prev", _vp), ("ret", _vp), ("done", _vp), ("nsteps", _vp),
        ("SCENE", _vp),
        ("Uref", _vp), ("OBS", _vp), ("nobs", _vp), ("U", _vp), ("status", _vp), ("active", _vp),
        ("track_flag", _vp), ("mpc_iters", _vp), ("mpc_ws", _vp), ("mpc_ws_bytes", C.c_uint64),
        ("mpc_fail", _vp),
    ]


_T = C.POINTER(ScbTrack)

# name -> (restype, argtypes); every symbol include/scb.h declares
PROTOTYPES = {
    "scb_version": (C.c_int, []),
    "scb_params_sizeof": (C.c_size_t, []),
    "scb_params_offsetof": (C.c_long, [C.c_char_p]),
    "scb_track_offsetof": (C.c_long, [C.c_char_p]),
    "scb_strerror": (C.c_char_p, [C.c_int]),
    "scb_last_cuda_error": (C.c_int, []),
    "scb_device_count": (C.c_int, []),
    "scb_params_default": (C.c_int, [_P, C.c_int, C.c_char_p]),
    "scb_model_dims": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "scb_active_words": (C.c_int, [C.c_int, C.c_int]),
    "scb_limits": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "scb_measure_fp64_peak": (C.c_int, [C.POINTER(C.c_double), _vp]),
    "scb_measure_latency_floor": (C.c_int, [C.POINTER(C.c_double)] * 3 + [_vp]),
    "scb_ctx_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "scb_ctx_destroy": (None, [_vp]),
    "scb_ctx_launches": (C.c_long, [_vp]),
    "scb_cbfqp_rows": (C.c_int, [_P, C.c_int, C.c_int, _vp, _vp, C.c_long, _vp, _vp, _vp, _vp]),
    "scb_cbfqp_solve": (C.c_int, [_P, C.c_int, C.c_int, _vp, _vp, _vp, C.c_long, _vp, _vp, _vp, _vp, _vp]),
    "scb_cbfqp_solve_host": (C.c_int, [_vp, _P, C.c_int, C.c_int, _vp, _vp, _vp, C.c_long, _vp, _vp, _vp, _vp]),
    "scb_odcbf_solve": (C.c_int, [_P, C.c_int, C.c_int, _vp, _vp, _vp, C.c_long, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "scb_odcbf_solve_host": (C.c_int, [_vp, _P, C.c_int, C.c_int, _vp, _vp, _vp, C.c_long, _vp, _vp, _vp, _vp, _vp, _vp]),
    "scb_mpc_active_words": (C.c_int, [_P, C.c_int, C.c_int]),
    "scb_mpccbf_solve": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.c_long, _vp,
                                   _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "scb_mpccbf_workspace_bytes": (C.c_size_t, [C.c_int]),
    "scb_mpccbf_launch_count": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "scb_mpccbf_solve_ws": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.c_long, _vp,
                                      _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "scb_mpccbf_solve_host": (C.c_int, [_vp, _P, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.c_long, _vp,
                                        _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "scb_backup_params_default": (None, [_B]),
    "scb_backup_params_sizeof": (C.c_size_t, []),
    "scb_backup_active_words": (C.c_int, [C.c_int]),
    "scb_backupcbf_solve": (C.c_int, [_B, C.c_int, C.c_int, _vp, _vp, _vp, C.c_long, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "scb_backupcbf_solve_host": (C.c_int, [_vp, _B, C.c_int, C.c_int, _vp, _vp, _vp, C.c_long, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "scb_shield_params_sizeof": (C.c_size_t, []),
    "scb_shield_step": (C.c_int, [C.POINTER(ScbShieldParams), C.POINTER(ScbShieldState), C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp,
                                  C.c_long, _vp, _vp, _vp, _vp]),
    "scb_select_obstacles": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_long, _vp, _vp, _vp, _vp]),
    "scb_track_sizeof": (C.c_size_t, []),
    "scb_control_step": (C.c_int, [_P, _T, _vp]),
    "scb_run_all_steps": (C.c_int, [_P, _T, C.c_int, _vp]),
    "scb_run_all_steps_launches": (C.c_long, [_P, _T, C.c_int]),
}


def bind(lib, names=None):
    """Attach restype/argtypes; raises AttributeError if a declared symbol is missing."""
    for name, (res, args) in PROTOTYPES.items():
        if names is not None and name not in names:
            continue
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
