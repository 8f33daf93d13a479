"""CPU-only: libscb.so builds (nvcc cross-compiles), loads, exports every function include/scb.h declares,
the ctypes struct mirror matches sizeof(scb_params), and the parameter defaults are the reference's.
No compute call is made here (there is no GPU and no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from safe_control_b200 import _abi


@pytest.fixture(scope="module")
def lib():
    from safe_control_b200 import build
    build.build()
    from safe_control_b200._lib import lib as load
    return load()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "scb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_functions()
    assert len(names) >= 18 and "scb_mpccbf_solve_host" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/scb.h but not exported by libscb.so"
        assert n in _abi.PROTOTYPES, f"{n} has no ctypes prototype in _abi.py"
    assert set(_abi.PROTOTYPES) == set(names)


def test_struct_mirror_and_helpers(lib):
    assert lib.scb_params_sizeof() == C.sizeof(_abi.ScbParams)
    assert lib.scb_track_sizeof() == C.sizeof(_abi.ScbTrack)
    assert lib.scb_version() == 210
    assert b"ok" == lib.scb_strerror(0)
    assert lib.scb_active_words(16, 2) == 1 and lib.scb_active_words(61, 2) == 2
    nx, nu = C.c_int(), C.c_int()
    assert lib.scb_model_dims(4, C.byref(nx), C.byref(nu)) == 0 and (nx.value, nu.value) == (12, 4)
    assert lib.scb_model_dims(99, None, None) == -1
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    assert lib.scb_limits(C.byref(a), C.byref(b), C.byref(c)) == 0 and a.value >= 64 and b.value >= 64 and c.value >= 10


def test_reference_defaults(lib):
    """cbf_qp.py:12-35, optimal_decay_cbf_qp.py:17-50, mpc_cbf.py:19-82, robots/*.py ctor defaults."""
    p = _abi.ScbParams()
    assert lib.scb_params_default(p, 1, b"cbf_qp") == 0
    assert (p.alpha1, p.alpha2, p.nx, p.nu, p.radius, p.dt) == (1.5, 1.5, 4, 2, 0.25, 0.05)
    assert list(p.u_ub)[:2] == [0.5, 0.5]
    assert lib.scb_params_default(p, 3, b"optimal_decay_cbf_qp") == 0
    assert (p.alpha, p.omega1_0, p.p_sb1) == (0.5, 1.0, 1e4) and abs(p.u_ub[1] - 0.30282535497070506) < 1e-15
    assert lib.scb_params_default(p, 2, b"mpc_cbf") == 0
    assert list(p.Q)[:4] == [50, 50, 1, 1] and list(p.R)[:2] == [0.5, 5000.0] and (p.alpha1, p.alpha2) == (0.1, 0.1)
    assert lib.scb_params_default(p, 4, b"mpc_cbf") == 0 and p.alpha == 0.15 and list(p.Q)[:3] == [30, 30, 5]
    assert lib.scb_params_default(p, 4, b"cbf_qp") == -2          # Quad3D: agent_barrier raises (quad3D.py:269-273)
    assert lib.scb_params_default(p, 0, b"optimal_decay_cbf_qp") == -2   # NotCompatibleError in the reference
    assert lib.scb_params_default(p, 1, b"nope") == -1


def test_robot_spec_overrides(lib):
    from safe_control_b200.params import resolve_params, NotCompatibleError
    p, s = resolve_params({"model": "DynamicUnicycle2D", "a_max": 1.0, "cbf_alpha1": 0.7, "cbf_mode": "hard", "radius": 0.3}, "cbf_qp")
    assert (p.u_ub[0], p.alpha1, p.alpha2, p.cbf_mode, p.radius) == (1.0, 0.7, 1.5, 1, 0.3)
    p, s = resolve_params({"model": "KinematicBicycle2D"}, "mpc_cbf")
    assert p.radius == 0.25 and s["v_min"] == 0.2          # robots/robot.py:49 beats the model's 0.3 default
    with pytest.raises(NotCompatibleError):
        resolve_params({"model": "SingleIntegrator2D"}, "optimal_decay_cbf_qp")
    p, s = resolve_params({"model": "VTOL2D", "pitch_max": 20.0}, "mpc_cbf")     # mpc_cbf.py:40-43, 83-87, 222-232
    assert (p.nx, p.nu, s["mpc_horizon"], p.pitch_max, p.alpha1, list(p.R)) == (6, 4, 30, 20.0, 0.05, [0.5, 0.5, 0.5, 50000.0])
    with pytest.raises(NotCompatibleError):
        resolve_params({"model": "VTOL2D"}, "cbf_qp")             # agent_barrier is not implemented (vtol2D.py:458-460)
    with pytest.raises(ValueError):
        resolve_params({"model": "Hovercraft"}, "mpc_cbf")
    p, s = resolve_params({"model": "KinematicBicycle2D"}, "optimal_decay_mpc_cbf")   # optimal_decay_mpc_cbf.py:37-39, 73-75, 87-90
    assert (p.od_mpc, p.od_sum_rterms, list(p.R)[:2], p.alpha1, p.p_sb1, p.omega1_0, s["mpc_horizon"]) == (1, 0, [0.5, 50.0], 0.05, 10.0, 1.0, 10)
    p, s = resolve_params({"model": "Unicycle2D", "w_max": 1.0}, "mpc_cbf")
    assert (p.nx, p.nu, p.alpha, p.u_ub[1], p.Q[2]) == (3, 2, 0.05, 1.0, 0.01)
    with pytest.raises(NotCompatibleError):
        resolve_params({"model": "Unicycle2D"}, "optimal_decay_cbf_qp")


def test_mpc_launch_count_and_workspace_are_host_arithmetic(lib):
    """scb_mpccbf_launch_count / scb_mpccbf_workspace_bytes need no device: 1 launch in index order, 3 with the
    hardest-first schedule once the batch exceeds one persistent wave (148 SMs x agents per CTA), +1 general-row launch
    when superellipsoid rows are enabled."""
    from safe_control_b200.params import resolve_params
    p, _ = resolve_params({"model": "DynamicUnicycle2D"}, "mpc_cbf")
    assert lib.scb_mpccbf_workspace_bytes(4096) == 4 * (16 + 1024 + 2 * 4096)
    assert lib.scb_mpccbf_launch_count(p, 4096, 16, 8, 0) == 1
    assert lib.scb_mpccbf_launch_count(p, 4096, 16, 8, 1) == 3
    assert lib.scb_mpccbf_launch_count(p, 64, 16, 8, 1) == 1            # fits the first wave: nothing to schedule
    p2, _ = resolve_params({"model": "DynamicUnicycle2D", "mpc_superellipsoid": True}, "mpc_cbf")
    assert lib.scb_mpccbf_launch_count(p2, 4096, 16, 8, 1) == 4
    assert lib.scb_mpccbf_launch_count(p, 4096, 16, 40, 1) < 0          # horizon beyond the compiled limit


def test_no_cpu_fallback():
    """Solves must refuse to run without a CUDA device instead of silently using the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from safe_control_b200 import BatchedCBFQP
    from safe_control_b200._lib import ScbError
    ctrl = BatchedCBFQP({"model": "DynamicUnicycle2D"}, num_obs=2)
    with pytest.raises(ScbError):
        ctrl.solve(torch.zeros(1, 4, dtype=torch.float64), torch.zeros(1, 2, dtype=torch.float64),
                   torch.zeros(1, 2, 7, dtype=torch.float64))


def test_tracker_initial_state_padding_matches_reference_rules():
    """pad_initial_state restates tracking.py:60-99 + robots/robot.py:65-131 for every model the loop drives."""
    import numpy as np
    from safe_control_b200.tracking import pad_initial_state
    X, yaw = pad_initial_state("SingleIntegrator2D", [[1.0, 2.0]])
    assert X.tolist() == [[1.0, 2.0]] and yaw.tolist() == [0.0]
    X, yaw = pad_initial_state("DoubleIntegrator2D", [[1.0, 2.0, 0.5]])                 # [x, y, theta] -> [x, y, 0, 0], yaw
    assert X.tolist() == [[1.0, 2.0, 0.0, 0.0]] and yaw.tolist() == [0.5]
    X, yaw = pad_initial_state("DoubleIntegrator2D", [[1.0, 2.0, 0.3, -0.2, 0.5]])
    assert X.tolist() == [[1.0, 2.0, 0.3, -0.2]] and yaw.tolist() == [0.5]
    X, yaw = pad_initial_state("DynamicUnicycle2D", [[1.0, 2.0, 0.5]])                  # velocity 0 appended
    assert X.tolist() == [[1.0, 2.0, 0.5, 0.0]] and yaw.tolist() == [0.5]
    X, yaw = pad_initial_state("Unicycle2D", [[1.0, 2.0, 0.5]])
    assert X.tolist() == [[1.0, 2.0, 0.5]] and yaw.tolist() == [0.5]
    X, yaw = pad_initial_state("Quad2D", [[1.0, 2.0, 0.5]])                             # only (x, z) kept (tracking.py:81-83)
    assert X.tolist() == [[1.0, 2.0, 0.0, 0.0, 0.0, 0.0]] and yaw.tolist() == [0.0]
    X, yaw = pad_initial_state("Quad3D", [[1.0, 2.0, 3.0, 0.5]])
    assert X.shape == (1, 12) and X[0, :3].tolist() == [1.0, 2.0, 3.0] and X[0, 5] == 0.5 and yaw.tolist() == [0.5]
    with pytest.raises(ValueError):
        pad_initial_state("Unicycle2D", [[1.0, 2.0]])


def test_struct_mirrors_field_by_field(lib):
    """offsetof of every field of scb_params / scb_track as the C compiler laid them out vs the ctypes mirrors (a renamed
    or reordered field of equal size would pass a sizeof-only check)."""
    for struct, fn in ((_abi.ScbParams, lib.scb_params_offsetof), (_abi.ScbTrack, lib.scb_track_offsetof)):
        for name, _ in struct._fields_:
            assert fn(name.encode()) == getattr(struct, name).offset, (struct.__name__, name)
    assert lib.scb_params_offsetof(b"no_such_field") == -1
