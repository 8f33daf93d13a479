"""The zero-change integration route, end to end on the CPU box: the UNMODIFIED reference tracking.py
(/root/reference, imported through oracle/refshim's stand-ins for casadi / matplotlib / shapely) constructs its position
controller through `from safe_control.position_control.cbf_qp import CBFQP` (tracking.py:140-148) -- and gets OUR class,
because safe_control_b200.shadow.install() put the shadow modules first on the package's search path.  The shim classes
are backed by the CPU build of the kernel bodies here (tests/_hostsim; on a GPU box they call libscb.so), and the closed
loop must reproduce the golden run that the reference produced with ITS OWN CBFQP (tests/golden/ref_tracking.npz).

Runs only where /root/reference exists (the build container); skipped on the GPU box."""
import contextlib
import io
import os
import sys
from unittest import mock

import numpy as np
import pytest

from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference checkout not present (GPU box)")


class HostSimCtx:
    """HostContext look-alike over the CPU host-sim (same per-agent source as the CUDA kernels)."""

    def cbfqp_solve(self, params, M, X, U_ref, OBS, nobs=None, out=None, want_active=True):
        from hostsim_util import hs_cbfqp_solve
        return hs_cbfqp_solve(params, X, U_ref, OBS, nobs)

    def odcbf_solve(self, params, M, X, U_ref, OBS, nobs=None, out=None):
        from hostsim_util import hs_odcbf_solve
        return hs_odcbf_solve(params, X, U_ref, OBS, nobs)

    def mpccbf_solve(self, params, M, H, X, goal, u_prev, OBS, nobs=None, U_ref=None, track=None, want_pred=False,
                     want_active=False):
        from hostsim_util import hs_mpccbf_solve
        return hs_mpccbf_solve(params, H, X, goal, u_prev, OBS, nobs)


@pytest.fixture()
def reference_tracking(monkeypatch):
    plt = mock.MagicMock(name="matplotlib.pyplot")
    plt.colormaps.get_cmap.return_value.colors = [(0.5, 0.5, 0.5)] * 9          # robots/robot.py:45-47 indexes the palette
    mpl = mock.MagicMock(name="matplotlib"); mpl.pyplot = plt
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.startswith(("safe_control.", "matplotlib", "shapely"))
             or k in ("safe_control", "casadi", "cvxpy", "do_mpc")}
    sys.modules["matplotlib.pyplot"] = plt; sys.modules["matplotlib"] = mpl
    for n in ("shapely", "shapely.geometry", "shapely.ops", "shapely.validation"):
        sys.modules[n] = mock.MagicMock(name=n)
    refshim.install()
    import safe_control_b200.shadow as shadow
    import safe_control_b200.position_control._common as common
    from hostsim_util import hostsim
    import safe_control_b200.params as params_mod
    monkeypatch.setattr(common, "host_ctx", lambda device=0: HostSimCtx())
    for modname in ("cbf_qp", "mpc_cbf", "optimal_decay_cbf_qp"):                # the shims bound host_ctx at import time
        m = sys.modules.get(f"safe_control_b200.position_control.{modname}")
        if m is not None:
            monkeypatch.setattr(m, "host_ctx", lambda device=0: HostSimCtx())
    # parameter resolution needs scb_params_default: on the CPU box take it from the host-sim build of scb_params.cc
    orig_resolve = params_mod.resolve_params
    patched = lambda spec, controller, dt=0.05, lib=None: orig_resolve(spec, controller, dt, lib=lib or hostsim())
    monkeypatch.setattr(params_mod, "resolve_params", patched)
    for modname in ("cbf_qp", "mpc_cbf", "optimal_decay_cbf_qp"):
        m = sys.modules.get(f"safe_control_b200.position_control.{modname}")
        if m is not None:
            monkeypatch.setattr(m, "resolve_params", patched)
    shadow.install()
    try:
        from safe_control.tracking import LocalTrackingController
        from safe_control.utils import env as envmod
        from safe_control.utils.headless_plot import NullArtist, NullAxes, NullFigure
        yield LocalTrackingController, envmod, NullArtist, NullAxes, NullFigure
    finally:
        shadow.uninstall()
        for k in [k for k in sys.modules if k.startswith("safe_control.") or k == "safe_control"]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


@pytest.mark.parametrize("name,steps", [("du_stop_rotate", 120), ("du_test_tracking", 150), ("si_test_tracking", 150)])
def test_unmodified_tracking_uses_the_shadow_controller(reference_tracking, name, steps):
    from track_util import load_tracking_golden
    LocalTrackingController, envmod, NullArtist, NullAxes, NullFigure = reference_tracking
    d = load_tracking_golden()[name]

    class Ax(NullAxes):
        patches = []

        def __getattr__(self, _n):
            return lambda *a, **k: NullArtist()

    spec = dict(d["spec"]); spec["num_constraints"] = int(d["M"])
    x0 = d["X"][0].copy()
    if spec["model"] == "SingleIntegrator2D":
        x0 = np.append(x0, d["yaw"][0])
    with contextlib.redirect_stdout(io.StringIO()):
        tc = LocalTrackingController(x0, spec, controller_type={"pos": "cbf_qp"}, dt=0.05, show_animation=False,
                                     enable_rotation=bool(d["enable_rotation"]), env=envmod.Env(), ax=Ax(NullFigure()),
                                     fig=NullFigure())
        tc.obs = d["scene"][0].copy()
        tc.set_waypoints(d["waypoints"])
    # the lazy import of tracking.py:140-148 resolved to OUR class
    import safe_control_b200.position_control.cbf_qp as ours
    assert type(tc.pos_controller) is ours.CBFQP
    assert type(tc.pos_controller).__module__ == "safe_control_b200.position_control.cbf_qp"
    assert sys.modules["safe_control.position_control.cbf_qp"].__file__.startswith(
        os.path.dirname(os.path.abspath(sys.modules["safe_control_b200.shadow"].__file__)))
    n = min(steps, len(d["ret"]))
    for k in range(n):
        np.testing.assert_allclose(tc.robot.X.reshape(-1), d["X"][k], rtol=0, atol=1e-9, err_msg=f"{name} step {k}")
        with contextlib.redirect_stdout(io.StringIO()):
            ret = tc.control_step()
        assert ret == d["ret"][k], (name, k, ret, d["ret"][k])
        assert (tc.pos_controller.status == "optimal") == (d["status"][k] == 0)
