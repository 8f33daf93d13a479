"""refshim -- run the *reference's own Python sources* in this container.

TEST INFRASTRUCTURE ONLY (fixture generation).  Never imported by the product
package, by `-m gpu` tests, by smoke() or by bench.py: /root/reference does
not exist on the GPU box.  The only consumer is tests/golden/gen_from_reference.py,
which is run HERE once and whose outputs (small .npz files) are committed.

The reference (tkkim-robot/safe_control @ 609c6bc) cannot be imported as-is:
every model file does `import casadi as ca` at module top
(robots/dynamic_unicycle2D.py:2), the QP controllers `import cvxpy as cp`
(position_control/cbf_qp.py:2) and solve with GUROBI (cbf_qp.py:190); none of
those are installed and there is no network.  install() registers

  * `casadi`  -> a numpy-backed *numeric* stand-in (fake_casadi.py): the
    reference's "casadi" code paths (`step(..., casadi=True)`,
    `agent_barrier_dt`) are then evaluated on plain float64 numbers by the
    reference's own formulas;
  * `cvxpy`   -> a minimal affine-expression layer (fake_cvxpy.py) that
    extracts (P, q, G, h) from the reference's own problem statement and
    hands it to oracle/qp_exact.py.  The QPs are strictly convex, so the
    optimum is unique and solver independent -- what this pins is the
    reference's row assembly, constants, bounds and objective;
  * `do_mpc`  -> a probing stand-in (fake_do_mpc.py): records what the
    reference's mpc_cbf.py hands to do-mpc (rhs, cost, CBF constraints, bounds,
    rterm, horizon, tvp padding) evaluated at numeric probe points; nothing is
    solved;
  * `matplotlib*` -> inert mocks (robots/kinematic_bicycle2D.py:4-5);
  * `safe_control` -> a namespace package rooted at the reference checkout
    (pyproject.toml:25-27 maps the package to the repo root).

Nothing is copied from the reference; its files are imported where they lie.
"""
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("SAFE_CONTROL_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "position_control", "cbf_qp.py"))


def install():
    """Register the stub modules and the `safe_control` namespace. Idempotent."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    from . import fake_casadi, fake_cvxpy, fake_do_mpc

    sys.modules.setdefault("casadi", fake_casadi)
    sys.modules.setdefault("cvxpy", fake_cvxpy)
    sys.modules.setdefault("do_mpc", fake_do_mpc)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.transforms",
                 "matplotlib.patches", "matplotlib.animation"):
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)
    if "safe_control" not in sys.modules:
        pkg = types.ModuleType("safe_control")
        pkg.__path__ = [REFERENCE_ROOT]
        sys.modules["safe_control"] = pkg
    return sys.modules["safe_control"]
