"""Shadow of the reference's position_control/mpc_cbf.py (see safe_control_b200/shadow/__init__.py)."""
from safe_control_b200.position_control.mpc_cbf import MPCCBF  # noqa: F401
