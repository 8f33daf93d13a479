"""Shadow of the reference's position_control/optimal_decay_mpc_cbf.py (see safe_control_b200/shadow/__init__.py)."""
from safe_control_b200.position_control.optimal_decay_mpc_cbf import OptimalDecayMPCCBF, NotCompatibleError  # noqa: F401
