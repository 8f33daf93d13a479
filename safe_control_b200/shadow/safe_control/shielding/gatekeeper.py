"""Shadow of the reference's shielding/gatekeeper.py (see safe_control_b200/shadow/__init__.py)."""
from safe_control_b200.shield import Gatekeeper  # noqa: F401
