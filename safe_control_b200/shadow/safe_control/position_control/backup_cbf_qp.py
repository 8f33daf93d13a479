"""Shadow of the reference's position_control/backup_cbf_qp.py (see safe_control_b200/shadow/__init__.py)."""
from safe_control_b200.position_control.backup_cbf_qp import BackupCBF  # noqa: F401
