"""Zero-change route into the reference: shadow its three controller modules.

`LocalTrackingController.__init__` imports its position controller lazily BY MODULE PATH
(`tracking.py:140-148`: `from safe_control.position_control.cbf_qp import CBFQP`, ... and
`dynamic_env/main.py:35-37`), so whichever file the import system finds first for
`safe_control.position_control.{cbf_qp, mpc_cbf, optimal_decay_cbf_qp}` is the controller the UNMODIFIED
`tracking.py` constructs and calls.  `install()` puts this package's own three files
(`shadow/safe_control/position_control/*.py`, each a one-line re-export of the B200-backed class) in front of the
reference's on `safe_control.position_control.__path__`:

    import safe_control_b200.shadow as shadow
    shadow.install()                                   # before the first LocalTrackingController(...)
    from safe_control.tracking import LocalTrackingController     # the reference's own file, untouched

Nothing of the reference is modified or copied; `uninstall()` restores the search path.
"""
import importlib
import os
import sys

SHADOW_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "safe_control", "position_control")
MODULES = ("cbf_qp", "mpc_cbf", "optimal_decay_cbf_qp", "optimal_decay_mpc_cbf", "backup_cbf_qp")
# the trajectory-rollout shields (examples/evade/test_evade.py:41-42 imports them by module path too)
SHIELD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "safe_control", "shielding")
SHIELD_MODULES = {"gatekeeper": "Gatekeeper", "mps": "MPS"}


def install():
    """Make `safe_control.position_control.{cbf_qp,mpc_cbf,optimal_decay_cbf_qp}` resolve to the shadow files.
    Needs the reference importable as `safe_control` (its own install, or a namespace rooted at its checkout)."""
    pc = importlib.import_module("safe_control.position_control")
    path = pc.__path__
    if SHADOW_DIR not in list(path):
        path.insert(0, SHADOW_DIR)
    for m in MODULES:                                  # drop already-imported reference modules of the same name
        sys.modules.pop(f"safe_control.position_control.{m}", None)
        if hasattr(pc, m):
            delattr(pc, m)
    return pc


def install_shielding():
    """The same for `safe_control.shielding.{gatekeeper,mps}`; the package's own `Gatekeeper` / `MPS` attributes
    (shielding/__init__.py:5-6) are re-pointed as well."""
    sp = importlib.import_module("safe_control.shielding")
    if SHIELD_DIR not in list(sp.__path__):
        sp.__path__.insert(0, SHIELD_DIR)
    for m, cls in SHIELD_MODULES.items():
        sys.modules.pop(f"safe_control.shielding.{m}", None)
        mod = importlib.import_module(f"safe_control.shielding.{m}")
        setattr(sp, m, mod); setattr(sp, cls, getattr(mod, cls))
    return sp


def uninstall_shielding():
    sp = sys.modules.get("safe_control.shielding")
    if sp is None:
        return
    if SHIELD_DIR in list(sp.__path__):
        sp.__path__.remove(SHIELD_DIR)
    for m, cls in SHIELD_MODULES.items():
        sys.modules.pop(f"safe_control.shielding.{m}", None)
        mod = importlib.import_module(f"safe_control.shielding.{m}")
        setattr(sp, m, mod); setattr(sp, cls, getattr(mod, cls))


def uninstall():
    pc = sys.modules.get("safe_control.position_control")
    if pc is None:
        return
    if SHADOW_DIR in list(pc.__path__):
        pc.__path__.remove(SHADOW_DIR)
    for m in MODULES:
        sys.modules.pop(f"safe_control.position_control.{m}", None)
        if hasattr(pc, m):
            delattr(pc, m)
