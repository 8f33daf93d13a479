"""safe_control_b200 -- B200-native batched CBF-QP / MPC-CBF safety filter.

Drop-in for the per-step position-controller solve of tkkim-robot/safe_control
(position_control/{cbf_qp,mpc_cbf,optimal_decay_cbf_qp}.py) for a batch of
independent agents.  Host code is Python over torch CUDA tensors; all arithmetic
runs in hand-written sm_100a kernels behind the C ABI in include/scb.h.
"""
from .params import resolve_params, NotCompatibleError  # noqa: F401
from .batched import (BatchedCBFQP, BatchedOptimalDecayCBFQP, BatchedMPCCBF, BatchedOptimalDecayMPCCBF,  # noqa: F401
                      HostContext)
from .tracking import BatchedTrackingController  # noqa: F401
from .backup import BatchedBackupCBF, EvadeSceneParams  # noqa: F401
from .shield import BatchedShield  # noqa: F401
from ._abi import MODEL_IDS, OPTIMAL, INFEASIBLE, MAXITER, NUMERICAL  # noqa: F401

__all__ = ["BatchedTrackingController", "BatchedCBFQP", "BatchedOptimalDecayCBFQP", "BatchedMPCCBF", "BatchedOptimalDecayMPCCBF", "BatchedBackupCBF", "EvadeSceneParams", "BatchedShield", "HostContext", "resolve_params",
           "NotCompatibleError", "MODEL_IDS", "OPTIMAL", "INFEASIBLE", "MAXITER", "NUMERICAL"]
