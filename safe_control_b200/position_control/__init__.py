"""Drop-in classes with the reference's constructor / solve_control_problem / .status surface
(position_control/{cbf_qp,mpc_cbf,optimal_decay_cbf_qp,optimal_decay_mpc_cbf,backup_cbf_qp}.py), backed by libscb.so."""
from .cbf_qp import CBFQP  # noqa: F401
from .optimal_decay_cbf_qp import OptimalDecayCBFQP, NotCompatibleError  # noqa: F401
from .mpc_cbf import MPCCBF  # noqa: F401
from .optimal_decay_mpc_cbf import OptimalDecayMPCCBF  # noqa: F401
from .backup_cbf_qp import BackupCBF  # noqa: F401
