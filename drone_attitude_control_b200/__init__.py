"""bnmpc: B200-native batched NMPC for the drone-attitude-control closed loop (CUDA only)."""
from ._lib import BnmpcError, default_config, lib  # noqa: F401
from .acados_shim import BatchedAcadosOcpSolver, BatchedAcadosSimSolver  # noqa: F401
from .closed_loop import BatchedClosedLoop, CircleRef, PhiloxNoise, follow_trajectory_batched  # noqa: F401
from .fleet import SolverFleet, group_bounds  # noqa: F401
