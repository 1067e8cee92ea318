"""ad_mpc_b200 -- B200-native batched SQP-RTI NMPC solver for the AD_MPC bicycle-model controller.

Host side: numpy + ctypes over the C ABI in include/admpc.h.  The compute path is the CUDA library
libadmpc_b200.so (ad_mpc_b200/csrc); nothing here falls back to the CPU.
"""
from . import _lib  # noqa: F401
from .solver import AcadosOcpSolverB200, BatchSolver, PinnedArray, PipelinedSolver, default_opts, kappa_pp_from_knots  # noqa: F401
from .optimizer import AD3DOptimizerB200  # noqa: F401
