"""cyrk_b200 -- B200-native batched IVP solver behind CyRK's cysolve_ivp call surface.

Only the hot path lives here: the host-side mirror of the reference interface (solver.py), the
ctypes binding of the C ABI (_lib.py) and the CUDA sources (csrc/).  There is no CPU fallback.
"""
from ._lib import CyrkB200Error, LIB_PATH, SIGNATURES, last_error, lib, status_message  # noqa: F401
from .solver import (  # noqa: F401
    BatchResult, Event, TrajectoryView, UserRHS, cysolve_ivp_batch, find_ode_method_int, EVENT_IDS,
    interpolate_from_solution_list)

__version__ = "0.1.0"
