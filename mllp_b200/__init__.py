"""mllp_b200 -- B200-native primal-dual LP iteration behind mllp's Python calling convention.

Only what the hot path needs: ``csrc/`` (sm_100a CUDA kernels + the C ABI of
include/mllp_b200.h), the ctypes binding, and the host-side mirror of the reference's
``linear_program_methods`` / ``linear_program_data`` interfaces for this path.
"""
from . import _cabi  # noqa: F401
from .linear_program_methods import (SCALAR_NAMES, BatchLP, DeviceLP, device_lp, estimate_step_size,  # noqa: F401
                                     pdhg_linear_program, pdhg_linear_program_batch, solve_linear_program,
                                     solve_linear_program_batch)
from .linear_program_data import get_netlib_dataset, load_csr, load_instance  # noqa: F401

__version__ = "0.1.0"
