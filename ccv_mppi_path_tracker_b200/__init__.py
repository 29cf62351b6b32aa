"""B200-native MPPI controller core: drop-in for the optimisation step of the ccv_mppi_path_tracker nodes.

The product is libmppi_b200.so (C ABI in include/mppi_b200.h, CUDA kernels in csrc/); this package is the thin
Python host mirror used by tests and bench.  Importing it does not need a GPU; creating a controller does.
"""
from . import _capi, params, paths  # noqa: F401
from .controllers import (CONTROLLERS, DiffDriveMPPI, FullBodyMPPI, SteeringDiffDriveMPPI, calc_ref_path,  # noqa: F401
                          comm_unique_id, merge_partials, philox4x32_10)
