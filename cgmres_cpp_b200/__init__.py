"""B200-native batched C/GMRES nonlinear-MPC controller (drop-in for blockahead/CGMRES_cpp's hot path).

The product is libcgmres_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/cgmres_b200.h) plus the header-only C++ wrapper include/cgmres.hpp; this
package is the thin Python mirror used by the tests and the benchmark.
"""
from .batched import (ARM, EXIT_BREAKDOWN, EXIT_CONVERGED, EXIT_FULL, EXIT_RHO0, MODE_EXACT, MODE_FAST, MODE_ONCHIP_EXACT, MODE_PIPELINED_EXACT, MSD,
                      SEMIACTIVE, BatchedCgmres, ModelDims, device_count, launch_count, measure_fp64_latency, measure_fp64_peak, model_dims, model_name,
                      model_params, plant_step_host)
from ._lib import LIB_PATH, CgmresB200Error

__all__ = [
    "ARM", "MSD", "SEMIACTIVE", "MODE_EXACT", "MODE_FAST", "MODE_ONCHIP_EXACT", "MODE_PIPELINED_EXACT", "EXIT_FULL", "EXIT_CONVERGED", "EXIT_RHO0",
    "EXIT_BREAKDOWN", "BatchedCgmres", "ModelDims", "model_dims", "model_params", "model_name", "launch_count", "measure_fp64_latency", "measure_fp64_peak", "plant_step_host",
    "device_count", "LIB_PATH", "CgmresB200Error",
]
