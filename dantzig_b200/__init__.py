"""dantzig_b200 -- B200-native (sm_100a) implementation of dantzig's simplex
hot path: the parametric self-dual pivot loop and the LU / FTRAN / BTRAN / CSC
routines under it, as CUDA kernels behind a C ABI (include/dantzig_b200.h).

The package holds only what that path needs:
  csrc/        CUDA kernels, host lowering, the C ABI, the ``dantzig.rust`` module
  _capi.py     ctypes binding of the C ABI
  model.py     array form of the model that crosses the boundary
  solver.py    templates, device-resident batches, solve entry points
  generate.py  deterministic synthetic LPs of the BASELINE.json shapes
  sharding.py  batch partitioning across ranks (one process per GPU)
There is no CPU fallback: importing works everywhere, solving needs a GPU.
"""
from .model import EQ, GE, LE, ModelArrays, ModelBuilder, dense_structure, dense_theta  # noqa: F401
from .solver import (  # noqa: F401
    BREAKDOWN, INFEASIBLE, OPTIMAL, PIVOT_CAP, UNBOUNDED, Batch, BatchResult, Solution, Template,
    device_count, device_info, measure_fp64_peak, solve_batch, solve_batch_multi, solve_dense_batch, solve_model,
)

__all__ = [
    "ModelArrays", "ModelBuilder", "dense_structure", "dense_theta", "LE", "GE", "EQ",
    "Template", "Batch", "BatchResult", "Solution", "solve_batch", "solve_batch_multi", "solve_dense_batch", "solve_model",
    "device_count", "device_info", "measure_fp64_peak",
    "OPTIMAL", "UNBOUNDED", "INFEASIBLE", "BREAKDOWN", "PIVOT_CAP",
]
