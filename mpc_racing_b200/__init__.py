"""mpc_racing_b200 - B200-native batched NMPC solver for the mpc-racing hot path.

Package layout (only what the path needs):
  csrc/      CUDA kernels (sm_100a) + the C ABI (include/mpcr.h) -> libmpcr.so
  _capi.py   ctypes binding of the C ABI
  solver.py  BatchedMPC: solve(states[B]) over host or device pointers
  pipeline.py  SolvePipeline: a stream of batches over several handles / streams (tail and copy overlap)
  rollout.py   ClosedLoopRollouts: device-resident receding-horizon loops (CUDA graph)
  GA/          the reference's GA fitness loop, multi-track and sharded over ranks
  control/   drop-in mirror of the reference's control.MPC / ControllerParameters
  models/    State / VehicleParameters with the reference's field names
  splines/   track geometry (centerline, lanes, error table) host side + device tables
  workload.py  synthetic problem batches for tests and bench
There is no CPU fallback: everything that computes needs libmpcr.so and a CUDA device.
"""
from ._capi import MpcrError, load_library, LIB_PATH  # noqa: F401

__all__ = ["MpcrError", "load_library", "LIB_PATH", "BatchedMPC", "MPC", "SolvePipeline", "ClosedLoopRollouts"]


def __getattr__(name):
    if name == "BatchedMPC":
        from .solver import BatchedMPC
        return BatchedMPC
    if name == "MPC":
        from .control.MPC import MPC
        return MPC
    if name == "SolvePipeline":
        from .pipeline import SolvePipeline
        return SolvePipeline
    if name == "ClosedLoopRollouts":
        from .rollout import ClosedLoopRollouts
        return ClosedLoopRollouts
    raise AttributeError(name)
