"""B200-native batched LinMPC / linear-MHE step behind the reference's API names.

The compute path is libbmpc.so (hand-written sm_100a CUDA behind the C ABI in include/bmpc.h);
this package is only the host-side mirror of the reference interface for that path.
"""
from . import _lib
from ._lib import (BmpcError, STATUS_INFEASIBLE, STATUS_ITERATION_LIMIT, STATUS_OPTIMAL)
from .batch import BatchLinMPC
from .host import InternalModel, KalmanFilter, LinModel, ManualEstimator, SteadyKalmanFilter, move_blocking
from .linmpc import LinMPC, sim
from .mhe import MovingHorizonEstimator
from . import workloads
from .shard import gather_moves, shard_range

__all__ = ["BatchLinMPC", "LinMPC", "MovingHorizonEstimator", "LinModel", "SteadyKalmanFilter", "KalmanFilter", "InternalModel", "ManualEstimator", "sim", "workloads", "shard_range", "gather_moves",
           "BmpcError", "move_blocking", "STATUS_OPTIMAL", "STATUS_ITERATION_LIMIT",
           "STATUS_INFEASIBLE"]
