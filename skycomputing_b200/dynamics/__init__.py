from .allocator import Allocator, group_units
from .benchmarker import BaseBenchmarker, DeviceBenchmarker, ModelBenchmarker
from .estimator import Estimator
from .parameter_server import ParameterServer
from .planner import Plan, SchedulePlanner, StageCosts, costs_from_single_gpu_step
from .worker import Worker
from .worker_manager import WorkerManager

__all__ = ["Allocator", "group_units", "BaseBenchmarker", "DeviceBenchmarker", "ModelBenchmarker",
           "Estimator", "ParameterServer", "Worker", "WorkerManager", "SchedulePlanner", "StageCosts",
           "Plan", "costs_from_single_gpu_step"]
