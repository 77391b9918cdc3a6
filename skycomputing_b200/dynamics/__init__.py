from .allocator import Allocator, group_units
from .benchmarker import BaseBenchmarker, DeviceBenchmarker, ModelBenchmarker
from .estimator import Estimator
from .parameter_server import ParameterServer
from .worker import Worker
from .worker_manager import WorkerManager

__all__ = ["Allocator", "group_units", "BaseBenchmarker", "DeviceBenchmarker", "ModelBenchmarker",
           "Estimator", "ParameterServer", "Worker", "WorkerManager"]
