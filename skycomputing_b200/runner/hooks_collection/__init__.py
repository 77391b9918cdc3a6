from .checkpoint_hook import CheckpointHook
from .distributed_timer_helper_hook import DistributedTimerHelperHook
from .reallocate_hook import ReallocateHook
from .stop_hook import StopHook

__all__ = ["CheckpointHook", "DistributedTimerHelperHook", "StopHook", "ReallocateHook"]
