from .hooks import Hook
from .hooks_collection import (CheckpointHook, DistributedTimerHelperHook, ReallocateHook,
                               StopHook)
from .runner import Runner, build_loss

__all__ = ["Hook", "CheckpointHook", "DistributedTimerHelperHook", "StopHook", "ReallocateHook",
           "Runner", "build_loss"]
