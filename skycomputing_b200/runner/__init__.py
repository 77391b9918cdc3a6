from .hooks import Hook
from .hooks_collection import CheckpointHook, DistributedTimerHelperHook, StopHook
from .runner import Runner, build_loss

__all__ = ["Hook", "CheckpointHook", "DistributedTimerHelperHook", "StopHook", "Runner", "build_loss"]
