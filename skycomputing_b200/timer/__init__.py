from .timer import DeviceTimer, DistributedTimer

__all__ = ["DistributedTimer", "DeviceTimer"]
