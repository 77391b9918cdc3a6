from .stimulator import Stimulator

__all__ = ["Stimulator"]
