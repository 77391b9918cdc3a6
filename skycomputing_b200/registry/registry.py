"""Name -> class registries used by the config system.

Capability parity with scaelum/registry/registry.py:8-30 (four global registries, decorator
registration keyed by ``cls.__name__``, duplicate assert, ``torch.nn`` fallback so that e.g.
``"Conv2d"`` in a device-benchmark config resolves).  Unlike the reference, ``register_module``
RETURNS the class, so decorated names stay importable.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterator, Optional

import torch.nn as nn


class Registry:
    def __init__(self, name: str):
        self.name = name
        self._registry: Dict[str, type] = {}

    def register_module(self, module_class: Optional[type] = None, *, name: Optional[str] = None,
                        force: bool = False) -> Callable | type:
        def _do(cls: type) -> type:
            key = name or cls.__name__
            assert force or key not in self._registry, (
                f"{key} is already registered in registry '{self.name}'"
            )
            self._registry[key] = cls
            return cls

        if module_class is None:  # used as @REG.register_module(name=...)
            return _do
        return _do(module_class)

    def get_module(self, module_name: str, include_torch: bool = True) -> type:
        if module_name in self._registry:
            return self._registry[module_name]
        if include_torch and hasattr(nn, module_name):
            return getattr(nn, module_name)
        raise NameError(f"Module {module_name} not found in registry '{self.name}'")

    def __contains__(self, module_name: str) -> bool:
        return module_name in self._registry

    def __iter__(self) -> Iterator[str]:
        return iter(self._registry)

    def __len__(self) -> int:
        return len(self._registry)

    def names(self) -> list[str]:
        return sorted(self._registry)


LAYER = Registry("layer")
DATASET = Registry("dataset")
HOOKS = Registry("hook")
DATA_GENERATOR = Registry("data_generator")
