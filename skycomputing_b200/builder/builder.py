"""Registry-driven factories (capability parity with scaelum/builder/builder.py:12-49)."""
from __future__ import annotations

from torch.utils.data import DataLoader

from ..registry import DATA_GENERATOR, DATASET, HOOKS, LAYER, Registry
from .module_wrapper import ModuleWrapper
from .sequential_wrapper import SequentialWrapper


def build_from_registry(module_name: str, registry: Registry, *args, **kwargs):
    return registry.get_module(module_name)(*args, **kwargs)


def build_layer(module_name: str, *args, **kwargs):
    return build_from_registry(module_name, LAYER, *args, **kwargs)


def build_hook(module_name: str, *args, **kwargs):
    return build_from_registry(module_name, HOOKS, *args, **kwargs)


def build_data_generator(module_name: str, *args, **kwargs):
    return build_from_registry(module_name, DATA_GENERATOR, *args, **kwargs)


def build_layers_from_cfg(model_cfg: list) -> list:
    layers = []
    for layer_cfg in model_cfg:
        cfg = dict(layer_cfg)
        layer_type = cfg.pop("layer_type")
        layers.append(build_layer(layer_type, **cfg))
    return layers


def build_module_from_cfg(rank, model_cfg: list, module_wrapper_cfg: dict) -> ModuleWrapper:
    """Build the stage a worker runs: layers -> SequentialWrapper -> ModuleWrapper.

    (The argument dicts are not mutated, unlike the reference.)
    """
    module = SequentialWrapper(*build_layers_from_cfg(model_cfg))
    cfg = dict(module_wrapper_cfg or {})
    cfg["record_forward_time"] = True
    return ModuleWrapper(rank=rank, module=module, **cfg)


def build_dataloader_from_cfg(dataset_cfg, dataloader_cfg):
    dataset_cfg = dict(dataset_cfg)
    dataset_type = dataset_cfg.pop("type")
    dataset = build_from_registry(dataset_type, DATASET, **dataset_cfg)
    return DataLoader(dataset, **dict(dataloader_cfg))
