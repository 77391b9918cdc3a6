from .builder import (build_data_generator, build_dataloader_from_cfg, build_from_registry,
                      build_hook, build_layer, build_layers_from_cfg, build_module_from_cfg)
from .module_wrapper import BackwardSlowdownFunction, BackwardSlowdownModule, ModuleWrapper
from .sequential_wrapper import SequentialWrapper

__all__ = ["build_dataloader_from_cfg", "build_from_registry", "build_hook", "build_layer",
           "build_layers_from_cfg", "build_module_from_cfg", "build_data_generator",
           "ModuleWrapper", "SequentialWrapper", "BackwardSlowdownFunction", "BackwardSlowdownModule"]
