"""Layer-indexed parameter store (capability parity with scaelum/dynamics/parameter_server.py:
14-39): the full, un-partitioned model as an ``nn.ModuleList`` on the host, keyed by GLOBAL layer
index so a checkpoint is independent of the allocation (resume under a different partition).

``lazy=True`` builds layers on demand, so rank 0 does not have to materialise a 2 B-parameter
model just to hold checkpoints; sharded save/load lives in ``runner.hooks_collection``.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional

import torch
import torch.nn as nn
from torch import Tensor

from ..builder import build_layer


class ParameterServer(nn.Module):
    def __init__(self, model_config: list, lazy: bool = False) -> None:
        super().__init__()
        self._model_config = model_config
        self._lazy = lazy
        self.module_list = nn.ModuleList()
        self._state: Dict[int, "OrderedDict[str, Tensor]"] = {}
        if not lazy:
            self._build_model()

    def _build_layer(self, idx: int) -> nn.Module:
        cfg = dict(self._model_config[idx])
        return build_layer(cfg.pop("layer_type"), **cfg)

    def _build_model(self) -> None:
        for idx in range(len(self._model_config)):
            self.module_list.append(self._build_layer(idx))

    def __len__(self) -> int:
        return len(self._model_config)

    def load_weights_from_file(self, checkpoint: str) -> None:
        sd = torch.load(checkpoint, map_location="cpu")
        if self._lazy:
            self._state = {}
            for key, val in sd.items():
                idx, name = key.split(".", 1)
                self._state.setdefault(int(idx), OrderedDict())[name] = val
        else:
            self.module_list.load_state_dict(sd)

    def save_weights_to_file(self, checkpoint: str) -> None:
        torch.save(self.full_state_dict(), checkpoint)

    def full_state_dict(self) -> "OrderedDict[str, Tensor]":
        if not self._lazy:
            return self.module_list.state_dict()
        out: "OrderedDict[str, Tensor]" = OrderedDict()
        for idx in sorted(self._state):
            for name, val in self._state[idx].items():
                out["{}.{}".format(idx, name)] = val
        return out

    def update_weights(self, state_dict: "OrderedDict[str, Tensor]", idx: int) -> None:
        if self._lazy:
            self._state[idx] = OrderedDict((k, v.detach().cpu()) for k, v in state_dict.items())
        else:
            self.module_list[idx].load_state_dict(state_dict)

    def get_state_dict(self, idx: int) -> Optional[Dict[str, Tensor]]:
        if not 0 <= idx < len(self._model_config):
            raise IndexError("layer index {} out of range (model has {} layers)".format(
                idx, len(self._model_config)))
        if self._lazy:
            # a layer without parameters / buffers never appears in a checkpoint file: its state
            # dict is empty, not missing (scaelum/dynamics/parameter_server.py:38-39 behaviour)
            return self._state.get(idx, OrderedDict())
        return self.module_list[idx].state_dict()
