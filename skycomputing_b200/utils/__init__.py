"""Small helpers (capability parity with scaelum/utils.py:14-87, minus the RPC plumbing that the
SPMD design does not need: ``call_method/remote_method/parameter_rrefs`` have no equivalent because
every rank owns its stage directly)."""
from __future__ import annotations

import time
from collections import OrderedDict
from typing import Dict, List

import torch
import torch.nn as nn

GPU = torch.cuda.is_available()


def synchronize() -> None:
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def get_time() -> float:
    synchronize()
    return time.time()


def _stage_layers(model: nn.Module) -> List[nn.Module]:
    """The per-layer modules of a stage (ModuleWrapper -> SequentialWrapper children)."""
    inner = getattr(model, "layers", None)
    if inner is None:
        inner = list(model.modules())[1]
    return list(inner.children()) if isinstance(inner, nn.Module) else list(inner)


def load_weights(model: nn.Module, state_dict: List[Dict]) -> None:
    modules = _stage_layers(model)
    assert len(modules) == len(state_dict), (
        "Weights do not match the model, model has {} modules while state dict has {}".format(
            len(modules), len(state_dict)))
    for mod, sd in zip(modules, state_dict):
        mod.load_state_dict(sd)


def weights_to_cpu(state_dict):
    out = OrderedDict()
    for key, val in state_dict.items():
        out[key] = val.detach().float().cpu()
    return out


def get_state_dict(model: nn.Module) -> List[Dict]:
    return [weights_to_cpu(mod.state_dict()) for mod in _stage_layers(model)]


def count_params(model, to_console: bool = False):
    num_params = sum(p.numel() for p in model.parameters()) / 1e6
    num_grad_params = sum(p.numel() for p in model.parameters() if p.requires_grad) / 1e6
    if to_console:
        print("Number of parameters: {:.5g} M".format(num_params))
        print("Number of parameters requiring grad: {:.5g} M".format(num_grad_params))
    return num_params, num_grad_params


def generate_worker_name(rank) -> str:
    return "worker{}".format(rank)
