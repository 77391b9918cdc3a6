"""Small helpers (capability parity with scaelum/utils.py:14-87).  The RPC helpers
``call_method / remote_method / parameter_rrefs`` keep their names; the RRef they operated on
becomes :class:`OwnerRef` (a value that lives on one rank of the SPMD job) and ``remote_method`` is a
collective (owner executes, result broadcast) instead of a point-to-point RPC."""
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
        # only floating-point entries are widened to fp32; integer buffers (e.g. BatchNorm's
        # num_batches_tracked) keep their dtype so the file stays format-identical
        val = val.detach()
        out[key] = (val.float() if val.is_floating_point() else val).cpu()
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


class OwnerRef:
    """SPMD stand-in for ``torch.distributed.rpc.RRef``: a value that lives on ``owner_rank``.
    Every rank can hold the handle; only the owner holds the value."""

    def __init__(self, value=None, owner_rank: int = 0):
        self._value = value
        self._owner = owner_rank

    def owner(self) -> int:
        return self._owner

    def is_owner(self) -> bool:
        return _my_rank() == self._owner

    def local_value(self):
        if not self.is_owner():
            raise RuntimeError(f"value lives on rank {self._owner}, this is rank {_my_rank()}")
        return self._value

    def to_here(self):
        """Collective: every rank receives a (pickled) copy of the owner's value."""
        return remote_method(lambda v: v, self)


def _my_rank() -> int:
    import torch.distributed as dist

    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def call_method(method, rref, *args, **kwargs):
    """``method(rref.local_value(), *args)`` (scaelum/utils.py:27-28); plain objects are accepted."""
    target = rref.local_value() if isinstance(rref, OwnerRef) else rref
    return method(target, *args, **kwargs)


def remote_method(method, rref, *args, **kwargs):
    """Run ``method`` on the owner of ``rref`` and return its result on EVERY rank
    (scaelum/utils.py:31-33 did a blocking RPC to the owner).  SPMD form: a collective - all ranks
    call it, the owner executes, the result is broadcast with ``broadcast_object_list``."""
    import torch.distributed as dist

    if not isinstance(rref, OwnerRef):
        return method(rref, *args, **kwargs)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    result = call_method(method, rref, *args, **kwargs) if rref.is_owner() else None
    if multi:
        box = [result]
        dist.broadcast_object_list(box, src=rref.owner())
        result = box[0]
    return result


def parameter_rrefs(module) -> List[OwnerRef]:
    """One owner handle per parameter of a locally built stage (scaelum/utils.py:66-70)."""
    me = _my_rank()
    return [OwnerRef(p, me) for p in module.parameters()]
