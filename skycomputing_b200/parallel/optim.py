"""Per-stage optimizers.

``FusedSGD`` - one multi-tensor kernel launch per step over the stage's flat fp32 parameter banks:
``p -= lr * (g/scale + wd * p)`` (optional momentum), refresh of the bf16 compute shadow and
gradient zeroing in the same pass (csrc/kernels/elementwise_sm100.cu: sgd_multi_kernel).
Parameters that do not live in a native bank (torch.nn fallback layers) are updated by a regular
``torch.optim.SGD`` next to it.

``FusedAdam`` - the same one-launch design for ``Adam`` / ``AdamW`` (adam_multi_kernel): both
moments, master, shadow and gradient zeroing in one pass, the step count lives in device memory so
the update replays correctly inside the whole-step CUDA graph.

``build_optimizer`` keeps the reference's config contract (``optim_cfg = {optim_type: <any
torch.optim class name>, **kwargs}``, experiment/launch.py:152-155): SGD / Adam / AdamW on the
native path map to the fused classes, everything else (or options the fused kernels do not cover:
nesterov, dampening, amsgrad, maximize) to the named ``torch.optim`` class over the fp32 masters
(the bf16 shadows are then refreshed lazily by version counter; such a step is NOT graph-safe and
the engines run it eagerly).  Optimizer state lives with the stage, as with the reference's
DistributedOptimizer.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn


class FusedSGD:
    def __init__(self, module: nn.Module, lr: float = 1e-3, momentum: float = 0.0,
                 weight_decay: float = 0.0, dampening: float = 0.0, nesterov: bool = False,
                 banks: Optional[list] = None):
        from ..models.bert_layers import native_param_banks
        from ..ops import native as nat

        if dampening != 0.0 or nesterov:
            raise ValueError("FusedSGD supports plain / momentum SGD only")
        self.lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self._nat = nat
        self.banks = list(banks) if banks is not None else native_param_banks(module)
        seen = set()
        uniq = []
        for b in self.banks:
            if id(b) not in seen:
                seen.add(id(b))
                uniq.append(b)
        self.banks = uniq
        for b in self.banks:
            b.ensure()
        bank_params = {id(p) for b in self.banks for p in b.params}
        rest = [p for p in module.parameters() if id(p) not in bank_params and p.requires_grad]
        self._rest_opt = (torch.optim.SGD(rest, lr=lr, momentum=momentum, weight_decay=weight_decay)
                          if rest else None)
        self._mom = ([torch.zeros_like(b.master()) for b in self.banks] if momentum else
                     [None] * len(self.banks))
        self._desc_dev: Optional[torch.Tensor] = None
        self._max_numel = 0
        import os as _os

        self._overwrite_ok = _os.environ.get("SKY_SGD_OVERWRITE", "1") != "0"
        self.param_groups = [dict(lr=lr, momentum=momentum, weight_decay=weight_decay)]

    @property
    def graph_safe(self) -> bool:
        """True when step() is pure device work with static addresses (CUDA-graph capturable)."""
        return self._rest_opt is None

    _mom2: list = []

    def _descriptors(self) -> torch.Tensor:
        if self._desc_dev is None and self.banks:
            # banks that received a wgrad GEMM in the first backward pass will receive one in every
            # step: their gradient need not be zeroed, the next step's first wgrad overwrites it
            for b in self.banks:
                b.overwrite_first = bool(getattr(b, "wgrad_target", False)) and self._overwrite_ok
            mom2 = self._mom2 or [None] * len(self.banks)
            descs = [b.sgd_descriptor(m, v) for b, m, v in zip(self.banks, self._mom, mom2)]
            raw = self._nat.ext().pack_sgd_descriptors(descs)
            host = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
            self._desc_dev = host.to(self.banks[0].master().device)
            self._max_numel = max(d[4] for d in descs)
        return self._desc_dev

    def step(self, grad_scale: float = 1.0) -> None:
        if self.banks:
            d = self._descriptors()
            self._nat.ext().sgd_multi(d_tensors=d.data_ptr(), n=len(self.banks),
                                      max_numel=self._max_numel, lr=float(self.param_groups[0]["lr"]),
                                      momentum=float(self.momentum),
                                      weight_decay=float(self.weight_decay),
                                      grad_scale=float(grad_scale), zero_grad=True,
                                      stream=torch.cuda.current_stream().cuda_stream)
            for b in self.banks:
                if b.overwrite_first:
                    b.fresh = True
        if self._rest_opt is not None:
            self._rest_opt.step()
            self._rest_opt.zero_grad(set_to_none=False)

    def zero_grad(self, set_to_none: bool = False) -> None:
        for b in self.banks:
            b.grad().zero_()
            b.fresh = False
        if self._rest_opt is not None:
            self._rest_opt.zero_grad(set_to_none=False)

    def state_dict(self) -> dict:
        return dict(momentum=[None if m is None else m.detach().cpu() for m in self._mom],
                    rest=None if self._rest_opt is None else self._rest_opt.state_dict(),
                    param_groups=self.param_groups)

    def load_state_dict(self, sd: dict) -> None:
        for m, s in zip(self._mom, sd.get("momentum", [])):
            if m is not None and s is not None:
                m.copy_(s)
        if self._rest_opt is not None and sd.get("rest") is not None:
            self._rest_opt.load_state_dict(sd["rest"])
        self.param_groups = sd.get("param_groups", self.param_groups)


class FusedAdam(FusedSGD):
    """``torch.optim.Adam`` / ``AdamW`` (amsgrad=False) as one multi-tensor launch per step."""

    def __init__(self, module: nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, decoupled: bool = False, amsgrad: bool = False,
                 maximize: bool = False, banks: Optional[list] = None, **unsupported):
        if amsgrad or maximize or unsupported:
            raise ValueError("FusedAdam covers plain Adam / AdamW only")
        super().__init__(module, lr=lr, momentum=0.0, weight_decay=weight_decay, banks=banks)
        self.betas, self.eps, self.decoupled = (float(betas[0]), float(betas[1])), eps, decoupled
        self._mom = [torch.zeros_like(b.master()) for b in self.banks]
        self._mom2 = [torch.zeros_like(b.master()) for b in self.banks]
        dev = self.banks[0].master().device if self.banks else torch.device("cpu")
        self._step = torch.zeros(1, dtype=torch.int64, device=dev)   # device-side step count
        if self._rest_opt is not None:
            rest = [p for g in self._rest_opt.param_groups for p in g["params"]]
            cls = torch.optim.AdamW if decoupled else torch.optim.Adam
            self._rest_opt = cls(rest, lr=lr, betas=self.betas, eps=eps, weight_decay=weight_decay)
        self.param_groups = [dict(lr=lr, betas=self.betas, eps=eps, weight_decay=weight_decay)]

    def step(self, grad_scale: float = 1.0) -> None:
        if self.banks:
            d = self._descriptors()
            ext = self._nat.ext()
            stream = torch.cuda.current_stream().cuda_stream
            ext.advance_counter(self._step.data_ptr(), 1, stream)
            ext.adam_multi(d_tensors=d.data_ptr(), n=len(self.banks), max_numel=self._max_numel,
                           lr=float(self.param_groups[0]["lr"]), beta1=self.betas[0],
                           beta2=self.betas[1], eps=float(self.eps),
                           weight_decay=float(self.weight_decay), decoupled=bool(self.decoupled),
                           step=self._step.data_ptr(), grad_scale=float(grad_scale),
                           zero_grad=True, stream=stream)
            for b in self.banks:
                if b.overwrite_first:
                    b.fresh = True
        if self._rest_opt is not None:
            self._rest_opt.step()
            self._rest_opt.zero_grad(set_to_none=False)

    def state_dict(self) -> dict:
        sd = super().state_dict()
        sd["second_moment"] = [v.detach().cpu() for v in self._mom2]
        sd["step"] = int(self._step.item())
        return sd

    def load_state_dict(self, sd: dict) -> None:
        super().load_state_dict(sd)
        for v, s_ in zip(self._mom2, sd.get("second_moment", [])):
            if s_ is not None:
                v.copy_(s_)
        if "step" in sd:
            self._step.fill_(int(sd["step"]))


class TorchOptimizerAdapter:
    """Any ``torch.optim`` class over the stage's parameters; same step()/zero_grad() surface."""

    # host-side bookkeeping (step counters, lazily created state, version-counter driven refresh of
    # the bf16 shadows): must not be captured into a CUDA graph
    graph_safe = False

    def __init__(self, optimizer: torch.optim.Optimizer):
        self.optimizer = optimizer
        self.param_groups = optimizer.param_groups

    def step(self, grad_scale: float = 1.0) -> None:
        if grad_scale != 1.0:
            for g in self.optimizer.param_groups:
                for p in g["params"]:
                    if p.grad is not None:
                        p.grad.mul_(grad_scale)
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=False)

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.optimizer.zero_grad(set_to_none=False)

    def state_dict(self):
        return self.optimizer.state_dict()

    def load_state_dict(self, sd):
        self.optimizer.load_state_dict(sd)


def build_optimizer(module: nn.Module, optim_cfg: dict, prefer_fused: bool = True):
    cfg = dict(optim_cfg)
    optim_type = cfg.pop("optim_type", "SGD")
    params = [p for p in module.parameters() if p.requires_grad]
    on_cuda = any(p.is_cuda for p in params)
    if optim_type in ("SGD", "Adam", "AdamW") and prefer_fused and on_cuda:
        from ..models.bert_layers import get_backend
        from ..ops import native as nat

        if get_backend() != "torch" and nat.available():
            try:
                if optim_type == "SGD":
                    return FusedSGD(module, **cfg)
                if optim_type == "AdamW":
                    cfg.setdefault("weight_decay", 1e-2)   # torch.optim.AdamW's default
                return FusedAdam(module, decoupled=optim_type == "AdamW", **cfg)
            except (ValueError, TypeError):
                cfg = dict(optim_cfg)
                cfg.pop("optim_type", None)
    if not params:
        return TorchOptimizerAdapter(torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=0.0))
    return TorchOptimizerAdapter(getattr(torch.optim, optim_type)(params, **cfg))
