"""Layer-indexed checkpoint save / resume.

Capability parity with scaelum/runner/hooks_collection/checkpoint_hook.py:13-74 (same constructor:
``load_checkpoint_from, save_path, save_interval``; ``before_run`` loads, ``after_epoch`` saves
every ``save_interval`` epochs into ``<save_path>/epoch_{n}.pth``) with the reference's broken
paths fixed and a sharded, collective implementation:

* checkpoints are keyed by GLOBAL layer index (``"<idx>.<param name>"``, the ModuleList format of
  the reference's ParameterServer), so they are independent of the allocation - resume under a
  different partition works;
* every rank contributes the state_dicts of ITS layers; rank 0 (the "central server") merges them
  through the ParameterServer and writes ONE reference-compatible file, plus (new) optimizer /
  iteration / RNG state in ``epoch_{n}.extra.rank{r}.pth`` shards;
* on load every rank reads the file and picks the layers of its own span.
"""
import os
import os.path as osp
from collections import OrderedDict

import torch

from ...registry import HOOKS
from ..hooks import Hook


def _dist():
    import torch.distributed as dist

    return dist if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 else None


@HOOKS.register_module
class CheckpointHook(Hook):
    def __init__(self, load_checkpoint_from: str = None, save_path: str = None,
                 save_interval: int = None, save_optimizer: bool = True,
                 resume_training_state: bool = False, strict_optimizer: bool = False):
        # resume_training_state=True also restores optimizer state and the epoch / iteration
        # counters from the per-rank ``.extra`` shards (same world size required); the default
        # restores weights only, like the reference.
        self._resume_training_state = resume_training_state
        self._strict_optimizer = strict_optimizer
        self._load_checkpoint_from = load_checkpoint_from
        self._save_interval = save_interval
        self._save_path = save_path
        self._save_optimizer = save_optimizer

    # ------------------------------------------------------------------ load
    def before_run(self, runner):
        if not self._load_checkpoint_from:
            return
        sd = torch.load(self._load_checkpoint_from, map_location="cpu")
        per_layer = {}
        for key, val in sd.items():
            idx, name = key.split(".", 1)
            per_layer.setdefault(int(idx), OrderedDict())[name] = val
        for module in runner.model.model:
            if not module.is_local:
                continue
            n_layers = (max(per_layer) + 1) if per_layer else 0
            if runner.parameter_server is not None:
                n_layers = max(n_layers, len(runner.parameter_server))
            b, e = module.layer_range if module.layer_range is not None else (0, n_layers)
            # layers without parameters / buffers (ReLU, Flatten, Dropout-only ...) have no keys in
            # the file: they load an empty state dict, as in the reference's ParameterServer path
            module.load_weights([per_layer.get(i, OrderedDict()) for i in range(b, e)])
        if runner.is_rank0:
            runner.parameter_server.load_weights_from_file(self._load_checkpoint_from)
        extra = self._extra_path(self._load_checkpoint_from, runner.rank)
        if self._resume_training_state and osp.exists(extra):
            st = torch.load(extra, map_location="cpu")
            if self._save_optimizer and st.get("optimizer") is not None:
                try:
                    runner.optimizer.load_state_dict(st["optimizer"])
                except Exception as exc:  # a resume with fresh optimizer state must not be silent
                    msg = "CheckpointHook: optimizer state of {} NOT restored ({}: {})".format(
                        extra, type(exc).__name__, exc)
                    if self._strict_optimizer:
                        raise RuntimeError(msg) from exc
                    import warnings

                    warnings.warn(msg)
                    runner._log(msg)
            runner.iter = st.get("iter", runner.iter)
            runner.epoch = st.get("epoch", runner.epoch)

    # ------------------------------------------------------------------ save
    @staticmethod
    def _extra_path(ckpt_path: str, rank: int) -> str:
        root, _ = osp.splitext(ckpt_path)
        return "{}.extra.rank{}.pth".format(root, rank)

    def after_epoch(self, runner):
        if not self._save_path or not self.every_n_epochs(runner, self._save_interval):
            return
        self.save(runner, osp.join(self._save_path, "epoch_{}.pth".format(runner.epoch)))

    def save(self, runner, file_name: str) -> None:
        if runner.is_rank0:
            os.makedirs(osp.dirname(osp.abspath(file_name)), exist_ok=True)
        local = {}
        for module in runner.model.model:
            if not module.is_local:
                continue
            b, _e = module.layer_range if module.layer_range is not None else (0, 0)
            for off, layer_sd in enumerate(module.get_state_dict()):
                local[b + off] = layer_sd
        d = _dist()
        if d is not None:
            gathered = [None] * d.get_world_size()
            d.all_gather_object(gathered, local)
        else:
            gathered = [local]
        if runner.is_rank0:
            for part in gathered:
                for idx, layer_sd in part.items():
                    try:
                        runner.parameter_server.update_weights(layer_sd, idx)
                    except Exception as exc:
                        raise Exception("have {} state dicts, have {} layers, error occurs at {}".format(
                            sum(len(p) for p in gathered), len(runner.parameter_server), idx)) from exc
            runner.parameter_server.save_weights_to_file(file_name)
        if self._save_optimizer:
            if d is not None:
                d.barrier()
            torch.save(dict(optimizer=runner.optimizer.state_dict() if runner.optimizer else None,
                            iter=runner.iter, epoch=runner.epoch),
                       self._extra_path(file_name, runner.rank))
