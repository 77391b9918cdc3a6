"""Benchmark-input generators (parity with scaelum/dataset/data_generator.py:10-34)."""
from __future__ import annotations

import abc

import torch

from ..registry import DATA_GENERATOR


class BaseGenerator(abc.ABC):
    @abc.abstractmethod
    def generate(self):
        ...


@DATA_GENERATOR.register_module
class RandomTensorGenerator(BaseGenerator):
    def __init__(self, generator_cfg):
        self.generator_cfg = dict(generator_cfg)

    def generate(self):
        return torch.rand(**self.generator_cfg)


@DATA_GENERATOR.register_module
class DataloaderGenerator(BaseGenerator):
    """First batch's INPUTS of a freshly built dataloader."""

    def __init__(self, generator_cfg):
        from ..builder import build_dataloader_from_cfg

        self.dataloader = build_dataloader_from_cfg(**generator_cfg)

    def generate(self):
        return next(iter(self.dataloader))[0]
