"""Multi-input Sequential: tuple/list outputs are splatted into the next layer.

This is the inter-layer calling convention of the layer-list config (capability parity with
scaelum/builder/sequential_wrapper.py:8-20).
"""
import torch.nn as nn


class SequentialWrapper(nn.Sequential):
    def forward(self, *inputs):
        for module in self._modules.values():
            if isinstance(inputs, (tuple, list)):
                inputs = module(*inputs)
            else:
                inputs = module(inputs)
        return inputs
