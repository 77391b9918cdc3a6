"""ResNet layer library (CIFAR-style 3x3 stem), registered to show that the layer-list mechanism
is model agnostic.  Capability parity with scaelum/model/layers.py:6-261; the reference's
self-referencing-through-``None`` bugs do not exist here because
``Registry.register_module`` returns the class.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..registry import LAYER


@LAYER.register_module
class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, in_channels, out_channels, stride=1):
        super().__init__()
        self.residual_function = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=stride, padding=1, bias=False),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels * BasicBlock.expansion, kernel_size=3, padding=1,
                      bias=False),
            nn.BatchNorm2d(out_channels * BasicBlock.expansion),
        )
        self.shortcut = nn.Sequential()
        if stride != 1 or in_channels != BasicBlock.expansion * out_channels:
            self.shortcut = nn.Sequential(
                nn.Conv2d(in_channels, out_channels * BasicBlock.expansion, kernel_size=1,
                          stride=stride, bias=False),
                nn.BatchNorm2d(out_channels * BasicBlock.expansion),
            )

    def forward(self, x):
        return torch.relu(self.residual_function(x) + self.shortcut(x))


@LAYER.register_module
class BottleNeck(nn.Module):
    expansion = 4

    def __init__(self, in_channels, out_channels, stride=1):
        super().__init__()
        e = BottleNeck.expansion
        self.residual_function = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=1, bias=False),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, stride=stride, kernel_size=3, padding=1, bias=False),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels * e, kernel_size=1, bias=False),
            nn.BatchNorm2d(out_channels * e),
        )
        self.shortcut = nn.Sequential()
        if stride != 1 or in_channels != out_channels * e:
            self.shortcut = nn.Sequential(
                nn.Conv2d(in_channels, out_channels * e, stride=stride, kernel_size=1, bias=False),
                nn.BatchNorm2d(out_channels * e),
            )

    def forward(self, x):
        return torch.relu(self.residual_function(x) + self.shortcut(x))


@LAYER.register_module
class ResHead(nn.Module):
    """3x3 stem: conv-bn-relu to 64 channels."""

    def __init__(self, in_channels=3, out_channels=64):
        super().__init__()
        self.conv1 = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1, bias=False),
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
        )

    def forward(self, x):
        return self.conv1(x)


@LAYER.register_module
class ResLayer(nn.Module):
    """`num_blocks` residual blocks of type `block_name`; the first one may down-sample."""

    def __init__(self, block_name, in_channels, out_channels, num_blocks, stride):
        super().__init__()
        block = LAYER.get_module(block_name)
        strides = [stride] + [1] * (num_blocks - 1)
        layers = []
        for s in strides:
            layers.append(block(in_channels, out_channels, s))
            in_channels = out_channels * block.expansion
        self.layers = nn.Sequential(*layers)
        self.out_channels = in_channels

    def forward(self, x):
        return self.layers(x)


@LAYER.register_module
class ResTail(nn.Module):
    def __init__(self, in_channels=512, num_classes=100):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(in_channels, num_classes)

    def forward(self, x):
        x = self.avg_pool(x)
        return self.fc(x.view(x.size(0), -1))


@LAYER.register_module
class ResNet(nn.Module):
    def __init__(self, block_name, num_block, num_classes=100):
        super().__init__()
        block = LAYER.get_module(block_name)
        self.head = ResHead(3, 64)
        chans = 64
        stages = []
        for out_c, n, s in zip((64, 128, 256, 512), num_block, (1, 2, 2, 2)):
            layer = ResLayer(block_name, chans, out_c, n, s)
            chans = layer.out_channels
            stages.append(layer)
        self.stages = nn.Sequential(*stages)
        self.tail = ResTail(512 * block.expansion, num_classes)

    def forward(self, x):
        return self.tail(self.stages(self.head(x)))


def resnet_layer_configs(block_name: str, num_block, num_classes: int = 100) -> list:
    """The same network as a flat layer-config list for the allocator (one entry per stage)."""
    block = LAYER.get_module(block_name)
    cfgs = [dict(layer_type="ResHead", in_channels=3, out_channels=64)]
    chans = 64
    for out_c, n, s in zip((64, 128, 256, 512), num_block, (1, 2, 2, 2)):
        cfgs.append(dict(layer_type="ResLayer", block_name=block_name, in_channels=chans,
                         out_channels=out_c, num_blocks=n, stride=s))
        chans = out_c * block.expansion
    cfgs.append(dict(layer_type="ResTail", in_channels=chans, num_classes=num_classes))
    return cfgs


def resnet18(num_classes=100):
    return ResNet("BasicBlock", [2, 2, 2, 2], num_classes)


def resnet34(num_classes=100):
    return ResNet("BasicBlock", [3, 4, 6, 3], num_classes)


def resnet50(num_classes=100):
    return ResNet("BottleNeck", [3, 4, 6, 3], num_classes)


def resnet101(num_classes=100):
    return ResNet("BottleNeck", [3, 4, 23, 3], num_classes)


def resnet152(num_classes=100):
    return ResNet("BottleNeck", [3, 8, 36, 3], num_classes)
