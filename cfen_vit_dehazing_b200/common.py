"""Primitive builders mirroring reference models/common.py (default_conv :11, MeanShift :16,
ResBlock :41, Upsampler :64).  They are parameter containers: the arithmetic runs inside the
fused CUDA kernels (csrc/), never through these modules' own forwards.
"""
import math

import torch
import torch.nn as nn

from .actnorm import ActNorm2d


def default_conv(in_channels, out_channels, kernel_size, bias=True):
    """Same-padded Conv2d (reference models/common.py:11-14)."""
    return nn.Conv2d(in_channels, out_channels, kernel_size, padding=kernel_size // 2, bias=bias)


class MeanShift(nn.Conv2d):
    """Frozen 1x1 colour shift (reference models/common.py:16-27); built, never executed by v3."""

    def __init__(self, rgb_range, rgb_mean=(0.4488, 0.4371, 0.4040), rgb_std=(1.0, 1.0, 1.0), sign=-1):
        super().__init__(3, 3, kernel_size=1)
        std = torch.tensor(rgb_std, dtype=torch.float32)
        mean = torch.tensor(rgb_mean, dtype=torch.float32)
        self.weight.data = torch.eye(3).view(3, 3, 1, 1) / std.view(3, 1, 1, 1)
        self.bias.data = sign * rgb_range * mean / std
        for p in self.parameters():
            p.requires_grad = False


class ResBlock(nn.Module):
    """x + conv(act(conv(x))) * res_scale (reference models/common.py:41-62)."""

    def __init__(self, conv, n_feats, kernel_size, bias=True, bn=False, act=None, res_scale=1):
        super().__init__()
        layers = []
        for i in range(2):
            layers.append(conv(n_feats, n_feats, kernel_size, bias=bias))
            if bn:
                layers.append(ActNorm2d(n_feats))
            if i == 0:
                layers.append(act if act is not None else nn.ReLU(True))
        self.body = nn.Sequential(*layers)
        self.res_scale = res_scale


class Upsampler(nn.Sequential):
    """The reference loops ``range(int(math.log(1, 2)))`` = 0 times (models/common.py:70), so
    this is an empty Sequential: no PixelShuffle is ever executed on the v3 path."""

    def __init__(self, conv, n_feats, bn=False, act=False, bias=True):
        stages = []
        for _ in range(int(math.log(1, 2))):  # == 0; kept to state the contract
            stages.append(conv(n_feats, 4 * n_feats, 3, bias))
            stages.append(nn.PixelShuffle(2))
        super().__init__(*stages)
