"""``MinkowskiEngine.utils`` surface: only ``kaiming_normal_`` is used (models/resnet_base.py:76)."""
import math

import torch
import torch.nn as nn


def _fans(t):
    # [K, Cin, Cout] kernels: fan_in = Cin*K, fan_out = Cout*K; 2-D kernels follow torch's Linear rule
    # (size(1), size(0)) -- recalled from MinkowskiEngine/utils/init.py, unverified offline.
    if t.dim() == 2:
        return t.size(1), t.size(0)
    return t.size(1) * t.size(0), t.size(2) * t.size(0)


def kaiming_normal_(tensor, a=0, mode='fan_in', nonlinearity='leaky_relu'):
    fan_in, fan_out = _fans(tensor)
    std = nn.init.calculate_gain(nonlinearity, a) / math.sqrt(fan_in if mode == 'fan_in' else fan_out)
    with torch.no_grad():
        return tensor.normal_(0, std)
