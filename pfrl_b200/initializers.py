"""Weight initialisers with the reference's distributions
(pfrl/initializers/lecun_normal.py, chainer_default.py)."""
import math

import torch
import torch.nn as nn


@torch.no_grad()
def init_lecun_normal(tensor, scale=1.0):
    """N(0, scale^2 / fan_in)."""
    fan_in = nn.init._calculate_correct_fan(tensor, "fan_in")
    return tensor.normal_(0, scale * math.sqrt(1.0 / fan_in))


@torch.no_grad()
def init_chainer_default(layer):
    """LeCun-normal weights and zero biases for Linear / Conv2d layers."""
    assert isinstance(layer, nn.Module)
    if isinstance(layer, (nn.Linear, nn.Conv2d)):
        init_lecun_normal(layer.weight)
        if layer.bias is not None:
            layer.bias.zero_()
    return layer


def constant_bias_initializer(bias=0.0):
    @torch.no_grad()
    def _fill(m):
        if isinstance(m, (nn.Linear, nn.Conv2d)):
            m.bias.fill_(bias)

    return _fill
