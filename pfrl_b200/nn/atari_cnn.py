"""Atari convolutional trunks (pfrl/nn/atari_cnn.py:17-82).  The first convolution and the
fully connected head have their own kernels (nn/fast_conv.py: exact-fp32 direct convolution;
ops/linear.py: tcgen05 3xTF32 product); the other convolutions run as implicit GEMMs on the
tensor cores (ops/conv.py), forward and backward."""
import torch.nn as nn
import torch.nn.functional as F

from pfrl_b200.nn.fast_conv import NatureConv1
from pfrl_b200.ops.conv import TCConv2d
from pfrl_b200.ops.linear import TCLinear
from pfrl_b200.initializers import constant_bias_initializer, init_chainer_default


class _AtariCNN(nn.Module):
    def __init__(self, convs, flat, n_output_channels, activation, bias):
        super().__init__()
        self.activation = activation
        self.n_output_channels = n_output_channels
        self.layers = nn.ModuleList(convs)
        self.output = TCLinear(flat, n_output_channels)
        self.apply(init_chainer_default)
        self.apply(constant_bias_initializer(bias=bias))

    def forward(self, state):
        h = state
        for conv in self.layers:
            h = self.activation(conv(h))
        return self.activation(self.output(h.reshape(h.shape[0], -1)))


class LargeAtariCNN(_AtariCNN):
    """Nature-DQN trunk: 8x8/4 -> 32, 4x4/2 -> 64, 3x3/1 -> 64, fc 3136 -> 512."""

    def __init__(self, n_input_channels=4, n_output_channels=512, activation=F.relu, bias=0.1):
        self.n_input_channels = n_input_channels
        super().__init__(
            [NatureConv1(n_input_channels), TCConv2d(32, 64, 4, stride=2),
             TCConv2d(64, 64, 3, stride=1)], 3136, n_output_channels, activation, bias)


class SmallAtariCNN(_AtariCNN):
    """NIPS-2013 DQN trunk: 8x8/4 -> 16, 4x4/2 -> 32, fc 2592 -> 256."""

    def __init__(self, n_input_channels=4, n_output_channels=256, activation=F.relu, bias=0.1):
        self.n_input_channels = n_input_channels
        super().__init__(
            [TCConv2d(n_input_channels, 16, 8, stride=4), TCConv2d(16, 32, 4, stride=2)],
            2592, n_output_channels, activation, bias)
