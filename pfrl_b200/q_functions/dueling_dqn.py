"""Dueling Q-networks (pfrl/q_functions/dueling_dqn.py:20-129).  Attribute
names (conv_layers, a_stream, v_stream, main_stream) follow the reference so
checkpoints are interchangeable."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from pfrl_b200 import action_value
from pfrl_b200.initializers import constant_bias_initializer, init_chainer_default
from pfrl_b200.nn.fast_conv import NatureConv1
from pfrl_b200.nn.mlp import MLP
from pfrl_b200.ops.conv import TCConv2d
from pfrl_b200.ops.linear import TCLinear
from pfrl_b200.q_function import StateQFunction


def _nature_convs(n_input_channels):
    return nn.ModuleList([
        NatureConv1(n_input_channels),
        TCConv2d(32, 64, 4, stride=2),  # nn.Conv2d run as implicit GEMMs on the tensor cores
        TCConv2d(64, 64, 3, stride=1),
    ])


class DuelingDQN(nn.Module, StateQFunction):
    """Q = V + (A - mean_a A) on the Nature trunk (arXiv:1511.06581)."""

    def __init__(self, n_actions, n_input_channels=4, activation=F.relu, bias=0.1):
        self.n_actions = n_actions
        self.n_input_channels = n_input_channels
        self.activation = activation
        super().__init__()
        self.conv_layers = _nature_convs(n_input_channels)
        self.a_stream = MLP(3136, n_actions, [512], linear_cls=TCLinear)
        self.v_stream = MLP(3136, 1, [512], linear_cls=TCLinear)
        self.conv_layers.apply(init_chainer_default)
        self.conv_layers.apply(constant_bias_initializer(bias=bias))

    def forward(self, x):
        h = x
        for conv in self.conv_layers:
            h = self.activation(conv(h))
        h = h.reshape(x.shape[0], -1)
        adv = self.a_stream(h)
        adv = adv - adv.sum(dim=1, keepdim=True) / self.n_actions
        return action_value.DiscreteActionValue(adv + self.v_stream(h))


class DistributionalDuelingDQN(nn.Module, StateQFunction):
    """Rainbow's network: Nature trunk -> Linear(3136, 1024) split into two
    512-wide halves -> advantage logits [nA, n_atoms] and value logits
    [n_atoms]; softmax over atoms of (V + A - mean_a A)."""

    def __init__(self, n_actions, n_atoms, v_min, v_max, n_input_channels=4,
                 activation=torch.relu, bias=0.1):
        assert n_atoms >= 2
        assert v_min < v_max
        self.n_actions = n_actions
        self.n_input_channels = n_input_channels
        self.activation = activation
        self.n_atoms = n_atoms
        super().__init__()
        self.z_values = torch.linspace(v_min, v_max, n_atoms, dtype=torch.float32)
        self.conv_layers = _nature_convs(n_input_channels)
        # TCLinear IS an nn.Linear; on CUDA its products run on the tensor cores (ops/linear.py)
        self.main_stream = TCLinear(3136, 1024)
        self.a_stream = TCLinear(512, n_actions * n_atoms)
        self.v_stream = TCLinear(512, n_atoms)
        self.apply(init_chainer_default)
        self.conv_layers.apply(constant_bias_initializer(bias=bias))

    def forward(self, x):
        h = x
        for conv in self.conv_layers:
            h = self.activation(conv(h))
        n = x.shape[0]
        h = self.activation(self.main_stream(h.reshape(n, -1)))
        h_a, h_v = torch.chunk(h, 2, dim=1)
        adv = self.a_stream(h_a).reshape(n, self.n_actions, self.n_atoms)
        adv = adv - adv.sum(dim=1, keepdim=True) / self.n_actions
        val = self.v_stream(h_v).reshape(n, 1, self.n_atoms)
        q = F.softmax(adv + val, dim=2)
        if self.z_values.device != x.device:
            self.z_values = self.z_values.to(x.device)
        return action_value.DistributionalDiscreteActionValue(q, self.z_values)
