"""Q-function heads for vector observations
(pfrl/q_functions/state_q_functions.py)."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from pfrl_b200.action_value import DiscreteActionValue, DistributionalDiscreteActionValue
from pfrl_b200.nn.mlp import MLP
from pfrl_b200.q_function import StateQFunction


def scale_by_tanh(x, low, high):
    scale = (high - low) / 2
    mean = (high + low) / 2
    return torch.tanh(x) * scale + mean


class SingleModelStateQFunctionWithDiscreteAction(nn.Module, StateQFunction):
    """Wrap a module whose output is [batch, n_actions] Q-values."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, x):
        return DiscreteActionValue(self.model(x))


class FCStateQFunctionWithDiscreteAction(SingleModelStateQFunctionWithDiscreteAction):
    """MLP Q-function: ndim_obs -> hidden^layers -> n_actions."""

    def __init__(self, ndim_obs, n_actions, n_hidden_channels, n_hidden_layers,
                 nonlinearity=F.relu, last_wscale=1.0):
        super().__init__(model=MLP(
            in_size=ndim_obs, out_size=n_actions,
            hidden_sizes=[n_hidden_channels] * n_hidden_layers, nonlinearity=nonlinearity,
            last_wscale=last_wscale))


class DistributionalSingleModelStateQFunctionWithDiscreteAction(nn.Module, StateQFunction):
    """Wrap a module producing [batch, n_actions, n_atoms] probabilities."""

    def __init__(self, model, z_values):
        super().__init__()
        self.model = model
        self.register_buffer("z_values", z_values)

    def forward(self, x):
        return DistributionalDiscreteActionValue(self.model(x), self.z_values)


class DistributionalFCStateQFunctionWithDiscreteAction(
        DistributionalSingleModelStateQFunctionWithDiscreteAction):
    """C51 MLP Q-function over ``n_atoms`` atoms in [v_min, v_max]."""

    def __init__(self, ndim_obs, n_actions, n_atoms, v_min, v_max, n_hidden_channels,
                 n_hidden_layers, nonlinearity=F.relu, last_wscale=1.0):
        assert n_atoms >= 2
        assert v_min < v_max
        # numpy linspace (fp64 then cast), as in the reference (state_q_functions.py:128)
        z_values = torch.from_numpy(np.linspace(v_min, v_max, num=n_atoms, dtype=np.float32))

        class _Head(nn.Module):
            def forward(self, h):
                return F.softmax(h.reshape(-1, n_actions, n_atoms), dim=2)

        model = nn.Sequential(
            MLP(in_size=ndim_obs, out_size=n_actions * n_atoms,
                hidden_sizes=[n_hidden_channels] * n_hidden_layers, nonlinearity=nonlinearity,
                last_wscale=last_wscale),
            _Head())
        super().__init__(model=model, z_values=z_values)


class DiscreteActionValueHead(nn.Module):
    def forward(self, q_values):
        return DiscreteActionValue(q_values)
