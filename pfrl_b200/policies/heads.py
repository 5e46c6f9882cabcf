"""Policy heads: map network outputs to torch.distributions objects.

Same classes as the reference's pfrl/policies/ (softmax_policy.py:5-7,
gaussian_policy.py:6-96), kept in one module.
"""
import numpy as np
import torch
from torch import nn


class DeterministicHead(nn.Module):
    """action -> point-mass distribution at that action (DDPG / TD3 policies;
    reference: pfrl/policies/deterministic_policy.py)."""

    def forward(self, loc):
        from pfrl_b200.distributions import Delta

        return torch.distributions.Independent(Delta(loc=loc), 1)


class SoftmaxCategoricalHead(nn.Module):
    """Unnormalised log-probabilities -> ``Categorical`` over discrete actions."""

    def forward(self, logits):
        dist = torch.distributions.Categorical(logits=logits)
        return dist

class GaussianHeadWithStateIndependentCovariance(nn.Module):
    """mean -> Independent(Normal(mean, sqrt(f(var_param)))) with a learned,
    state-independent diagonal variance parameter of size ``action_size``."""

    def __init__(self, action_size, var_type="spherical",
                 var_func=nn.functional.softplus, var_param_init=0):
        super().__init__()
        self.var_func = var_func
        var_size = {"spherical": 1, "diagonal": action_size}[var_type]
        self.var_param = nn.Parameter(
            torch.tensor(np.broadcast_to(var_param_init, var_size), dtype=torch.float))

    def forward(self, mean):
        var = self.var_func(self.var_param)
        return torch.distributions.Independent(
            torch.distributions.Normal(loc=mean, scale=torch.sqrt(var)), 1)


class GaussianHeadWithDiagonalCovariance(nn.Module):
    """[mean | pre_var] -> Independent(Normal(mean, sqrt(f(pre_var))))."""

    def __init__(self, var_func=nn.functional.softplus):
        super().__init__()
        self.var_func = var_func

    def forward(self, mean_and_var):
        assert mean_and_var.ndim == 2
        mean, pre_var = mean_and_var.chunk(2, dim=1)
        scale = self.var_func(pre_var).sqrt()
        return torch.distributions.Independent(
            torch.distributions.Normal(loc=mean, scale=scale), 1)


class GaussianHeadWithFixedCovariance(nn.Module):
    """mean -> Independent(Normal(mean, scale)) with a constant scale."""

    def __init__(self, scale=1):
        super().__init__()
        self.scale = scale

    def forward(self, mean):
        return torch.distributions.Independent(
            torch.distributions.Normal(loc=mean, scale=self.scale), 1)
