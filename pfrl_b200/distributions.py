"""Degenerate distribution for deterministic policies (DDPG / TD3 heads).

``DeterministicHead`` wraps the policy output in ``Independent(Delta(loc), 1)``
so that agents can treat stochastic and deterministic policies alike:
``sample()`` / ``rsample()`` return the location, ``mode_of_distribution`` its
mean; a density does not exist, so ``log_prob`` and ``entropy`` refuse.
(Reference counterpart: pfrl/distributions/delta.py.)
"""
import numbers

import torch
from torch.distributions import Distribution, constraints


def _no_density(*_args, **_kwargs):
    raise RuntimeError("Not defined")


class Delta(Distribution):
    has_rsample = True
    support = constraints.real
    arg_constraints = {"loc": constraints.real}

    def __init__(self, loc, validate_args=None):
        scalar = isinstance(loc, numbers.Number)
        self.loc = loc
        super().__init__(torch.Size() if scalar else loc.size(), validate_args=validate_args)

    # moments: all mass at loc
    mean = property(lambda self: self.loc)
    stddev = property(lambda self: torch.zeros_like(self.loc))
    variance = stddev

    def rsample(self, sample_shape=torch.Size()):
        """The location itself (differentiable), broadcast to the sample shape."""
        return self.loc.expand(self._extended_shape(sample_shape))

    def sample(self, sample_shape=torch.Size()):
        return self.rsample(sample_shape).detach()

    def expand(self, batch_shape, _instance=None):
        return Delta(self.loc.expand(torch.Size(batch_shape)), validate_args=False)

    log_prob = _no_density
    entropy = _no_density
