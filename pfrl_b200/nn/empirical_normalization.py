"""Observation normaliser with running moments (PPO's ``obs_normalizer``).

Buffer names (``_mean``, ``_var``, ``count``) and the arithmetic follow the
reference (pfrl/nn/empirical_normalization.py:6-109) so that its checkpoints
load and seeded PPO runs agree; pinned by tests/golden/empirical_normalization.npz.
"""
import numpy as np
import torch
from torch import nn


def _unit_buffer(value, shape, dtype, axis):
    return torch.tensor(np.expand_dims(np.full(shape, value, dtype=dtype), axis))


class EmpiricalNormalization(nn.Module):
    """``forward(x, update=True)``: optionally fold the batch ``x`` into the
    running mean / variance, then return ``clip((x - mean) / sqrt(var + eps))``.

    Moments are merged with the pairwise (Chan et al.) update, one batch at a
    time, and frozen once ``until`` samples have been seen.
    """

    def __init__(self, shape, batch_axis=0, eps=1e-2, dtype=np.float32, until=None,
                 clip_threshold=None):
        super().__init__()
        self.batch_axis = batch_axis
        self.eps = eps
        self.until = until
        self.clip_threshold = clip_threshold
        np_dtype = np.dtype(dtype)
        self.register_buffer("_mean", _unit_buffer(0, shape, np_dtype, batch_axis))
        self.register_buffer("_var", _unit_buffer(1, shape, np_dtype, batch_axis))
        self.register_buffer("count", torch.tensor(0))

    # read-only views without the batch axis
    @property
    def mean(self):
        return self._mean.squeeze(self.batch_axis).clone()

    @property
    def std(self):
        return self._var.squeeze(self.batch_axis).sqrt().clone()

    def _frozen(self):
        return self.until is not None and self.count >= self.until

    def experience(self, x):
        """Merge the moments of batch ``x`` into the running ones."""
        n = x.shape[self.batch_axis]
        if n == 0 or self._frozen():
            return
        self.count += n
        weight = n / self.count.float()          # share of the new batch in the total
        assert 0 < weight <= 1
        batch_var, batch_mean = torch.var_mean(x, dim=self.batch_axis, keepdim=True,
                                               unbiased=False)
        shift = batch_mean - self._mean
        self._mean += weight * shift
        self._var += weight * (batch_var - self._var + shift * (batch_mean - self._mean))

    def forward(self, x, update=True):
        if update:
            self.experience(x)
        out = (x - self._mean) * (self._var + self.eps) ** -0.5
        if self.clip_threshold is None:
            return out
        return out.clamp(-self.clip_threshold, self.clip_threshold)

    def inverse(self, y):
        return y * (self._var + self.eps).sqrt() + self._mean
