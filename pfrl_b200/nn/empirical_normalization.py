import numpy as np
import torch
from torch import nn


class EmpiricalNormalization(nn.Module):
    """Normalise by running mean / variance (pfrl/nn/empirical_normalization.py:
    6-109): ``experience(x)`` merges a batch's moments into the running ones
    (Chan's parallel update) until ``until`` samples were seen;
    ``forward(x, update)`` returns clip((x - mean) / sqrt(var + eps))."""

    def __init__(self, shape, batch_axis=0, eps=1e-2, dtype=np.float32, until=None,
                 clip_threshold=None):
        super().__init__()
        dtype = np.dtype(dtype)
        self.batch_axis = batch_axis
        self.eps = eps
        self.until = until
        self.clip_threshold = clip_threshold
        self.register_buffer(
            "_mean", torch.tensor(np.expand_dims(np.zeros(shape, dtype=dtype), batch_axis)))
        self.register_buffer(
            "_var", torch.tensor(np.expand_dims(np.ones(shape, dtype=dtype), batch_axis)))
        self.register_buffer("count", torch.tensor(0))
        self._cached_std_inverse = None

    @property
    def mean(self):
        return torch.squeeze(self._mean, self.batch_axis).clone()

    @property
    def std(self):
        return torch.sqrt(torch.squeeze(self._var, self.batch_axis)).clone()

    @property
    def _std_inverse(self):
        if self._cached_std_inverse is None:
            self._cached_std_inverse = (self._var + self.eps) ** -0.5
        return self._cached_std_inverse

    def experience(self, x):
        if self.until is not None and self.count >= self.until:
            return
        n = x.shape[self.batch_axis]
        if n == 0:
            return
        self.count += n
        rate = n / self.count.float()
        assert rate > 0 and rate <= 1
        var_x, mean_x = torch.var_mean(x, dim=self.batch_axis, keepdim=True, unbiased=False)
        delta = mean_x - self._mean
        self._mean += rate * delta
        self._var += rate * (var_x - self._var + delta * (mean_x - self._mean))
        self._cached_std_inverse = None

    def forward(self, x, update=True):
        if update:
            self.experience(x)
        y = (x - self._mean) * self._std_inverse
        if self.clip_threshold is not None:
            y = torch.clamp(y, -self.clip_threshold, self.clip_threshold)
        return y

    def inverse(self, y):
        return y * torch.sqrt(self._var + self.eps) + self._mean
