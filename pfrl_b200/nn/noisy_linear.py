"""Factorised NoisyNet linear layer (pfrl/nn/noisy_linear.py:25-70)."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from pfrl_b200.ops.linear import linear as _linear


@torch.no_grad()
def _lecun_uniform_(w, scale):
    fan_in = nn.init._calculate_correct_fan(w, "fan_in")
    bound = scale * math.sqrt(3.0 / fan_in)
    return w.uniform_(-bound, bound)


@torch.no_grad()
def _variance_scaling_constant_(t, scale):
    fan = t.shape[0] if t.ndim == 1 else nn.init._calculate_correct_fan(t, "fan_in")
    return t.fill_(scale / math.sqrt(fan))


class FactorizedNoisyLinear(nn.Module):
    """y = (mu_W + sigma_W * (f(eps_out) f(eps_in)^T)) x + mu_b + sigma_b f(eps_out),
    f(e) = sign(e) sqrt|e|, fresh N(0,1) noise on EVERY forward call.
    Sub-modules are called ``mu`` and ``sigma`` like the reference's, so
    state_dicts are interchangeable."""

    def __init__(self, mu_link, sigma_scale=0.4):
        super().__init__()
        self._kernel = None
        self.out_size = mu_link.out_features
        self.hasbias = mu_link.bias is not None
        in_size = mu_link.weight.shape[1]
        device = mu_link.weight.device
        self.mu = nn.Linear(in_size, self.out_size, bias=self.hasbias)
        _lecun_uniform_(self.mu.weight, scale=1 / math.sqrt(3))
        self.sigma = nn.Linear(in_size, self.out_size, bias=self.hasbias)
        _variance_scaling_constant_(self.sigma.weight, sigma_scale)
        if self.hasbias:
            _variance_scaling_constant_(self.sigma.bias, sigma_scale)
        self.mu.to(device)
        self.sigma.to(device)

    def _eps(self, n, dtype, device):
        r = torch.normal(mean=0.0, std=1.0, size=(n,), dtype=dtype, device=device)
        return torch.abs(torch.sqrt(torch.abs(r))) * torch.sign(r)

    def forward(self, x):
        sw = self.sigma.weight
        n_out, n_in = sw.shape
        eps = self._eps(n_in + n_out, sw.dtype, sw.device)
        eps_in, eps_out = eps[:n_in], eps[n_in:]
        weight = torch.addcmul(self.mu.weight, sw, torch.outer(eps_out, eps_in))
        if not self.hasbias:
            return _linear(x, weight)
        # CUDA: tcgen05 product with fp32 results (csrc/gemm.cu); elsewhere F.linear
        return _linear(x, weight, torch.addcmul(self.mu.bias, self.sigma.bias, eps_out))
