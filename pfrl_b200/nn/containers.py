"""Small structural modules (reference: pfrl/nn/lmbda.py, branched.py,
concat_obs_and_action.py)."""
import torch
from torch import nn


class Lambda(nn.Module):
    """Turn any callable into a Module, e.g. a distribution constructor at the
    end of an ``nn.Sequential`` policy."""

    def __init__(self, lambd):
        super().__init__()
        self.lambd = lambd

    def forward(self, *inputs):
        return self.lambd(*inputs)


class Branched(nn.Module):
    """Apply several child modules to the same input and return the tuple of
    their outputs, e.g. (policy head, value head) for PPO."""

    def __init__(self, *modules):
        super().__init__()
        self.child_modules = nn.ModuleList(modules)

    def forward(self, *args, **kwargs):
        return tuple(child(*args, **kwargs) for child in self.child_modules)


class ConcatObsAndAction(nn.Module):
    """(obs, action) -> one tensor, concatenated on the last axis (Q(s, a) MLPs);
    the lower-rank operand gets trailing singleton axes first."""

    def forward(self, obs_and_action):
        obs, action = obs_and_action
        rank = max(obs.ndim, action.ndim)
        obs = obs.reshape(obs.shape + (1,) * (rank - obs.ndim))
        action = action.reshape(action.shape + (1,) * (rank - action.ndim))
        return torch.cat((obs, action), dim=-1)


class BoundByTanh(nn.Module):
    """Squash unbounded outputs into the action box: ``tanh(x) * (high - low) / 2
    + (high + low) / 2`` (reference: pfrl/nn/bound_by_tanh.py, functions/
    bound_by_tanh.py).  The bounds are converted once per device / dtype instead
    of on every call; the module has no state_dict entries, like the reference's."""

    def __init__(self, low, high):
        super().__init__()
        assert low is not None and high is not None
        self.low, self.high = low, high
        self._cache = None

    def forward(self, x):
        key = (x.device, x.dtype)
        if self._cache is None or self._cache[0] != key:
            low = torch.as_tensor(self.low, dtype=x.dtype, device=x.device)
            high = torch.as_tensor(self.high, dtype=x.dtype, device=x.device)
            self._cache = (key, (high - low) / 2, (high + low) / 2)
        _, scale, loc = self._cache
        return torch.tanh(x) * scale + loc
