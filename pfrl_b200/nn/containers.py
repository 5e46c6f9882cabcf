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
