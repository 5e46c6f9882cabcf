import torch


class ConcatObsAndAction(torch.nn.Module):
    """(obs, action) -> cat on the last axis, for Q(s, a) MLPs
    (pfrl/nn/concat_obs_and_action.py)."""

    def forward(self, obs_and_action):
        obs, action = obs_and_action
        if obs.ndim > action.ndim:
            action = action.reshape(action.shape + (1,) * (obs.ndim - action.ndim))
        elif action.ndim > obs.ndim:
            obs = obs.reshape(obs.shape + (1,) * (action.ndim - obs.ndim))
        return torch.cat((obs, action), dim=-1)
