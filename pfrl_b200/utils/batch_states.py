"""batch_states: list of observations -> model input on the device.

Mirrors pfrl/utils/batch_states.py:18-36 (phi per observation, collate, move
to device).  Observations that already live on the device (tensors produced
by the GPU vector envs) take a fused path with no host round trip.
"""
import torch
from torch.utils.data._utils.collate import default_collate


def _move(batched, device):
    if isinstance(batched, torch.Tensor):
        return batched.to(device)
    if isinstance(batched, list):
        return [x.to(device) for x in batched]
    if isinstance(batched, tuple):
        return tuple(x.to(device) for x in batched)
    raise TypeError("Unsupported type of data")


def batch_states(states, device, phi):
    # device-resident batch (one tensor holding every env's observation)
    whole = getattr(states, "batch", None)
    if isinstance(whole, torch.Tensor) and whole.is_cuda:
        states = whole
    if isinstance(states, torch.Tensor) and states.is_cuda:
        mode = getattr(phi, "b2rl_obs_mode", None)
        if mode == 1:
            return states.to(torch.float32) * phi.b2rl_obs_scale
        if mode == 0:
            return states
        return phi(states)
    # host batch in one (possibly page-locked) slab: envs.MultiprocessVectorEnv
    slab = getattr(states, "host_batch", None)
    mode = getattr(phi, "b2rl_obs_mode", None)
    if isinstance(slab, torch.Tensor) and mode in (0, 1) and torch.device(device).type == "cuda":
        whole = slab.to(device)          # one copy for all environments (the slab is reused)
        return whole.to(torch.float32) * phi.b2rl_obs_scale if mode == 1 else whole
    features = [phi(s) for s in states]
    collated = default_collate(features)
    if isinstance(features[0], tuple):
        collated = tuple(collated)
    return _move(collated, device)
