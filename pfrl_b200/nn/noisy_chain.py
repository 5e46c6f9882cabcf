import torch.nn as nn

from pfrl_b200.nn.noisy_linear import FactorizedNoisyLinear


def to_factorized_noisy(module, *args, **kwargs):
    """Replace every nn.Linear inside ``module`` (recursively, in place) by a
    FactorizedNoisyLinear built from it (pfrl/nn/noisy_chain.py:11-32)."""
    for name, child in list(module.named_children()):
        if isinstance(child, nn.Linear):
            module._modules[name] = FactorizedNoisyLinear(child, *args, **kwargs)
        else:
            to_factorized_noisy(child, *args, **kwargs)
