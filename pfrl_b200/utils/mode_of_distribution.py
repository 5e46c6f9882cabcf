"""The action a policy takes when it acts deterministically."""
import torch


def mode_of_distribution(distrib):
    """Most probable action (pfrl/utils/mode_of_distribution.py)."""
    if isinstance(distrib, torch.distributions.Independent):
        return mode_of_distribution(distrib.base_dist)
    if isinstance(distrib, torch.distributions.Categorical):
        return distrib.probs.argmax(dim=-1)
    if isinstance(distrib, (torch.distributions.Normal, torch.distributions.MultivariateNormal)):
        return distrib.mean
    if isinstance(distrib, torch.distributions.TransformedDistribution):
        x = mode_of_distribution(distrib.base_dist)
        for transform in distrib.transforms:
            x = transform(x)
        return x
    raise RuntimeError("{} is not supported".format(distrib))
