"""The action a stochastic policy takes when asked to act deterministically
(reference: pfrl/utils/mode_of_distribution.py)."""
import torch
from torch import distributions as D


def _through_transforms(distrib):
    x = mode_of_distribution(distrib.base_dist)
    for transform in distrib.transforms:
        x = transform(x)
    return x


# most specific wrapper types first
_RULES = (
    (D.Independent, lambda d: mode_of_distribution(d.base_dist)),
    (D.TransformedDistribution, _through_transforms),
    (D.Categorical, lambda d: d.probs.argmax(dim=-1)),
    ((D.Normal, D.MultivariateNormal), lambda d: d.mean),
)


def mode_of_distribution(distrib):
    """Most probable value of ``distrib``: argmax for Categorical, the mean for
    (multivariate) normals, pushed through the transforms of a
    TransformedDistribution (e.g. tanh-squashed Gaussians), looking inside
    ``Independent``."""
    assert isinstance(distrib, torch.distributions.Distribution)
    for kinds, rule in _RULES:
        if isinstance(distrib, kinds):
            return rule(distrib)
    raise RuntimeError("{} is not supported".format(distrib))
