"""Module path of the reference (pfrl/explorers/additive_ou.py)."""
from pfrl_b200.explorers.stochastic import AdditiveOU  # NOQA
