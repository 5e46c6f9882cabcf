"""Module path of the reference (pfrl/explorers/additive_gaussian.py)."""
from pfrl_b200.explorers.epsilon_greedy import AdditiveGaussian  # NOQA
