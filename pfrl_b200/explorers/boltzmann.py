"""Module path of the reference (pfrl/explorers/boltzmann.py)."""
from pfrl_b200.explorers.stochastic import Boltzmann  # NOQA
