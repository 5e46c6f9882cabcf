"""Module path of the reference (pfrl/nn/lmbda.py)."""
from pfrl_b200.nn.containers import Lambda  # NOQA
