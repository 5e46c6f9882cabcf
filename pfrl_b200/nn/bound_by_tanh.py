"""Module path of the reference (pfrl/nn/bound_by_tanh.py)."""
from pfrl_b200.nn.containers import BoundByTanh  # NOQA
