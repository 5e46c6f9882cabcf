"""Module path of the reference (pfrl/nn/branched.py)."""
from pfrl_b200.nn.containers import Branched  # NOQA
