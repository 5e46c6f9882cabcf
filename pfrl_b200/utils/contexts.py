"""Module path of the reference (pfrl/utils/contexts.py)."""
from pfrl_b200.utils.modes import evaluating  # NOQA
