"""Module path of the reference (pfrl/nn/concat_obs_and_action.py)."""
from pfrl_b200.nn.containers import ConcatObsAndAction  # NOQA
