"""Module path of the reference (pfrl/policies/deterministic_policy.py)."""
from pfrl_b200.policies.heads import DeterministicHead  # NOQA
