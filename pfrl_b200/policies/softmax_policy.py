"""Module path of the reference (pfrl/policies/softmax_policy.py)."""
from pfrl_b200.policies.heads import SoftmaxCategoricalHead  # NOQA
