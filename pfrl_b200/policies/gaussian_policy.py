"""Module path of the reference (pfrl/policies/gaussian_policy.py)."""
from pfrl_b200.policies.heads import (  # NOQA
    GaussianHeadWithDiagonalCovariance,
    GaussianHeadWithFixedCovariance,
    GaussianHeadWithStateIndependentCovariance,
)
