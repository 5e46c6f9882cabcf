from pfrl_b200.policies.gaussian_policy import (  # NOQA
    GaussianHeadWithDiagonalCovariance,
    GaussianHeadWithFixedCovariance,
    GaussianHeadWithStateIndependentCovariance,
)
from pfrl_b200.policies.softmax_policy import SoftmaxCategoricalHead  # NOQA
