from pfrl_b200.policies.heads import (  # NOQA
    GaussianHeadWithDiagonalCovariance,
    GaussianHeadWithFixedCovariance,
    GaussianHeadWithStateIndependentCovariance,
    SoftmaxCategoricalHead,
)
