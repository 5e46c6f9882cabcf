from pfrl_b200.policies.heads import (  # NOQA
    DeterministicHead,
    GaussianHeadWithDiagonalCovariance,
    GaussianHeadWithFixedCovariance,
    GaussianHeadWithStateIndependentCovariance,
    SoftmaxCategoricalHead,
)
