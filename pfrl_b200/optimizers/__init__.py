from pfrl_b200.optimizers.rmsprop_eps_inside_sqrt import RMSpropEpsInsideSqrt  # NOQA
from pfrl_b200.optimizers.rmsprop_eps_inside_sqrt import SharedRMSpropEpsInsideSqrt  # NOQA
