from pfrl_b200.explorers.epsilon_greedy import ConstantEpsilonGreedy  # NOQA
from pfrl_b200.explorers.epsilon_greedy import Greedy  # NOQA
from pfrl_b200.explorers.epsilon_greedy import LinearDecayEpsilonGreedy  # NOQA
from pfrl_b200.explorers.epsilon_greedy import AdditiveGaussian  # NOQA
