from pfrl_b200.explorers.epsilon_greedy import ConstantEpsilonGreedy  # NOQA
from pfrl_b200.explorers.epsilon_greedy import Greedy  # NOQA
from pfrl_b200.explorers.epsilon_greedy import LinearDecayEpsilonGreedy  # NOQA
from pfrl_b200.explorers.epsilon_greedy import AdditiveGaussian  # NOQA
from pfrl_b200.explorers.stochastic import AdditiveOU  # NOQA
from pfrl_b200.explorers.stochastic import Boltzmann  # NOQA
from pfrl_b200.explorers.stochastic import ExponentialDecayEpsilonGreedy  # NOQA
