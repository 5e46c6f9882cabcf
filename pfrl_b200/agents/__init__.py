from pfrl_b200.agents.categorical_double_dqn import CategoricalDoubleDQN  # NOQA
from pfrl_b200.agents.categorical_dqn import CategoricalDQN  # NOQA
from pfrl_b200.agents.double_dqn import DoubleDQN  # NOQA
from pfrl_b200.agents.dqn import DQN  # NOQA
