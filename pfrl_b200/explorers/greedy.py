"""Module path of the reference (pfrl/explorers/greedy.py)."""
from pfrl_b200.explorers.epsilon_greedy import Greedy  # NOQA
