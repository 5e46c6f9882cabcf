from pfrl_b200.collections.prioritized import PrioritizedBuffer  # NOQA
from pfrl_b200.collections.random_access_queue import RandomAccessQueue  # NOQA
