from pfrl_b200.collections.prioritized import PrioritizedBuffer  # NOQA
