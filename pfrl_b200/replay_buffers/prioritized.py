"""Module path of the reference (pfrl/replay_buffers/prioritized.py)."""
from pfrl_b200.replay_buffers.device_buffer import PrioritizedReplayBuffer, PriorityWeightError  # NOQA
