"""Module path of the reference (pfrl/replay_buffers/replay_buffer.py)."""
from pfrl_b200.replay_buffers.device_buffer import ReplayBuffer  # NOQA
