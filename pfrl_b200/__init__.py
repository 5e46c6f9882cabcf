"""pfrl_b200 -- a B200-native (sm_100a) batched-RL hot path behind PFRL's
public surface.

Python/PyTorch host code mirrors the reference's class names and signatures
(`pfrl_b200.replay_buffers.PrioritizedReplayBuffer`, `pfrl_b200.agents.DQN`,
`pfrl_b200.experiments.train_agent_batch`, ...) and calls hand-written CUDA
through the C ABI of include/b2rl.h (libb2rl.so, loaded with ctypes).
"""
__version__ = "0.1.0"
