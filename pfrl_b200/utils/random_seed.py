"""Seeding of every RNG stream the hot path consumes."""
import random

import numpy as np
import torch


def set_random_seed(seed):
    """Seed Python's ``random`` (PPO minibatch shuffles), numpy's GLOBAL legacy
    RandomState (prioritized sampling, sample_n_k, epsilon-greedy -- all on one
    stream, in call order) and torch (CPU + CUDA: NoisyNet noise, IQN taus,
    policy sampling); cf. pfrl/utils/random_seed.py:7-22."""
    for seeder in (random.seed, np.random.seed, torch.manual_seed):
        seeder(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
