"""Epsilon-greedy explorers.

They draw from numpy's *global* legacy RandomState in the reference's order
(one ``np.random.rand()`` per decision, pfrl/explorers/epsilon_greedy.py:8-12),
because prioritized sampling shares that stream: seeded runs consume the same
random numbers as the reference.
"""
from logging import getLogger

import numpy as np

from pfrl_b200 import explorer


def select_action_epsilon_greedily(epsilon, random_action_func, greedy_action_func):
    """Returns (action, was_greedy)."""
    if np.random.rand() < epsilon:
        return random_action_func(), False
    return greedy_action_func(), True


class ConstantEpsilonGreedy(explorer.Explorer):
    """Fixed epsilon (pfrl/explorers/epsilon_greedy.py:15-39)."""

    def __init__(self, epsilon, random_action_func, logger=getLogger(__name__)):
        assert 0 <= epsilon <= 1
        self.epsilon = epsilon
        self.random_action_func = random_action_func
        self.logger = logger

    def select_action(self, t, greedy_action_func, action_value=None):
        a, greedy = select_action_epsilon_greedily(
            self.epsilon, self.random_action_func, greedy_action_func)
        self.logger.debug("t:%s a:%s %s", t, a, "greedy" if greedy else "non-greedy")
        return a

    def __repr__(self):
        return "ConstantEpsilonGreedy(epsilon={})".format(self.epsilon)


class LinearDecayEpsilonGreedy(explorer.Explorer):
    """Epsilon annealed linearly from start to end over decay_steps
    (pfrl/explorers/epsilon_greedy.py:42-88)."""

    def __init__(self, start_epsilon, end_epsilon, decay_steps, random_action_func,
                 logger=getLogger(__name__)):
        assert 0 <= start_epsilon <= 1
        assert 0 <= end_epsilon <= 1
        assert decay_steps >= 0
        self.start_epsilon = start_epsilon
        self.end_epsilon = end_epsilon
        self.decay_steps = decay_steps
        self.random_action_func = random_action_func
        self.logger = logger
        self.epsilon = start_epsilon

    def compute_epsilon(self, t):
        if t > self.decay_steps:
            return self.end_epsilon
        span = self.end_epsilon - self.start_epsilon
        return self.start_epsilon + span * (t / self.decay_steps)

    def select_action(self, t, greedy_action_func, action_value=None):
        self.epsilon = self.compute_epsilon(t)
        a, greedy = select_action_epsilon_greedily(
            self.epsilon, self.random_action_func, greedy_action_func)
        self.logger.debug("t:%s a:%s %s", t, a, "greedy" if greedy else "non-greedy")
        return a

    def __repr__(self):
        return "LinearDecayEpsilonGreedy(epsilon={})".format(self.epsilon)


class Greedy(explorer.Explorer):
    """Always the greedy action: no exploration (pfrl/explorers/greedy.py)."""

    def select_action(self, t, greedy_action_func, action_value=None):
        return greedy_action_func()

    def __repr__(self):
        return "Greedy()"


class AdditiveGaussian(explorer.Explorer):
    """greedy action + N(0, scale^2) noise drawn from numpy's global stream,
    optionally clipped to [low, high] (pfrl/explorers/additive_gaussian.py)."""

    def __init__(self, scale, low=None, high=None):
        self.scale = scale
        self.low = low
        self.high = high

    def select_action(self, t, greedy_action_func, action_value=None):
        a = greedy_action_func()
        noisy = a + np.random.normal(scale=self.scale, size=a.shape).astype(np.float32)
        if self.low is None and self.high is None:
            return noisy
        return np.clip(noisy, self.low, self.high)

    def __repr__(self):
        return "AdditiveGaussian(scale={}, low={}, high={})".format(self.scale, self.low, self.high)
