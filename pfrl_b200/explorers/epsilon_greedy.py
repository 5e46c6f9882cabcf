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


class _EpsilonGreedy(explorer.Explorer):
    """Shared mechanics: with probability ``compute_epsilon(t)`` call
    ``random_action_func()``, otherwise take the greedy action.  ``epsilon``
    always holds the value used by the last decision."""

    label = "EpsilonGreedy"

    def __init__(self, epsilon, random_action_func, logger):
        self.epsilon = epsilon
        self.random_action_func = random_action_func
        self.logger = logger

    def compute_epsilon(self, t):
        return self.epsilon

    def select_action(self, t, greedy_action_func, action_value=None):
        self.epsilon = self.compute_epsilon(t)
        action, was_greedy = select_action_epsilon_greedily(
            self.epsilon, self.random_action_func, greedy_action_func)
        self.logger.debug("t:%s a:%s %s", t, action, "greedy" if was_greedy else "non-greedy")
        return action

    def __repr__(self):
        return "{}(epsilon={})".format(self.label, self.epsilon)


def _check_probability(*values):
    for v in values:
        assert 0 <= v <= 1


class ConstantEpsilonGreedy(_EpsilonGreedy):
    """Fixed epsilon (pfrl/explorers/epsilon_greedy.py:15-39)."""

    label = "ConstantEpsilonGreedy"

    def __init__(self, epsilon, random_action_func, logger=getLogger(__name__)):
        _check_probability(epsilon)
        super().__init__(epsilon, random_action_func, logger)


class LinearDecayEpsilonGreedy(_EpsilonGreedy):
    """Epsilon annealed linearly from start to end over decay_steps
    (pfrl/explorers/epsilon_greedy.py:42-88)."""

    label = "LinearDecayEpsilonGreedy"

    def __init__(self, start_epsilon, end_epsilon, decay_steps, random_action_func,
                 logger=getLogger(__name__)):
        _check_probability(start_epsilon, end_epsilon)
        assert decay_steps >= 0
        super().__init__(start_epsilon, random_action_func, logger)
        self.start_epsilon = start_epsilon
        self.end_epsilon = end_epsilon
        self.decay_steps = decay_steps

    def compute_epsilon(self, t):
        if t > self.decay_steps:
            return self.end_epsilon
        span = self.end_epsilon - self.start_epsilon
        return self.start_epsilon + span * (t / self.decay_steps)


class ExponentialDecayEpsilonGreedy(_EpsilonGreedy):
    """epsilon_t = max(start * decay**t, end) (pfrl/explorers/epsilon_greedy.py:91-134)."""

    label = "ExponentialDecayEpsilonGreedy"

    def __init__(self, start_epsilon, end_epsilon, decay, random_action_func,
                 logger=getLogger(__name__)):
        _check_probability(start_epsilon, end_epsilon)
        assert 0 < decay < 1
        super().__init__(start_epsilon, random_action_func, logger)
        self.start_epsilon = start_epsilon
        self.end_epsilon = end_epsilon
        self.decay = decay

    def compute_epsilon(self, t):
        return max(self.start_epsilon * (self.decay ** t), self.end_epsilon)


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
