"""Step hooks for the ``step_hooks=`` argument of the training loops
(reference: pfrl/experiments/hooks.py).  Any callable ``(env, agent, step)``
works; these two exist so that scripts written for the reference (learning
rate / clip-range decay in the Rainbow and PPO examples) run unchanged."""
import abc

import numpy as np


class StepHook(abc.ABC):
    """Interface of a hook: called once per environment step with
    ``(env, agent, step)``."""

    @abc.abstractmethod
    def __call__(self, env, agent, step):
        raise NotImplementedError


class LinearInterpolationHook(StepHook):
    """Calls ``setter(env, agent, value)`` with ``value`` moving linearly from
    ``start_value`` at step 1 to ``stop_value`` at ``total_steps`` (and staying
    there), e.g. to decay a learning rate."""

    def __init__(self, total_steps, start_value, stop_value, setter):
        self.total_steps = total_steps
        self.start_value = start_value
        self.stop_value = stop_value
        self.setter = setter

    def __call__(self, env, agent, step):
        value = np.interp(step, [1, self.total_steps], [self.start_value, self.stop_value])
        self.setter(env, agent, value)
