"""Environment interfaces the training loops rely on (reference: pfrl/env.py)."""
import abc


class Env(abc.ABC):
    """One environment: ``reset() -> obs``, ``step(a) -> (obs, r, done, info)``."""

    @abc.abstractmethod
    def reset(self):
        ...

    @abc.abstractmethod
    def step(self, action):
        ...

    @abc.abstractmethod
    def close(self):
        ...


class VectorEnv(abc.ABC):
    """``num_envs`` environments stepped in lock-step.

    ``step(actions)`` returns ``(observations, rewards, dones, infos)``, one
    entry per environment.  ``reset(mask)`` restarts the environments whose
    mask entry is False (all of them for ``mask=None``) and returns the current
    observation of every environment.  ``infos[i].get("needs_reset")`` asks the
    training loop to reset environment i although it is not terminal.
    """

    num_envs = None

    @abc.abstractmethod
    def reset(self, mask=None):
        ...

    @abc.abstractmethod
    def step(self, actions):
        ...

    @abc.abstractmethod
    def seed(self, seeds):
        ...

    @abc.abstractmethod
    def close(self):
        ...

    @property
    def unwrapped(self):
        return self
