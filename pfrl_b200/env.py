from abc import ABCMeta, abstractmethod


class Env(object, metaclass=ABCMeta):
    """Single environment (pfrl/env.py:4-20)."""

    @abstractmethod
    def step(self, action):
        raise NotImplementedError()

    @abstractmethod
    def reset(self):
        raise NotImplementedError()

    @abstractmethod
    def close(self):
        raise NotImplementedError()


class VectorEnv(object, metaclass=ABCMeta):
    """Batch of environments stepped together (pfrl/env.py:23-55):
    ``step(actions) -> (obss, rewards, dones, infos)``, ``reset(mask)`` resets
    the environments whose mask entry is False (all if mask is None)."""

    @abstractmethod
    def step(self, action):
        raise NotImplementedError()

    @abstractmethod
    def reset(self, mask):
        raise NotImplementedError()

    @abstractmethod
    def seed(self, seeds):
        raise NotImplementedError()

    @abstractmethod
    def close(self):
        raise NotImplementedError()

    @property
    def unwrapped(self):
        return self
