from abc import ABCMeta, abstractmethod


class StateQFunction(object, metaclass=ABCMeta):
    """Q(s, .) -> ActionValue (pfrl/q_function.py)."""

    @abstractmethod
    def __call__(self, x):
        raise NotImplementedError()
