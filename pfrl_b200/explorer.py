from abc import ABCMeta, abstractmethod


class Explorer(object, metaclass=ABCMeta):
    """Exploration strategy interface (pfrl/explorer.py)."""

    @abstractmethod
    def select_action(self, t, greedy_action_func, action_value=None):
        """Choose an action at step t given a callable that returns the
        greedy one."""
        raise NotImplementedError()
