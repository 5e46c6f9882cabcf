"""What the agents ask of an exploration strategy (reference: pfrl/explorer.py)."""
import abc


class Explorer(abc.ABC):
    @abc.abstractmethod
    def select_action(self, t, greedy_action_func, action_value=None):
        """Return the action to take at global step ``t``.

        ``greedy_action_func()`` yields the greedy action on demand (so that a
        purely random choice costs no forward pass); ``action_value`` is the
        ActionValue of the current observation when the agent has one.
        """
        raise NotImplementedError
