from pfrl_b200 import explorer


class Greedy(explorer.Explorer):
    """No exploration (pfrl/explorers/greedy.py:4-13)."""

    def select_action(self, t, greedy_action_func, action_value=None):
        return greedy_action_func()

    def __repr__(self):
        return "Greedy()"
