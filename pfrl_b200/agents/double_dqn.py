from pfrl_b200.agents import dqn
from pfrl_b200.utils.modes import evaluating


class DoubleDQN(dqn.DQN):
    """Double DQN (arXiv:1509.06461): the online network picks the next
    action, the target network evaluates it (pfrl/agents/double_dqn.py:12-40)."""

    def _next_q(self, exp_batch):
        next_state = exp_batch["next_state"]
        with evaluating(self.model):
            next_qout = self.model(next_state)
        target_next_qout = self.target_model(next_state)
        return target_next_qout.evaluate_actions(next_qout.greedy_actions)
