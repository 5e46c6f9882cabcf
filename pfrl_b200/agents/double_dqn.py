from pfrl_b200.agents import dqn
from pfrl_b200.utils.contexts import evaluating


class DoubleDQN(dqn.DQN):
    """Double DQN (arXiv:1509.06461): the online network picks the next
    action, the target network evaluates it (pfrl/agents/double_dqn.py:12-40)."""

    def _compute_target_values(self, exp_batch):
        next_state = exp_batch["next_state"]
        with evaluating(self.model):
            next_qout = self.model(next_state)
        target_next_qout = self.target_model(next_state)
        next_q_max = target_next_qout.evaluate_actions(next_qout.greedy_actions)
        return exp_batch["reward"] + exp_batch["discount"] * (
            1.0 - exp_batch["is_state_terminal"]) * next_q_max
