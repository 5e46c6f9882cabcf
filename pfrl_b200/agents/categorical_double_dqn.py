from pfrl_b200.agents import categorical_dqn
from pfrl_b200.utils.modes import evaluating


class CategoricalDoubleDQN(categorical_dqn.CategoricalDQN):
    """Categorical Double DQN -- the agent of the Rainbow reproduction
    (pfrl/agents/categorical_double_dqn.py:10-52): the online network (in
    eval mode) picks a*, the target network (eval mode) supplies p(s', a*)."""

    def _next_distribution(self, exp_batch):
        next_state = exp_batch["next_state"]
        with evaluating(self.target_model), evaluating(self.model):
            target_next_qout = self.target_model(next_state)
            next_qout = self.model(next_state)
        next_p = target_next_qout.evaluate_actions_as_distribution(
            next_qout.greedy_actions.detach())
        return next_p.detach(), target_next_qout.z_values
