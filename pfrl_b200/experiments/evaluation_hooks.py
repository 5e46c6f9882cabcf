"""Hooks called by the Evaluator after each evaluation (reference:
pfrl/experiments/evaluation_hooks.py:12-53).  A hook declares which training
loops it may be used with; the loops refuse hooks that do not support them."""
import abc


class EvaluationHook(abc.ABC):
    support_train_agent = False
    support_train_agent_batch = False
    support_train_agent_async = False

    @abc.abstractmethod
    def __call__(self, env, agent, evaluator, step, eval_stats, agent_stats, env_stats):
        """``step`` is the training step of the evaluation; ``eval_stats`` comes
        from eval_performance, ``agent_stats`` / ``env_stats`` from the
        respective ``get_statistics()``."""
        raise NotImplementedError
