"""Single-environment training loop, the ``num_envs == 1`` sibling of
train_agent_batch (reference: pfrl/experiments/train_agent.py:24-260; used by
the CPU quickstart configuration).

Order of events per step, which the reference's tests pin
(tests/experiments_tests/test_train_agent.py): act, env.step, observe, step
hooks (step counted from 1), episode bookkeeping, evaluation (at episode ends,
or at any step with ``eval_during_episode``), early stop on
``successful_score``, env.reset, checkpoint.
"""
import logging
import os

from pfrl_b200.experiments.evaluator import Evaluator, save_agent


class _Episode(object):
    __slots__ = ("ret", "length")

    def __init__(self):
        self.ret = 0
        self.length = 0


def train_agent(agent, env, steps, outdir, checkpoint_freq=None, max_episode_len=None,
                step_offset=0, evaluator=None, successful_score=None, step_hooks=(),
                eval_during_episode=False, logger=None):
    """Returns the list of statistics dicts recorded at each evaluation."""
    logger = logger or logging.getLogger(__name__)
    history = []
    t = step_offset
    if hasattr(agent, "t"):
        agent.t = step_offset
    finished_episodes = 0
    ep = _Episode()
    obs = env.reset()

    def evaluate():
        """True when training should stop (successful_score reached)."""
        score = evaluator.evaluate_if_necessary(t=t, episodes=finished_episodes)
        if score is not None:
            record = dict(agent.get_statistics())
            record["eval_score"] = score
            history.append(record)
        return successful_score is not None and evaluator.max_score >= successful_score

    try:
        while t < steps:
            obs, reward, done, info = env.step(agent.act(obs))
            t += 1
            ep.ret += reward
            ep.length += 1
            reset = ep.length == max_episode_len or info.get("needs_reset", False)
            agent.observe(obs, reward, done, reset)
            for hook in step_hooks:
                hook(env, agent, t)
            last_step = t == steps
            ended = done or reset or last_step
            if ended:
                logger.info("outdir:%s step:%s episode:%s R:%s", outdir, t, finished_episodes,
                            ep.ret)
                logger.info("statistics:%s", agent.get_statistics())
                finished_episodes += 1
            if evaluator is not None and (ended or eval_during_episode) and evaluate():
                break
            if ended:
                if last_step:
                    break
                ep = _Episode()
                obs = env.reset()
            if checkpoint_freq and t % checkpoint_freq == 0:
                save_agent(agent, t, outdir, logger, suffix="_checkpoint")
    except (Exception, KeyboardInterrupt):
        save_agent(agent, t, outdir, logger, suffix="_except")  # keep what was learnt so far
        raise
    save_agent(agent, t, outdir, logger, suffix="_finish")
    return history


def train_agent_with_evaluation(agent, env, steps, eval_n_steps, eval_n_episodes, eval_interval,
                                outdir, checkpoint_freq=None, train_max_episode_len=None,
                                step_offset=0, eval_max_episode_len=None, eval_env=None,
                                successful_score=None, step_hooks=(), evaluation_hooks=(),
                                save_best_so_far_agent=True, use_tensorboard=False,
                                eval_during_episode=False, logger=None):
    """train_agent + an Evaluator that runs every ``eval_interval`` steps on
    ``eval_env`` (default: the training env, then only between episodes).
    Returns ``(agent, eval_stats_history)``."""
    logger = logger or logging.getLogger(__name__)
    for hook in evaluation_hooks:
        if not getattr(hook, "support_train_agent", False):
            raise ValueError("{} does not support train_agent_with_evaluation().".format(hook))
    os.makedirs(outdir, exist_ok=True)
    if eval_env is None:
        assert not eval_during_episode, (
            "To run evaluation during training episodes, you need to specify `eval_env`"
            " that is independent from `env`.")
        eval_env = env
    evaluator = Evaluator(
        agent=agent, env=eval_env, n_steps=eval_n_steps, n_episodes=eval_n_episodes,
        eval_interval=eval_interval, outdir=outdir,
        max_episode_len=train_max_episode_len if eval_max_episode_len is None
        else eval_max_episode_len,
        step_offset=step_offset, evaluation_hooks=evaluation_hooks,
        save_best_so_far_agent=save_best_so_far_agent, use_tensorboard=use_tensorboard,
        logger=logger)
    history = train_agent(
        agent, env, steps, outdir, checkpoint_freq=checkpoint_freq,
        max_episode_len=train_max_episode_len, step_offset=step_offset, evaluator=evaluator,
        successful_score=successful_score, step_hooks=step_hooks,
        eval_during_episode=eval_during_episode, logger=logger)
    return agent, history
