"""Single-environment training loop (pfrl/experiments/train_agent.py:24-221),
used by the CPU quickstart configuration."""
import logging
import os

from pfrl_b200.experiments.evaluator import save_agent


def train_agent(agent, env, steps, outdir, checkpoint_freq=None, max_episode_len=None,
                step_offset=0, evaluator=None, successful_score=None, step_hooks=(),
                logger=None):
    logger = logger or logging.getLogger(__name__)
    episode_r = 0
    episode_idx = 0
    obs = env.reset()
    t = step_offset
    if hasattr(agent, "t"):
        agent.t = step_offset
    eval_stats_history = []
    episode_len = 0
    try:
        while t < steps:
            action = agent.act(obs)
            obs, r, done, info = env.step(action)
            t += 1
            episode_r += r
            episode_len += 1
            reset = episode_len == max_episode_len or info.get("needs_reset", False)
            agent.observe(obs, r, done, reset)
            for hook in step_hooks:
                hook(env, agent, t)
            episode_end = done or reset or t == steps
            if episode_end:
                logger.info("outdir:%s step:%s episode:%s R:%s", outdir, t, episode_idx, episode_r)
                logger.info("statistics:%s", agent.get_statistics())
                episode_idx += 1
            if evaluator is not None and (episode_end or not getattr(evaluator, "eval_after_episode", True)):
                eval_score = evaluator.evaluate_if_necessary(t=t, episodes=episode_idx)
                if eval_score is not None:
                    eval_stats = dict(agent.get_statistics())
                    eval_stats["eval_score"] = eval_score
                    eval_stats_history.append(eval_stats)
                if successful_score is not None and evaluator.max_score >= successful_score:
                    break
            if episode_end:
                if t == steps:
                    break
                episode_r = 0
                episode_len = 0
                obs = env.reset()
            if checkpoint_freq and t % checkpoint_freq == 0:
                save_agent(agent, t, outdir, logger, suffix="_checkpoint")
    except (Exception, KeyboardInterrupt):
        save_agent(agent, t, outdir, logger, suffix="_except")
        raise
    save_agent(agent, t, outdir, logger, suffix="_finish")
    return eval_stats_history


def train_agent_with_evaluation(agent, env, steps, eval_n_steps, eval_n_episodes, eval_interval,
                                outdir, **kwargs):
    os.makedirs(outdir, exist_ok=True)
    return agent, train_agent(agent, env, steps, outdir, **kwargs)
