"""Periodic evaluation during training.

Keeps the contract train_agent_batch relies on (pfrl/experiments/evaluator.py:
396-522): ``evaluate_if_necessary(t, episodes)`` returns a score when an
evaluation ran, ``max_score`` tracks the best, ``scores.txt`` gets one TSV row
per evaluation (steps, episodes, elapsed, mean, median, stdev, max, min +
the agent's statistics) and the best agent is saved to ``<outdir>/best``.
"""
import logging
import os
import statistics
import time

import numpy as np


def save_agent(agent, t, outdir, logger, suffix=""):
    dirname = os.path.join(outdir, "{}{}".format(t, suffix))
    agent.save(dirname)
    logger.info("Saved the agent to %s", dirname)


def _batch_run_episodes(env, agent, n_steps, n_episodes, max_episode_len=None, logger=None):
    """Run evaluation episodes on a vector env until n_steps or n_episodes."""
    assert (n_steps is None) != (n_episodes is None)
    logger = logger or logging.getLogger(__name__)
    num_envs = env.num_envs
    episode_r = np.zeros(num_envs, dtype=np.float64)
    episode_len = np.zeros(num_envs, dtype="i")
    scores, lengths = [], []
    total = 0
    obss = env.reset()
    while True:
        actions = agent.batch_act(obss)
        obss, rs, dones, infos = env.step(actions)
        rs = np.asarray(rs, dtype=np.float64)
        dones = np.asarray(dones, dtype=bool)
        episode_r += rs
        episode_len += 1
        total += num_envs
        resets = np.zeros(num_envs, dtype=bool) if max_episode_len is None \
            else episode_len == max_episode_len
        resets = np.logical_or(resets, [info.get("needs_reset", False) for info in infos])
        agent.batch_observe(obss, rs, dones, resets)
        end = np.logical_or(resets, dones)
        for i in np.nonzero(end)[0]:
            scores.append(float(episode_r[i]))
            lengths.append(int(episode_len[i]))
        episode_r[end] = 0
        episode_len[end] = 0
        if n_episodes is not None and len(scores) >= n_episodes:
            scores, lengths = scores[:n_episodes], lengths[:n_episodes]
            break
        if n_steps is not None and total >= n_steps:
            break
        obss = env.reset(np.logical_not(end))
    if not scores:  # no episode finished within n_steps: report the partial ones
        scores = [float(x) for x in episode_r]
        lengths = [int(x) for x in episode_len]
    return scores, lengths


def eval_performance(env, agent, n_steps, n_episodes, max_episode_len=None, logger=None):
    with agent.eval_mode():
        scores, lengths = _batch_run_episodes(env, agent, n_steps, n_episodes, max_episode_len,
                                              logger)
    return dict(
        episodes=len(scores), mean=statistics.mean(scores), median=statistics.median(scores),
        stdev=statistics.stdev(scores) if len(scores) >= 2 else 0.0, max=np.max(scores),
        min=np.min(scores), length_mean=statistics.mean(lengths),
        length_median=statistics.median(lengths),
        length_stdev=statistics.stdev(lengths) if len(lengths) >= 2 else 0.0,
        length_max=np.max(lengths), length_min=np.min(lengths))


class Evaluator(object):
    def __init__(self, agent, env, n_steps, n_episodes, eval_interval, outdir,
                 max_episode_len=None, step_offset=0, save_best_so_far_agent=True, logger=None):
        assert (n_steps is None) != (n_episodes is None), \
            "One of n_steps or n_episodes must be None."
        self.agent = agent
        self.env = env
        self.max_score = np.finfo(np.float32).min
        self.start_time = time.time()
        self.n_steps = n_steps
        self.n_episodes = n_episodes
        self.eval_interval = eval_interval
        self.outdir = outdir
        self.max_episode_len = max_episode_len
        self.step_offset = step_offset
        self.prev_eval_t = self.step_offset - self.step_offset % self.eval_interval
        self.save_best_so_far_agent = save_best_so_far_agent
        self.logger = logger or logging.getLogger(__name__)
        os.makedirs(outdir, exist_ok=True)
        with open(os.path.join(outdir, "scores.txt"), "w") as f:
            cols = ("steps", "episodes", "elapsed", "mean", "median", "stdev", "max", "min")
            cols += tuple(name for name, _ in self.agent.get_statistics())
            print("\t".join(cols), file=f)

    def evaluate_and_update_max_score(self, t, episodes):
        stats = eval_performance(self.env, self.agent, self.n_steps, self.n_episodes,
                                 max_episode_len=self.max_episode_len, logger=self.logger)
        elapsed = time.time() - self.start_time
        row = (t, episodes, elapsed, stats["mean"], stats["median"], stats["stdev"],
               stats["max"], stats["min"]) + tuple(v for _, v in self.agent.get_statistics())
        with open(os.path.join(self.outdir, "scores.txt"), "a+") as f:
            print("\t".join(str(x) for x in row), file=f)
        mean = stats["mean"]
        if mean > self.max_score:
            self.logger.info("The best score is updated %s -> %s", self.max_score, mean)
            self.max_score = mean
            if self.save_best_so_far_agent:
                save_agent(self.agent, "best", self.outdir, self.logger)
        return mean

    def evaluate_if_necessary(self, t, episodes):
        if t >= self.prev_eval_t + self.eval_interval:
            score = self.evaluate_and_update_max_score(t, episodes)
            self.prev_eval_t = t - t % self.eval_interval
            return score
        return None
