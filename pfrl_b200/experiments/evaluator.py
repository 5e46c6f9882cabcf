"""Periodic evaluation during training.

Same contract as the reference (pfrl/experiments/evaluator.py:12-522), which
its tests pin (tests/experiments_tests/test_evaluator.py:13-369):

* ``run_evaluation_episodes`` (one env) / ``batch_run_evaluation_episodes``
  (a VectorEnv) run the agent in ``eval_mode`` for ``n_episodes`` episodes or
  ``n_steps`` time steps and return ``(scores, lengths)``.
* On a VectorEnv, episodes are numbered in the order they START and only a
  gap-free prefix of that numbering is ever reported, so that short episodes
  finishing early on other environments cannot crowd out long ones; the last
  ``batch_observe`` of an evaluation carries ``reset=True`` for every env.
* ``Evaluator.evaluate_if_necessary(t, episodes)`` evaluates once per
  ``eval_interval`` steps, appends a row to ``<outdir>/scores.txt`` (steps,
  episodes, elapsed, mean, median, stdev, max, min, the agent's statistics,
  the env's statistics), calls the evaluation hooks, tracks ``max_score`` and
  saves the best agent to ``<outdir>/best``.
"""
import logging
import os
import statistics
import time

import numpy as np

from pfrl_b200.env import VectorEnv


def save_agent(agent, t, outdir, logger, suffix=""):
    dirname = os.path.join(outdir, "{}{}".format(t, suffix))
    agent.save(dirname)
    logger.info("Saved the agent to %s", dirname)


# --------------------------------------------------------------------- one env
def _run_episodes(env, agent, n_steps, n_episodes, max_episode_len, logger):
    scores, lengths = [], []
    steps_taken = 0
    ret, length = 0, 0
    obs = env.reset()
    while True:
        obs, r, done, info = env.step(agent.act(obs))
        ret += r
        length += 1
        steps_taken += 1
        ended = done or length == max_episode_len or info.get("needs_reset", False)
        agent.observe(obs, r, done, ended)
        if ended:
            logger.info("evaluation episode %s length:%s R:%s", len(scores), length, ret)
            scores.append(float(ret))       # plain floats: `statistics` dislikes numpy scalars
            lengths.append(float(length))
        budget_spent = len(scores) >= n_episodes if n_steps is None else steps_taken >= n_steps
        if budget_spent:
            break
        if ended:
            ret, length = 0, 0
            obs = env.reset()
    if not scores:  # the whole step budget went into one unfinished episode
        scores.append(float(ret))
        lengths.append(float(length))
    return scores, lengths


def run_evaluation_episodes(env, agent, n_steps, n_episodes, max_episode_len=None, logger=None):
    """Evaluate on a single env; returns ``(scores, lengths)``."""
    assert (n_steps is None) != (n_episodes is None)
    with agent.eval_mode():
        return _run_episodes(env, agent, n_steps, n_episodes, max_episode_len,
                             logger or logging.getLogger(__name__))


# ------------------------------------------------------------------ vector env
def _on_host(values, dtype):
    """Rewards / dones of a vector env as a numpy array (device envs may hand out tensors)."""
    if hasattr(values, "detach"):
        values = values.detach().cpu().numpy()
    return np.asarray(values, dtype=dtype)


class _StartOrderLedger(object):
    """Episodes of a vector env, identified by the order in which they start."""

    def __init__(self, num_envs):
        self.running = list(range(num_envs))  # id of the episode each env is in
        self.next_id = num_envs
        self.finished = {}                    # id -> (return, length)

    def close(self, env_index, ret, length):
        self.finished[self.running[env_index]] = (ret, length)
        self.running[env_index] = self.next_id
        self.next_id += 1

    def prefix(self):
        """Finished episodes 0, 1, ... up to the first one still running."""
        out = []
        while len(out) in self.finished:
            out.append(self.finished[len(out)])
        return out


def _batch_run_episodes(env, agent, n_steps, n_episodes, max_episode_len, logger):
    num_envs = env.num_envs
    ledger = _StartOrderLedger(num_envs)
    ret = np.zeros(num_envs, dtype=np.float64)
    length = np.zeros(num_envs, dtype="i")
    obss = env.reset()
    while True:
        obss, rs, dones, infos = env.step(agent.batch_act(obss))
        ret += _on_host(rs, np.float64)
        dones = _on_host(dones, bool)
        length += 1
        resets = np.zeros(num_envs, dtype=bool) if max_episode_len is None \
            else length == max_episode_len
        resets = np.logical_or(resets, [info.get("needs_reset", False) for info in infos])
        ended = np.logical_or(resets, dones)
        for i in np.flatnonzero(ended):
            ledger.close(i, ret[i], length[i])
        ret[ended] = 0
        length[ended] = 0

        done_in_order = ledger.prefix()
        report = []
        if n_steps is None:
            stop = len(done_in_order) >= n_episodes
            if stop:
                report = done_in_order[:n_episodes]
        else:
            used = 0
            for episode in done_in_order:
                used += episode[1]
                if used > n_steps:      # this one does not fit any more
                    break
                report.append(episode)
            stop = used >= n_steps
            if not stop:
                # the oldest unfinished episode: would it exhaust the budget as it stands?
                i = ledger.running.index(len(done_in_order))
                if used + length[i] >= n_steps:
                    stop = True
                    if not done_in_order:   # nothing finished at all: report the partial episode
                        report.append((ret[i], length[i]))
        if stop:
            resets.fill(True)           # the agent sees every episode end here
        agent.batch_observe(obss, rs, dones, resets)
        if stop:
            break
        obss = env.reset(np.logical_not(ended))
    for k, (r, n) in enumerate(report):
        logger.info("evaluation episode %s length: %s R: %s", k, n, r)
    return [float(r) for r, _ in report], [float(n) for _, n in report]


def batch_run_evaluation_episodes(env, agent, n_steps, n_episodes, max_episode_len=None,
                                  logger=None):
    """Evaluate on a VectorEnv; returns ``(scores, lengths)``."""
    assert (n_steps is None) != (n_episodes is None)
    with agent.eval_mode():
        return _batch_run_episodes(env, agent, n_steps, n_episodes, max_episode_len,
                                   logger or logging.getLogger(__name__))


def eval_performance(env, agent, n_steps, n_episodes, max_episode_len=None, logger=None):
    """Statistics (mean / median / stdev / max / min of returns and lengths)
    of one evaluation."""
    assert (n_steps is None) != (n_episodes is None)
    run = batch_run_evaluation_episodes if isinstance(env, VectorEnv) else run_evaluation_episodes
    scores, lengths = run(env, agent, n_steps, n_episodes, max_episode_len=max_episode_len,
                          logger=logger)

    def spread(xs):
        return statistics.stdev(xs) if len(xs) >= 2 else 0.0

    return dict(
        episodes=len(scores), mean=statistics.mean(scores), median=statistics.median(scores),
        stdev=spread(scores), max=np.max(scores), min=np.min(scores),
        length_mean=statistics.mean(lengths), length_median=statistics.median(lengths),
        length_stdev=spread(lengths), length_max=np.max(lengths), length_min=np.min(lengths))


_SCORE_COLUMNS = ("steps", "episodes", "elapsed", "mean", "median", "stdev", "max", "min")


class Evaluator(object):
    def __init__(self, agent, env, n_steps, n_episodes, eval_interval, outdir,
                 max_episode_len=None, step_offset=0, evaluation_hooks=(),
                 save_best_so_far_agent=True, logger=None, use_tensorboard=False):
        assert (n_steps is None) != (n_episodes is None), \
            "One of n_steps or n_episodes must be None."
        if use_tensorboard:
            raise NotImplementedError("tensorboard logging is outside the scope of pfrl_b200")
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
        self.evaluation_hooks = evaluation_hooks
        self.save_best_so_far_agent = save_best_so_far_agent
        self.logger = logger or logging.getLogger(__name__)
        self.env_get_stats = getattr(self.env, "get_statistics", lambda: [])
        self.env_clear_stats = getattr(self.env, "clear_statistics", lambda: None)
        assert callable(self.env_get_stats) and callable(self.env_clear_stats)
        os.makedirs(outdir, exist_ok=True)
        names = _SCORE_COLUMNS + tuple(n for n, _ in self.agent.get_statistics()) \
            + tuple(n for n, _ in self.env_get_stats())
        with open(os.path.join(outdir, "scores.txt"), "w") as f:
            print("\t".join(names), file=f)

    def evaluate_and_update_max_score(self, t, episodes):
        self.env_clear_stats()
        stats = eval_performance(self.env, self.agent, self.n_steps, self.n_episodes,
                                 max_episode_len=self.max_episode_len, logger=self.logger)
        elapsed = time.time() - self.start_time
        agent_stats = self.agent.get_statistics()
        env_stats = self.env_get_stats()
        mean = stats["mean"]
        row = (t, episodes, elapsed, mean, stats["median"], stats["stdev"], stats["max"],
               stats["min"]) + tuple(v for _, v in agent_stats) + tuple(v for _, v in env_stats)
        with open(os.path.join(self.outdir, "scores.txt"), "a+") as f:
            print("\t".join(str(x) for x in row), file=f)
        for hook in self.evaluation_hooks:
            hook(env=self.env, agent=self.agent, evaluator=self, step=t, eval_stats=stats,
                 agent_stats=agent_stats, env_stats=env_stats)
        if mean > self.max_score:
            self.logger.info("The best score is updated %s -> %s", self.max_score, mean)
            self.max_score = mean
            if self.save_best_so_far_agent:
                save_agent(self.agent, "best", self.outdir, self.logger)
        return mean

    def evaluate_if_necessary(self, t, episodes):
        if t < self.prev_eval_t + self.eval_interval:
            return None
        score = self.evaluate_and_update_max_score(t, episodes)
        self.prev_eval_t = t - t % self.eval_interval
        return score
