"""Synchronous vector-env training driver.

Contract of pfrl/experiments/train_agent_batch.py:10-263, as pinned by the reference's
tests/experiments_tests/test_train_agent_batch.py (run against this module by
tools/run_reference_tests.py): one `agent.batch_act` / `env.step` /
`agent.batch_observe` per vector step; the global step counter advances by `num_envs`
per vector step and every single value of it is shown to the step hooks and to the
checkpoint rule; an environment is reset (through `env.reset(mask)`, mask = "keep
going") when it reports done, when `info["needs_reset"]` is set or when its episode
reaches `max_episode_len`; the agent is saved as `<t>_finish` at the end and
`<t>_except` when anything raises.

What is organised differently from the reference: the per-environment bookkeeping lives
in `EpisodeLedger`, which keeps its counters in the memory space the environment reports
in.  A device-resident vector env that returns CUDA tensors for rewards / dones gets
device-side accounting (returns, lengths, episode count, the window of recent returns) and
the driver reads them back only when a log line or an evaluation needs a number; the one
small transfer per step is the packed (done | reset) mask the replay windows of the agent
need on the host.
"""
import logging
import os

import numpy as np

from pfrl_b200.experiments.evaluator import Evaluator, save_agent

try:  # torch is needed only for device-resident environments
    import torch
except ImportError:  # pragma: no cover
    torch = None


def _is_device_tensor(x):
    return torch is not None and isinstance(x, torch.Tensor) and x.is_cuda


class EpisodeLedger:
    """Returns, lengths and episode counts of `num_envs` environments, plus a ring of
    the last `window` finished returns -- on the host (numpy) or on the device (torch),
    decided by the first reward batch it sees."""

    def __init__(self, num_envs, window):
        self.num_envs = num_envs
        self.window = window
        self.on_device = None
        self._ret = self._len = self._ring = None
        self._episodes = 0
        self._finished = 0  # returns ever written to the ring

    # -- lazily bound storage ------------------------------------------------
    def _bind(self, rewards):
        self.on_device = _is_device_tensor(rewards)
        if self.on_device:
            dev = rewards.device
            self._ret = torch.zeros(self.num_envs, dtype=torch.float64, device=dev)
            self._len = torch.zeros(self.num_envs, dtype=torch.int32, device=dev)
            # one spare slot at the end swallows the writes of unfinished environments
            self._ring = torch.zeros(max(self.window, 1) + 1, dtype=torch.float64, device=dev)
            self._episodes = torch.zeros((), dtype=torch.int64, device=dev)
            self._finished = torch.zeros((), dtype=torch.int64, device=dev)
        else:
            self._ret = np.zeros(self.num_envs, dtype=np.float64)
            self._len = np.zeros(self.num_envs, dtype=np.int32)
            self._ring = np.zeros(max(self.window, 1), dtype=np.float64)

    # -- one vector step -----------------------------------------------------
    def advance(self, rewards, dones, infos, max_episode_len):
        """Account one step; returns (rewards, dones, resets) in the form the agent
        takes (host arrays; device inputs are copied once, packed)."""
        if self._ret is None:
            self._bind(rewards)
        wants_reset = [bool(i.get("needs_reset", False)) for i in infos]
        if self.on_device:
            r = rewards.to(torch.float64)
            d = dones.to(torch.bool)
            self._ret += r
            self._len += 1
            resets = torch.as_tensor(wants_reset, device=r.device)
            if max_episode_len is not None:
                resets = resets | (self._len == max_episode_len)
            self._end = d | resets
            self._close_episodes_device()
            # what the agent's host-side replay windows need: one packed D2H copy
            packed = torch.stack([r, d.to(torch.float64), resets.to(torch.float64)]).cpu().numpy()
            self._end_host = (packed[1] != 0) | (packed[2] != 0)
            return packed[0], packed[1] != 0, packed[2] != 0
        r = np.asarray(rewards, dtype=np.float64)
        d = np.asarray(dones, dtype=bool)
        self._ret += r
        self._len += 1
        resets = np.asarray(wants_reset, dtype=bool)
        if max_episode_len is not None:
            resets = resets | (self._len == max_episode_len)
        self._end = d | resets
        for ret in self._ret[self._end]:
            self._ring[self._finished % len(self._ring)] = ret
            self._finished += 1
        self._episodes += int(self._end.sum())
        return r, d, resets

    def _close_episodes_device(self):
        end = self._end
        n_end = end.sum()
        # finished returns go to consecutive ring slots; fixed shapes, no host round trip
        w = self._ring.numel() - 1
        order = torch.cumsum(end.to(torch.int64), 0) - 1
        slots = torch.where(end, (self._finished + order) % w, torch.full_like(order, w))
        self._ring.scatter_(0, slots, self._ret)
        self._finished += n_end
        self._episodes += n_end

    def restart_finished(self):
        """Zero the counters of the environments whose episode just ended; returns the
        `not_end` mask for `env.reset(mask)` (host bools)."""
        end = self._end
        if self.on_device:
            self._ret = torch.where(end, torch.zeros_like(self._ret), self._ret)
            self._len = torch.where(end, torch.zeros_like(self._len), self._len)
            return np.logical_not(self._end_host)
        self._ret[end] = 0
        self._len[end] = 0
        return np.logical_not(end)

    # -- numbers, read on demand ---------------------------------------------
    @property
    def episodes(self):
        return int(self._episodes)

    def recent_returns(self):
        n = int(self._finished)
        if n == 0:
            return np.zeros(0)
        ring = self._ring[:-1].cpu().numpy() if self.on_device else self._ring
        k = min(n, len(ring))
        newest = (n - 1) % len(ring)
        return np.roll(ring, -(newest + 1))[-k:]  # oldest ... newest


class _Run:
    """One call of train_agent_batch: the loop and its side effects."""

    def __init__(self, agent, env, steps, outdir, checkpoint_freq, log_interval, max_episode_len,
                 step_offset, evaluator, successful_score, step_hooks, return_window_size, logger):
        self.agent, self.env, self.steps, self.outdir = agent, env, steps, outdir
        self.checkpoint_freq, self.log_interval = checkpoint_freq, log_interval
        self.max_episode_len, self.evaluator = max_episode_len, evaluator
        self.successful_score, self.step_hooks = successful_score, step_hooks
        self.logger = logger or logging.getLogger(__name__)
        self.ledger = EpisodeLedger(env.num_envs, return_window_size)
        self.t = step_offset
        self.history = []

    def _count_env_steps(self):
        # every value of the global counter is visible to checkpoints and hooks
        for _ in range(self.env.num_envs):
            self.t += 1
            if self.checkpoint_freq and self.t % self.checkpoint_freq == 0:
                save_agent(self.agent, self.t, self.outdir, self.logger, suffix="_checkpoint")
            for hook in self.step_hooks:
                hook(self.env, self.agent, self.t)

    def _log_due(self):
        li = self.log_interval
        return li is not None and self.t >= li and self.t % li < self.env.num_envs

    def _log(self):
        recent = self.ledger.recent_returns()
        self.logger.info("outdir:%s step:%s episode:%s last_R: %s average_R:%s", self.outdir,
                         self.t, self.ledger.episodes, recent[-1] if len(recent) else np.nan,
                         np.mean(recent) if len(recent) else np.nan)
        self.logger.info("statistics: %s", self.agent.get_statistics())

    def _evaluate(self):
        """True when the success criterion stops the run."""
        score = self.evaluator.evaluate_if_necessary(t=self.t, episodes=self.ledger.episodes)
        if score is None:
            return False
        stats = dict(self.agent.get_statistics())
        stats["eval_score"] = score
        self.history.append(stats)
        return (self.successful_score is not None
                and self.evaluator.max_score >= self.successful_score)

    def loop(self):
        agent, env, ledger = self.agent, self.env, self.ledger
        if hasattr(agent, "t"):
            agent.t = self.t
        obss = env.reset()
        while True:
            obss, rewards, dones, infos = env.step(agent.batch_act(obss))
            rewards, dones, resets = ledger.advance(rewards, dones, infos, self.max_episode_len)
            agent.batch_observe(obss, rewards, dones, resets)
            self._count_env_steps()
            if self._log_due():
                self._log()
            if self.evaluator and self._evaluate():
                return
            if self.t >= self.steps:
                return
            obss = env.reset(ledger.restart_finished())


def train_agent_batch(agent, env, steps, outdir, checkpoint_freq=None, log_interval=None,
                      max_episode_len=None, step_offset=0, evaluator=None,
                      successful_score=None, step_hooks=(), return_window_size=100,
                      logger=None):
    """Train ``agent`` on the vector env ``env`` for ``steps`` env steps.
    Returns the list of evaluation statistics dicts."""
    run = _Run(agent, env, steps, outdir, checkpoint_freq, log_interval, max_episode_len,
               step_offset, evaluator, successful_score, step_hooks, return_window_size, logger)
    try:
        run.loop()
    except (Exception, KeyboardInterrupt):
        save_agent(agent, run.t, outdir, run.logger, suffix="_except")
        env.close()
        if evaluator:
            evaluator.env.close()
        raise
    save_agent(agent, run.t, outdir, run.logger, suffix="_finish")
    return run.history


def train_agent_batch_with_evaluation(
        agent, env, steps, eval_n_steps, eval_n_episodes, eval_interval, outdir,
        checkpoint_freq=None, max_episode_len=None, step_offset=0, eval_max_episode_len=None,
        return_window_size=100, eval_env=None, log_interval=None, successful_score=None,
        step_hooks=(), evaluation_hooks=(), save_best_so_far_agent=True, use_tensorboard=False,
        logger=None):
    """train_agent_batch + periodic evaluation; returns (agent, history)."""
    logger = logger or logging.getLogger(__name__)
    unsupported = [h for h in evaluation_hooks
                   if not getattr(h, "support_train_agent_batch", False)]
    if unsupported:
        raise ValueError(
            "{} does not support train_agent_batch_with_evaluation().".format(unsupported[0]))
    os.makedirs(outdir, exist_ok=True)
    evaluator = Evaluator(
        agent=agent, n_steps=eval_n_steps, n_episodes=eval_n_episodes,
        eval_interval=eval_interval, outdir=outdir,
        max_episode_len=max_episode_len if eval_max_episode_len is None else eval_max_episode_len,
        env=env if eval_env is None else eval_env, step_offset=step_offset,
        evaluation_hooks=evaluation_hooks, save_best_so_far_agent=save_best_so_far_agent,
        use_tensorboard=use_tensorboard, logger=logger)
    history = train_agent_batch(
        agent, env, steps, outdir, checkpoint_freq=checkpoint_freq,
        max_episode_len=max_episode_len, step_offset=step_offset, evaluator=evaluator,
        successful_score=successful_score, return_window_size=return_window_size,
        log_interval=log_interval, step_hooks=step_hooks, logger=logger)
    return agent, history
