"""Evaluation contract, after the reference's
tests/experiments_tests/test_evaluator.py:13-369: when the Evaluator runs,
what it saves, how many agent calls an evaluation makes, and -- on a vector
env -- which episodes are reported (start order, gap-free prefix) under an
episode budget and under a step budget."""
from unittest import mock

import numpy as np
import pytest

from pfrl_b200.envs import SerialVectorEnv
from pfrl_b200.experiments import evaluator


@pytest.mark.parametrize("save_best_so_far_agent", [True, False])
@pytest.mark.parametrize("n_steps", [None, 1, 2])
@pytest.mark.parametrize("n_episodes", [None, 1, 2])
def test_evaluate_if_necessary(save_best_so_far_agent, n_steps, n_episodes, tmp_path):
    agent = mock.MagicMock()            # MagicMock supports the eval_mode() context manager
    agent.act.return_value = "action"
    agent.get_statistics.return_value = []
    env = mock.Mock()
    env.reset.return_value = "obs"
    env.step.return_value = ("obs", 0, True, {})
    env.get_statistics.return_value = []
    hook = mock.Mock()
    kw = dict(agent=agent, env=env, n_steps=n_steps, n_episodes=n_episodes, eval_interval=3,
              outdir=str(tmp_path), max_episode_len=None, step_offset=0,
              evaluation_hooks=[hook], save_best_so_far_agent=save_best_so_far_agent)
    if (n_steps is None) == (n_episodes is None):
        with pytest.raises(AssertionError):
            evaluator.Evaluator(**kw)
        return
    budget = n_steps or n_episodes      # every episode is one step long here
    ev = evaluator.Evaluator(**kw)
    for t in (1, 2):
        assert ev.evaluate_if_necessary(t=t, episodes=t) is None
    assert agent.act.call_count == 0 and hook.call_count == 0
    saves = int(save_best_so_far_agent)
    assert ev.evaluate_if_necessary(t=3, episodes=3) == 0          # first evaluation
    assert agent.act.call_count == budget and agent.observe.call_count == budget
    assert hook.call_count == 1 and agent.save.call_count == saves
    assert set(hook.call_args[1]) == {"env", "agent", "evaluator", "step", "eval_stats",
                                      "agent_stats", "env_stats"}
    ev.evaluate_if_necessary(t=6, episodes=6)                      # same score: no new best
    assert agent.act.call_count == 2 * budget
    assert hook.call_count == 2 and agent.save.call_count == saves
    env.step.return_value = ("obs", 1, True, {})                   # better score
    assert ev.evaluate_if_necessary(t=9, episodes=9) == 1
    assert agent.act.call_count == 3 * budget
    assert hook.call_count == 3 and agent.save.call_count == 2 * saves
    rows = open(str(tmp_path / "scores.txt")).read().strip().splitlines()
    assert len(rows) == 4 and rows[0].split("\t")[:3] == ["steps", "episodes", "elapsed"]


def _one_env():
    env = mock.Mock()
    # episode A: 0 -> 1 -> 2 -> 3 (needs_reset); episode B: 4 -> 5 -> 6 -> 7 (done)
    env.reset.side_effect = [("state", 0), ("state", 4)]
    env.step.side_effect = [
        (("state", 1), 0.1, False, {}), (("state", 2), 0.2, False, {}),
        (("state", 3), 0.3, False, {"needs_reset": True}),
        (("state", 5), -0.5, False, {}), (("state", 6), 0, False, {}),
        (("state", 7), 1, True, {})]
    return env


@pytest.mark.parametrize("n_steps", [2, 5, 6])
def test_single_env_step_budget(n_steps):
    agent = mock.MagicMock()
    with pytest.raises(AssertionError):
        evaluator.run_evaluation_episodes(_one_env(), agent, n_steps=n_steps, n_episodes=1)
    scores, lengths = evaluator.run_evaluation_episodes(_one_env(), agent, n_steps=n_steps,
                                                        n_episodes=None)
    assert agent.act.call_count == n_steps and agent.observe.call_count == n_steps
    want = {2: ([0.3], [2]), 5: ([0.6], [3]), 6: ([0.6, 0.5], [3, 3])}[n_steps]
    np.testing.assert_allclose(scores, want[0])
    np.testing.assert_allclose(lengths, want[1])


def test_single_env_episode_budget():
    agent = mock.MagicMock()
    scores, lengths = evaluator.run_evaluation_episodes(_one_env(), agent, n_steps=None,
                                                        n_episodes=2)
    np.testing.assert_allclose(scores, [0.6, 0.5])
    np.testing.assert_allclose(lengths, [3, 3])
    assert agent.act.call_count == 6 and agent.observe.call_count == 6
    # the reset flag reaches the agent at the needs_reset step
    third = agent.observe.call_args_list[2][0]
    assert third[2] is False and third[3] is True


def _two_envs(first_rewards):
    def make(idx):
        env = mock.Mock()
        if idx == 0:
            # A: 0 -> 1 -> 2 -> 3 (needs_reset); B: 4 -> 5 -> 6 -> 7 (done)
            env.reset.side_effect = [("state", 0), ("state", 4)]
            r = first_rewards
            env.step.side_effect = [
                (("state", 1), r[0], False, {}), (("state", 2), r[1], False, {}),
                (("state", 3), r[2], False, {"needs_reset": True}),
                (("state", 5), -0.5, False, {}), (("state", 6), 0, False, {}),
                (("state", 7), 1, True, {})]
        else:
            # a: 0 -> 1 (needs_reset); b: 2 -> 3 (needs_reset); c: 4 -> 5 -> 6 -> 7 (done)
            env.reset.side_effect = [("state", 0), ("state", 2), ("state", 4)]
            env.step.side_effect = [
                (("state", 1), 2, False, {"needs_reset": True}),
                (("state", 3), 3, False, {"needs_reset": True}),
                (("state", 5), -0.6, False, {}), (("state", 6), 0, False, {}),
                (("state", 7), 1, True, {})]
        return env

    return SerialVectorEnv([make(0), make(1)])


@pytest.mark.parametrize("n_steps", [2, 5, 6])
def test_vector_env_step_budget(n_steps):
    agent = mock.MagicMock()
    agent.batch_act.side_effect = [[1, 1]] * 5
    with pytest.raises(AssertionError):
        evaluator.batch_run_evaluation_episodes(_two_envs((0, 0.1, 0.2)), agent, n_steps=n_steps,
                                                n_episodes=1)
    # start order: A (env 0), a (env 1), b, B/c ...; timeline
    #   env 0: [1   2  (3_A)  5  6 (7_B)]
    #   env 1: [(1_a) (3_b) 5  6 (7_c)]
    scores, lengths = evaluator.batch_run_evaluation_episodes(
        _two_envs((0, 0.1, 0.2)), agent, n_steps=n_steps, n_episodes=None)
    if n_steps == 2:
        # A is still running after 2 steps and nothing before it has finished:
        # its partial return is the result, although a and b are complete
        np.testing.assert_allclose(scores, [0.1])
        np.testing.assert_allclose(lengths, [2])
        assert agent.batch_observe.call_count == 2
    else:
        np.testing.assert_allclose(scores, [0.3, 2.0, 3.0])
        np.testing.assert_allclose(lengths, [3, 1, 1])
    assert all(agent.batch_observe.call_args[0][3])    # final batch_reset is all True


def test_vector_env_episode_budget():
    agent = mock.MagicMock()
    agent.batch_act.side_effect = [[1, 1]] * 5
    scores, lengths = evaluator.batch_run_evaluation_episodes(
        _two_envs((0, 0, 0)), agent, n_steps=None, n_episodes=4)
    # reported in START order: A, a, b, c (c started before B)
    np.testing.assert_allclose(scores, [0, 2, 3, 0.4])
    np.testing.assert_allclose(lengths, [3, 1, 1, 3])
    assert all(agent.batch_observe.call_args[0][3])


def test_eval_performance_dispatches_on_vector_env():
    agent = mock.MagicMock()
    agent.batch_act.side_effect = [[1, 1]] * 5
    stats = evaluator.eval_performance(_two_envs((0, 0, 0)), agent, None, 4)
    assert stats["episodes"] == 4 and abs(stats["mean"] - 1.35) < 1e-12
    assert stats["length_max"] == 3 and stats["length_min"] == 1
    agent = mock.MagicMock()
    stats = evaluator.eval_performance(_one_env(), agent, None, 2)
    assert stats["episodes"] == 2 and abs(stats["mean"] - 0.55) < 1e-12
