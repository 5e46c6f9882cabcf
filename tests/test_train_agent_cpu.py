"""The single-env training loop's call contract, after the reference's
tests/experiments_tests/test_train_agent.py:10-175 (mock agent / env /
evaluator), plus train_agent_with_evaluation writing scores."""
import os
from unittest import mock

import numpy as np
import pytest


def _scripted_env(resets, steps):
    env = mock.Mock()
    env.reset.side_effect = resets
    env.step.side_effect = steps
    return env


def _check_hooks(hook, env, agent, n):
    assert hook.call_count == n
    for i, call in enumerate(hook.call_args_list):
        assert call[0] == (env, agent, i + 1)       # steps are counted from 1


def test_call_counts_one_episode(tmp_path):
    from pfrl_b200.experiments import train_agent

    agent, hook = mock.Mock(), mock.Mock()
    env = _scripted_env([("state", 0)], [
        (("state", 1), 0, False, {}), (("state", 2), 0, False, {}),
        (("state", 3), -0.5, False, {}), (("state", 4), 0, False, {}),
        (("state", 5), 1, True, {})])
    history = train_agent(agent=agent, env=env, steps=5, outdir=str(tmp_path), step_hooks=[hook])
    assert history == []
    assert agent.act.call_count == 5 and agent.observe.call_count == 5
    assert agent.observe.call_args_list[4][0][2] is True        # done at state 5
    assert env.reset.call_count == 1 and env.step.call_count == 5
    _check_hooks(hook, env, agent, 5)


def test_needs_reset_starts_a_new_episode(tmp_path):
    from pfrl_b200.experiments import train_agent

    agent, hook = mock.Mock(), mock.Mock()
    env = _scripted_env([("state", 0), ("state", 4)], [
        (("state", 1), 0, False, {}), (("state", 2), 0, False, {}),
        (("state", 3), 0, False, {"needs_reset": True}), (("state", 5), -0.5, False, {}),
        (("state", 6), 0, False, {}), (("state", 7), 1, True, {})])
    train_agent(agent=agent, env=env, steps=5, outdir=str(tmp_path), step_hooks=[hook])
    assert agent.act.call_count == 5 and agent.observe.call_count == 5
    third = agent.observe.call_args_list[2][0]
    assert third[2] is False and third[3] is True               # not done, but reset
    assert env.reset.call_count == 2 and env.step.call_count == 5
    _check_hooks(hook, env, agent, 5)


@pytest.mark.parametrize("eval_during_episode", [False, True])
def test_evaluator_is_consulted_at_the_right_steps(eval_during_episode, tmp_path):
    from pfrl_b200.experiments import train_agent

    agent, evaluator = mock.MagicMock(), mock.Mock()
    env = _scripted_env([("state", 0)] * 2, [
        (("state", 1), 0, False, {}), (("state", 2), 0, False, {}),
        (("state", 3), -0.5, True, {}), (("state", 4), 0, False, {}),
        (("state", 5), 1, True, {})])
    train_agent(agent=agent, env=env, steps=5, outdir=str(tmp_path), evaluator=evaluator,
                eval_during_episode=eval_during_episode)
    calls = [c[1] for c in evaluator.evaluate_if_necessary.call_args_list]
    if eval_during_episode:
        assert [c["t"] for c in calls] == [1, 2, 3, 4, 5]
        assert [c["episodes"] for c in calls] == [0, 0, 1, 1, 2]
    else:
        assert [(c["t"], c["episodes"]) for c in calls] == [(3, 1), (5, 2)]


def test_unsupported_evaluation_hook_is_rejected(tmp_path):
    from pfrl_b200.experiments import train_agent_with_evaluation

    class Hook:
        support_train_agent = False

    hook = Hook()
    with pytest.raises(ValueError) as err:
        train_agent_with_evaluation(
            agent=mock.Mock(), env=mock.Mock(), steps=1, eval_n_steps=1, eval_n_episodes=None,
            eval_interval=1, outdir=str(tmp_path), evaluation_hooks=[hook])
    assert str(err.value) == "{} does not support train_agent_with_evaluation().".format(hook)


def test_train_agent_with_evaluation_scores_and_early_stop(tmp_path):
    import torch

    from pfrl_b200 import agents, explorers, q_functions
    from pfrl_b200.envs import ChainEnv
    from pfrl_b200.experiments import train_agent_with_evaluation
    from pfrl_b200.replay_buffers import HostReplayBuffer
    from pfrl_b200.utils import set_random_seed

    set_random_seed(0)
    q = q_functions.FCStateQFunctionWithDiscreteAction(5, 2, 32, 2)
    agent = agents.DoubleDQN(
        q, torch.optim.Adam(q.parameters(), lr=3e-3), HostReplayBuffer(5000), 0.95,
        explorers.LinearDecayEpsilonGreedy(1.0, 0.05, 600, lambda: np.random.randint(2)),
        replay_start_size=50, minibatch_size=32, update_interval=1, target_update_interval=50,
        phi=lambda x: x.astype(np.float32, copy=False))
    out = str(tmp_path / "run")
    agent2, history = train_agent_with_evaluation(
        agent, ChainEnv(seed=0), steps=4000, eval_n_steps=None, eval_n_episodes=3,
        eval_interval=500, outdir=out, eval_env=ChainEnv(seed=1), successful_score=0.9)
    assert agent2 is agent and history and history[-1]["eval_score"] >= 0.9
    assert "average_loss" in history[-1]
    rows = open(os.path.join(out, "scores.txt")).read().strip().splitlines()
    assert rows[0].split("\t")[:4] == ["steps", "episodes", "elapsed", "mean"]
    assert len(rows) == len(history) + 1
    assert os.path.isdir(os.path.join(out, "best"))
    assert history[-1]["cumulative_steps"] < 4000        # stopped early on success
