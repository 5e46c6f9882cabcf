"""End-to-end parity of the DQN family with seeded runs of the REAL reference
(tests/golden/agent_trace_*.npz, oracle/gen_golden_losses.py:agent_traces;
the construction code is shared so both sides are configured identically):
DoubleDQN, DQN (sum accumulator, MSE, soft target updates, gradient clipping),
CategoricalDQN and "Rainbow" (CategoricalDoubleDQN + factorised noisy layers),
all on 3-step prioritised replay with capacity wrap-around.

Checked at every vector step: the chosen actions (epsilon-greedy / noisy
greedy, so the exploration stream, the noise stream and the argmax all agree),
the number of updates and buffer length, average_q / average_loss; at the end
the parameters, the beta schedule and the priority trees' total / max.  The
runs are chaotic: one different sampled index would show up in every later
number, so agreement here means every minibatch was the same.

CPU: agents in their torch formulation, device buffers running their real
host logic over tests/fake_store.OracleBackedStore.
"""
import os
import sys
from unittest import mock

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from fake_store import OracleBackedStore  # noqa: E402

from oracle.gen_golden_losses import TRACE_KINDS, TRACE_PER, _make_trace_agent, _run_trace  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("kind", TRACE_KINDS)
def test_agent_reproduces_the_reference_run(kind):
    import pfrl_b200
    from pfrl_b200.replay_buffers import PrioritizedReplayBuffer

    g = np.load(os.path.join(GOLD, "agent_trace_%s.npz" % kind))
    with mock.patch("pfrl_b200.replay_buffers.device_buffer.DeviceReplayStore", OracleBackedStore):
        rbuf = PrioritizedReplayBuffer(150, device=0, **TRACE_PER)
        torch.manual_seed(3)
        q, agent = _make_trace_agent(pfrl_b200, kind, rbuf)
        q.load_state_dict({k: torch.tensor(g["init_" + k]) for k in q.state_dict()})
        agent.target_model.load_state_dict(q.state_dict())
        np.random.seed(9)
        torch.manual_seed(9)

        def check(t, a):
            assert a == g["actions"][t].tolist(), "actions diverge at step %d" % t

        actions, stats = _run_trace(agent, rbuf, check=check)
        rbuf._flush()
        info = rbuf.store.info()
    assert np.array_equal(stats[:, 2:], g["stats"][:, 2:])           # n_updates, rlen
    live = stats[:, 2] > 0
    np.testing.assert_allclose(stats[live, 0], g["stats"][live, 0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(stats[live, 1], g["stats"][live, 1], rtol=1e-5, atol=1e-7)
    for k, v in q.state_dict().items():
        np.testing.assert_allclose(v.numpy(), g["final_" + k], rtol=1e-5, atol=1e-6, err_msg=k)
    # priorities are functions of fp32 TD errors: equal to fp32 round-off (bit-equal for ddqn)
    np.testing.assert_allclose(info["total"], float(g["total"]), rtol=1e-5)
    np.testing.assert_allclose(info["min"], float(g["min"]), rtol=1e-5)
    np.testing.assert_allclose(info["max_priority"], float(g["max_priority"]), rtol=1e-5)
    assert rbuf.beta == float(g["beta"])


from oracle.gen_golden_losses import (MORE_TRACE_KINDS, _make_more_agent, _module_attrs,  # noqa: E402
                                      _run_more_trace)


@pytest.mark.parametrize("kind", MORE_TRACE_KINDS)
def test_more_agents_reproduce_the_reference_run(kind):
    """IQN / SAC / TD3 / DDPG on the uniform device buffer (index stream of
    sample_n_k, capacity wrap-around) and PPO (dataset order, minibatch schedule,
    GAE, normaliser): actions and statistics at every step, final parameters."""
    import random

    import pfrl_b200
    from pfrl_b200.replay_buffers import ReplayBuffer

    g = np.load(os.path.join(GOLD, "agent_trace_%s.npz" % kind))
    with mock.patch("pfrl_b200.replay_buffers.device_buffer.DeviceReplayStore", OracleBackedStore):
        rbuf = None if kind == "ppo" else ReplayBuffer(150, device=0)
        torch.manual_seed(3)
        agent = _make_more_agent(pfrl_b200, kind, rbuf)
        for name, mod in _module_attrs(agent):
            mod.load_state_dict({k: torch.tensor(g["init_%s__%s" % (name, k)])
                                 for k in mod.state_dict()})
        assert [n for n, _ in agent.get_statistics()] == g["stat_names"].tolist()
        np.random.seed(9)
        torch.manual_seed(9)
        random.seed(9)
        discrete = g["actions"].dtype.kind in "iu"

        def check(t, a):
            if discrete:
                assert a.tolist() == g["actions"][t].tolist(), "actions diverge at step %d" % t
            else:
                np.testing.assert_allclose(a, g["actions"][t], rtol=1e-5, atol=2e-6,
                                           err_msg="actions diverge at step %d" % t)

        actions, stats = _run_more_trace(agent, kind, g["actions"].shape[0], check=check)
    want = g["stats"]
    both_nan = np.isnan(stats) & np.isnan(want)
    np.testing.assert_allclose(np.where(both_nan, 0.0, stats), np.where(both_nan, 0.0, want),
                               rtol=5e-5, atol=1e-6)   # measured: <= 1.2e-5 relative
    for name, mod in _module_attrs(agent):
        for k, v in mod.state_dict().items():
            # measured: bit-equal for IQN, <= 4e-7 absolute for the others
            np.testing.assert_allclose(v.numpy(), g["final_%s__%s" % (name, k)], rtol=1e-5,
                                       atol=2e-6, err_msg="%s.%s" % (name, k))


def test_whole_training_driver_reproduces_the_reference_run(tmp_path):
    """train_agent_batch_with_evaluation end to end (vector-env loop, per-env step
    counting, max_episode_len resets, evaluator schedule and episode accounting,
    scores.txt, saved directories) against the reference's seeded run
    (tests/golden/driver_trace.npz)."""
    import pfrl_b200
    from oracle.gen_golden_losses import _run_driver
    from pfrl_b200.replay_buffers import PrioritizedReplayBuffer

    g = np.load(os.path.join(GOLD, "driver_trace.npz"))
    with mock.patch("pfrl_b200.replay_buffers.device_buffer.DeviceReplayStore", OracleBackedStore):
        rbuf = PrioritizedReplayBuffer(150, device=0, **TRACE_PER)
        q, agent, history, rows = _run_driver(pfrl_b200, rbuf, str(tmp_path))
    header, body = rows[0], rows[1:]
    keep = [i for i, name in enumerate(header) if name != "elapsed"]
    assert [header[i] for i in keep] == g["header"].tolist()
    got = np.array([[float(r[i]) for i in keep] for r in body], dtype=np.float64)
    assert got.shape == g["scores"].shape
    np.testing.assert_allclose(got, g["scores"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose([h["eval_score"] for h in history], g["eval_scores"], rtol=1e-12)
    assert sorted(os.listdir(str(tmp_path))) == g["saved"].tolist()
    assert agent.t == int(g["t"]) and agent.optim_t == int(g["optim_t"])
    for k, v in q.state_dict().items():
        np.testing.assert_allclose(v.numpy(), g["final_" + k], rtol=1e-5, atol=1e-6, err_msg=k)
