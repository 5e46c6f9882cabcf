"""A2C against a seeded run of the real reference (tests/golden/a2c_trace.npz,
oracle/gen_golden_losses.py:a2c_trace): identical sampled actions, running
statistics and final parameters within fp32 tolerance.  CPU only."""
import os

import numpy as np
import pytest
import torch
from torch import nn

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "a2c_trace.npz"))
N, OBS = 6, 9


def _model(kind):
    from pfrl_b200.nn import Branched
    from pfrl_b200.policies import (GaussianHeadWithStateIndependentCovariance,
                                    SoftmaxCategoricalHead)

    if kind == "discrete":
        head = nn.Sequential(nn.Linear(16, 4), SoftmaxCategoricalHead())
    else:
        head = nn.Sequential(nn.Linear(16, 3), GaussianHeadWithStateIndependentCovariance(
            action_size=3, var_type="diagonal", var_func=lambda x: torch.exp(2 * x),
            var_param_init=0))
    model = nn.Sequential(nn.Linear(OBS, 16), nn.Tanh(), Branched(head, nn.Linear(16, 1)))
    model.load_state_dict({k: torch.tensor(G["%s_init_%s" % (kind, k)])
                           for k in model.state_dict()})
    return model


@pytest.mark.parametrize("kind,kw", [
    ("discrete", dict(use_gae=False, max_grad_norm=0.5)),
    ("gaussian", dict(use_gae=True, tau=0.9, max_grad_norm=None)),
])
def test_a2c_matches_reference_trace(kind, kw):
    from pfrl_b200.agents import A2C

    model = _model(kind)
    opt = torch.optim.RMSprop(model.parameters(), lr=7e-3, eps=1e-5, alpha=0.99)
    agent = A2C(model, opt, gamma=0.97, num_processes=N, update_steps=4,
                average_actor_loss_decay=0.0, average_entropy_decay=0.0,
                average_value_decay=0.0, **kw)
    torch.manual_seed(77)
    steps = G["reward"].shape[0]
    for t in range(steps):
        a = agent.batch_act(list(G["obs"][t]))
        ref = G[kind + "_actions"][t]
        if kind == "discrete":
            assert np.array_equal(a, ref), t
        else:
            np.testing.assert_allclose(a, ref, rtol=1e-5, atol=1e-6)
        agent.batch_observe(list(G["obs"][t + 1]), list(G["reward"][t]), list(G["done"][t]),
                            [False] * N)
        got = [v for _, v in agent.get_statistics()]
        np.testing.assert_allclose(got, G[kind + "_stats"][t], rtol=1e-5, atol=1e-6)
    for k, v in model.state_dict().items():
        np.testing.assert_allclose(v.numpy(), G["%s_final_%s" % (kind, k)], rtol=1e-4, atol=1e-5)
    assert [n for n, _ in agent.get_statistics()] == ["average_actor", "average_value",
                                                      "average_entropy"]


def test_a2c_reset_counts_as_done_and_eval_mode():
    from pfrl_b200.agents import A2C

    model = _model("discrete")
    agent = A2C(model, torch.optim.SGD(model.parameters(), lr=0.0), gamma=0.9, num_processes=N,
                update_steps=3, act_deterministically=True)
    obs = list(G["obs"][0])
    agent.batch_act(obs)
    with pytest.warns(UserWarning):
        agent.batch_observe(obs, [0.0] * N, [False] * N, [True] + [False] * (N - 1))
    assert agent.window.alive[0].tolist() == [0.0] + [1.0] * (N - 1)
    t = agent.t
    with agent.eval_mode():
        a1 = agent.batch_act(obs)
        agent.batch_observe(obs, [0.0] * N, [False] * N, [False] * N)
        a2 = agent.batch_act(obs)
    assert agent.t == t and np.array_equal(a1, a2)


def test_a2c_save_load(tmp_path):
    from pfrl_b200.agents import A2C

    m1, m2 = _model("gaussian"), _model("discrete")
    a1 = A2C(m1, torch.optim.SGD(m1.parameters(), lr=0.1), 0.9, N)
    a1.save(str(tmp_path))
    assert sorted(os.listdir(tmp_path)) == ["model.pt", "optimizer.pt"]
    m3 = _model("gaussian")
    with torch.no_grad():
        for p in m3.parameters():
            p.add_(1.0)
    a3 = A2C(m3, torch.optim.SGD(m3.parameters(), lr=0.1), 0.9, N)
    a3.load(str(tmp_path))
    for p, q in zip(m1.parameters(), m3.parameters()):
        assert torch.equal(p, q)
    del m2


def test_a2c_learns_chain(tmp_path):
    from pfrl_b200 import experiments
    from pfrl_b200.agents import A2C
    from pfrl_b200.envs import ChainEnv, SerialVectorEnv
    from pfrl_b200.nn import Branched
    from pfrl_b200.policies import SoftmaxCategoricalHead
    from pfrl_b200.utils import set_random_seed

    set_random_seed(0)
    n_envs = 4
    model = nn.Sequential(nn.Linear(5, 32), nn.Tanh(), Branched(
        nn.Sequential(nn.Linear(32, 2), SoftmaxCategoricalHead()), nn.Linear(32, 1)))
    agent = A2C(model, torch.optim.Adam(model.parameters(), lr=1e-2), gamma=0.95,
                num_processes=n_envs, update_steps=5, use_gae=True, tau=0.95,
                act_deterministically=True, max_grad_norm=1.0)
    # no time-limit resets: A2C folds resets into terminals (a2c.py:274-284)
    env = SerialVectorEnv([ChainEnv(seed=i, max_steps=10 ** 9) for i in range(n_envs)])
    experiments.train_agent_batch(agent, env, 6000, str(tmp_path), log_interval=None)
    eval_env = ChainEnv(max_steps=10 ** 9)
    with agent.eval_mode():
        obs, total = eval_env.reset(), 0.0
        for _ in range(30):
            obs, r, done, _ = eval_env.step(agent.act(obs))
            total += r
            if done:
                break
    assert total > 0.9
    assert all(np.isfinite(v) for _, v in agent.get_statistics())
