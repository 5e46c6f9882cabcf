"""CPU: host-side agent logic (schedules, losses in their torch formulation,
save/load, the training loops) on a toy env with the host replay buffer."""
import os
import random
from unittest import mock

import numpy as np
import pytest
import torch
from torch import nn

from pfrl_b200 import agents, experiments, explorers, nn as pnn, policies, q_functions
from pfrl_b200.envs import ChainEnv, SerialVectorEnv
from pfrl_b200.replay_buffers import HostReplayBuffer
from pfrl_b200.utils import set_random_seed


def make_dqn(cls=agents.DoubleDQN, **kw):
    q = q_functions.FCStateQFunctionWithDiscreteAction(5, 2, 32, 2)
    opt = torch.optim.Adam(q.parameters(), lr=3e-3)
    ex = explorers.LinearDecayEpsilonGreedy(1.0, 0.05, 600, lambda: np.random.randint(2))
    return cls(q, opt, HostReplayBuffer(5000), 0.95, ex, replay_start_size=50,
               minibatch_size=32, update_interval=1, target_update_interval=50,
               phi=lambda x: x.astype(np.float32, copy=False), **kw)


def greedy_return(agent, continuous=False):
    env = ChainEnv(continuous=continuous)
    with agent.eval_mode():
        obs, total = env.reset(), 0.0
        for _ in range(30):
            obs, r, done, info = env.step(agent.act(obs))
            total += r
            agent.observe(obs, r, done, False)
            if done:
                break
    return total


@pytest.mark.parametrize("cls", [agents.DQN, agents.DoubleDQN])
def test_dqn_learns_chain_batch(cls, tmp_path):
    set_random_seed(0)
    agent = make_dqn(cls)
    env = SerialVectorEnv([ChainEnv(seed=i) for i in range(2)])
    experiments.train_agent_batch(agent, env, 2500, str(tmp_path), log_interval=None)
    assert greedy_return(agent) > 0.9
    stats = dict(agent.get_statistics())
    assert stats["n_updates"] > 1000 and np.isfinite(stats["average_loss"])
    assert os.path.exists(os.path.join(str(tmp_path), "2500_finish", "model.pt"))
    # save / load round trip (agent.py:81-137 layout)
    other = make_dqn(cls)
    other.load(os.path.join(str(tmp_path), "2500_finish"))
    for a, b in zip(agent.model.parameters(), other.model.parameters()):
        assert torch.equal(a, b)


def test_single_env_quickstart_loop(tmp_path):
    """configs[0]: DoubleDQN, uniform replay, 1 env, CPU only."""
    set_random_seed(1)
    agent = make_dqn(agents.DoubleDQN)
    env = ChainEnv()
    experiments.train_agent(agent, env, 1500, str(tmp_path))
    assert greedy_return(agent) > 0.9


def test_categorical_dqn_runs_and_projection_matches_reference_kat():
    import os as _os

    G = np.load(_os.path.join(_os.path.dirname(__file__), "golden", "losses.npz"))
    from pfrl_b200.agents.categorical_dqn import _apply_categorical_projection

    out = _apply_categorical_projection(torch.tensor(G["proj_y"]), torch.tensor(G["proj_p"]),
                                        torch.tensor(G["proj_z"]))
    np.testing.assert_allclose(out.numpy(), G["proj_out"], rtol=1e-6, atol=1e-7)
    set_random_seed(0)
    q = q_functions.DistributionalFCStateQFunctionWithDiscreteAction(5, 2, 21, -1, 2, 32, 2)
    opt = torch.optim.Adam(q.parameters(), lr=3e-3)
    ex = explorers.ConstantEpsilonGreedy(0.3, lambda: np.random.randint(2))
    agent = agents.CategoricalDoubleDQN(
        q, opt, HostReplayBuffer(2000, num_steps=3), 0.95, ex, replay_start_size=40,
        minibatch_size=16, target_update_interval=40,
        phi=lambda x: x.astype(np.float32, copy=False))
    env = SerialVectorEnv([ChainEnv(seed=i) for i in range(2)])
    obs = env.reset()
    for _ in range(150):
        a = agent.batch_act(obs)
        obs, r, d, info = env.step(a)
        resets = [i["needs_reset"] for i in info]
        agent.batch_observe(obs, r, d, resets)
        obs = env.reset(np.logical_not(np.logical_or(d, resets)))
    assert agent.optim_t > 50 and np.isfinite(dict(agent.get_statistics())["average_loss"])


def test_ppo_learns_chain():
    set_random_seed(0)
    model = nn.Sequential(
        nn.Linear(5, 32), nn.Tanh(),
        pnn.Branched(nn.Sequential(nn.Linear(32, 2), policies.SoftmaxCategoricalHead()),
                     nn.Linear(32, 1)))
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    agent = agents.PPO(model, opt, gamma=0.95, lambd=0.9, update_interval=128,
                       minibatch_size=32, epochs=4, entropy_coef=0.0,
                       phi=lambda x: x.astype(np.float32, copy=False))
    env = SerialVectorEnv([ChainEnv(seed=i) for i in range(4)])
    obs = env.reset()
    for _ in range(400):
        a = agent.batch_act(obs)
        obs, r, d, info = env.step(a)
        resets = [i["needs_reset"] for i in info]
        agent.batch_observe(obs, r, d, resets)
        obs = env.reset(np.logical_not(np.logical_or(d, resets)))
    assert agent.n_updates > 0
    agent.act_deterministically = True
    assert greedy_return(agent) > 0.9
    stats = dict(agent.get_statistics())
    assert set(stats) == {"average_value", "average_entropy", "average_value_loss",
                          "average_policy_loss", "n_updates", "explained_variance"}


def test_ppo_minibatch_schedule_matches_reference_shape():
    from pfrl_b200.agents.ppo import _yield_minibatch_indices

    random.seed(0)
    order = list(range(10))
    mbs = list(_yield_minibatch_indices(order, 4, 3))
    assert len(mbs) == 8 and all(len(m) == 4 for m in mbs)  # ceil(30 / 4)
    flat = [i for m in mbs for i in m]
    assert all(flat.count(i) >= 3 for i in range(10))
    # identical stream consumption to `random.sample(dataset, k=len(dataset))`
    random.seed(0)
    first = random.sample(order, k=10)
    assert mbs[0] == first[-4:]


def test_sac_runs():
    set_random_seed(0)
    from torch import distributions

    def squashed(x):
        mean, log_scale = torch.chunk(x, 2, dim=1)
        base = distributions.Independent(
            distributions.Normal(mean, torch.exp(torch.clamp(log_scale, -5, 2))), 1)
        return distributions.transformed_distribution.TransformedDistribution(
            base, [distributions.transforms.TanhTransform(cache_size=1)])

    policy = nn.Sequential(nn.Linear(5, 32), nn.ReLU(), nn.Linear(32, 2), pnn.Lambda(squashed))

    def qf():
        return nn.Sequential(pnn.ConcatObsAndAction(), nn.Linear(6, 32), nn.ReLU(),
                             nn.Linear(32, 1))

    q1, q2 = qf(), qf()
    agent = agents.SoftActorCritic(
        policy, q1, q2, torch.optim.Adam(policy.parameters(), lr=3e-3),
        torch.optim.Adam(q1.parameters(), lr=3e-3), torch.optim.Adam(q2.parameters(), lr=3e-3),
        HostReplayBuffer(5000), gamma=0.95, replay_start_size=64, minibatch_size=32,
        entropy_target=-1.0, temperature_optimizer_lr=3e-3,
        phi=lambda x: x.astype(np.float32, copy=False),
        burnin_action_func=lambda: np.random.uniform(-1, 1, size=1).astype(np.float32))
    env = SerialVectorEnv([ChainEnv(continuous=True, seed=i) for i in range(2)])
    obs = env.reset()
    for _ in range(120):
        a = agent.batch_act(obs)
        obs, r, d, info = env.step(a)
        resets = [i["needs_reset"] for i in info]
        agent.batch_observe(obs, r, d, resets)
        obs = env.reset(np.logical_not(np.logical_or(d, resets)))
    stats = dict(agent.get_statistics())
    assert stats["n_updates"] > 50 and np.isfinite(stats["average_q1"])


def test_train_agent_batch_call_contract(tmp_path):
    """Call counts pinned by the reference's
    tests/experiments_tests/test_train_agent_batch.py:11-138."""
    steps, num_envs = 7, 2
    agent = mock.Mock()
    agent.batch_act.side_effect = lambda obs: [0] * num_envs
    agent.get_statistics.return_value = []

    class Env:
        def __init__(self):
            self.num_envs = num_envs
            self.t = 0
            self.resets = []

        def reset(self, mask=None):
            self.resets.append(None if mask is None else list(mask))
            return [np.zeros(1)] * num_envs

        def step(self, actions):
            self.t += 1
            done = [self.t == 2, False]
            return [np.zeros(1)] * num_envs, [1.0, 0.5], done, [{}, {"needs_reset": self.t == 3}]

        def close(self):
            pass

    env = Env()
    hook = mock.Mock()
    experiments.train_agent_batch(agent, env, steps, str(tmp_path), step_hooks=[hook])
    # ceil(7 / 2) = 4 vector steps
    assert agent.batch_act.call_count == 4
    assert agent.batch_observe.call_count == 4
    assert hook.call_count == 8  # one call per env step, t = 1..8
    assert [c[0][2] for c in hook.call_args_list] == list(range(1, 9))
    # resets: initial + after every step but the last; done env 0 at t=2, reset env 1 at t=3
    assert env.resets[0] is None and len(env.resets) == 4
    assert env.resets[2] == [False, True] and env.resets[3] == [True, False]
    assert agent.save.call_count == 1
    args = agent.batch_observe.call_args_list[2][0]
    assert list(args[3]) == [False, True]


@pytest.mark.parametrize("kind", ["td3", "ddpg"])
def test_td3_and_ddpg_learn_chain(kind, tmp_path):
    set_random_seed(0)

    def pol():
        return nn.Sequential(nn.Linear(5, 32), nn.ReLU(), nn.Linear(32, 1), nn.Tanh(),
                             policies.DeterministicHead())

    def qf():
        return nn.Sequential(pnn.ConcatObsAndAction(), nn.Linear(6, 32), nn.ReLU(),
                             nn.Linear(32, 1))

    ex = explorers.AdditiveGaussian(scale=0.3, low=-1, high=1)
    phi = lambda x: x.astype(np.float32, copy=False)  # noqa: E731
    burn = lambda: np.random.uniform(-1, 1, size=1).astype(np.float32)  # noqa: E731
    if kind == "td3":
        p, q1, q2 = pol(), qf(), qf()
        agent = agents.TD3(p, q1, q2, torch.optim.Adam(p.parameters(), lr=3e-3),
                           torch.optim.Adam(q1.parameters(), lr=3e-3),
                           torch.optim.Adam(q2.parameters(), lr=3e-3), HostReplayBuffer(5000),
                           0.95, ex, replay_start_size=64, minibatch_size=32, phi=phi,
                           burnin_action_func=burn)
    else:
        p, q = pol(), qf()
        agent = agents.DDPG(p, q, torch.optim.Adam(p.parameters(), lr=3e-3),
                            torch.optim.Adam(q.parameters(), lr=3e-3), HostReplayBuffer(5000),
                            0.95, ex, replay_start_size=64, minibatch_size=32, phi=phi,
                            target_update_method="soft", target_update_interval=1,
                            soft_update_tau=0.05, burnin_action_func=burn)
    env = SerialVectorEnv([ChainEnv(continuous=True, seed=i) for i in range(2)])
    experiments.train_agent_batch(agent, env, 1500, str(tmp_path))
    assert greedy_return(agent, continuous=True) > 0.9
    stats = dict(agent.get_statistics())
    assert all(np.isfinite(v) for v in stats.values())
    agent.save(str(tmp_path / "ckpt"))
    assert os.path.exists(str(tmp_path / "ckpt" / ("policy.pt" if kind == "td3" else "model.pt")))
