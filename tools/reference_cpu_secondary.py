"""The REAL reference (pfnet/pfrl, imported from /root/reference through
oracle/refimport.py) on the secondary workloads of tools/bench_secondary.py,
on the host CPU: same networks, same hyper-parameters, same synthetic envs
(host numpy variants), same act / observe loop.  Build-container tool (the
reference does not exist on the GPU box); prints one JSON line per workload.

    python tools/reference_cpu_secondary.py
"""
import json
import os
import sys
import time

import numpy as np
import torch
from torch import distributions, nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.refimport import import_reference  # noqa: E402

pfrl = import_reference()
from pfrl_b200.envs import SyntheticAtariVectorEnv, SyntheticContinuousVectorEnv  # noqa: E402

CORES = len(os.sched_getaffinity(0))
torch.set_num_threads(CORES)


def loop(agent, env, steps):
    obs = env.reset()
    for _ in range(steps):
        a = agent.batch_act(obs)
        obs, r, d, info = env.step(a)
        agent.batch_observe(obs, r, d, np.zeros(env.num_envs, dtype=bool))
        obs = env.reset(np.logical_not(d))


def timed(agent, env, steps, warm):
    loop(agent, env, warm)
    t0 = time.perf_counter()
    loop(agent, env, steps)
    return time.perf_counter() - t0


def dqn():
    q = nn.Sequential(pfrl.nn.LargeAtariCNN(), nn.Linear(512, 18),
                      pfrl.q_functions.DiscreteActionValueHead())
    rbuf = pfrl.replay_buffers.PrioritizedReplayBuffer(200000, alpha=0.6, beta0=0.4,
                                                       betasteps=10 ** 6, num_steps=1)
    agent = pfrl.agents.DQN(
        q, torch.optim.RMSprop(q.parameters(), lr=2.5e-4, alpha=0.95, eps=1e-2), rbuf, 0.99,
        pfrl.explorers.ConstantEpsilonGreedy(0.1, lambda: np.random.randint(18)), gpu=None,
        replay_start_size=1000, minibatch_size=32, update_interval=4,
        target_update_interval=10000, phi=lambda x: np.asarray(x, dtype=np.float32) / 255)
    env = SyntheticAtariVectorEnv(16, device="cpu", seed=0)
    loop(agent, env, 1000 // 16 + 8)
    steps = 30
    n0 = agent.optim_t
    dt = timed(agent, env, steps, 2)
    print(json.dumps({"impl": "reference (CPU, %d cores)" % CORES,
                      "workload": "DQN configs[1]: Nature CNN, PER, batch 32, update_interval 4, "
                                  "16 host envs", "env_steps_per_sec": steps * 16 / dt,
                      "updates_per_sec": (agent.optim_t - n0) / dt}), flush=True)


def ppo():
    obs_dim, act_dim, E, T = 376, 17, 256, 8
    model = nn.Sequential(
        pfrl.nn.Branched(
            nn.Sequential(nn.Linear(obs_dim, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh(),
                          nn.Linear(64, act_dim),
                          pfrl.policies.GaussianHeadWithStateIndependentCovariance(
                              action_size=act_dim, var_type="diagonal",
                              var_func=lambda x: torch.exp(2 * x), var_param_init=0)),
            nn.Sequential(nn.Linear(obs_dim, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh(),
                          nn.Linear(64, 1))))
    agent = pfrl.agents.PPO(
        model, torch.optim.Adam(model.parameters(), lr=3e-4, eps=1e-5),
        obs_normalizer=pfrl.nn.EmpiricalNormalization(obs_dim, clip_threshold=5), gpu=None,
        gamma=0.995, lambd=0.95, update_interval=E * T, minibatch_size=64, epochs=10,
        clip_eps=0.2, clip_eps_vf=None, entropy_coef=0.0)
    env = SyntheticContinuousVectorEnv(E, obs_dim, act_dim, device="cpu", seed=0)
    steps = 2 * T
    dt = timed(agent, env, steps, T)
    print(json.dumps({"impl": "reference (CPU, %d cores)" % CORES,
                      "workload": "PPO configs[3]: obs 376 act 17, 256 envs, T=8 (2048/update), "
                                  "minibatch 64 x 10 epochs", "env_steps_per_sec": steps * E / dt,
                      "updates": agent.n_updates, "seconds": dt}), flush=True)


def sac():
    obs_dim, act_dim, E = 17, 6, 16

    def squashed(x):
        mean, log_scale = torch.chunk(x, 2, dim=1)
        base = distributions.Independent(
            distributions.Normal(mean, torch.exp(torch.clamp(log_scale, -20, 2))), 1)
        return distributions.transformed_distribution.TransformedDistribution(
            base, [distributions.transforms.TanhTransform(cache_size=1)])

    policy = nn.Sequential(nn.Linear(obs_dim, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(),
                           nn.Linear(256, 2 * act_dim), pfrl.nn.lmbda.Lambda(squashed))

    def qf():
        return nn.Sequential(pfrl.nn.ConcatObsAndAction(), nn.Linear(obs_dim + act_dim, 256),
                             nn.ReLU(), nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 1))

    q1, q2 = qf(), qf()
    agent = pfrl.agents.SoftActorCritic(
        policy, q1, q2, torch.optim.Adam(policy.parameters(), lr=3e-4),
        torch.optim.Adam(q1.parameters(), lr=3e-4), torch.optim.Adam(q2.parameters(), lr=3e-4),
        pfrl.replay_buffers.ReplayBuffer(10 ** 6), gamma=0.99, gpu=None, replay_start_size=2048,
        minibatch_size=1024, entropy_target=-act_dim, temperature_optimizer_lr=3e-4)
    env = SyntheticContinuousVectorEnv(E, obs_dim, act_dim, device="cpu", seed=1)
    loop(agent, env, 2048 // E + 4)
    steps = 20
    dt = timed(agent, env, steps, 2)
    print(json.dumps({"impl": "reference (CPU, %d cores)" % CORES,
                      "workload": "SAC configs[4]: obs 17 act 6, 1M uniform replay, batch 1024, "
                                  "update every env step", "env_steps_per_sec": steps * E / dt,
                      "updates_per_sec": steps * E / dt, "seconds": dt}), flush=True)


if __name__ == "__main__":
    dqn()
    ppo()
    sac()
