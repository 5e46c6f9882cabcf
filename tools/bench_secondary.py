"""Secondary workloads (BASELINE configs[1], [3], [4]) on one GPU: Nature-DQN
with PER at batch 32 (configs[1]), PPO on a
synthetic Humanoid-shaped env (obs 376, act 17, 256 envs, lambda .95) and SAC
(obs 17, act 6, 1M replay, batch 1024).  Prints one JSON line per workload.
Each can run eagerly or with its update captured as a CUDA graph.  bench.py folds
`run_all()` into its JSON line ("secondary"); stand-alone:
python tools/bench_secondary.py  (GPU box)."""
import json
import os
import sys
import time

import numpy as np
import torch
from torch import distributions, nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfrl_b200 import agents, nn as pnn, policies  # noqa: E402
from pfrl_b200.envs import SyntheticContinuousVectorEnv  # noqa: E402
from pfrl_b200.replay_buffers import ReplayBuffer  # noqa: E402
from pfrl_b200.utils.phi import Identity  # noqa: E402


def loop(agent, env, steps):
    obs = env.reset()
    for _ in range(steps):
        a = agent.batch_act(obs)
        obs, r, d, info = env.step(a)
        agent.batch_observe(obs, r, d, np.zeros(env.num_envs, dtype=bool))
        obs = env.reset(np.logical_not(d))


def timed(agent, env, steps, warm):
    loop(agent, env, warm)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loop(agent, env, steps)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


def dqn():
    """configs[1]: DQN, synthetic 84x84x4 frames, PER 1M capacity (prefilled), batch 32."""
    from pfrl_b200 import explorers, q_functions
    from pfrl_b200.envs import SyntheticAtariVectorEnv
    from pfrl_b200.replay_buffers import PrioritizedReplayBuffer
    from pfrl_b200.utils.phi import ScaleU8

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    q = nn.Sequential(pnn.LargeAtariCNN(), nn.Linear(512, 18),
                      q_functions.DiscreteActionValueHead()).cuda()
    cap = 10 ** 6
    rbuf = PrioritizedReplayBuffer(cap, alpha=0.6, beta0=0.4, betasteps=10 ** 6, num_steps=1,
                                   part_capacity=cap + 16384)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    rng = np.random.RandomState(5)
    done = 0
    while done < cap:  # prefill to capacity (untimed), like bench.py's shard
        m = min(1 << 17, cap - done)
        fr = torch.randint(0, 256, (m + 4, 84, 84), dtype=torch.uint8, device="cuda", generator=gen)
        tm = rng.rand(m) < 1e-3
        tm[-1] = True
        rbuf.append_trajectory(fr, rng.randint(0, 18, size=m).astype(np.int64),
                               rng.randint(-1, 2, size=m).astype(np.float64), tm)
        done += m
    agent = agents.DQN(q, torch.optim.RMSprop(q.parameters(), lr=2.5e-4, alpha=0.95, eps=1e-2),
                       rbuf, 0.99, explorers.ConstantEpsilonGreedy(0.1, lambda: np.random.randint(18)),
                       gpu=0, replay_start_size=1000, minibatch_size=32, update_interval=4,
                       target_update_interval=10000, phi=ScaleU8(), cuda_graph=GRAPH)
    env = SyntheticAtariVectorEnv(16, device="cuda", seed=0)
    loop(agent, env, 1000 // 16 + 8)
    steps = 250
    n0 = agent.optim_t
    dt = timed(agent, env, steps, 4)
    return {"workload": "DQN configs[1]: Nature CNN, PER 1M cap (full), batch 32, "
            "update_interval 4, 16 GPU envs", "env_steps_per_sec": steps * 16 / dt,
            "updates_per_sec": (agent.optim_t - n0) / dt, "seconds": dt, "cuda_graph": GRAPH}


def ppo():
    obs_dim, act_dim, E, T = 376, 17, 256, 8
    model = nn.Sequential(
        pnn.Branched(
            nn.Sequential(nn.Linear(obs_dim, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh(),
                          nn.Linear(64, act_dim),
                          policies.GaussianHeadWithStateIndependentCovariance(
                              action_size=act_dim, var_type="diagonal",
                              var_func=lambda x: torch.exp(2 * x), var_param_init=0)),
            nn.Sequential(nn.Linear(obs_dim, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh(),
                          nn.Linear(64, 1))))
    agent = agents.PPO(model, torch.optim.Adam(model.parameters(), lr=3e-4, eps=1e-5),
                       obs_normalizer=pnn.EmpiricalNormalization(obs_dim, clip_threshold=5),
                       gpu=0, gamma=0.995, lambd=0.95, update_interval=E * T, minibatch_size=64,
                       epochs=10, clip_eps=0.2, clip_eps_vf=None, entropy_coef=0.0,
                       cuda_graph=GRAPH)
    env = SyntheticContinuousVectorEnv(E, obs_dim, act_dim, device="cuda", seed=0)
    steps = (12 if GRAPH else 4) * T
    dt = timed(agent, env, steps, T)
    return {"workload": "PPO configs[3]: obs 376 act 17, 256 envs, T=8 (2048/update), "
            "minibatch 64 x 10 epochs, lambda .95", "env_steps_per_sec": steps * E / dt,
            "updates": agent.n_updates, "seconds": dt, "cuda_graph": GRAPH}


def sac():
    obs_dim, act_dim, E = 17, 6, 16

    def squashed(x):
        mean, log_scale = torch.chunk(x, 2, dim=1)
        base = distributions.Independent(
            distributions.Normal(mean, torch.exp(torch.clamp(log_scale, -20, 2))), 1)
        return distributions.transformed_distribution.TransformedDistribution(
            base, [distributions.transforms.TanhTransform(cache_size=1)])

    policy = nn.Sequential(nn.Linear(obs_dim, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(),
                           nn.Linear(256, 2 * act_dim), pnn.Lambda(squashed))

    def qf():
        return nn.Sequential(pnn.ConcatObsAndAction(), nn.Linear(obs_dim + act_dim, 256), nn.ReLU(),
                             nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 1))

    q1, q2 = qf(), qf()
    agent = agents.SoftActorCritic(
        policy, q1, q2, torch.optim.Adam(policy.parameters(), lr=3e-4),
        torch.optim.Adam(q1.parameters(), lr=3e-4), torch.optim.Adam(q2.parameters(), lr=3e-4),
        ReplayBuffer(10 ** 6), gamma=0.99, gpu=0, replay_start_size=2048, minibatch_size=1024,
        entropy_target=-act_dim, temperature_optimizer_lr=3e-4, phi=Identity(), cuda_graph=GRAPH)
    env = SyntheticContinuousVectorEnv(E, obs_dim, act_dim, device="cuda", seed=1)
    loop(agent, env, 2048 // E + 4)
    steps = 100 if GRAPH else 30
    dt = timed(agent, env, steps, 4)
    return {"workload": "SAC configs[4]: obs 17 act 6, 1M uniform replay, batch 1024, "
            "update every env step", "env_steps_per_sec": steps * E / dt,
            "updates_per_sec": steps * E / dt, "seconds": dt, "cuda_graph": GRAPH}


GRAPH = False


def run_all(graph=True):
    """{"c2_dqn": ..., "c4_ppo": ..., "c5_sac": ...} for bench.py's "secondary" key."""
    global GRAPH
    GRAPH = bool(graph)
    out = {}
    for key, fn in (("c2_dqn", dqn), ("c4_ppo", ppo), ("c5_sac", sac)):
        out[key] = fn()
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    for g in (False, True):
        for v in run_all(g).values():
            print(json.dumps(v), flush=True)
