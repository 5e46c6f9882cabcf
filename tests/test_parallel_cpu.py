"""CPU, world_size 2 over gloo: the data-parallel plumbing (gradient bucket
all-reduce, parameter broadcast, PPO advantage-moment sync)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out, n_step=1, steps=40):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from pfrl_b200 import agents, explorers, parallel, q_functions
    from pfrl_b200.envs import ChainEnv, SerialVectorEnv
    from pfrl_b200.replay_buffers import HostReplayBuffer

    parallel.init(backend="gloo")
    torch.manual_seed(100 + rank)  # different init per rank on purpose
    np.random.seed(rank)
    q = q_functions.FCStateQFunctionWithDiscreteAction(5, 2, 16, 1)
    parallel.broadcast_parameters(q)
    agent = agents.DoubleDQN(
        q, torch.optim.SGD(q.parameters(), lr=0.05), HostReplayBuffer(1000, num_steps=n_step), 0.9,
        explorers.ConstantEpsilonGreedy(0.3, lambda: np.random.randint(2)),
        replay_start_size=20, minibatch_size=8, target_update_interval=10,
        phi=lambda x: x.astype(np.float32, copy=False), grad_sync=parallel.GradSync())
    env = SerialVectorEnv([ChainEnv(seed=10 * rank + i) for i in range(2)])  # own env shard
    obs = env.reset()
    for _ in range(steps):
        a = agent.batch_act(obs)
        obs, r, d, info = env.step(a)
        resets = [i["needs_reset"] for i in info]
        agent.batch_observe(obs, r, d, resets)
        obs = env.reset(np.logical_not(np.logical_or(d, resets)))
    flat = torch.cat([p.detach().reshape(-1) for p in agent.model.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    # the advantage moments
    adv = torch.arange(6, dtype=torch.float32) + 10 * rank
    stats = parallel.sync_advantage_stats(adv)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([agent.optim_t], dtype=torch.int64))
    if rank == 0:
        out.put((agent.optim_t, [g.numpy() for g in gathered], stats.numpy(),
                 [int(c) for c in counts]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gradient_allreduce_keeps_replicas_identical():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    optim_t, params, stats, counts = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert optim_t > 10
    # different data per rank, averaged gradients: replicas must stay bit-identical
    assert np.array_equal(params[0], params[1])
    both = np.concatenate([np.arange(6), np.arange(6) + 10]).astype(np.float64)
    np.testing.assert_allclose(stats, [both.mean(), both.std()], rtol=1e-6)


@pytest.mark.timeout(300)
def test_ranks_update_at_the_same_steps_with_nstep_windows():
    """3-step returns: each rank's buffer fills at its own pace (episode boundaries
    differ per shard), so the ranks cross replay_start_size at different steps; the
    update decision is collective, hence the same number of optimizer steps everywhere
    and no unmatched all-reduce at the end (parallel.py)."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out, 3, 60)) for r in range(2)]
    for p in procs:
        p.start()
    optim_t, params, stats, counts = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert counts[0] == counts[1] and counts[0] > 10
    assert np.array_equal(params[0], params[1])
