"""The reference's own CPU implementation of the hot path, timed on host cores.

BENCH INFRASTRUCTURE ONLY (bench.py's `cpu_baseline` leg and `--impl reference`).
When the unmodified reference is importable -- /root/reference in the build container,
or the archive oracle/build_ref.py made of it (oracle/_ref/pfrl_ref.zip, which travels to
the GPU box) -- these functions drive pfnet/pfrl's OWN classes through its public API
(`kind = "reference"`); otherwise they fall back to the pure-Python port
(oracle/pyport.py, `kind = "port"`).

Replay path timed (one pass = one minibatch):
    PrioritizedReplayBuffer.sample(B)            pfrl/replay_buffers/prioritized.py:117-123
    batch_experiences(exps, cpu, phi=/255, g)    pfrl/replay_buffer.py:157-212
    PrioritizedReplayBuffer.update_errors(err)   pfrl/replay_buffers/prioritized.py:125-126
"""
import os
import time

import numpy as np

from . import refimport

FRAME = (84, 84)
STACK = 4


def kind():
    return "reference" if refimport.available() else "port"


def _threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def build_reference_buffer(capacity, n_step, alpha, beta0, normalize_by_max, pool=65536, seed=0):
    """A full reference PrioritizedReplayBuffer of `capacity` n-step experiences whose
    observations are LazyFrames over a pool of distinct 84x84 uint8 frames (the
    reference stores references to frame arrays, pfrl/wrappers/atari_wrappers.py:251-272)."""
    pfrl = refimport.import_reference()
    from pfrl.wrappers.atari_wrappers import LazyFrames

    rng = np.random.RandomState(seed)
    pool = min(pool, capacity + STACK + n_step)
    frames = rng.randint(0, 256, size=(pool, 1) + FRAME, dtype=np.uint8)
    flist = [frames[i] for i in range(pool)]
    buf = pfrl.replay_buffers.PrioritizedReplayBuffer(
        capacity, alpha=alpha, beta0=beta0, betasteps=None, num_steps=n_step,
        normalize_by_max=normalize_by_max)
    T = capacity + n_step - 1
    acts = rng.randint(0, 18, size=T)
    rews = rng.randint(-1, 2, size=T)
    t0 = time.perf_counter()
    prev = LazyFrames([flist[j % pool] for j in range(STACK)], stack_axis=0)
    for t in range(T):
        nxt = LazyFrames([flist[(t + 1 + j) % pool] for j in range(STACK)], stack_axis=0)
        buf.append(prev, int(acts[t]), float(rews[t]), nxt, None, False)
        prev = nxt
    fill_s = time.perf_counter() - t0
    # non-uniform priorities, like the GPU arm's prefill
    np.random.seed(seed)
    for _ in range(8):
        buf.sample(512)
        buf.update_errors([float(x) for x in rng.rand(512) * 2])
    return buf, flist, fill_s


def replay_run(capacity, batch, steps, warmup, seconds=None, n_step=3, alpha=0.5, beta0=0.4,
               normalize_by_max="memory", gamma=0.99, seed=0, buf=None):
    """Time `steps` passes (or as many as fit in `seconds`).  Returns a dict incl. the
    buffer (so the Rainbow loop can reuse the 1M fill)."""
    import torch

    if not refimport.available():
        raise ImportError("reference not importable")
    refimport.import_reference()
    from pfrl.replay_buffer import batch_experiences

    torch.set_num_threads(min(16, _threads()))  # collate only; the replay code is one thread
    fill_s = 0.0
    flist = None
    if buf is None:
        buf, flist, fill_s = build_reference_buffer(capacity, n_step, alpha, beta0,
                                                    normalize_by_max, seed=seed)
    phi = lambda x: np.asarray(x, dtype=np.float32) / 255  # noqa: E731
    dev = torch.device("cpu")
    rng = np.random.RandomState(seed + 1)
    np.random.seed(seed)

    def one():
        exps = buf.sample(batch)
        b = batch_experiences(exps, dev, phi, gamma)
        buf.update_errors([float(x) for x in np.abs(rng.randn(batch))])
        return b

    for _ in range(warmup):
        one()
    done = 0
    t0 = time.perf_counter()
    while done < steps:
        one()
        done += 1
        if seconds is not None and time.perf_counter() - t0 > seconds:
            break
    dt = time.perf_counter() - t0
    return {"samples_per_sec": done * batch / dt, "steps": done, "seconds": dt,
            "ms_per_step": 1e3 * dt / done, "setup_s": fill_s, "buffer": buf, "frames": flist,
            "kind": "reference", "cores": 1}


def rainbow_run(buf, flist, batch, seconds, num_envs=16, update_interval=4, gamma=0.99, seed=0):
    """The reference's CategoricalDoubleDQN (Rainbow) training loop on the host:
    batch_act -> (synthetic host env) -> batch_observe (append / sample / update), the
    loop of pfrl/experiments/train_agent_batch.py:65-141 without its bookkeeping."""
    import torch

    pfrl = refimport.import_reference()
    from pfrl.wrappers.atari_wrappers import LazyFrames

    threads = min(32, _threads())
    torch.set_num_threads(threads)
    torch.manual_seed(seed)
    q = pfrl.q_functions.DistributionalDuelingDQN(18, 51, -10, 10)
    pfrl.nn.to_factorized_noisy(q, sigma_scale=0.5)
    opt = torch.optim.Adam(q.parameters(), 6.25e-5, eps=1.5e-4)
    agent = pfrl.agents.CategoricalDoubleDQN(
        q, opt, buf, gpu=None, gamma=gamma, explorer=pfrl.explorers.Greedy(),
        minibatch_size=batch, replay_start_size=batch, target_update_interval=32000,
        update_interval=update_interval, batch_accumulator="mean",
        phi=lambda x: np.asarray(x, dtype=np.float32) / 255)
    rng = np.random.RandomState(seed)
    pool = len(flist)
    cur = [LazyFrames([flist[rng.randint(pool)]] * STACK, stack_axis=0) for _ in range(num_envs)]
    env_steps = 0

    def vec_step():
        nonlocal cur, env_steps
        acts = agent.batch_act(cur)
        nxt = [LazyFrames(c._frames[1:] + [flist[rng.randint(pool)]], stack_axis=0) for c in cur]
        rews = [float(rng.randint(-1, 2)) for _ in range(num_envs)]
        agent.batch_observe(nxt, rews, [False] * num_envs, [False] * num_envs)
        cur = nxt
        env_steps += num_envs
        return acts

    vec_step()  # warm-up
    n0 = agent.optim_t
    env_steps = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        vec_step()
    dt = time.perf_counter() - t0
    upd = agent.optim_t - n0
    return {"env_steps_per_sec": env_steps / dt, "updates_per_sec": upd / dt,
            "ms_per_update": 1e3 * dt / max(upd, 1), "seconds": dt, "num_envs": num_envs,
            "threads": threads, "kind": "reference"}
