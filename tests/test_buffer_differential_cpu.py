"""Randomised differential test of the device buffers' host logic against
the REAL reference buffers (build container only; skipped when
/root/reference is absent): random interleavings of append (several env ids,
terminals), stop_current_episode, sample and update_errors, with small
capacities so that eviction, n-step tails and the sample / update protocol all
interact.  Compared: lengths after every op, sampled batches (through each
side's batch_experiences), importance weights, and the trees' total at the
end.  The store is tests/fake_store.OracleBackedStore."""
import os
import sys
from unittest import mock

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from fake_store import OracleBackedStore  # noqa: E402

from oracle import refimport  # noqa: E402

pytestmark = pytest.mark.skipif(not refimport.available(), reason="reference tree not present")


def _run(seed, prioritized, num_steps, capacity, n_envs, n_ops):
    pfrl = refimport.import_reference()
    import pfrl_b200

    rng = np.random.RandomState(seed)
    if prioritized:
        kw = dict(alpha=0.7, beta0=0.5, betasteps=50, num_steps=num_steps,
                  normalize_by_max=[True, "memory", False][seed % 3])
        ref = pfrl.replay_buffers.PrioritizedReplayBuffer(capacity, **kw)
        raw = ref.update_errors
        ref.update_errors = lambda e: raw([float(x) for x in e])
        mine = pfrl_b200.replay_buffers.PrioritizedReplayBuffer(capacity, device=0, **kw)
    else:
        ref = pfrl.replay_buffers.ReplayBuffer(capacity, num_steps)
        mine = pfrl_b200.replay_buffers.ReplayBuffer(capacity, num_steps, device=0)
    phi = lambda x: np.asarray(x, dtype=np.float32)  # noqa: E731
    cur = [rng.randn(3).astype(np.float32) for _ in range(n_envs)]
    cpu = torch.device("cpu")
    for step in range(n_ops):
        op = rng.rand()
        if op < 0.7:
            e = rng.randint(n_envs)
            nxt = rng.randn(3).astype(np.float32)
            done = rng.rand() < 0.15
            args = (cur[e], int(rng.randint(4)), float(rng.randn()), nxt, None, bool(done))
            ref.append(*args, env_id=e)
            mine.append(*args, env_id=e)
            cur[e] = rng.randn(3).astype(np.float32) if done else nxt
            if done or rng.rand() < 0.05:
                ref.stop_current_episode(env_id=e)
                mine.stop_current_episode(env_id=e)
        elif op < 0.8:
            e = rng.randint(n_envs)
            ref.stop_current_episode(env_id=e)
            mine.stop_current_episode(env_id=e)
        elif len(ref) > 0:
            n = int(rng.randint(1, min(len(ref), 6) + 1))
            state = np.random.get_state()
            a = ref.sample(n)
            np.random.set_state(state)
            b = mine.sample(n)
            ba = pfrl.replay_buffer.batch_experiences(a, cpu, phi, 0.9)
            bb = pfrl_b200.replay_buffer.batch_experiences(b, cpu, pfrl_b200.utils.phi.Identity(), 0.9)
            for k in ("state", "next_state", "action", "discount", "is_state_terminal"):
                assert torch.equal(ba[k].float(), bb[k].float()), (seed, step, k)
            torch.testing.assert_close(ba["reward"], bb["reward"], rtol=1e-6, atol=1e-7)
            if prioritized:
                wa = np.asarray([x[0]["weight"] for x in a], dtype=np.float32)
                np.testing.assert_allclose(bb["weights"].numpy(), wa, rtol=2e-6)
                errs = [float(x) for x in np.abs(rng.randn(n)) * 2]
                ref.update_errors(errs)
                mine.update_errors(errs)
        assert len(ref) == len(mine), (seed, step)
    if prioritized and len(ref) > 0:
        mine._flush()
        info = mine.store.info()
        assert info["total"] == ref.memory.priority_sums.sum()
        assert info["max_priority"] == ref.memory.max_priority


@pytest.mark.parametrize("seed", range(12))
def test_random_interleavings_match_the_reference(seed):
    prioritized = seed % 2 == 0
    num_steps = [1, 2, 3, 5][seed % 4]
    capacity = [7, 16, 33, 50][(seed // 2) % 4]
    with mock.patch("pfrl_b200.replay_buffers.device_buffer.DeviceReplayStore", OracleBackedStore):
        _run(seed, prioritized, num_steps, capacity, n_envs=1 + seed % 3, n_ops=400)
