"""The device buffers' HOST logic (n-step windows, stop_current_episode,
LazyFrames de-duplication into the part ring, staged appends, eviction
bookkeeping, weights / beta schedule, update_errors' priority formula) run on
the reference's golden traces without a GPU: DeviceReplayStore is replaced by
tests/fake_store.OracleBackedStore, whose trees are the CPU oracle.  The same
traces run against the CUDA store in tests/test_replay_buffers_gpu.py."""
import os
import sys
from unittest import mock

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
import golden_replay as gr  # noqa: E402
from fake_store import OracleBackedStore  # noqa: E402

PATCH = "pfrl_b200.replay_buffers.device_buffer.DeviceReplayStore"


def _batch(exps, gamma):
    from pfrl_b200.replay_buffer import batch_experiences
    from pfrl_b200.utils.phi import ScaleU8

    b = batch_experiences(exps, torch.device("cpu"), ScaleU8(), gamma)
    return {k: v.numpy() for k, v in b.items()}


@pytest.mark.parametrize("name", ["per_trace_1step", "per_trace_3step_memory",
                                  "per_trace_lazyframes"])
def test_prioritized_host_logic_on_reference_trace(name):
    from pfrl_b200.replay_buffers import PrioritizedReplayBuffer
    from pfrl_b200.utils.lazy_frames import LazyFrames

    g = gr.load(name)
    with mock.patch(PATCH, OracleBackedStore):
        buf = gr.replay_per_trace(
            g, lambda **kw: PrioritizedReplayBuffer(device=0, **kw), _batch, LazyFrames,
            indices_of=lambda b, e: e.index.numpy())
        buf._flush()
        info = buf.store.info()
    assert info["max_priority"] == float(g["final_max_priority"])
    assert info["total"] == float(g["final_total"])
    assert info["min"] == float(g["final_min"])
    if name == "per_trace_lazyframes":
        # every frame uploaded once: parts ~ appends + 3 per episode start, not 8 per append
        n_app = int((g["ops"][:, 0] == gr.OP_APPEND).sum())
        assert buf.store.part_head < 2 * n_app


@pytest.mark.parametrize("name", ["uniform_trace_sac", "uniform_trace_3step"])
def test_uniform_host_logic_on_reference_trace(name):
    from pfrl_b200.replay_buffer import batch_experiences
    from pfrl_b200.replay_buffers import ReplayBuffer
    from pfrl_b200.utils.phi import Identity

    g = gr.load(name)
    seed, capacity, num_steps, steps, batch = [int(x) for x in g["meta"]]
    gamma = float(g["gamma"])
    with mock.patch(PATCH, OracleBackedStore):
        buf = ReplayBuffer(capacity, num_steps=num_steps, device=0)
        np.random.seed(seed)
        sample_at = set(int(t) for t in g["sample_at"])
        off = 0
        for t in range(steps):
            buf.append(g["obs"][t], g["acts"][t], float(g["rews"][t]), g["obs"][t + 1], None,
                       bool(g["terms"][t]))
            if g["terms"][t]:
                buf.stop_current_episode()
            if t in sample_at:
                b = batch_experiences(buf.sample(batch), torch.device("cpu"), Identity(), gamma)
                b = {k: v.numpy() for k, v in b.items()}
                sl = slice(off, off + batch)
                assert np.array_equal(b["state"], g["state"][sl])
                assert np.array_equal(b["next_state"], g["next_state"][sl])
                assert np.array_equal(b["action"], g["action"][sl])
                np.testing.assert_allclose(b["reward"], g["reward"][sl], rtol=1e-6, atol=1e-7)
                assert np.array_equal(b["discount"], g["discount"][sl])
                assert np.array_equal(b["is_state_terminal"], g["terminal"][sl])
                off += batch
    assert off == len(g["reward"])
