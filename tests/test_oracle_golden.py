"""CPU: the oracle (oracle/per_oracle.c + oracle/replay.py) reproduces what
the real reference produced for the committed scripted traces."""
import numpy as np
import pytest

import golden_replay as gr
from oracle.replay import (OraclePrioritizedReplayBuffer, OracleReplayBuffer,
                           batch_experiences_np)
from pfrl_b200.utils.lazy_frames import LazyFrames

PER_TRACES = ["per_trace_1step", "per_trace_3step_memory", "per_trace_lazyframes"]


def _oracle_batch(exps, gamma):
    phi = lambda x: np.asarray(x, dtype=np.float32) / 255  # noqa: E731
    b = batch_experiences_np(exps, phi, gamma)
    b["weights"] = np.asarray([e[0]["weight"] for e in exps], dtype=np.float32)
    return b


@pytest.mark.parametrize("name", PER_TRACES)
def test_oracle_per_matches_reference_golden(name):
    g = gr.load(name)
    buf = gr.replay_per_trace(
        g, lambda **kw: OraclePrioritizedReplayBuffer(**kw), _oracle_batch, LazyFrames,
        indices_of=lambda b, e: b.memory.sampled_indices or b._last_idx, rtol=1e-7)
    assert buf.memory.max_priority == float(g["final_max_priority"])
    assert buf.memory.total() == float(g["final_total"])
    assert buf.memory.min() == float(g["final_min"])


@pytest.mark.parametrize("name", ["uniform_trace_sac", "uniform_trace_3step"])
def test_oracle_uniform_matches_reference_golden(name):
    g = gr.load(name)
    seed, capacity, num_steps, steps, batch = [int(x) for x in g["meta"]]
    gamma = float(g["gamma"])
    buf = OracleReplayBuffer(capacity, num_steps=num_steps)
    np.random.seed(seed)
    sample_at = set(int(t) for t in g["sample_at"])
    off = 0
    for t in range(steps):
        buf.append(g["obs"][t], g["acts"][t], float(g["rews"][t]), g["obs"][t + 1], None,
                   bool(g["terms"][t]))
        if g["terms"][t]:
            buf.stop_current_episode()
        if t in sample_at:
            exps = buf.sample(batch)
            b = batch_experiences_np(exps, lambda x: x, gamma)
            sl = slice(off, off + batch)
            assert np.array_equal(b["state"], g["state"][sl])
            assert np.array_equal(b["next_state"], g["next_state"][sl])
            assert np.array_equal(b["action"], g["action"][sl])
            assert np.array_equal(b["reward"], g["reward"][sl])
            assert np.array_equal(b["discount"], g["discount"][sl])
            assert np.array_equal(b["is_state_terminal"], g["terminal"][sl])
            off += batch
    assert off == len(g["reward"])
