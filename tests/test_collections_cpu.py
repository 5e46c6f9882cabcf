"""pfrl_b200.collections.PrioritizedBuffer (the reference's class boundary
for the priority trees) on the reference's golden 1-step PER trace: same
sampled values, probabilities -> weights, max_priority.  CPU: the device store
is tests/fake_store.OracleBackedStore; the CUDA variant is in
tests/test_zz_pending_validation_gpu.py."""
import os
import sys
from unittest import mock

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import golden_replay as gr  # noqa: E402
from fake_store import OracleBackedStore  # noqa: E402


def replay_collections_trace(make_buffer):
    g = gr.load("per_trace_1step")
    kw, seed, gamma, lazy = gr.per_kwargs(g)
    buf = make_buffer(kw["capacity"])
    alpha, eps, beta = kw["alpha"], 0.01, kw["beta0"]
    beta_add = (1.0 - beta) / kw["betasteps"]
    np.random.seed(seed)
    n_appended = off = 0
    for row in g["ops"]:
        op, n = int(row[0]), int(row[6])
        if op == gr.OP_APPEND:
            buf.append(n_appended)              # value = absolute index of the transition
            n_appended += 1
        elif op == gr.OP_SAMPLE:
            sampled, probs, min_prob = buf.sample(n)
            first_live = n_appended - len(buf)
            sl = slice(off, off + n)
            assert [v - first_live for v in sampled] == g["idx"][sl].tolist()
            # normalize_by_max=True means "batch": the global min is replaced by the batch's
            # (replay_buffers/prioritized.py:40-41, 57-66)
            assert str(g["normalize_by_max"]) == "True"
            floor = min(probs)
            assert floor >= min_prob * (1 - 1e-12)
            got = np.asarray([(p / floor) ** -beta for p in probs], dtype=np.float32)
            np.testing.assert_allclose(got, g["weight"][sl].astype(np.float32), rtol=2e-6)
            beta = min(1.0, beta + beta_add)
            errs = [float(x) for x in g["errors"][sl]]
            buf.set_last_priority([(min(max(e, 0), 1) + eps) ** alpha for e in errs])
            off += n
    assert buf.max_priority == float(g["final_max_priority"])
    return buf


def test_prioritized_buffer_on_reference_trace():
    from pfrl_b200.collections import PrioritizedBuffer

    with mock.patch("pfrl_b200.collections.prioritized.DeviceReplayStore", OracleBackedStore):
        buf = replay_collections_trace(lambda cap: PrioritizedBuffer(capacity=cap))
        with pytest.raises(NotImplementedError):
            buf.sample(2, uniform_ratio=0.5)
        with pytest.raises(NotImplementedError):
            PrioritizedBuffer(capacity=4, wait_priority_after_sampling=False)
        # protocol asserts of the reference (:98,:108-110)
        buf.sample(2)
        with pytest.raises(AssertionError):
            buf.sample(2)
        with pytest.raises(AssertionError):
            buf.set_last_priority([1.0])
        with pytest.raises(AssertionError):
            buf.set_last_priority([1.0, 0.0])
        buf.set_last_priority([1.0, 2.0])
        assert buf.max_priority >= 2.0
