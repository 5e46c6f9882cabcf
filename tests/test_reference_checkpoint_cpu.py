"""Replay checkpoints written by the reference (tests/golden/ref_*.pkl, made by
oracle/gen_golden.py:gen_reference_checkpoints with the real pfrl) load into
our buffers without the pfrl package, and hold what the reference's own
batch_experiences says they hold.  CPU only: the device buffer runs against
tests/fake_store.py."""
import os
import sys
from unittest import mock

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, os.path.dirname(__file__))
from fake_store import FakeStore  # noqa: E402


def _expected(name):
    return np.load(os.path.join(GOLD, name + "_expected.npz"))


def test_reader_needs_no_pfrl_and_recovers_priorities():
    from pfrl_b200.replay_buffers import reference_pickle
    from pfrl_b200.utils.lazy_frames import LazyFrames

    path = os.path.join(GOLD, "ref_per_lazyframes.pkl")
    assert reference_pickle.looks_like_pickle(path)
    with mock.patch.dict(sys.modules, {"pfrl": None}):      # `import pfrl` would raise
        ref = reference_pickle.read(path)
    exp = _expected("ref_per_lazyframes")
    assert len(ref.experiences) == int(exp["n"]) and ref.capacity == int(exp["capacity"])
    assert np.array_equal(ref.priorities, exp["priority"])          # bit-exact leaves
    assert ref.max_priority == float(exp["max_priority"])
    assert abs(ref.priorities.sum() - float(exp["total"])) < 1e-9
    first = ref.experiences[0][0]["state"]
    assert isinstance(first, LazyFrames) and np.asarray(first).shape == (4, 6, 6)
    # frame sharing survives: next_state of k and state of k+1 share 3 of 4 frames
    a, b = ref.experiences[0][0]["next_state"], ref.experiences[1][0]["state"]
    if not ref.experiences[0][0]["is_state_terminal"]:
        assert all(x is y for x, y in zip(a._frames, b._frames))

    uni = reference_pickle.read(os.path.join(GOLD, "ref_uniform_3step.pkl"))
    assert uni.priorities is None and uni.capacity == 50
    assert len(uni.experiences) == int(_expected("ref_uniform_3step")["n"])
    assert max(len(e) for e in uni.experiences) == 3


def test_host_buffer_loads_reference_checkpoint(tmp_path):
    from pfrl_b200.replay_buffer import batch_experiences
    from pfrl_b200.replay_buffers import HostReplayBuffer

    exp = _expected("ref_uniform_3step")
    buf = HostReplayBuffer(capacity=50, num_steps=3)
    buf.load(os.path.join(GOLD, "ref_uniform_3step.pkl"))
    assert len(buf) == int(exp["n"])
    b = batch_experiences(list(buf.memory), torch.device("cpu"),
                          lambda x: np.asarray(x, dtype=np.float32), 0.9)
    for k in ("state", "next_state", "action", "reward", "is_state_terminal", "discount"):
        np.testing.assert_allclose(b[k].numpy(), exp[k], rtol=1e-6, atol=1e-7, err_msg=k)
    # and its own format still round-trips
    buf.save(str(tmp_path / "own.pkl"))
    again = HostReplayBuffer(capacity=50, num_steps=3)
    again.load(str(tmp_path / "own.pkl"))
    assert len(again) == len(buf)


def _check_records(buf, store, exp, gamma):
    lay = buf.layout
    np_dtype = np.dtype(lay.part_dtype)
    n = int(exp["n"])
    assert len(store.records) == n
    for k, rec in enumerate(store.records):
        s = store.obs(rec["sp"], np_dtype, lay.part_nbytes, lay.part_shape)
        ns = store.obs(rec["nx"], np_dtype, lay.part_nbytes, lay.part_shape)
        np.testing.assert_array_equal(s.astype(np.float32), exp["state"][k])
        np.testing.assert_array_equal(ns.astype(np.float32), exp["next_state"][k])
        act = rec["action"].view(lay.action_dtype).reshape(lay.action_shape)
        np.testing.assert_array_equal(act, exp["action"][k])
        L = rec["len"]
        ret = sum(gamma ** i * rec["rewards"][i] for i in range(L))
        assert abs(np.float32(ret) - exp["reward"][k]) <= 1e-6 * max(1.0, abs(ret))
        assert abs(np.float32(gamma ** L) - exp["discount"][k]) < 1e-7
        assert float(rec["terminal"]) == exp["is_state_terminal"][k]


def test_device_buffers_restore_reference_checkpoints():
    from pfrl_b200.replay_buffers import PrioritizedReplayBuffer, ReplayBuffer

    FakeStore.instances.clear()
    with mock.patch("pfrl_b200.replay_buffers.device_buffer.DeviceReplayStore", FakeStore):
        exp = _expected("ref_per_lazyframes")
        per = PrioritizedReplayBuffer(capacity=40, alpha=0.6, beta0=0.4, betasteps=100, device=0)
        per.load(os.path.join(GOLD, "ref_per_lazyframes.pkl"))
        store = per.store
        assert len(per) == int(exp["n"]) and not per._waiting
        _check_records(per, store, exp, 0.99)
        assert np.array_equal([r["priority"] for r in store.records], exp["priority"])
        assert store.max_priority == float(exp["max_priority"])
        # LazyFrames de-duplication: far fewer parts than 2 * 4 frames per experience
        assert store.part_head < 2.2 * int(exp["n"])
        # appending afterwards continues the same ring
        from pfrl_b200.utils.lazy_frames import LazyFrames

        frames = [np.full((1, 6, 6), i, np.uint8) for i in range(5)]
        per.append(LazyFrames(frames[:4], stack_axis=0), 1, 0.5,
                   LazyFrames(frames[1:], stack_axis=0))
        per._flush()
        assert store.records[-1]["priority"] == float(exp["max_priority"])

        exp = _expected("ref_uniform_3step")
        uni = ReplayBuffer(capacity=50, num_steps=3, device=0)
        uni.load(os.path.join(GOLD, "ref_uniform_3step.pkl"))
        _check_records(uni, uni.store, exp, 0.9)

        # a uniform checkpoint cannot silently become a prioritised buffer
        per2 = PrioritizedReplayBuffer(capacity=50, num_steps=3, device=0)
        with pytest.raises(TypeError):
            per2.load(os.path.join(GOLD, "ref_uniform_3step.pkl"))
        # nor a 3-step checkpoint a 1-step buffer
        one = ReplayBuffer(capacity=50, num_steps=1, device=0)
        with pytest.raises(ValueError):
            one.load(os.path.join(GOLD, "ref_uniform_3step.pkl"))


def test_hwc_lazy_frames_are_refused_loudly():
    from pfrl_b200.replay_buffers import ReplayBuffer
    from pfrl_b200.utils.lazy_frames import LazyFrames

    frames = [np.zeros((6, 6, 1), np.uint8) for _ in range(5)]
    with mock.patch("pfrl_b200.replay_buffers.device_buffer.DeviceReplayStore", FakeStore):
        buf = ReplayBuffer(10, device=0)
        with pytest.raises(TypeError, match="stack_axis"):
            buf.append(LazyFrames(frames[:4]), 0, 0.0, LazyFrames(frames[1:]))   # default axis 2
