"""GPU: the reference's own replay-buffer unit tests
(tests/replay_buffers_test/test_replay_buffer.py), re-expressed against the
device buffers.  The reference compares whole lists of transition dicts; the
device buffers keep what batch_experiences reads (first state / action, per
step rewards, last next_state, any-terminal), so the assertions check exactly
those fields, the lengths and the counts."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _t(state, terminal=False, next_state=3, reward=2, action=1):
    return dict(state=np.int64(state), action=action, reward=reward,
                next_state=np.int64(next_state), next_action=None, is_state_terminal=terminal)


@pytest.mark.parametrize("capacity", [100, None])
@pytest.mark.parametrize("num_steps", [1, 3])
class TestReplayBuffer:
    def test_append_and_sample(self, capacity, num_steps):  # reference :22-65
        from pfrl_b200.replay_buffers import ReplayBuffer

        rbuf = ReplayBuffer(capacity, num_steps)
        assert len(rbuf) == 0
        for _ in range(num_steps):
            rbuf.append(**_t(0))
        assert len(rbuf) == 1
        s1 = rbuf.sample(1)
        assert len(s1) == 1 and len(s1[0]) == num_steps
        assert int(s1[0][0]["state"]) == 0 and s1[0][0]["action"] == 1
        assert [tr["reward"] for tr in s1[0]] == [2.0] * num_steps
        assert int(s1[0][-1]["next_state"]) == 3
        rbuf.append(**_t(1))
        assert len(rbuf) == 2
        s2 = rbuf.sample(2)
        assert len(s2) == 2
        firsts = sorted(int(e[0]["state"]) for e in s2)
        # windows [0,0,0] and [0,0,1] (n = 3) or [0] and [1] (n = 1): sampled without repetition
        assert firsts == ([0, 1] if num_steps == 1 else [0, 0])
        assert sorted(int(e[-1]["next_state"]) for e in s2) == [3, 3]

    def test_append_and_terminate(self, capacity, num_steps):  # reference :67-118
        from pfrl_b200.replay_buffers import ReplayBuffer

        rbuf = ReplayBuffer(capacity, num_steps)
        for _ in range(num_steps):
            rbuf.append(**_t(0))
        assert len(rbuf) == 1
        rbuf.append(**_t(1, terminal=True))
        assert len(rbuf) == num_steps + 1  # the terminal flushes the n-1 shorter tails
        s = rbuf.sample(num_steps + 1)
        lens = sorted(len(e) for e in s)
        assert lens == ([1, 1] if num_steps == 1 else [1, 2, 3, 3])
        for e in s:
            ends_with_terminal = e[-1]["is_state_terminal"]
            if len(e) < num_steps or (num_steps > 1 and int(e[0]["state"]) == 0 and ends_with_terminal):
                assert ends_with_terminal
            assert all(not tr["is_state_terminal"] for tr in e[:-1])

    def test_stop_current_episode(self, capacity, num_steps):  # reference :120-144
        from pfrl_b200.replay_buffers import ReplayBuffer

        rbuf = ReplayBuffer(capacity, num_steps)
        for _ in range(num_steps - 1):
            rbuf.append(**_t(0))
        assert len(rbuf) == 0
        rbuf.stop_current_episode()
        assert len(rbuf) == num_steps - 1

    def test_save_and_load(self, capacity, num_steps, tmp_path):  # reference :146-205
        from pfrl_b200.replay_buffers import ReplayBuffer

        rbuf = ReplayBuffer(capacity, num_steps)
        for _ in range(num_steps):
            rbuf.append(**_t(0))
        rbuf.append(**_t(1))
        assert len(rbuf) == 2
        fn = os.path.join(str(tmp_path), "rbuf.pkl")
        rbuf.save(fn)
        rbuf = ReplayBuffer(capacity, num_steps)
        assert len(rbuf) == 0
        rbuf.load(fn)
        assert len(rbuf) == 2
        s2 = rbuf.sample(2)
        assert sorted(len(e) for e in s2) == [num_steps, num_steps]
        assert sorted(int(e[0]["state"]) for e in s2) == ([0, 1] if num_steps == 1 else [0, 0])


def test_capacity_drops_the_oldest():  # reference :636-702 (capacity semantics)
    from pfrl_b200.replay_buffers import ReplayBuffer

    rbuf = ReplayBuffer(10)
    for i in range(15):
        rbuf.append(**_t(i, next_state=i + 1))
    assert len(rbuf) == 10
    states = sorted(int(e[0]["state"]) for e in rbuf.sample(10))
    assert states == list(range(5, 15))


def test_env_id_windows_do_not_mix():  # reference :523-544
    from pfrl_b200.replay_buffers import ReplayBuffer

    rbuf = ReplayBuffer(100, num_steps=2)
    rbuf.append(**_t(0, next_state=1), env_id=0)
    rbuf.append(**_t(100, next_state=101), env_id=1)
    assert len(rbuf) == 0  # one transition per env: no 2-step window yet
    rbuf.append(**_t(1, next_state=2), env_id=0)
    rbuf.append(**_t(101, next_state=102), env_id=1)
    assert len(rbuf) == 2
    pairs = sorted((int(e[0]["state"]), int(e[-1]["next_state"])) for e in rbuf.sample(2))
    assert pairs == [(0, 2), (100, 102)]
    rbuf.stop_current_episode(env_id=0)
    assert len(rbuf) == 3  # the tail [1] of env 0 only


@pytest.mark.parametrize("normalize_by_max", ["batch", "memory"])
@pytest.mark.parametrize("num_steps", [1, 3])
def test_normalize_by_max(normalize_by_max, num_steps):  # reference :400-449
    from pfrl_b200.replay_buffers import PrioritizedReplayBuffer

    rbuf = PrioritizedReplayBuffer(200, normalize_by_max=normalize_by_max, error_max=1000,
                                   num_steps=num_steps)
    for i in range(100 + num_steps - 1):
        rbuf.append(**_t(i, next_state=i + 1))
    assert len(rbuf) == 100

    def set_errors_based_on_state(samples):
        rbuf.update_errors([float(s[0]["state"]) for s in samples])

    np.random.seed(0)
    set_errors_based_on_state(rbuf.sample(100))
    for i in range(0, 100, 7):
        samples = rbuf.sample(i + 1)
        weights = [s[0]["weight"] for s in samples]
        assert len(set(weights)) == len(samples)  # all errors differ -> all weights differ
        max_w = max(weights)
        if normalize_by_max == "batch":
            np.testing.assert_allclose(max_w, 1)
        elif any(int(s[0]["state"]) == 0 for s in samples):
            np.testing.assert_allclose(max_w, 1)
        else:
            assert max_w < 1
        set_errors_based_on_state(samples)
