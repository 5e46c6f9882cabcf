"""SerialVectorEnv / MultiprocessVectorEnv against plain envs, after the
reference's tests/envs_tests/test_vector_envs.py:8-90 (gym is not installed:
the env is a small seeded random walk with array observations), plus the
shared-memory observation slab."""
import numpy as np
import pytest

from pfrl_b200.envs import MultiprocessVectorEnv, SerialVectorEnv

# subprocess tests: a stuck worker must fail the test, not hang the run
pytestmark = pytest.mark.timeout(180)


class WalkEnv:
    action_space = "Discrete(3)"
    observation_space = "Box(6)"
    spec = "Walk-v0"

    def __init__(self, ragged=False):
        self.rng = np.random.RandomState(0)
        self.ragged = ragged

    def seed(self, s):
        self.rng = np.random.RandomState(s)
        return [s]

    def _obs(self):
        n = 6 if not self.ragged else int(self.rng.randint(3, 7))
        return self.rng.randn(n).astype(np.float32)

    def reset(self):
        self.t = 0
        return self._obs()

    def step(self, action):
        self.t += 1
        done = bool(self.rng.rand() < 0.2)
        return self._obs(), float(action) + self.rng.rand(), done, {"t": self.t}

    def close(self):
        pass


def _make(kind, n, **kw):
    if kind == "serial":
        return SerialVectorEnv([WalkEnv(**kw) for _ in range(n)])
    return MultiprocessVectorEnv([(lambda: WalkEnv(**kw)) for _ in range(n)],
                                 shared_obs=(kind == "multiprocess-slab"))


@pytest.mark.parametrize("num_envs", [1, 3])
@pytest.mark.parametrize("kind", ["serial", "multiprocess-slab", "multiprocess-pipe"])
def test_seed_reset_and_step(kind, num_envs):
    vec = _make(kind, num_envs)
    envs = [WalkEnv() for _ in range(num_envs)]
    try:
        assert vec.num_envs == num_envs
        assert vec.action_space == WalkEnv.action_space
        assert vec.observation_space == WalkEnv.observation_space
        assert vec.spec == WalkEnv.spec
        seeds = [100 + i for i in range(num_envs)]
        vec.seed(seeds)
        for e, s in zip(envs, seeds):
            e.seed(s)
        obss = vec.reset()
        real = [e.reset() for e in envs]
        np.testing.assert_array_equal(np.stack(obss), np.stack(real))
        for _ in range(4):
            actions = [i % 3 for i in range(num_envs)]
            real, rr, rd, ri = zip(*[e.step(a) for e, a in zip(envs, actions)])
            obss, rews, dones, infos = vec.step(actions)
            np.testing.assert_array_equal(np.stack(obss), np.stack(real))
            assert tuple(rews) == rr and tuple(dones) == rd and tuple(infos) == ri
        # full mask: nothing is reset
        obss = vec.reset(np.ones(num_envs))
        np.testing.assert_array_equal(np.stack(obss), np.stack(real))
        # partial mask: every env except the last restarts
        mask = np.zeros(num_envs)
        mask[-1] = 1
        obss = vec.reset(mask)
        real = list(real)
        for i in range(num_envs):
            if not mask[i]:
                real[i] = envs[i].reset()
        np.testing.assert_array_equal(np.stack(obss), np.stack(real))
    finally:
        vec.close()


def test_shared_slab_semantics():
    vec = _make("multiprocess-slab", 3)
    try:
        vec.seed([1, 2, 3])
        first = vec.reset()
        assert first.host_batch is not None and tuple(first.host_batch.shape) == (3, 6)
        np.testing.assert_array_equal(first.host_batch.numpy(), np.stack(first))
        kept = [o.copy() for o in first]
        second, _, _, _ = vec.step([0, 1, 2])
        # entries handed out earlier are private copies: the next step must not touch them
        np.testing.assert_array_equal(np.stack(first), np.stack(kept))
        np.testing.assert_array_equal(second.host_batch.numpy(), np.stack(second))
        assert not np.array_equal(np.stack(second), np.stack(kept))
        # the slab feeds batch_states only on a CUDA device; on CPU the per-env path is used
        import torch

        from pfrl_b200.utils.batch_states import batch_states
        from pfrl_b200.utils.phi import Identity

        b = batch_states(second, torch.device("cpu"), Identity())
        assert b.data_ptr() != second.host_batch.data_ptr()
        np.testing.assert_array_equal(b.numpy(), np.stack(second))
    finally:
        vec.close()
    with pytest.raises(AssertionError):
        vec.step([0, 0, 0])


def test_ragged_observations_fall_back_to_the_pipe():
    vec = _make("multiprocess-slab", 2, ragged=True)
    envs = [WalkEnv(ragged=True) for _ in range(2)]
    try:
        vec.seed([5, 6])
        for e, s in zip(envs, (5, 6)):
            e.seed(s)
        obss = vec.reset()
        real = [e.reset() for e in envs]
        for _ in range(6):
            for a, b in zip(obss, real):
                np.testing.assert_array_equal(a, b)
            obss, _, _, _ = vec.step([1, 1])
            real = [e.step(1)[0] for e in envs]
            if any(len(o) != len(real[0]) for o in real) or len(real[0]) != vec._slab.shape[1]:
                assert obss.host_batch is None
    finally:
        vec.close()


def _broken_env():
    raise ValueError("cannot build this env")


class _FailsOnStep(WalkEnv):
    def step(self, action):
        raise KeyError("boom")


def test_worker_failures_surface_in_the_parent():
    with pytest.raises(RuntimeError, match="cannot build this env"):
        MultiprocessVectorEnv([_broken_env])
    vec = MultiprocessVectorEnv([WalkEnv, _FailsOnStep])
    try:
        vec.reset()
        with pytest.raises(RuntimeError, match="boom"):
            vec.step([0, 0])
    finally:
        vec.close()
