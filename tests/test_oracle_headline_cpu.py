"""CPU: the oracle (oracle/per_oracle.c + oracle/replay.py) reproduces what the REAL
reference produced at the HEADLINE shapes -- 1 M-transition PER, 84x84x4 frames,
3-step / B = 512 and 1-step / B = 32 (tests/golden/headline_*.npz, written by
oracle/gen_golden_headline.py).  Pins the oracle at the sizes the GPU parity tests
(tests/test_headline_shapes_gpu.py) use."""
import os

import numpy as np
import pytest

from oracle.gen_golden_headline import FrameRef, STACK, checksums, script
from oracle.replay import OraclePrioritizedReplayBuffer, batch_experiences_np

GOLD = os.path.join(os.path.dirname(__file__), "golden")


class _Lazy:
    """LazyFrames stand-in: frames concatenated on demand (stack_axis 0)."""

    __slots__ = ("_frames",)

    def __init__(self, frames):
        self._frames = frames

    def __array__(self, dtype=None, copy=None):
        a = np.concatenate([np.asarray(f) for f in self._frames], axis=0)
        return a if dtype is None else a.astype(dtype)


@pytest.mark.timeout(600)
@pytest.mark.parametrize("name", ["headline_c3_rainbow", "headline_c2_dqn"])
def test_oracle_matches_reference_at_headline_shape(name):
    g = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    seed, capacity, num_steps, batch, extra, rounds, chunk, between = [int(x) for x in g["meta"]]
    alpha, beta0, betasteps, gamma = [float(x) for x in g["params"]]
    nb = str(g["normalize_by_max"])
    nb = {"True": True, "False": False}.get(nb, nb)
    buf = OraclePrioritizedReplayBuffer(
        capacity, alpha=alpha, beta0=beta0, betasteps=None if np.isnan(betasteps) else betasteps,
        normalize_by_max=nb, num_steps=num_steps)
    total = capacity + extra
    acts, rews, term = script(seed, total, chunk)
    fid = t = 0
    while t < total:
        m = min(chunk, total - t)
        refs = [FrameRef(fid + j) for j in range(m + STACK)]
        obs = [_Lazy(refs[j:j + STACK]) for j in range(m + 1)]
        for j in range(m):
            buf.append(obs[j], int(acts[t + j]), float(rews[t + j]), obs[j + 1], None,
                       bool(term[t + j]))
        fid += m + STACK
        t += m
    assert len(buf) == capacity
    np.random.seed(seed)
    phi = lambda x: np.asarray(x, dtype=np.float32) / 255  # noqa: E731
    frames = [FrameRef(fid + j) for j in range(STACK)]
    fid += STACK
    cur = _Lazy(list(frames))
    bi = 0
    for r in range(rounds):
        exps = buf.sample(batch)
        assert np.array_equal(np.asarray(buf.memory.sampled_indices), g["idx"][r]), r
        w = np.array([e[0]["weight"] for e in exps])
        np.testing.assert_allclose(w, g["weight"][r], rtol=1e-12)
        b = batch_experiences_np(exps, phi, gamma)
        assert np.array_equal(b["reward"], g["reward"][r])
        assert np.array_equal(b["discount"], g["discount"][r])
        assert np.array_equal(b["is_state_terminal"], g["terminal"][r])
        assert np.array_equal(b["action"], g["action"][r])
        s1, s2 = checksums(b["state"])
        n1, n2 = checksums(b["next_state"])
        assert np.array_equal(s1, g["s1"][r]) and np.array_equal(s2, g["s2"][r])
        assert np.array_equal(n1, g["n1"][r]) and np.array_equal(n2, g["n2"][r])
        buf.update_errors([float(x) for x in g["errors"][r]])
        for j in range(between):
            a, rw, tm = int(g["between_actions"][bi]), float(g["between_rewards"][bi]), bool(
                g["between_terminals"][bi])
            bi += 1
            frames = frames[1:] + [FrameRef(fid)]
            fid += 1
            nxt = _Lazy(list(frames))
            buf.append(cur, a, rw, nxt, None, tm)
            if tm:
                frames = [FrameRef(fid)] * STACK
                fid += 1
                cur = _Lazy(list(frames))
            else:
                cur = nxt
    assert len(buf) == int(g["final_len"])
    assert buf.memory.max_priority == float(g["final_max_priority"])
    assert buf.memory.total() == float(g["final_total"])
    assert buf.memory.min() == float(g["final_min"])
