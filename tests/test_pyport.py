"""CPU: the pure-Python port used as the timed CPU baseline (oracle/pyport.py)
computes the same thing as the oracle (and hence the reference)."""
import numpy as np

from oracle.pyport import PyPrioritizedReplayBuffer
from oracle.replay import OraclePrioritizedReplayBuffer


def test_pyport_matches_oracle_indices_and_weights():
    kw = dict(capacity=300, alpha=0.5, beta0=0.4, betasteps=50, num_steps=3,
              normalize_by_max="memory")
    a, b = PyPrioritizedReplayBuffer(**kw), OraclePrioritizedReplayBuffer(**kw)
    rng = np.random.RandomState(0)
    for t in range(900):
        term = bool(rng.rand() < 0.05)
        args = (t, int(rng.randint(4)), float(rng.randn()), t + 1, None, term)
        a.append(*args)
        b.append(*args)
        if term:
            a.stop_current_episode()
            b.stop_current_episode()
        if len(a) >= 16 and t % 5 == 0:
            np.random.seed(t)
            ea = a.sample(16)
            np.random.seed(t)
            eb = b.sample(16)
            assert [e[0]["state"] for e in ea] == [e[0]["state"] for e in eb]
            assert [e[0]["weight"] for e in ea] == [e[0]["weight"] for e in eb]
            err = [float(x) for x in rng.rand(16)]
            a.update_errors(err)
            b.update_errors(err)
    assert len(a) == len(b) == 300
    assert a.memory.sum[1] == b.memory.total()


def test_pyport_bulk_load_equals_appends():
    from oracle.pyport import PyPrioritizedBuffer

    pr = np.random.RandomState(1).rand(100) + 0.1
    x, y = PyPrioritizedBuffer(128), PyPrioritizedBuffer(128)
    x.bulk_load(list(range(100)), pr)
    for i in range(100):
        y.append(i, float(pr[i]))
    assert x.sum == y.sum and x.min == y.min
