"""CPU, build container only: re-validate the oracle live against the REAL
reference (skipped on the GPU box, where /root/reference does not exist)."""
import numpy as np
import pytest

from oracle import refimport
from oracle.replay import OraclePrioritizedBuffer

pytestmark = pytest.mark.skipif(not refimport.available(), reason="reference tree not present")


@pytest.mark.parametrize("cap", [1, 2, 3, 7, 64, 100, 777, 1024, 1025])
def test_dense_heap_oracle_is_bit_identical_to_reference_tree(cap):
    refimport.import_reference()
    from pfrl.collections.prioritized import PrioritizedBuffer

    rng = np.random.RandomState(cap)
    ref, ora = PrioritizedBuffer(capacity=cap), OraclePrioritizedBuffer(cap)
    draws = 0
    for t in range(1200):
        if rng.rand() < 0.6 or len(ref) < min(4, cap):
            for _ in range(int(rng.randint(1, 8))):
                pr = None if rng.rand() < 0.7 else float(rng.rand() * 3 + 1e-3)
                ref.append(t, pr)
                ora.append(t, pr)
        else:
            n = int(rng.randint(1, min(len(ref), 40) + 1))
            np.random.seed(int(rng.randint(1 << 30)))
            state = np.random.get_state()
            ri, rp, rmin = ref._sample_indices_and_probabilities(n, 0)
            ref.sampled_indices, ref.flag_wait_priority = ri, True
            np.random.set_state(state)
            oi, op, tot, mn = ora.sample_indices(n)
            assert list(oi) == list(ri)
            assert [p / tot for p in op.tolist()] == rp
            assert mn / tot == rmin
            newp = [float(x) for x in (rng.rand(n) * 2 + 1e-6) ** 0.6]
            ref.set_last_priority(newp)
            ora.set_last_priority(newp)
            assert ref.max_priority == ora.max_priority
            draws += n
        assert len(ref) == len(ora)
    assert draws > 0 or cap == 1
    assert ref.priority_sums.sum() == ora.total()
