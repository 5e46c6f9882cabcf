"""Host-side index sampling with the reference's RNG stream consumption."""
import numpy as np


def sample_n_k(n, k):
    """k distinct integers drawn uniformly from ``range(n)``.

    Same contract and -- importantly for seeded parity -- the same draws from
    numpy's global legacy RandomState as the reference
    (pfrl/utils/random.py:4-28): for ``3k >= n`` one
    ``choice(n, k, replace=False)``; otherwise one ``choice(n, 2k)`` whose
    first k entries are kept, collisions being replaced from the spare half
    in order (a fresh ``choice(n, k)`` refills the spare half if it runs out).
    """
    if k < 0 or k > n:
        raise ValueError("Sample larger than population or is negative")
    if k == 0:
        return np.empty((0,), dtype=np.int64)
    if 3 * k >= n:
        return np.random.choice(n, k, replace=False)
    pool = np.random.choice(n, 2 * k)
    taken = set()
    cursor = k
    for pos in range(k):
        v = pool[pos]
        while v in taken:
            v = pool[pos] = pool[cursor]
            cursor += 1
            if cursor == 2 * k:
                pool[k:] = np.random.choice(n, k)
                cursor = k
        taken.add(v)
    return pool[:k]
