"""gym.spaces stand-in (test infrastructure only; see gym/__init__.py)."""
import numpy as np


class Space:
    pass


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(
            np.all(x >= self.low) and np.all(x <= self.high)
        )

    def __eq__(self, other):
        return (isinstance(other, Box) and self.shape == other.shape
                and self.dtype == other.dtype and np.array_equal(self.low, other.low)
                and np.array_equal(self.high, other.high))

    __hash__ = None


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def sample(self):
        return int(np.random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n

    def __hash__(self):
        return hash(("Discrete", self.n))
