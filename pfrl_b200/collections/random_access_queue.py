"""FIFO queue with O(1) random access (the uniform replay buffer's container;
reference: pfrl/collections/random_access_queue.py, row a6 of SURVEY section 8).

One Python list plus the offset of the current front: ``popleft`` advances the
offset (amortised O(1): the dead prefix is dropped once it outgrows the live
part), ``q[i]`` is a single list index.  The reference keeps two lists and
reverses one of them lazily; behaviour (indexing incl. negative indices,
``maxlen`` eviction, ``sample`` consuming ``sample_n_k``'s stream) is the same.
"""
from pfrl_b200.utils.random import sample_n_k


class RandomAccessQueue(object):
    def __init__(self, *args, **kwargs):
        self.maxlen = kwargs.pop("maxlen", None)
        assert self.maxlen is None or self.maxlen >= 0
        self._items = list(*args, **kwargs)
        self._head = 0
        self._trim()

    # -- size management ------------------------------------------------------
    def __len__(self):
        return len(self._items) - self._head

    def _trim(self):
        if self.maxlen is not None:
            excess = len(self) - self.maxlen
            if excess > 0:
                self._drop(excess)

    def _drop(self, n):
        end = self._head + n
        self._items[self._head:end] = [None] * n      # release the references now
        self._head = end
        if self._head > 32 and self._head * 2 > len(self._items):
            del self._items[:self._head]
            self._head = 0

    # -- access ---------------------------------------------------------------
    def _position(self, i):
        n = len(self)
        if not -n <= i < n:
            raise IndexError("RandomAccessQueue index out of range")
        return self._head + (i if i >= 0 else n + i)

    def __getitem__(self, i):
        return self._items[self._position(i)]

    def __setitem__(self, i, x):
        self._items[self._position(i)] = x

    def __iter__(self):
        for k in range(self._head, len(self._items)):
            yield self._items[k]

    def __repr__(self):
        return "RandomAccessQueue({})".format(list(self))

    # -- queue operations -------------------------------------------------------
    def append(self, x):
        self._items.append(x)
        self._trim()

    def extend(self, xs):
        self._items.extend(xs)
        self._trim()

    def popleft(self):
        if not len(self):
            raise IndexError("pop from empty RandomAccessQueue")
        x = self._items[self._head]
        self._drop(1)
        return x

    def sample(self, k):
        return [self[i] for i in sample_n_k(len(self), k)]
