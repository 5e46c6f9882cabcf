"""``PrioritizedBuffer`` of the reference (pfrl/collections/prioritized.py:21-116)
with its two segment trees in HBM.

This is the reference's own class boundary for SURVEY rows a1-a3: the values
stay Python objects in a host deque (any object, as in the reference), while
the priorities live in the device sum / min trees of ``libb2rl`` and every
method maps onto one C-ABI call:

    append(value, priority)   -> b2rl_replay_append  (batched: staged, flushed before use)
    sample(n)                 -> b2rl_per_sample     (exact mode: same indices as the
                                 reference under the same numpy seed), probabilities
                                 p_i / total and min_prob from the roots read before the draws
    set_last_priority(p)      -> b2rl_per_update_priorities (raises max_priority)

The replay-buffer classes in pfrl_b200.replay_buffers do not go through this
class (they also keep the observations in HBM); it exists for code written
against ``pfrl.collections`` directly.  Not supported: ``uniform_ratio > 0``,
``wait_priority_after_sampling=False`` and ``popleft()`` (used only by the
episodic buffers, which are out of scope).
"""
import collections

import numpy as np

from pfrl_b200.store import DeviceReplayStore

_UNBOUNDED_DEFAULT = 1 << 20


class PrioritizedBuffer(object):
    def __init__(self, capacity=None, wait_priority_after_sampling=True,
                 initial_max_priority=1.0, *, device=0, max_batch=4096,
                 unbounded_capacity=_UNBOUNDED_DEFAULT):
        if not wait_priority_after_sampling:
            raise NotImplementedError(
                "wait_priority_after_sampling=False is only used by the episodic buffers")
        self.capacity = capacity
        self.wait_priority_after_sampling = True
        self.data = collections.deque(maxlen=capacity)
        self._alloc = capacity if capacity is not None else unbounded_capacity
        # trees only: one dummy 16-byte part, no observations behind the records
        self._store = DeviceReplayStore(self._alloc, 16, stack=1, n_step=1, action_bytes=8,
                                        prioritized=True, part_capacity=64, device=device,
                                        max_batch=max_batch)
        if initial_max_priority != 1.0:
            self._store.set_max_priority(initial_max_priority)
        self._max_priority_hint = float(initial_max_priority)
        self._staged = []              # priorities (or None) of appends not yet on the device
        self.flag_wait_priority = False
        self.sampled_indices = []

    def __len__(self):
        return len(self.data)

    @property
    def max_priority(self):
        self._flush()
        return self._store.info()["max_priority"]

    def append(self, value, priority=None):
        if self.capacity is None and len(self.data) >= self._alloc:
            raise MemoryError("unbounded PrioritizedBuffer outgrew its %d-entry allocation; "
                              "pass unbounded_capacity=..." % self._alloc)
        self.data.append(value)        # a bounded deque drops the oldest, like the device ring
        self._staged.append(priority)
        if len(self._staged) >= 4096:
            self._flush()

    def popleft(self):
        raise NotImplementedError("popleft() is only used by the episodic buffers")

    def _flush(self):
        staged, self._staged = self._staged, []
        # runs of "default priority" and "explicit priority" go down separately: the C ABI
        # takes either an array of priorities or NULL (= current max_priority) per call
        start = 0
        while start < len(staged):
            explicit = staged[start] is not None
            stop = start
            while stop < len(staged) and (staged[stop] is not None) == explicit \
                    and stop - start < self._alloc:
                stop += 1
            n = stop - start
            zeros = np.zeros(n, dtype=np.int32)
            self._store.append(zeros, zeros, np.zeros(n, dtype=np.int64), np.zeros(n),
                               np.ones(n, dtype=np.uint8), np.zeros(n, dtype=np.uint8),
                               priority=np.asarray(staged[start:stop], dtype=np.float64)
                               if explicit else None)
            start = stop

    def sample(self, n, uniform_ratio=0):
        """-> (sampled values, their probabilities, min_prob); priorities of the
        sampled entries read 0 until ``set_last_priority`` (collections/prioritized.py:86-105)."""
        assert not self.flag_wait_priority
        if uniform_ratio != 0:
            raise NotImplementedError("uniform_ratio > 0 is only used by the episodic buffers")
        assert 0 < n <= len(self.data)
        self._flush()
        roots = self._store.info()                     # total and min BEFORE the draws (:59-60)
        u = np.random.random_sample(n)                 # == n x np.random.uniform(0.0, root), :302
        index, priority = self._store.sample(u, want_priority=True)
        indices = index.cpu().numpy().tolist()
        total = roots["total"]
        probabilities = [p / total for p in priority.cpu().numpy().tolist()]
        self.sampled_indices = indices
        self.flag_wait_priority = True
        return [self.data[i] for i in indices], probabilities, roots["min"] / total

    def set_last_priority(self, priority):
        assert self.flag_wait_priority
        assert all(p > 0.0 for p in priority)
        assert len(self.sampled_indices) == len(priority)
        self._store.update_priorities(np.asarray(priority, dtype=np.float64))
        self.flag_wait_priority = False
        self.sampled_indices = []
