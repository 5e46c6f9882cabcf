"""CPU restatement of the reference replay path (TEST INFRASTRUCTURE ONLY).

Restates, on top of per_oracle.c:
  pfrl/collections/prioritized.py:21-116      PrioritizedBuffer
  pfrl/replay_buffers/replay_buffer.py:24-80  ReplayBuffer (n-step window)
  pfrl/replay_buffers/prioritized.py:31-126   PriorityWeightError,
                                              PrioritizedReplayBuffer
  pfrl/collections/random_access_queue.py:100 RandomAccessQueue.sample
  pfrl/utils/random.py:4-28                   sample_n_k
  pfrl/replay_buffer.py:157-212               batch_experiences
  pfrl/utils/batch_states.py:18-36            batch_states

Random numbers come from numpy's *global legacy* RandomState exactly where the
reference draws them, so a seeded run consumes the identical stream.
"""
import collections
import ctypes

import numpy as np

from . import lib as _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class OraclePrioritizedBuffer:
    """Dense-heap restatement of PrioritizedBuffer (prioritized.py:21-116)."""

    def __init__(self, capacity):
        assert capacity is not None and capacity > 0
        self.capacity = capacity
        self._L = _lib()
        self._h = self._L.ora_per_create(int(capacity))
        self.data = collections.deque()
        self.sampled_indices = []

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ora_per_destroy(self._h)
            self._h = None

    def __len__(self):
        return int(self._L.ora_per_len(self._h))

    @property
    def max_priority(self):
        return float(self._L.ora_per_max_priority(self._h))

    def total(self):
        return float(self._L.ora_per_total(self._h))

    def min(self):
        return float(self._L.ora_per_min(self._h))

    def append(self, value, priority=None):
        # prioritized.py:39-48 (eviction happens inside the C append)
        if len(self.data) == self.capacity:
            self.data.popleft()
        self.data.append(value)
        self._L.ora_per_append(self._h, -1.0 if priority is None else float(priority))

    def sample_indices(self, n, u=None):
        """prioritized.py:56-84 with uniform_ratio == 0.

        Returns (indices, priorities, total, min_tree_root)."""
        if u is None:
            # np.random.uniform(0.0, root) == 0.0 + root * random_sample()
            u = np.random.random_sample(n)
        u = np.ascontiguousarray(u, dtype=np.float64)
        idx = np.empty(n, dtype=np.int64)
        pri = np.empty(n, dtype=np.float64)
        tot = np.empty(1, dtype=np.float64)
        mn = np.empty(1, dtype=np.float64)
        rc = self._L.ora_per_sample(
            self._h, n, _ptr(u), _ptr(idx), _ptr(pri), _ptr(tot), _ptr(mn)
        )
        assert rc == 0, "sample() called while waiting for priorities"
        self.sampled_indices = idx.tolist()
        return idx, pri, float(tot[0]), float(mn[0])

    def sample(self, n, u=None):
        idx, pri, total, mn = self.sample_indices(n, u)
        probs = [p / total for p in pri.tolist()]  # :79-82, uniform_ratio 0
        min_prob = mn / total  # :60
        sampled = [self.data[i] for i in idx.tolist()]  # :102
        return sampled, probs, min_prob

    def set_last_priority(self, priority):
        p = np.ascontiguousarray(priority, dtype=np.float64)
        rc = self._L.ora_per_set_last_priority(self._h, len(p), _ptr(p))
        assert rc == 0, "set_last_priority protocol/positivity violated (rc=%d)" % rc
        self.sampled_indices = []


def sample_n_k(n, k):
    """k distinct uniform indices from range(n); pfrl/utils/random.py:4-28."""
    if not 0 <= k <= n:
        raise ValueError("Sample larger than population or is negative")
    if k == 0:
        return np.empty((0,), dtype=np.int64)
    if 3 * k >= n:
        return np.random.choice(n, k, replace=False)
    draw = np.random.choice(n, 2 * k)
    seen = set()
    spare = k
    for i in range(k):
        x = draw[i]
        while x in seen:
            x = draw[i] = draw[spare]
            spare += 1
            if spare == 2 * k:
                draw[k:] = np.random.choice(n, k)
                spare = k
        seen.add(x)
    return draw[:k]


class OracleReplayBuffer:
    """Uniform n-step buffer; replay_buffers/replay_buffer.py:24-80."""

    def __init__(self, capacity=None, num_steps=1):
        assert num_steps > 0
        self.capacity = capacity
        self.num_steps = num_steps
        self.memory = collections.deque(maxlen=capacity)
        self.last_n = collections.defaultdict(
            lambda: collections.deque([], maxlen=num_steps)
        )

    def _emit(self, experience):
        self.memory.append(experience)

    def append(self, state, action, reward, next_state=None, next_action=None,
               is_state_terminal=False, env_id=0, **kwargs):
        # :33-62
        window = self.last_n[env_id]
        window.append(dict(state=state, action=action, reward=reward,
                           next_state=next_state, next_action=next_action,
                           is_state_terminal=is_state_terminal, **kwargs))
        if is_state_terminal:
            while window:
                self._emit(list(window))
                window.popleft()
        elif len(window) == self.num_steps:
            self._emit(list(window))

    def stop_current_episode(self, env_id=0):
        # :64-76
        window = self.last_n[env_id]
        if 0 < len(window) < self.num_steps:
            self._emit(list(window))
        if 0 < len(window) <= self.num_steps:
            window.popleft()
        while window:
            self._emit(list(window))
            window.popleft()

    def sample(self, n):
        assert len(self.memory) >= n
        return [self.memory[int(i)] for i in sample_n_k(len(self.memory), n)]

    def __len__(self):
        return len(self.memory)


class OraclePrioritizedReplayBuffer(OracleReplayBuffer):
    """replay_buffers/prioritized.py:69-126 over OraclePrioritizedBuffer."""

    def __init__(self, capacity, alpha=0.6, beta0=0.4, betasteps=2e5, eps=0.01,
                 normalize_by_max=True, error_min=0, error_max=1, num_steps=1):
        assert num_steps > 0
        self.capacity = capacity
        self.num_steps = num_steps
        self.memory = OraclePrioritizedBuffer(capacity)
        self.last_n = collections.defaultdict(
            lambda: collections.deque([], maxlen=num_steps)
        )
        assert 0.0 <= alpha and 0.0 <= beta0 <= 1.0
        self.alpha, self.beta, self.eps = alpha, beta0, eps
        self.beta_add = 0 if betasteps is None else (1.0 - beta0) / betasteps
        if normalize_by_max is True:
            normalize_by_max = "batch"
        assert normalize_by_max in (False, "batch", "memory")
        self.normalize_by_max = normalize_by_max
        self.error_min, self.error_max = error_min, error_max

    def priority_from_errors(self, errors):
        # prioritized.py:47-55
        out = []
        for d in errors:
            if self.error_min is not None:
                d = max(self.error_min, d)
            if self.error_max is not None:
                d = min(self.error_max, d)
            out.append((d + self.eps) ** self.alpha)
        return out

    def weights_from_probabilities(self, probabilities, min_probability):
        # prioritized.py:57-66
        if self.normalize_by_max == "batch":
            min_probability = np.min(probabilities)
        if self.normalize_by_max:
            w = [(p / min_probability) ** -self.beta for p in probabilities]
        else:
            w = [(len(self.memory) * p) ** -self.beta for p in probabilities]
        self.beta = min(1.0, self.beta + self.beta_add)
        return w

    def sample(self, n, u=None):
        assert len(self.memory) >= n
        sampled, probs, min_prob = self.memory.sample(n, u)
        self.last_probabilities = probs
        self.last_min_probability = min_prob
        weights = self.weights_from_probabilities(probs, min_prob)
        for e, w in zip(sampled, weights):
            e[0]["weight"] = w
        return sampled

    def update_errors(self, errors):
        self.memory.set_last_priority(self.priority_from_errors(errors))


def batch_states_np(states, phi):
    """utils/batch_states.py:18-36 without the device move (numpy stack)."""
    return np.stack([np.asarray(phi(s)) for s in states])


def batch_experiences_np(experiences, phi, gamma):
    """pfrl/replay_buffer.py:157-212 as numpy arrays (fp32 where torch is)."""
    out = {
        "state": batch_states_np([e[0]["state"] for e in experiences], phi),
        "action": np.asarray([e[0]["action"] for e in experiences]),
        # Python-float arithmetic, then cast to float32 (:183-190)
        "reward": np.asarray(
            [sum((gamma ** i) * e[i]["reward"] for i in range(len(e)))
             for e in experiences], dtype=np.float32),
        "next_state": batch_states_np(
            [e[-1]["next_state"] for e in experiences], phi),
        "is_state_terminal": np.asarray(
            [any(t["is_state_terminal"] for t in e) for e in experiences],
            dtype=np.float32),
        "discount": np.asarray(
            [gamma ** len(e) for e in experiences], dtype=np.float32),
    }
    return out
