"""Pure-Python port of the reference's prioritized replay path, used ONLY as
the timed "reference CPU path" of bench.py (cpu_baseline / --impl reference,
kind="port").  TEST / BENCH INFRASTRUCTURE: never imported by pfrl_b200.

The real reference (pfnet/pfrl) cannot travel to the GPU box, so this port
keeps its cost profile instead: the same interpreter-level work per tree node
(one Python-level read-add-write per level and per tree, ~21 levels at 1M
capacity), the same per-sample Python list handling, the same numpy
concatenate / divide / collate per observation in batch_experiences, and a
single thread.  Restates
  pfrl/collections/prioritized.py:39-116,154-180,245-258,294-312
  pfrl/replay_buffers/prioritized.py:47-66,117-126
  pfrl/replay_buffer.py:157-212, pfrl/utils/batch_states.py:18-36
on a dense list-backed heap (see oracle/per_oracle.c for why that is
bit-identical to the reference's nested-list tree).  Checked against the
oracle in tests/test_pyport.py.
"""
import collections
import math

import numpy as np
import torch
from torch.utils.data._utils.collate import default_collate


class PyPrioritizedBuffer:
    def __init__(self, capacity):
        self.capacity = capacity
        P = 1
        while P < capacity:
            P *= 2
        self.P = P
        self.ns = 2 * P
        self.sum = [0.0] * (2 * self.ns)
        self.min = [math.inf] * (2 * self.ns)
        self.data = collections.deque()
        self.napp = 0
        self.npop = 0
        self.max_priority = 1.0
        self.sampled_slots = []
        self.flag_wait_priority = False

    def __len__(self):
        return self.napp - self.npop

    def _write(self, slot, s_val, m_val):
        s, m = self.sum, self.min
        n = self.ns + slot
        s[n] = s_val
        m[n] = m_val
        n >>= 1
        while n >= 1:
            s[n] = s[2 * n] + s[2 * n + 1]
            a, b = m[2 * n], m[2 * n + 1]
            m[n] = a if a < b else b
            n >>= 1

    def _write_sum(self, slot, val):
        s = self.sum
        n = self.ns + slot
        s[n] = val
        n >>= 1
        while n >= 1:
            s[n] = s[2 * n] + s[2 * n + 1]
            n >>= 1

    def bulk_load(self, values, priorities):
        """Setup helper (untimed): load len(values) <= capacity elements."""
        k = len(values)
        assert self.napp == 0 and k <= self.capacity
        self.data.extend(values)
        leaves = np.zeros(self.ns)
        leaves[:k] = priorities
        mins = np.full(self.ns, np.inf)
        mins[:k] = priorities
        s_levels, m_levels = [leaves], [mins]
        while len(s_levels[-1]) > 1:
            a = s_levels[-1]
            s_levels.append(a[0::2] + a[1::2])
            b = m_levels[-1]
            m_levels.append(np.minimum(b[0::2], b[1::2]))
        s = [0.0]
        m = [math.inf]
        for lv_s, lv_m in zip(reversed(s_levels), reversed(m_levels)):
            s.extend(lv_s.tolist())
            m.extend(lv_m.tolist())
        self.sum, self.min = s, m
        self.napp = k

    def append(self, value, priority=None):
        if len(self) == self.capacity:
            self._write(self.npop % self.ns, 0.0, math.inf)
            self.npop += 1
            self.data.popleft()
        if priority is None:
            priority = self.max_priority
        self.data.append(value)
        self._write(self.napp % self.ns, priority, priority)
        self.napp += 1

    def sample(self, n):
        assert not self.flag_wait_priority
        s = self.sum
        total = s[1]
        min_prob = self.min[1] / total
        ns, P = self.ns, self.P
        older = 3 if (self.npop % ns) >= P else 2
        slots, vals = [], []
        for _ in range(n):
            pos = np.random.uniform(0.0, s[1])
            node = older
            left = s[older]
            if not pos < left:
                pos -= left
                node = older ^ 1
            while node < ns:
                left = s[2 * node]
                if pos < left:
                    node = 2 * node
                else:
                    pos -= left
                    node = 2 * node + 1
            slot = node - ns
            vals.append(s[node])
            self._write_sum(slot, 0.0)
            slots.append(slot)
        self.sampled_slots = slots
        self.flag_wait_priority = True
        base = self.npop % ns
        sampled = [self.data[(sl - base) % ns] for sl in slots]
        probs = [v / total for v in vals]
        return sampled, probs, min_prob

    def set_last_priority(self, priority):
        assert self.flag_wait_priority
        assert all([p > 0.0 for p in priority])
        assert len(self.sampled_slots) == len(priority)
        for sl, p in zip(self.sampled_slots, priority):
            self._write(sl, p, p)
            self.max_priority = max(self.max_priority, p)
        self.flag_wait_priority = False
        self.sampled_slots = []


class PyPrioritizedReplayBuffer:
    def __init__(self, capacity, alpha=0.6, beta0=0.4, betasteps=2e5, eps=0.01,
                 normalize_by_max=True, error_min=0, error_max=1, num_steps=1):
        self.capacity = capacity
        self.num_steps = num_steps
        self.memory = PyPrioritizedBuffer(capacity)
        self.last_n = collections.defaultdict(lambda: collections.deque([], maxlen=num_steps))
        self.alpha, self.beta, self.eps = alpha, beta0, eps
        self.beta_add = 0 if betasteps is None else (1.0 - beta0) / betasteps
        self.normalize_by_max = "batch" if normalize_by_max is True else normalize_by_max
        self.error_min, self.error_max = error_min, error_max

    def __len__(self):
        return len(self.memory)

    def append(self, state, action, reward, next_state=None, next_action=None,
               is_state_terminal=False, env_id=0):
        w = self.last_n[env_id]
        w.append(dict(state=state, action=action, reward=reward, next_state=next_state,
                      next_action=next_action, is_state_terminal=is_state_terminal))
        if is_state_terminal:
            while w:
                self.memory.append(list(w))
                w.popleft()
        elif len(w) == self.num_steps:
            self.memory.append(list(w))

    def stop_current_episode(self, env_id=0):
        w = self.last_n[env_id]
        if 0 < len(w) < self.num_steps:
            self.memory.append(list(w))
        if 0 < len(w) <= self.num_steps:
            w.popleft()
        while w:
            self.memory.append(list(w))
            w.popleft()

    def sample(self, n):
        assert len(self.memory) >= n
        sampled, probs, min_prob = self.memory.sample(n)
        if self.normalize_by_max == "batch":
            min_prob = np.min(probs)
        if self.normalize_by_max:
            weights = [(p / min_prob) ** -self.beta for p in probs]
        else:
            weights = [(len(self.memory) * p) ** -self.beta for p in probs]
        self.beta = min(1.0, self.beta + self.beta_add)
        for e, w in zip(sampled, weights):
            e[0]["weight"] = w
        return sampled

    def update_errors(self, errors):
        out = []
        for d in errors:
            if self.error_min is not None:
                d = max(self.error_min, d)
            if self.error_max is not None:
                d = min(self.error_max, d)
            out.append((d + self.eps) ** self.alpha)
        self.memory.set_last_priority(out)


def py_batch_states(states, device, phi):
    feats = [phi(s) for s in states]
    return default_collate(feats).to(device)


def py_batch_experiences(experiences, device, phi, gamma):
    return {
        "state": py_batch_states([e[0]["state"] for e in experiences], device, phi),
        "action": torch.as_tensor([e[0]["action"] for e in experiences], device=device),
        "reward": torch.as_tensor(
            [sum((gamma ** i) * e[i]["reward"] for i in range(len(e))) for e in experiences],
            dtype=torch.float32, device=device),
        "next_state": py_batch_states([e[-1]["next_state"] for e in experiences], device, phi),
        "is_state_terminal": torch.as_tensor(
            [any(t["is_state_terminal"] for t in e) for e in experiences],
            dtype=torch.float32, device=device),
        "discount": torch.as_tensor(
            [gamma ** len(e) for e in experiences], dtype=torch.float32, device=device),
    }
