"""Host-memory uniform replay buffer for CPU-only agents (``gpu=None``), e.g.
the CartPole quickstart configuration.  This is an explicit, separate class:
the device buffers never fall back to it.  Semantics follow
pfrl/replay_buffers/replay_buffer.py:11-94 (n-step window per env_id, uniform
``sample_n_k`` sampling, pickle save/load)."""
import collections
import pickle

from pfrl_b200.collections.random_access_queue import RandomAccessQueue


class HostReplayBuffer:
    def __init__(self, capacity=None, num_steps=1):
        assert num_steps > 0
        self._capacity = capacity
        self.num_steps = num_steps
        self.memory = RandomAccessQueue(maxlen=capacity)   # O(1) indexing for sample()
        self.last_n_transitions = collections.defaultdict(
            lambda: collections.deque([], maxlen=num_steps))

    @property
    def capacity(self):
        return self._capacity

    def append(self, state, action, reward, next_state=None, next_action=None,
               is_state_terminal=False, env_id=0, **kwargs):
        window = self.last_n_transitions[env_id]
        window.append(dict(state=state, action=action, reward=reward, next_state=next_state,
                           next_action=next_action, is_state_terminal=is_state_terminal,
                           **kwargs))
        if is_state_terminal:
            while window:
                self.memory.append(list(window))
                window.popleft()
        elif len(window) == self.num_steps:
            self.memory.append(list(window))

    def stop_current_episode(self, env_id=0):
        window = self.last_n_transitions[env_id]
        if 0 < len(window) < self.num_steps:
            self.memory.append(list(window))
        if 0 < len(window) <= self.num_steps:
            window.popleft()
        while window:
            self.memory.append(list(window))
            window.popleft()

    def sample(self, num_experiences):
        assert len(self.memory) >= num_experiences
        return self.memory.sample(num_experiences)

    def __len__(self):
        return len(self.memory)

    def save(self, filename):
        with open(filename, "wb") as f:
            pickle.dump(list(self.memory), f)

    def load(self, filename):
        """Own format (a pickled list) or a checkpoint written by the
        reference's ReplayBuffer.save (replay_buffer.py:85-94)."""
        from pfrl_b200.replay_buffers import reference_pickle

        items = reference_pickle.read(filename).experiences
        self.memory = RandomAccessQueue(items, maxlen=self._capacity)
