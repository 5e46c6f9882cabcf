""""ABC": the N-step toy problem the reference trains its agents on in its
test-suite (pfrl/envs/abc.py), restated.

States 0 .. N-1 plus a terminal state; only action n is correct in state n.
N correct actions in a row earn +1 and end the episode (episodic) or restart
at state 0 (continuing); a wrong action ends the episode (episodic) or changes
nothing (continuing).  Observations are one-hot over N + 2 slots: the states,
the terminal state and one spare slot used when ``partially_observable`` shifts
a whole episode's observations by one (at random, or on alternate episodes when
``deterministic``).  With ``discrete=False`` an action is a vector of N logits
in [-1, 1]; the inner discrete action is its argmax (``deterministic``) or a
softmax sample drawn from numpy's global stream.
"""
import numpy as np

from pfrl_b200 import env

try:  # real gym (or the test shim) when present, so isinstance checks of callers work
    from gym import spaces as _spaces
except ImportError:  # minimal stand-ins with the attributes agents' tests read
    class _spaces(object):
        class Box(object):
            def __init__(self, low, high, shape, dtype):
                self.shape, self.dtype = tuple(shape), np.dtype(dtype)
                self.low = np.full(shape, low, dtype=dtype)
                self.high = np.full(shape, high, dtype=dtype)

            def sample(self):
                return np.random.uniform(self.low, self.high).astype(self.dtype)

        class Discrete(object):
            def __init__(self, n):
                self.n = n

            def sample(self):
                return int(np.random.randint(self.n))


class ABC(env.Env):
    MAX_SHIFT = 1

    def __init__(self, size=2, discrete=True, partially_observable=False, episodic=True,
                 deterministic=False):
        self.size = size
        self.terminal_state = size
        self.episodic = episodic
        self.partially_observable = partially_observable
        self.deterministic = deterministic
        self.n_max_offset = self.MAX_SHIFT
        self.n_dim_obs = size + 1 + self.MAX_SHIFT
        self.observation_space = _spaces.Box(low=-np.inf, high=np.inf, shape=(self.n_dim_obs,),
                                             dtype=np.float32)
        self._continuous = not discrete
        if discrete:
            self.action_space = _spaces.Discrete(size)
        else:
            self.action_space = _spaces.Box(low=-1.0, high=1.0, shape=(size,), dtype=np.float32)
        self._state = 0
        self._offset = 0

    def observe(self):
        one_hot = np.zeros(self.n_dim_obs, dtype=np.float32)
        one_hot[self._state + self._offset] = 1.0
        return one_hot

    def reset(self):
        self._state = 0
        if not self.partially_observable:
            self._offset = 0
        elif self.deterministic:
            self._offset = (self._offset + 1) % (self.MAX_SHIFT + 1)
        else:
            self._offset = np.random.randint(self.MAX_SHIFT + 1)
        return self.observe()

    def _inner_action(self, action):
        if not self._continuous:
            return action
        assert isinstance(action, np.ndarray)
        logits = np.clip(action, self.action_space.low, self.action_space.high)
        if self.deterministic:
            return np.argmax(logits)
        weights = np.exp(logits)
        return np.random.choice(range(self.size), p=weights / weights.sum())

    def step(self, action):
        correct = self._inner_action(action) == self._state
        reward, done = 0, False
        if correct and self._state == self.size - 1:      # the last correct action: goal
            reward = 1.0
            if self.episodic:
                done, self._state = True, self.terminal_state
            else:
                self._state = 0
        elif correct:
            self._state += 1
        elif self.episodic:                               # a mistake ends the episode
            done, self._state = True, self.terminal_state
        return self.observe(), reward, done, {}

    def close(self):
        pass
