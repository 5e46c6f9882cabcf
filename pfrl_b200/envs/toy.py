"""A tiny deterministic-dynamics test environment (vector observations,
discrete or continuous actions) used by the agents' training tests."""
import numpy as np


class ChainEnv:
    """Walk on a chain of ``size`` cells; action 1 moves right (+1 reward at
    the end, episode terminates), action 0 moves left.  Continuous variant:
    action[0] > 0 means right.  Observation: one-hot cell (float32)."""

    def __init__(self, size=5, continuous=False, seed=0, max_steps=30):
        self.size = size
        self.continuous = continuous
        self.rng = np.random.RandomState(seed)
        self.max_steps = max_steps
        self.n_actions = 2
        self.obs_dim = size
        self.act_dim = 1
        # vector-env wrappers (ours and the reference's) read these from their first env
        self.action_space = None
        self.observation_space = None
        self.spec = None

    def _obs(self):
        o = np.zeros(self.size, dtype=np.float32)
        o[self.pos] = 1
        return o

    def reset(self):
        self.pos = 0
        self.steps = 0
        return self._obs()

    def step(self, action):
        right = (np.asarray(action).reshape(-1)[0] > 0) if self.continuous else int(action) == 1
        self.pos = min(self.size - 1, self.pos + 1) if right else max(0, self.pos - 1)
        self.steps += 1
        done = self.pos == self.size - 1
        reward = 1.0 if done else -0.01
        info = {"needs_reset": self.steps >= self.max_steps and not done}
        return self._obs(), reward, done, info

    def seed(self, s):
        self.rng = np.random.RandomState(s)

    def close(self):
        pass
