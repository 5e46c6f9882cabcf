import numpy as np

from pfrl_b200 import env


class SerialVectorEnv(env.VectorEnv):
    """Step a list of ordinary envs one after the other in this process
    (pfrl/envs/serial_vector_env.py:6-44)."""

    def __init__(self, envs):
        self.envs = envs
        self.last_obs = [None] * self.num_envs
        self.action_space = getattr(envs[0], "action_space", None)
        self.observation_space = getattr(envs[0], "observation_space", None)
        self.spec = getattr(envs[0], "spec", None)

    def step(self, actions):
        results = [e.step(a) for e, a in zip(self.envs, actions)]
        self.last_obs, rews, dones, infos = zip(*results)
        return self.last_obs, rews, dones, infos

    def reset(self, mask=None):
        if mask is None:
            mask = np.zeros(self.num_envs)
        self.last_obs = tuple(
            o if keep else e.reset() for keep, e, o in zip(mask, self.envs, self.last_obs))
        return self.last_obs

    def seed(self, seeds):
        for e, s in zip(self.envs, seeds):
            e.seed(s)

    def close(self):
        for e in self.envs:
            e.close()

    @property
    def num_envs(self):
        return len(self.envs)
