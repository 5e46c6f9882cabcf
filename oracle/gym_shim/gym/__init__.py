"""Minimal stand-in for the `gym` package (TEST INFRASTRUCTURE ONLY).

The reference (pfnet/pfrl) imports `gym` in pfrl/wrappers/*.py and
pfrl/envs/abc.py:2, but gym is not installed in this image and there is no
network.  This shim provides only the names those modules touch at import
time, so that `import pfrl` works when the reference is put on PYTHONPATH by
oracle/refimport.py.  Nothing in the product package imports this.
"""
from . import spaces  # noqa: F401


class Env:
    metadata = {}
    reward_range = (-float("inf"), float("inf"))
    spec = None
    action_space = None
    observation_space = None

    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return [seed]

    @property
    def unwrapped(self):
        return self


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.action_space = getattr(env, "action_space", None)
        self.observation_space = getattr(env, "observation_space", None)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)

    def close(self):
        return self.env.close()

    @property
    def unwrapped(self):
        return self.env.unwrapped


class ObservationWrapper(Wrapper):
    def reset(self, **kwargs):
        return self.observation(self.env.reset(**kwargs))

    def step(self, action):
        o, r, d, i = self.env.step(action)
        return self.observation(o), r, d, i


class RewardWrapper(Wrapper):
    def step(self, action):
        o, r, d, i = self.env.step(action)
        return o, self.reward(r), d, i


class ActionWrapper(Wrapper):
    def step(self, action):
        return self.env.step(self.action(action))


from . import wrappers  # noqa: E402,F401
