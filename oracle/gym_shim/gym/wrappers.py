"""gym.wrappers stand-in (test infrastructure only)."""
from . import Wrapper


class TimeLimit(Wrapper):
    def __init__(self, env, max_episode_steps=None):
        super().__init__(env)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = 0

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)

    def step(self, action):
        o, r, d, i = self.env.step(action)
        self._elapsed_steps += 1
        if self._max_episode_steps is not None and self._elapsed_steps >= self._max_episode_steps:
            i = dict(i)
            i["TimeLimit.truncated"] = not d
            d = True
        return o, r, d, i
