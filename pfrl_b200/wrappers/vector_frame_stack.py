"""Frame stacking for vector envs with shared frames.

Same behaviour as the reference's pfrl/wrappers/vector_frame_stack.py:54-105:
the observation of env i is a ``LazyFrames`` over its last ``k`` frame OBJECTS,
on reset the first frame is repeated ``k`` times.  Because consecutive
observations share frame objects, the device replay buffers (which
de-duplicate parts by object identity) upload and store every frame once --
the ingestion path for real, host-side environments.
"""
from collections import deque

import numpy as np

from pfrl_b200 import env
from pfrl_b200.utils.lazy_frames import LazyFrames


class VectorEnvWrapper(env.VectorEnv):
    """Forwards everything to the wrapped vector env."""

    def __init__(self, env):
        self.env = env
        self.action_space = getattr(env, "action_space", None)
        self.observation_space = getattr(env, "observation_space", None)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError("attempted to get missing private attribute '{}'".format(name))
        return getattr(self.env, name)

    @property
    def num_envs(self):
        return self.env.num_envs

    def step(self, action):
        return self.env.step(action)

    def reset(self, mask=None):
        return self.env.reset(mask)

    def seed(self, seed=None):
        return self.env.seed(seed)

    def render(self, mode="human", **kwargs):
        return self.env.render(mode, **kwargs)

    def compute_reward(self, achieved_goal, desired_goal, info):
        return self.env.compute_reward(achieved_goal, desired_goal, info)

    def close(self):
        return self.env.close()

    def __str__(self):
        return "<{}{}>".format(type(self).__name__, self.env)

    __repr__ = __str__


class VectorFrameStack(VectorEnvWrapper):
    def __init__(self, env, k, stack_axis=0):
        super().__init__(env)
        self.k = k
        self.stack_axis = stack_axis
        self.frames = [deque([], maxlen=k) for _ in range(env.num_envs)]
        from pfrl_b200.wrappers.atari_wrappers import _stacked_space

        self.observation_space = _stacked_space(self.observation_space, k, stack_axis)

    def _observations(self):
        assert all(len(f) == self.k for f in self.frames)
        return [LazyFrames(list(f), stack_axis=self.stack_axis) for f in self.frames]

    def reset(self, mask=None):
        batch_ob = self.env.reset(mask)
        if mask is None:
            mask = np.zeros(self.env.num_envs)
        for keep, frames, ob in zip(mask, self.frames, batch_ob):
            if not keep:
                frames.extend([ob] * self.k)
        return self._observations()

    def step(self, action):
        batch_ob, rewards, dones, infos = self.env.step(action)
        for frames, ob in zip(self.frames, batch_ob):
            frames.append(ob)
        return self._observations(), rewards, dones, infos
