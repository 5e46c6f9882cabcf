"""Frame stacking for a single env, at the reference's module path
(pfrl/wrappers/atari_wrappers.py: ``LazyFrames`` :251-272, ``FrameStack``
:214-248).  The emulator-specific wrappers of that module (no-op resets, fire
reset, episodic life, max-and-skip, 84x84 warping, reward clipping) need gym /
ALE, which this image does not have; they sit before the path rebuilt here and
are out of scope.
"""
import collections

import numpy as np

from pfrl_b200.utils.lazy_frames import LazyFrames  # NOQA


def _stacked_space(space, k, axis):
    """Observation space of k stacked frames, for Box-like spaces (duck-typed:
    ``low`` / ``high`` / ``dtype``); anything else is passed through."""
    if space is None or not hasattr(space, "low") or not hasattr(space, "high"):
        return space
    try:
        return type(space)(low=np.repeat(space.low, k, axis=axis),
                           high=np.repeat(space.high, k, axis=axis), dtype=space.dtype)
    except Exception:
        return space


class FrameStack(object):
    """Observation = the last ``k`` frames as one LazyFrames (frames shared
    between consecutive observations, so a replay buffer stores each once).
    ``channel_order``: "hwc" stacks along the last axis, "chw" along the first."""

    def __init__(self, env, k, channel_order="hwc"):
        self.env = env
        self.k = k
        self.stack_axis = {"hwc": 2, "chw": 0}[channel_order]
        self.frames = collections.deque([], maxlen=k)
        self.action_space = getattr(env, "action_space", None)
        self.observation_space = _stacked_space(
            getattr(env, "observation_space", None), k, self.stack_axis)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def _observation(self):
        assert len(self.frames) == self.k
        return LazyFrames(list(self.frames), stack_axis=self.stack_axis)

    def reset(self):
        first = self.env.reset()
        self.frames.extend([first] * self.k)
        return self._observation()

    def step(self, action):
        frame, reward, done, info = self.env.step(action)
        self.frames.append(frame)
        return self._observation(), reward, done, info
