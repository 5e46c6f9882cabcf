import numpy as np


class LazyFrames(object):
    """Observation made of shared frame arrays, concatenated on demand.

    Same role as pfrl/wrappers/atari_wrappers.py:251-272: consecutive
    frame-stacked observations share their frame objects, so a replay buffer
    stores every frame once.  The device buffers recognise the ``_frames``
    attribute and de-duplicate frames by identity.
    """

    def __init__(self, frames, stack_axis=2):  # the reference's default (hwc frames)
        self.stack_axis = stack_axis
        self._frames = frames

    def __array__(self, dtype=None, copy=None):
        out = np.concatenate(self._frames, axis=self.stack_axis)
        if dtype is not None:
            out = out.astype(dtype)
        return out
