"""Feature extractors ("phi") that the device gather understands.

The reference applies an arbitrary Python ``phi`` to every observation on the
host (pfrl/utils/batch_states.py:29).  A phi object carrying
``b2rl_obs_mode`` / ``b2rl_obs_scale`` declares the same map in a form the
fused gather kernel applies while it writes the minibatch; any other callable
still works through the host path.
"""
import numpy as np

from .._lib import OBS_RAW, OBS_U8_TO_F32


class ScaleU8:
    """phi(x) = float32(x) * scale for uint8 observations, e.g. the Atari
    ``np.asarray(x, dtype=np.float32) / 255`` of
    examples/atari/train_dqn_batch_ale.py:229-231 (scale = 1/255; the
    float32 product differs from the float32 quotient by <= 1 ulp)."""

    b2rl_obs_mode = OBS_U8_TO_F32

    def __init__(self, scale=1.0 / 255.0):
        self.b2rl_obs_scale = float(np.float32(scale))

    def __call__(self, x):
        return np.asarray(x, dtype=np.float32) * np.float32(self.b2rl_obs_scale)


class RawU8:
    """phi(x) = x for uint8 observations whose ``/ 255`` is folded into the network's
    first layer (nn.fast_conv.NatureConv1 reads uint8 as float(x) * 1/255): the gather
    writes bytes (2.6x less HBM traffic than f32 batches) and the
    convolution expands them in shared memory.  Same numbers as ScaleU8 + f32 conv."""

    b2rl_obs_mode = OBS_RAW
    b2rl_obs_scale = 1.0

    def __call__(self, x):
        return np.asarray(x, dtype=np.uint8)


class Identity:
    """phi(x) = x (already-float observations; raw byte copy on the device)."""

    b2rl_obs_mode = OBS_RAW
    b2rl_obs_scale = 1.0

    def __call__(self, x):
        return x


class AsFloat32(Identity):
    """phi(x) = x.astype(float32) for observations that are stored as float32
    (quickstart's ``phi = lambda x: x.astype(numpy.float32, copy=False)``)."""

    def __call__(self, x):
        return np.asarray(x, dtype=np.float32)
