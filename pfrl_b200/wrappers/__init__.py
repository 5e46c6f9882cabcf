from pfrl_b200.wrappers.vector_frame_stack import VectorEnvWrapper, VectorFrameStack  # NOQA
from pfrl_b200.wrappers import atari_wrappers  # NOQA
