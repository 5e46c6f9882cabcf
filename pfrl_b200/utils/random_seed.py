import random

import numpy as np
import torch


def set_random_seed(seed):
    """Seed ``random``, ``numpy.random`` (global legacy stream) and torch
    (CPU and, if present, CUDA); same three streams the reference seeds in
    pfrl/utils/random_seed.py:7-22."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
