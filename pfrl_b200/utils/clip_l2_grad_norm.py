import torch


def clip_l2_grad_norm_(parameters, max_norm):
    """Clip the global L2 norm of the gradients in place and return the norm
    before clipping (pfrl/utils/clip_l2_grad_norm.py:5-31)."""
    return torch.nn.utils.clip_grad_norm_(parameters, max_norm, norm_type=2)
