"""Parameter copies between networks (pfrl/utils/copy_param.py)."""
import torch


def copy_param(target_link, source_link):
    """Hard copy of every state_dict entry (copy_param.py:4-6)."""
    target_link.load_state_dict(source_link.state_dict())


@torch.no_grad()
def soft_copy_param(target_link, source_link, tau):
    """Polyak averaging ``target = (1 - tau) * target + tau * source`` over
    floating state_dict entries; integer buffers are copied
    (copy_param.py:9-22).  One launch for all tensors on CUDA (b2rl_polyak)."""
    tgt = target_link.state_dict()
    src = source_link.state_dict()
    f_t, f_s = [], []
    for k, tv in tgt.items():
        sv = src[k]
        if tv.dtype in (torch.float16, torch.bfloat16, torch.float32, torch.float64):
            assert tv.shape == sv.shape
            f_t.append(tv)
            f_s.append(sv)
        else:
            tv.copy_(sv)
    if not f_t:
        return
    if all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and s.is_cuda
           and s.dtype == torch.float32 and s.is_contiguous() for t, s in zip(f_t, f_s)):
        # one launch for the whole network (csrc/sac.cu, K9), rounded like the
        # reference's target.mul_(1 - tau); target.add_(tau * source)
        from pfrl_b200.ops.sac import polyak_

        polyak_(f_t, f_s, tau)
    else:
        torch._foreach_mul_(f_t, 1.0 - tau)
        torch._foreach_add_(f_t, f_s, alpha=tau)


def copy_grad(target_link, source_link):
    """Give ``target_link``'s parameters clones of ``source_link``'s gradients
    (``None`` where the source has none)."""
    pairs = zip(target_link.parameters(), source_link.parameters())
    for dst, src in pairs:
        assert dst.shape == src.shape
        dst.grad = None if src.grad is None else src.grad.clone()


def synchronize_parameters(src, dst, method, tau=None):
    """copy_param.py:37-41"""
    if method == "hard":
        copy_param(dst, src)
    elif method == "soft":
        soft_copy_param(dst, src, tau)
    else:
        raise ValueError("unknown target update method: %r" % (method,))
