"""RMSprop with epsilon INSIDE the square root, ``g / sqrt(E[g^2] + eps)``:
the optimizer of the Nature-DQN and A3C/A2C papers, used by the reference's
DQN and A2C reproduction scripts (pfrl/optimizers/rmsprop_eps_inside_sqrt.py).

State and hyper-parameter names are ``torch.optim.RMSprop``'s (the class is a
subclass, so optimizer checkpoints interchange); only ``step`` differs.  The
update runs as multi-tensor (``torch._foreach_*``) operations over all
parameters of a group: a handful of launches per step instead of ~6 per
parameter tensor, which is what matters for the small networks of this path
on a GPU.
"""
import torch


class RMSpropEpsInsideSqrt(torch.optim.RMSprop):
    def _state_of(self, p, group):
        state = self.state[p]
        if len(state) == 0:
            state["step"] = 0
            state["square_avg"] = torch.zeros_like(p)
            if group["momentum"] > 0:
                state["momentum_buffer"] = torch.zeros_like(p)
            if group["centered"]:
                state["grad_avg"] = torch.zeros_like(p)
        return state

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            if any(p.grad.is_sparse for p in params):
                raise RuntimeError("RMSprop does not support sparse gradients")
            states = [self._state_of(p, group) for p in params]
            for st in states:
                st["step"] += 1
            grads = [p.grad for p in params]
            alpha, eps, lr = group["alpha"], group["eps"], group["lr"]
            if group["weight_decay"] != 0:
                grads = torch._foreach_add(grads, params, alpha=group["weight_decay"])
            square_avg = [st["square_avg"] for st in states]
            torch._foreach_mul_(square_avg, alpha)
            torch._foreach_addcmul_(square_avg, grads, grads, value=1 - alpha)
            if group["centered"]:
                grad_avg = [st["grad_avg"] for st in states]
                torch._foreach_mul_(grad_avg, alpha)
                torch._foreach_add_(grad_avg, grads, alpha=1 - alpha)
                denom = torch._foreach_addcmul(square_avg, grad_avg, grad_avg, value=-1)
                torch._foreach_add_(denom, eps)
            else:
                denom = torch._foreach_add(square_avg, eps)
            torch._foreach_sqrt_(denom)
            if group["momentum"] > 0:
                bufs = [st["momentum_buffer"] for st in states]
                torch._foreach_mul_(bufs, group["momentum"])
                torch._foreach_addcdiv_(bufs, grads, denom)
                torch._foreach_add_(params, bufs, alpha=-lr)
            else:
                torch._foreach_addcdiv_(params, grads, denom, value=-lr)
        return loss


class SharedRMSpropEpsInsideSqrt(RMSpropEpsInsideSqrt):
    """Same, with the state allocated at construction (so that it can be moved
    to shared memory before worker processes start)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        for group in self.param_groups:
            for p in group["params"]:
                self._state_of(p, group)
