"""One-process-per-GPU data parallelism for the agents.

The reference has no distributed training at all.
Here every rank owns its own vector-env shard, HBM replay shard, priority
trees and sampler RNG stream; replay contents, priorities and indices are
never exchanged.  The only collectives are
  * one all-reduce (sum, then / world) of the flat gradient bucket per
    optimizer step (NCCL over NVLink / NVSwitch; gloo on CPU for tests),
  * for PPO, an all-reduce of the advantage moments (count, sum, sum of
    squares) so that every rank standardises with the global mean / std,
  * a parameter broadcast from rank 0 at start-up.
"""
import os

import torch
import torch.distributed as dist


def init(backend=None, device=None):
    """Initialise torch.distributed from the torchrun environment
    (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).  Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, **kw)
    return rank, world


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def broadcast_parameters(module, src=0):
    """Make every rank start from rank ``src``'s parameters and buffers."""
    if world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


class GradSync:
    """Average gradients across ranks with ONE all-reduce of a flat bucket.

    Usage: ``agent.grad_sync = GradSync()``; the agents call it with the
    module between ``backward()`` and ``optimizer.step()``.  Gradients are
    packed with multi-tensor copies into a persistent flat buffer, reduced,
    scaled by 1 / world and unpacked (two small launches around the
    collective; at <= 30 MB the all-reduce is latency-bound, so one bucket
    beats per-tensor calls)."""

    def __init__(self):
        self._buckets = {}

    def _bucket(self, module):
        key = id(module)
        params = [p for p in module.parameters() if p.requires_grad]
        b = self._buckets.get(key)
        n = sum(p.numel() for p in params)
        if b is None or b[0].numel() != n:
            flat = torch.zeros(n, dtype=params[0].dtype, device=params[0].device)
            views, off = [], 0
            for p in params:
                views.append(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
            b = (flat, views, params)
            self._buckets[key] = b
        return b

    def all_ready(self, flag):
        """True iff `flag` is true on EVERY rank (all-reduce MIN of one scalar).  The
        replay updater calls it at update opportunities until it returns True once, so
        that all ranks take their first (and every later) optimizer step together."""
        if world_size() == 1:
            return bool(flag)
        dev = (torch.device("cuda", torch.cuda.current_device())
               if dist.get_backend() == "nccl" else torch.device("cpu"))
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(int(t.item()))

    # The three stages are separate so that an agent can replay pack / unpack inside CUDA graphs
    # and issue only the collective itself eagerly between them (agents/dqn.py).
    def pack(self, module):
        """Gradients -> the flat bucket (one multi-tensor copy)."""
        flat, views, params = self._bucket(module)
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
        torch._foreach_copy_(views, grads)

    def reduce(self, module):
        """The collective: SUM over ranks of the flat bucket."""
        dist.all_reduce(self._bucket(module)[0], op=dist.ReduceOp.SUM)

    def unpack(self, module):
        """Bucket / world -> the gradients."""
        flat, views, params = self._bucket(module)
        flat.div_(world_size())
        for p, v in zip(params, views):
            if p.grad is None:
                p.grad = v.clone()
        torch._foreach_copy_([p.grad for p in params], views)

    def __call__(self, module):
        if world_size() == 1:
            return
        self.pack(module)
        self.reduce(module)
        self.unpack(module)


def sync_advantage_stats(adv):
    """Global (mean, std unbiased=False) of the advantages over all ranks,
    from all-reduced (count, sum, sum of squares) in fp64."""
    a = adv.double().reshape(-1)
    m = torch.stack([torch.tensor(float(a.numel()), device=a.device, dtype=torch.float64),
                     a.sum(), (a * a).sum()])
    if world_size() > 1:
        dist.all_reduce(m, op=dist.ReduceOp.SUM)
    mean = m[1] / m[0]
    var = torch.clamp(m[2] / m[0] - mean * mean, min=0.0)
    return torch.stack([mean, var.sqrt()]).float()
