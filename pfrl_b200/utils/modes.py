import contextlib


@contextlib.contextmanager
def evaluating(net):
    """``with evaluating(net):`` runs the block with ``net`` in eval mode and
    puts it back in training mode afterwards if that is where it was
    (reference: pfrl/utils/contexts.py)."""
    restore = bool(net.training)
    net.eval()
    try:
        yield net
    finally:
        if restore:
            net.train()


@contextlib.contextmanager
def no_distribution_validation():
    """torch.distributions validates constructor arguments and samples with a
    host-side ``.all()`` -- a device synchronisation, illegal inside CUDA graph
    capture.  Switch it off for the block (the values come from our own
    networks and were validated during the eager warm-up steps)."""
    import torch

    previous = torch.distributions.Distribution._validate_args
    torch.distributions.Distribution.set_default_validate_args(False)
    try:
        yield
    finally:
        torch.distributions.Distribution.set_default_validate_args(previous)
