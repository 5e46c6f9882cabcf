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
