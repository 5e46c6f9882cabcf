from contextlib import contextmanager


@contextmanager
def evaluating(net):
    """Put ``net`` in eval mode for the block and restore its previous mode
    afterwards (pfrl/utils/contexts.py:4-13)."""
    was_training = net.training
    try:
        net.eval()
        yield net
    finally:
        if was_training:
            net.train()
