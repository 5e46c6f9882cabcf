from torch import nn


class Lambda(nn.Module):
    """Wrap a callable as a Module (pfrl/nn/lmbda.py)."""

    def __init__(self, lambd):
        super().__init__()
        self.lambd = lambd

    def forward(self, x):
        return self.lambd(x)
