import torch


class Branched(torch.nn.Module):
    """Feed one input to several child modules and return the tuple of their
    outputs (pfrl/nn/branched.py), e.g. (policy head, value head) for PPO."""

    def __init__(self, *modules):
        super().__init__()
        self.child_modules = torch.nn.ModuleList(modules)

    def forward(self, *args, **kwargs):
        return tuple(mod(*args, **kwargs) for mod in self.child_modules)
