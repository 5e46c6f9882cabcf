import torch.nn as nn
import torch.nn.functional as F

from pfrl_b200.initializers import init_chainer_default, init_lecun_normal


class MLP(nn.Module):
    """Stack of Linear layers with one nonlinearity between them.

    Hidden layers: LeCun-normal weights, zero bias; output layer: LeCun-normal
    scaled by ``last_wscale``, zero bias (the reference's initialisation,
    pfrl/nn/mlp.py:7-36).  Attribute names ``hidden_layers`` / ``output`` match
    the reference so that state_dicts are interchangeable.
    """

    def __init__(self, in_size, out_size, hidden_sizes, nonlinearity=F.relu, last_wscale=1,
                 linear_cls=nn.Linear):
        super().__init__()
        self.in_size, self.out_size = in_size, out_size
        self.hidden_sizes = hidden_sizes
        self.nonlinearity = nonlinearity
        widths = [in_size] + list(hidden_sizes)
        if hidden_sizes:
            # all hidden Linear layers are constructed (torch's default init draws from
            # the global generator) BEFORE any of them is re-initialised, and the output
            # layer after that: the reference's order, so that the same torch seed gives
            # the same initial weights (tests/test_checkpoint_interchange_cpu.py)
            self.hidden_layers = nn.ModuleList(
                [linear_cls(a, b) for a, b in zip(widths, widths[1:])])
            for layer in self.hidden_layers:
                init_chainer_default(layer)
        self.output = linear_cls(widths[-1], out_size)
        init_lecun_normal(self.output.weight, scale=last_wscale)
        nn.init.zeros_(self.output.bias)

    def forward(self, x):
        for layer in (self.hidden_layers if self.hidden_sizes else ()):
            x = self.nonlinearity(layer(x))
        return self.output(x)
