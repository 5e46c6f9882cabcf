import torch.nn as nn
import torch.nn.functional as F

from pfrl_b200.initializers import init_chainer_default, init_lecun_normal


class MLP(nn.Module):
    """Fully-connected stack (pfrl/nn/mlp.py:7-36): hidden layers
    LeCun-normal / zero bias, output layer LeCun-normal scaled by
    ``last_wscale``."""

    def __init__(self, in_size, out_size, hidden_sizes, nonlinearity=F.relu, last_wscale=1):
        super().__init__()
        self.in_size = in_size
        self.out_size = out_size
        self.hidden_sizes = hidden_sizes
        self.nonlinearity = nonlinearity
        widths = [in_size] + list(hidden_sizes)
        if hidden_sizes:
            self.hidden_layers = nn.ModuleList(
                nn.Linear(a, b) for a, b in zip(widths[:-1], widths[1:]))
            self.hidden_layers.apply(init_chainer_default)
        self.output = nn.Linear(widths[-1], out_size)
        init_lecun_normal(self.output.weight, scale=last_wscale)
        nn.init.zeros_(self.output.bias)

    def forward(self, x):
        h = x
        if self.hidden_sizes:
            for layer in self.hidden_layers:
                h = self.nonlinearity(layer(h))
        return self.output(h)
