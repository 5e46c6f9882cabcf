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

    def __init__(self, in_size, out_size, hidden_sizes, nonlinearity=F.relu, last_wscale=1):
        super().__init__()
        self.in_size, self.out_size = in_size, out_size
        self.hidden_sizes = hidden_sizes
        self.nonlinearity = nonlinearity
        fan_in = in_size
        if hidden_sizes:
            layers = []
            for width in hidden_sizes:
                layers.append(init_chainer_default(nn.Linear(fan_in, width)))
                fan_in = width
            self.hidden_layers = nn.ModuleList(layers)
        self.output = nn.Linear(fan_in, out_size)
        init_lecun_normal(self.output.weight, scale=last_wscale)
        nn.init.zeros_(self.output.bias)

    def forward(self, x):
        for layer in (self.hidden_layers if self.hidden_sizes else ()):
            x = self.nonlinearity(layer(x))
        return self.output(x)
