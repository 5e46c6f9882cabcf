import torch
from torch import nn


class SoftmaxCategoricalHead(nn.Module):
    """logits -> Categorical (pfrl/policies/softmax_policy.py:5-7)."""

    def forward(self, logits):
        return torch.distributions.Categorical(logits=logits)
