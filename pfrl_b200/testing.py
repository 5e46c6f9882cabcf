"""Test helper shared by the suites (reference: pfrl/testing.py): an
``assert_allclose`` that accepts tensors and nested lists / tuples of them."""
import numpy as np
import torch


def _to_numpy(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    if isinstance(x, (list, tuple)):
        return np.asarray([_to_numpy(item) for item in x])
    return x


def torch_assert_allclose(actual, desired, *args, **kwargs):
    np.testing.assert_allclose(_to_numpy(actual), _to_numpy(desired), *args, **kwargs)
