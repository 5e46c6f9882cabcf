"""Known-answer tests the reference holds for the functions on the hot path
(SURVEY section 8c), re-expressed against pfrl_b200 on the CPU:

  tests/utils_tests/test_batch_states.py:11-53           batch_states on tuple states
  tests/agents_tests/test_dqn.py:138-208                 compute_[weighted_]value_loss
  tests/agents_tests/test_categorical_dqn.py:59-216      categorical projection
  tests/agents_tests/test_categorical_dqn.py:292-364     categorical value losses
  tests/agents_tests/test_iqn.py:104-167                 quantile-Huber loss, cosine basis
  tests/agents_tests/test_ppo.py:551-586                 minibatch schedule

The fused CUDA versions of the same functions are compared with the
reference's outputs in tests/test_losses_gpu.py / test_agent_loss_parity.py.
"""
import numpy as np
import pytest
import torch
from torch import nn


# ---------------------------------------------------------------- batch_states
def test_batch_states_tuple_observations():
    from pfrl_b200.utils.batch_states import batch_states

    base = np.arange(4).reshape(2, 2)
    states = [(base, 0, np.zeros(1)), (base + 1, 1, np.ones(1))]
    out = batch_states(states, torch.device("cpu"), lambda s: (s[0] * 2, s[1], s[2] * 3))
    assert isinstance(out, tuple) and len(out) == 3
    np.testing.assert_allclose(out[0], [[[0, 2], [4, 6]], [[2, 4], [6, 8]]])
    np.testing.assert_allclose(out[1], [0, 1])
    np.testing.assert_allclose(out[2], [[0], [3]])


# ------------------------------------------------------------ scalar TD losses
def _huber(d):
    return 0.5 * d * d if abs(d) < 1 else abs(d) - 0.5


@pytest.mark.parametrize("batch_accumulator", ["mean", "sum"])
@pytest.mark.parametrize("clip_delta", [True, False])
def test_dqn_value_losses(clip_delta, batch_accumulator):
    from pfrl_b200.agents.dqn import compute_value_loss, compute_weighted_value_loss

    y = torch.tensor([1.0, 2.0, 3.0, 4.0])
    t = torch.tensor([2.1, 2.2, 2.3, 2.4])
    per_sample = torch.tensor([_huber(float(d)) if clip_delta else 0.5 * float(d) ** 2
                               for d in y - t])
    reduce = torch.mean if batch_accumulator == "mean" else torch.sum
    kw = dict(clip_delta=clip_delta, batch_accumulator=batch_accumulator)
    assert abs(float(compute_value_loss(y, t, **kw)) - float(reduce(per_sample))) < 1e-5
    ones = torch.ones(4)
    assert abs(float(compute_weighted_value_loss(y, t, ones, **kw))
               - float(reduce(per_sample))) < 1e-5
    torch.manual_seed(0)
    w = torch.rand(4) * 2
    assert abs(float(compute_weighted_value_loss(y, t, w, **kw))
               - float(reduce(per_sample * w))) < 1e-5


# -------------------------------------------------------- categorical projection
def _project_one_by_one(y, p, z):
    """Atom-by-atom projection onto the support z (the definition, O(B n^2))."""
    out = np.zeros_like(p)
    n = len(z)
    for b in range(len(y)):
        for i in range(n):
            v, mass = y[b, i], p[b, i]
            if v <= z[0]:
                out[b, 0] += mass
            elif v > z[-1]:
                out[b, -1] += mass
            else:
                j = int(np.searchsorted(z, v, side="left")) - 1   # z[j] < v <= z[j+1]
                width = z[j + 1] - z[j]
                out[b, j] += (z[j + 1] - v) / width * mass
                out[b, j + 1] += (v - z[j]) / width * mass
    return out


@pytest.mark.parametrize("v_range", [(-3, -1), (-2, 0), (-2, 1), (0, 1), (1, 5)])
@pytest.mark.parametrize("n_atoms", [2, 5])
@pytest.mark.parametrize("batch_size", [1, 7])
def test_categorical_projection_random(batch_size, n_atoms, v_range):
    from pfrl_b200.agents.categorical_dqn import _apply_categorical_projection

    rng = np.random.RandomState(batch_size * 100 + n_atoms * 10 + v_range[0] + 5)
    z = np.linspace(v_range[0], v_range[1], num=n_atoms, dtype=np.float32)
    y = rng.normal(size=(batch_size, n_atoms)).astype(np.float32)
    p = rng.dirichlet(np.ones(n_atoms), size=batch_size).astype(np.float32)
    want = _project_one_by_one(y, p, z)
    np.testing.assert_allclose(want.sum(axis=1), np.ones(batch_size), atol=1e-5)
    got = _apply_categorical_projection(torch.tensor(y), torch.tensor(p), torch.tensor(z)).numpy()
    np.testing.assert_allclose(got.sum(axis=1), np.ones(batch_size), atol=1e-5)
    np.testing.assert_allclose(got, want, atol=1e-5)


def test_categorical_projection_manual_cases():
    from pfrl_b200.agents.categorical_dqn import _apply_categorical_projection

    z = torch.linspace(-1, 1, 3)
    y = torch.tensor([[-1, 0, 1], [1, -1, 0], [1, 1, 1], [-1, -1, -1], [0, 0, 0],
                      [-0.5, 0, 1], [-0.5, 0, 0.5]], dtype=torch.float32)
    p = torch.tensor([[0.5, 0.2, 0.3]] * 7)
    want = [[0.5, 0.2, 0.3], [0.2, 0.3, 0.5], [0.0, 0.0, 1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0],
            [0.25, 0.45, 0.3], [0.25, 0.6, 0.15]]
    np.testing.assert_allclose(_apply_categorical_projection(y, p, z).numpy(), want, atol=1e-5)
    # delta_z = 2/3 is not exactly representable
    z = torch.linspace(-1, 1, 4)
    y = torch.tensor([[-1, -1, 1, 1], [-1, 0, 1, 1]], dtype=torch.float32)
    p = torch.tensor([[0.5, 0.1, 0.1, 0.3], [0.5, 0.2, 0.0, 0.3]])
    want = [[0.6, 0.0, 0.0, 0.4], [0.5, 0.1, 0.1, 0.3]]
    np.testing.assert_allclose(_apply_categorical_projection(y, p, z).numpy(), want, atol=1e-5)


@pytest.mark.parametrize("batch_accumulator", ["mean", "sum"])
def test_categorical_value_losses(batch_accumulator):
    from pfrl_b200.agents.categorical_dqn import compute_value_loss, compute_weighted_value_loss

    y = np.asarray([[0.1, 0.2, 0.3, 0.4], [0.05, 0.1, 0.2, 0.65]], dtype="f")
    t = np.asarray([[0.2, 0.2, 0.2, 0.4], [0.1, 0.3, 0.3, 0.3]], dtype="f")
    elt = -t * np.log(np.clip(y, 1e-10, 1.0))
    plain = elt.sum(axis=1).mean() if batch_accumulator == "mean" else elt.sum()
    got = compute_value_loss(torch.tensor(elt), batch_accumulator=batch_accumulator)
    assert abs(float(got) - plain) < 1e-5
    got = compute_weighted_value_loss(torch.tensor(elt), 2, torch.ones(2),
                                      batch_accumulator=batch_accumulator)
    assert abs(float(got) - plain) < 1e-5
    w = np.random.RandomState(0).uniform(0, 2, size=2).astype("f")
    want = (elt.sum(axis=1) * w).mean() if batch_accumulator == "mean" else \
        (elt * w[:, None]).sum()
    got = compute_weighted_value_loss(torch.tensor(elt), 2, torch.tensor(w),
                                      batch_accumulator=batch_accumulator)
    assert abs(float(got) - want) < 1e-5


# ----------------------------------------------------------------------- IQN
@pytest.mark.parametrize("n_prime", [1, 7])
@pytest.mark.parametrize("n", [1, 5])
@pytest.mark.parametrize("batch_size", [1, 3])
def test_eltwise_huber_quantile_loss(batch_size, n, n_prime):
    """Over-estimates are charged (1 - tau) x Huber, under-estimates tau x Huber,
    values and gradients."""
    from pfrl_b200.agents.iqn import compute_eltwise_huber_quantile_loss

    torch.manual_seed(batch_size * 100 + n * 10 + n_prime)
    y = torch.randn(batch_size, n, requires_grad=True)
    t = torch.randn(batch_size, n_prime)
    tau = torch.rand(batch_size, n)
    loss = compute_eltwise_huber_quantile_loss(y, t, tau)
    assert loss.shape == (batch_size, n, n_prime)
    yb, tb = torch.broadcast_tensors(y[:, :, None], t[:, None, :])
    huber = nn.functional.smooth_l1_loss(yb, tb, reduction="none")
    scale = torch.where(yb > tb, 1 - tau[:, :, None], tau[:, :, None]).detach()
    want = scale * huber
    assert bool((loss > 0).all())
    torch.testing.assert_close(loss, want, atol=1e-5, rtol=0)
    g_got, = torch.autograd.grad(loss.sum(), y, retain_graph=True)
    g_want, = torch.autograd.grad(want.sum(), y)
    torch.testing.assert_close(g_got, g_want, atol=1e-5, rtol=0)


@pytest.mark.parametrize("n_basis", [1, 7])
@pytest.mark.parametrize("m", [1, 5])
def test_cosine_basis_functions(m, n_basis):
    from pfrl_b200.agents.iqn import cosine_basis_functions

    x = torch.rand(3, m)
    y = cosine_basis_functions(x, n_basis_functions=n_basis)
    assert y.shape == (3, m, n_basis)
    k = torch.arange(1, n_basis + 1, dtype=torch.float32)
    torch.testing.assert_close(y, torch.cos(x[..., None] * k * np.pi), atol=1e-5, rtol=0)


# ----------------------------------------------------------------------- PPO
def _minibatches(dataset, size, epochs):
    from pfrl_b200.agents.ppo import _yield_minibatch_indices

    return list(_yield_minibatch_indices(dataset, size, epochs))


def test_minibatch_schedule_divisible():
    mbs = _minibatches([1, 2, 3, 4], 2, 3)
    flat = sum(mbs, [])
    assert len(mbs) == 6 and len(flat) == 12
    for lo in (0, 4, 8):
        assert set(flat[lo:lo + 4]) == {1, 2, 3, 4}


def test_minibatch_schedule_indivisible():
    mbs = _minibatches([1, 2, 3], 2, 3)
    flat = sum(mbs, [])
    assert len(mbs) == 5 and len(flat) == 10
    assert all(flat[:6].count(v) == 2 for v in (1, 2, 3))      # two full epochs
    assert all(1 <= flat[6:].count(v) <= 2 for v in (1, 2, 3))  # the last one, rounded up


def test_minibatch_schedule_dataset_smaller_than_minibatch():
    mbs = _minibatches([1, 2], 4, 3)
    flat = sum(mbs, [])
    assert len(mbs) == 2 and len(flat) == 8
    assert flat.count(1) == 4 and flat.count(2) == 4
