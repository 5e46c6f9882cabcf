"""RMSpropEpsInsideSqrt: the multi-tensor implementation against the formula
written out per element in numpy (plain, centered, momentum, weight decay),
state-dict compatibility with torch.optim.RMSprop, and -- when the reference
tree is present -- the reference's own class on a torch that still runs it."""
import numpy as np
import pytest
import torch


def _numpy_reference(w, grads, lr, alpha, eps, weight_decay, momentum, centered):
    w = w.astype(np.float64).copy()
    sq = np.zeros_like(w)
    ga = np.zeros_like(w)
    buf = np.zeros_like(w)
    for g in grads:
        g = g.astype(np.float64) + weight_decay * w
        sq = alpha * sq + (1 - alpha) * g * g
        if centered:
            ga = alpha * ga + (1 - alpha) * g
            denom = np.sqrt(sq - ga * ga + eps)
        else:
            denom = np.sqrt(sq + eps)
        if momentum > 0:
            buf = momentum * buf + g / denom
            w = w - lr * buf
        else:
            w = w - lr * g / denom
    return w


@pytest.mark.parametrize("centered", [False, True])
@pytest.mark.parametrize("momentum", [0.0, 0.9])
@pytest.mark.parametrize("weight_decay", [0.0, 0.01])
def test_matches_the_formula(centered, momentum, weight_decay):
    from pfrl_b200.optimizers import RMSpropEpsInsideSqrt

    rng = np.random.RandomState(0)
    shapes = [(5, 3), (7,), ()]
    params = [torch.nn.Parameter(torch.tensor(rng.randn(*s), dtype=torch.float64)) for s in shapes]
    start = [p.detach().numpy().copy() for p in params]
    opt = RMSpropEpsInsideSqrt(params, lr=2.5e-4, alpha=0.95, eps=1e-2, momentum=momentum,
                               centered=centered, weight_decay=weight_decay)
    history = [[] for _ in shapes]
    for step in range(6):
        for i, p in enumerate(params):
            g = rng.randn(*shapes[i])
            history[i].append(np.asarray(g))
            p.grad = torch.tensor(g, dtype=torch.float64)
        opt.step()
    for i, p in enumerate(params):
        want = _numpy_reference(start[i], history[i], 2.5e-4, 0.95, 1e-2, weight_decay, momentum,
                                centered)
        np.testing.assert_allclose(p.detach().numpy(), want, rtol=1e-12, atol=1e-14)
    assert all(int(s["step"]) == 6 for s in opt.state.values())


def test_eps_placement_and_state_dict_compatibility():
    from pfrl_b200.optimizers import RMSpropEpsInsideSqrt, SharedRMSpropEpsInsideSqrt

    w = torch.nn.Parameter(torch.ones(3))
    opt = RMSpropEpsInsideSqrt([w], lr=0.1, alpha=0.0, eps=1.0)
    w.grad = torch.zeros(3)
    w.grad[0] = 3.0
    opt.step()
    # alpha = 0: E[g^2] = g^2 -> step = lr * g / sqrt(g^2 + eps) = 0.1 * 3 / sqrt(10)
    np.testing.assert_allclose(w.detach().numpy(), [1 - 0.3 / np.sqrt(10.0), 1, 1], rtol=1e-6)
    stock = torch.optim.RMSprop([torch.nn.Parameter(torch.ones(3))], lr=0.1, alpha=0.0, eps=1.0)
    stock.load_state_dict(opt.state_dict())       # same state layout as torch's RMSprop
    assert set(stock.state_dict()["state"][0]) >= {"step", "square_avg"}
    shared = SharedRMSpropEpsInsideSqrt([torch.nn.Parameter(torch.ones(2))], lr=0.1, centered=True,
                                        momentum=0.5)
    st = list(shared.state.values())[0]
    assert set(st) == {"step", "square_avg", "momentum_buffer", "grad_avg"}
    p = torch.nn.Parameter(torch.ones(2))
    skip = RMSpropEpsInsideSqrt([p], lr=0.1)
    assert skip.step(lambda: torch.tensor(1.5)) == 1.5 and torch.equal(p, torch.ones(2))
