"""GPU: K9 (csrc/sac.cu) -- multi-tensor Polyak averaging and the SAC TD target are bit
for bit the reference's sequence of separate fp32 operations
(pfrl/utils/copy_param.py:9-22, pfrl/agents/soft_actor_critic.py:225-240)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_polyak_matches_reference_sequence():
    from pfrl_b200.utils.copy_param import soft_copy_param

    torch.manual_seed(0)
    mk = lambda: torch.nn.Sequential(  # noqa: E731
        torch.nn.Linear(23, 256), torch.nn.ReLU(), torch.nn.BatchNorm1d(256),
        torch.nn.Linear(256, 256), torch.nn.ReLU(), torch.nn.Linear(256, 1)).cuda()
    src, dst, ref = mk(), mk(), mk()
    ref.load_state_dict(dst.state_dict())
    for tau in (5e-3, 0.01, 0.37, 1.0, 0.0):
        # the reference's loop
        td, sd = ref.state_dict(), src.state_dict()
        for k, tv in td.items():
            sv = sd[k]
            if sv.dtype in (torch.float32, torch.float64, torch.float16):
                tv.mul_(1 - tau)
                tv.add_(tau * sv)
        soft_copy_param(dst, src, tau)
        for (k, a), b in zip(dst.state_dict().items(), ref.state_dict().values()):
            if a.dtype == torch.float32:
                assert torch.equal(a, b), (tau, k)


def test_many_small_tensors_and_large_one():
    from pfrl_b200.ops.sac import polyak_

    g = torch.Generator(device="cuda").manual_seed(1)
    shapes = [(1,), (3,), (2049,), (4096,), (7, 13)] * 45 + [(1 << 20,)]   # > 96 pairs
    t = [torch.randn(s, device="cuda", generator=g) for s in shapes]
    s_ = [torch.randn(s, device="cuda", generator=g) for s in shapes]
    want = [a.clone().mul_(1 - 0.005).add_(0.005 * b) for a, b in zip(t, s_)]
    polyak_(t, s_, 0.005)
    for a, b in zip(t, want):
        assert torch.equal(a, b)


@pytest.mark.parametrize("temp_kind", ["float", "tensor"])
def test_sac_target_matches_eager_expression(temp_kind):
    from pfrl_b200.ops.sac import sac_target

    g = torch.Generator(device="cuda").manual_seed(2)
    n = 1024
    r, q1, q2, lp = (torch.randn(n, device="cuda", generator=g) for _ in range(4))
    disc = torch.full((n,), 0.99, device="cuda") ** torch.randint(1, 4, (n,), device="cuda")
    term = (torch.rand(n, device="cuda", generator=g) < 0.1).float()
    temp = 0.2371 if temp_kind == "float" else torch.tensor(0.2371, device="cuda")
    next_q = torch.min(q1[:, None], q2[:, None])
    entropy_term = temp * lp[..., None]
    want = r + disc * (1.0 - term) * torch.flatten(next_q - entropy_term)
    got = sac_target(r, disc, term, q1[:, None], q2[:, None], lp, temp)
    assert torch.equal(got, want)
