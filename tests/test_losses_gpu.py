"""GPU: fused loss / GAE kernels (through the C ABI) vs the reference's own
outputs (tests/golden/losses.npz).  fp32 tolerance 1e-5 (north_star)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "losses.npz"))
TOL = dict(rtol=1e-5, atol=1e-6)


def cu(x, dtype=None):
    t = torch.as_tensor(np.asarray(x)).cuda()
    return t if dtype is None else t.to(dtype)


@pytest.mark.parametrize("clip", [True, False])
@pytest.mark.parametrize("acc", ["mean", "sum"])
@pytest.mark.parametrize("use_w", [True, False])
def test_td_loss_fwd_bwd(clip, acc, use_w):
    from pfrl_b200.ops.losses import td_loss

    q = cu(G["td_q"]).requires_grad_(True)
    loss, delta, y, t = td_loss(q, cu(G["td_action"]), cu(G["td_next_q"]), cu(G["td_reward"]),
                                cu(G["td_discount"]), cu(G["td_terminal"]),
                                cu(G["td_weights"]) if use_w else None, clip, acc == "mean")
    key = "td_%d_%s_%d" % (clip, acc, use_w)
    np.testing.assert_allclose(loss.item(), G[key + "_loss"], **TOL)
    np.testing.assert_allclose(delta.cpu().numpy(), G["td_delta"], **TOL)
    (loss * 1.0).backward()
    np.testing.assert_allclose(q.grad.cpu().numpy(), G[key + "_grad"], **TOL)


def test_projection_kat():
    from pfrl_b200.ops.losses import c51_loss

    n = G["proj_z"].shape[0]
    y = torch.full((1, n), 1.0 / n, device="cuda")
    # reward 0, discount 1, z replaced by the atoms to project: feed Tz = proj_y
    # through reward = 0 and an explicit support is not possible, so use the
    # identity Tz = r + disc * z with z = proj_z and compare the target row of
    # the shifted problem instead: r = 0.1, disc = 0.9
    z = cu(G["proj_z"])
    r, d = 0.1, 0.9
    p = cu(G["proj_p"])
    _, _, t = c51_loss(y, p, cu([r], torch.float32), cu([d], torch.float32),
                       cu([0.0], torch.float32), None, z=z, return_target=True)
    from oracle.losses import categorical_projection

    Tz = (np.float32(r) + np.float32(d) * G["proj_z"]).astype(np.float32)[None]
    ref = categorical_projection(Tz, G["proj_p"], G["proj_z"])
    np.testing.assert_allclose(t.cpu().numpy(), ref, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("acc", ["mean", "sum"])
@pytest.mark.parametrize("use_w", [True, False])
def test_c51_fwd_bwd(acc, use_w):
    from pfrl_b200.ops.losses import c51_loss

    y = cu(G["c51_y"]).requires_grad_(True)
    loss, delta, t = c51_loss(y, cu(G["c51_next_p"]), cu(G["c51_reward"]), cu(G["c51_discount"]),
                              cu(G["c51_terminal"]), cu(G["c51_weights"]) if use_w else None,
                              z=cu(G["c51_z"]), mean=acc == "mean", return_target=True)
    key = "c51_%s_%d" % (acc, use_w)
    np.testing.assert_allclose(t.cpu().numpy(), G["c51_target"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(delta.cpu().numpy(), G["c51_delta"], **TOL)
    np.testing.assert_allclose(loss.item(), G[key + "_loss"], **TOL)
    loss.backward()
    np.testing.assert_allclose(y.grad.cpu().numpy(), G[key + "_grad"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("acc", ["mean", "sum"])
@pytest.mark.parametrize("use_w", [True, False])
def test_quantile_huber_fwd_bwd(acc, use_w):
    from pfrl_b200.ops.losses import quantile_huber_loss

    y = cu(G["qh_y"]).requires_grad_(True)
    loss, err = quantile_huber_loss(y, cu(G["qh_t"]), cu(G["qh_taus"]),
                                    cu(G["qh_weights"]) if use_w else None, acc == "mean")
    key = "qh_%s_%d" % (acc, use_w)
    np.testing.assert_allclose(loss.item(), G[key + "_loss"], **TOL)
    np.testing.assert_allclose(err.cpu().numpy(), G["qh_delta"], **TOL)
    loss.backward()
    np.testing.assert_allclose(y.grad.cpu().numpy(), G[key + "_grad"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_gae(tag):
    from pfrl_b200.ops.ppo import gae

    gamma, lambd = [float(x) for x in G["gae_%s_params" % tag]]
    adv, vt, stats = gae(cu(G["gae_reward"], torch.float32), cu(G["gae_nonterminal"], torch.float32),
                         cu(G["gae_v"]), cu(G["gae_v_next"]), cu(G["gae_cut"]), gamma, lambd)
    # rewards were fp64 in the reference run; the kernel takes fp32 rewards
    np.testing.assert_allclose(adv.cpu().numpy(), G["gae_%s_adv" % tag], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(vt.cpu().numpy(), G["gae_%s_vt" % tag], rtol=1e-5, atol=2e-5)
    a = adv.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(stats.cpu().numpy(), [a.mean(), a.std()], rtol=1e-5)


def test_gae_full_size_properties():
    """C4 shape (256 envs x 2048 steps): linearity in the rewards and the
    lambda=0 / lambda=1 closed forms (size-independent properties)."""
    from pfrl_b200.ops.ppo import gae

    T, E = 2048, 256
    g = torch.Generator(device="cuda").manual_seed(0)
    r = torch.randn(T, E, device="cuda", generator=g)
    v = torch.randn(T, E, device="cuda", generator=g)
    vn = torch.randn(T, E, device="cuda", generator=g)
    nt = (torch.rand(T, E, device="cuda", generator=g) > 0.001).float()
    cut = (nt == 0)
    cut[-1] = True
    gamma = 0.995
    a0, vt0, _ = gae(r, nt, v, vn, cut, gamma, 0.0)
    torch.testing.assert_close(a0, r + gamma * nt * vn - v, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(vt0, a0 + v, rtol=1e-5, atol=1e-5)
    a1, _, _ = gae(r, nt, v, vn, cut, gamma, 0.95)
    a2, _, _ = gae(2 * r, nt, 2 * v, 2 * vn, cut, gamma, 0.95)
    torch.testing.assert_close(a2, 2 * a1, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("tag,clip_vf", [("a", None), ("b", 0.2)])
def test_ppo_loss_fwd_bwd(tag, clip_vf):
    from pfrl_b200.ops.ppo import ppo_loss

    lp = cu(G["ppo_lp"]).requires_grad_(True)
    ent = cu(G["ppo_ent"]).requires_grad_(True)
    v = cu(G["ppo_v"]).requires_grad_(True)
    loss, parts = ppo_loss(lp, ent, v, cu(G["ppo_lp_old"]), cu(G["ppo_v_old"]), cu(G["ppo_adv"]),
                           cu(G["ppo_vt"]), cu(G["ppo_mean_std"]), 0.2, clip_vf, 0.5, 0.01)
    np.testing.assert_allclose(loss.item(), G["ppo_%s_loss" % tag], **TOL)
    np.testing.assert_allclose(parts[1].item(), G["ppo_%s_policy" % tag], **TOL)
    np.testing.assert_allclose(parts[2].item(), G["ppo_%s_value" % tag], **TOL)
    loss.backward()
    np.testing.assert_allclose(lp.grad.cpu().numpy(), G["ppo_%s_g_lp" % tag], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(ent.grad.cpu().numpy(), G["ppo_%s_g_ent" % tag], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(v.grad.cpu().numpy(), G["ppo_%s_g_v" % tag], rtol=1e-4, atol=1e-7)
