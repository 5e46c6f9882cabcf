"""CPU: oracle/losses.py reproduces the reference's loss / GAE outputs
(tests/golden/losses.npz, generated from the real reference)."""
import os

import numpy as np
import pytest

from oracle import losses as ol

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "losses.npz"))


@pytest.mark.parametrize("clip", [True, False])
@pytest.mark.parametrize("acc", ["mean", "sum"])
@pytest.mark.parametrize("use_w", [True, False])
def test_td_loss(clip, acc, use_w):
    loss, delta, y, t = ol.td_loss(G["td_q"], G["td_action"], G["td_next_q"], G["td_reward"],
                                   G["td_discount"], G["td_terminal"],
                                   G["td_weights"] if use_w else None, clip, acc == "mean")
    np.testing.assert_allclose(loss, G["td_%d_%s_%d_loss" % (clip, acc, use_w)], rtol=1e-6)
    np.testing.assert_allclose(delta, G["td_delta"], rtol=1e-6, atol=1e-6)


def test_projection_kat():
    out = ol.categorical_projection(G["proj_y"], G["proj_p"], G["proj_z"])
    np.testing.assert_allclose(out, G["proj_out"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("acc", ["mean", "sum"])
@pytest.mark.parametrize("use_w", [True, False])
def test_c51(acc, use_w):
    loss, per, t = ol.c51_loss(G["c51_y"], G["c51_next_p"], G["c51_reward"], G["c51_discount"],
                               G["c51_terminal"], G["c51_weights"] if use_w else None,
                               G["c51_z"], acc == "mean")
    np.testing.assert_allclose(t, G["c51_target"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(per, G["c51_delta"], rtol=1e-5)
    np.testing.assert_allclose(loss, G["c51_%s_%d_loss" % (acc, use_w)], rtol=1e-5)


@pytest.mark.parametrize("acc", ["mean", "sum"])
@pytest.mark.parametrize("use_w", [True, False])
def test_quantile_huber(acc, use_w):
    loss, err = ol.quantile_huber(G["qh_y"], G["qh_t"], G["qh_taus"],
                                  G["qh_weights"] if use_w else None, acc == "mean")
    np.testing.assert_allclose(loss, G["qh_%s_%d_loss" % (acc, use_w)], rtol=1e-5)
    np.testing.assert_allclose(err, G["qh_delta"], rtol=1e-5)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_gae(tag):
    gamma, lambd = G["gae_%s_params" % tag]
    adv, vt = ol.gae_segments(G["gae_reward"], G["gae_nonterminal"], G["gae_v"],
                              G["gae_v_next"], G["gae_cut"], float(gamma), float(lambd))
    np.testing.assert_allclose(adv, G["gae_%s_adv" % tag], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(vt, G["gae_%s_vt" % tag], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("tag,clip_vf", [("a", None), ("b", 0.2)])
def test_ppo_loss(tag, clip_vf):
    tot, lp, lv, le = ol.ppo_loss(G["ppo_lp"], G["ppo_ent"], G["ppo_v"], G["ppo_lp_old"],
                                  G["ppo_v_old"], G["ppo_adv"], G["ppo_vt"], G["ppo_mean_std"],
                                  0.2, clip_vf, 0.5, 0.01)
    np.testing.assert_allclose(tot, G["ppo_%s_loss" % tag], rtol=1e-5)
    np.testing.assert_allclose(lp, G["ppo_%s_policy" % tag], rtol=1e-5)
    np.testing.assert_allclose(lv, G["ppo_%s_value" % tag], rtol=1e-5)
