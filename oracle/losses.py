"""numpy restatement of the loss / GAE arithmetic of the reference hot path
(TEST INFRASTRUCTURE ONLY; checked against tests/golden/losses_*.npz which
oracle/gen_golden_losses.py produced with the real reference)."""
import numpy as np

f32 = np.float32


def huber(d):
    ad = np.abs(d)
    return np.where(ad < 1, f32(0.5) * d * d, ad - f32(0.5)).astype(f32)


def td_loss(q, action, next_q, reward, discount, terminal, weights, clip_delta, mean):
    """pfrl/agents/dqn.py:388-470 + :44-104.  Returns loss, |y-t|, y, t."""
    q = q.astype(f32)
    y = q[np.arange(len(action)), action]
    t = (reward + discount * (f32(1) - terminal) * next_q).astype(f32)
    d = y - t
    l = huber(d) if clip_delta else (d * d / f32(2)).astype(f32)
    if weights is not None:
        s = np.sum(l * weights, dtype=np.float64)
        loss = s / len(y) if mean else s
    else:
        loss = np.mean(l, dtype=np.float64) if mean else np.sum(l, dtype=np.float64)
    return f32(loss), np.abs(d), y, t


def categorical_projection(y, y_probs, z):
    """pfrl/agents/categorical_dqn.py:7-57"""
    B, n = y.shape
    z = z.astype(f32)
    dz = z[1] - z[0]
    yy = np.clip(y.astype(f32), z[0], z[-1])
    bj = np.clip(((yy - z[0]) / dz).astype(f32), 0, n - 1)
    lo, up = np.floor(bj), np.ceil(bj)
    frac = (bj - lo).astype(f32)
    out = np.zeros((B, n), dtype=np.float64)
    for i in range(B):
        np.add.at(out[i], lo[i].astype(int), (y_probs[i] * (f32(1) - frac[i])).astype(f32))
        np.add.at(out[i], up[i].astype(int), (y_probs[i] * frac[i]).astype(f32))
    return out.astype(f32)


def c51_loss(y, next_p, reward, discount, terminal, weights, z, mean):
    """pfrl/agents/categorical_dqn.py:100-204.  Returns loss, per-sample, target."""
    z = z.astype(f32)
    Tz = (reward[:, None] + (f32(1) - terminal[:, None]) * discount[:, None] * z[None]).astype(f32)
    t = categorical_projection(Tz, next_p, z)
    elt = (-t * np.log(np.clip(y, f32(1e-10), f32(1.0)))).astype(f32)
    per = elt.sum(1, dtype=np.float64).astype(f32)
    if weights is not None:
        s = np.dot(per.astype(np.float64), weights.astype(np.float64))
        loss = s / len(per) if mean else s
    else:
        loss = np.mean(per, dtype=np.float64) if mean else np.sum(per, dtype=np.float64)
    return f32(loss), per, t


def quantile_huber(y, t, taus, weights, mean):
    """pfrl/agents/iqn.py:176-250.  y,taus [B,N]; t [B,N'].  Returns loss, mean error."""
    yy, tt, ta = y[:, :, None], t[:, None, :], taus[:, :, None]
    ind = (tt < yy).astype(f32)
    elt = (np.abs(ta - ind) * huber(yy - tt)).astype(np.float64)
    per = elt.mean(2).sum(1)
    err = elt.mean((1, 2))
    if weights is not None:
        s = np.dot(per, weights.astype(np.float64))
        loss = s / len(per) if mean else s
    else:
        loss = per.mean() if mean else per.sum()
    return f32(loss), err.astype(f32)


def gae_segments(reward, nonterminal, v, v_next, cut, gamma, lambd):
    """pfrl/agents/ppo.py:36-53 on [T, E] arrays; cut marks segment ends."""
    T, E = reward.shape
    adv = np.zeros((T, E), dtype=np.float64)
    for e in range(E):
        a = 0.0
        for t in range(T - 1, -1, -1):
            if cut[t, e]:
                a = 0.0
            td = float(reward[t, e]) + gamma * float(nonterminal[t, e]) * float(v_next[t, e]) \
                - float(v[t, e])
            a = td + gamma * lambd * a
            adv[t, e] = a
    return adv.astype(f32), (adv + v.astype(np.float64)).astype(f32)


def ppo_loss(log_prob, entropy, v_pred, log_prob_old, v_pred_old, adv, v_teacher, mean_std,
             clip_eps, clip_eps_vf, value_coef, entropy_coef):
    """pfrl/agents/ppo.py:495,634-671.  Returns total, policy, value, entropy losses."""
    a = adv.astype(f32)
    if mean_std is not None:
        a = ((a - f32(mean_std[0])) / (f32(mean_std[1]) + f32(1e-8))).astype(f32)
    ratio = np.exp((log_prob - log_prob_old).astype(f32))
    lp = -np.mean(np.minimum(ratio * a, np.clip(ratio, 1 - clip_eps, 1 + clip_eps) * a),
                  dtype=np.float64)
    if clip_eps_vf is None:
        lv = np.mean((v_pred - v_teacher) ** 2, dtype=np.float64)
    else:
        vc = np.minimum(np.maximum(v_pred, v_pred_old - clip_eps_vf), v_pred_old + clip_eps_vf)
        lv = np.mean(np.maximum((v_pred - v_teacher) ** 2, (vc - v_teacher) ** 2),
                     dtype=np.float64)
    le = -np.mean(entropy, dtype=np.float64)
    return lp + value_coef * lv + entropy_coef * le, lp, lv, le
