"""GPU: the CUDA replay path at the HEADLINE shapes (BASELINE configs[2] and [1]:
1 M-transition PER, 84x84x4 uint8 frames, 3-step / B = 512 and 1-step / B = 32)
against what the REAL reference produced (tests/golden/headline_*.npz, written by
oracle/gen_golden_headline.py): sampled indices bit-identical, importance weights,
n-step reward / discount / terminal / action, and integer checksums of every
gathered state / next_state -- through the public API + C ABI, with the separate
launches and with the fused single-launch step."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
NPIX = 84 * 84
STACK = 4


def _frame_pixels(fid):
    i = np.arange(NPIX, dtype=np.int64)
    return ((i * 31 + fid * 17 + (fid * fid) % 251 + (i * fid) % 7) % 256).astype(np.uint8)


def _frames_cuda(f0, n, dev):
    out = torch.empty((n, 84, 84), dtype=torch.uint8, device=dev)
    i = torch.arange(NPIX, device=dev, dtype=torch.int64)[None, :]
    for lo in range(0, n, 8192):
        hi = min(n, lo + 8192)
        fid = torch.arange(f0 + lo, f0 + hi, device=dev, dtype=torch.int64)[:, None]
        x = (i * 31 + fid * 17 + (fid * fid) % 251 + (i * fid) % 7) % 256
        out[lo:hi] = x.to(torch.uint8).view(-1, 84, 84)
    return out


def _script(seed, total, chunk):
    rng = np.random.RandomState(seed)
    acts = rng.randint(0, 18, size=total).astype(np.int64)
    rews = rng.randint(-1, 2, size=total).astype(np.float64)
    term = rng.rand(total) < 1e-3
    for end in range(chunk, total + chunk, chunk):
        term[min(end, total) - 1] = True
    return acts, rews, term


def _checksums_u8(x):
    x = x.reshape(x.shape[0], -1).to(torch.int64)
    w = (torch.arange(x.shape[1], device=x.device, dtype=torch.int64) % 251) + 1
    return x.sum(1).cpu().numpy(), (x * w).sum(1).cpu().numpy()


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("name", ["headline_c3_rainbow", "headline_c2_dqn"])
def test_headline_shape_matches_reference(name, fused):
    from pfrl_b200.replay_buffer import batch_experiences
    from pfrl_b200.replay_buffers import PrioritizedReplayBuffer
    from pfrl_b200.utils.lazy_frames import LazyFrames
    from pfrl_b200.utils.phi import ScaleU8

    path = os.path.join(GOLD, name + ".npz")
    g = np.load(path, allow_pickle=False)
    seed, capacity, num_steps, batch, extra, rounds, chunk, between = [int(x) for x in g["meta"]]
    alpha, beta0, betasteps, gamma = [float(x) for x in g["params"]]
    nb = str(g["normalize_by_max"])
    nb = {"True": True, "False": False}.get(nb, nb)
    dev = torch.device("cuda")
    buf = PrioritizedReplayBuffer(
        capacity, alpha=alpha, beta0=beta0, betasteps=None if np.isnan(betasteps) else betasteps,
        normalize_by_max=nb, num_steps=num_steps, part_capacity=capacity + 16384,
        max_batch=max(batch, 512), fused=fused)
    total = capacity + extra
    acts, rews, term = _script(seed, total, chunk)
    fid = t = 0
    while t < total:
        m = min(chunk, total - t)
        buf.append_trajectory(_frames_cuda(fid, m + STACK, dev), acts[t:t + m], rews[t:t + m],
                              term[t:t + m])
        fid += m + STACK
        t += m
    assert len(buf) == capacity

    np.random.seed(seed)
    phi = ScaleU8()

    def host_frame(f):
        return _frame_pixels(f).reshape(1, 84, 84)

    frames = [host_frame(fid + j) for j in range(STACK)]
    fid += STACK
    cur = LazyFrames(list(frames), stack_axis=0)
    bi = 0
    for r in range(rounds):
        exps = buf.sample(batch)
        b = batch_experiences(exps, dev, phi, gamma)
        assert np.array_equal(exps.index.cpu().numpy(), g["idx"][r]), "round %d indices" % r
        np.testing.assert_allclose(b["weights"].cpu().numpy(), g["weight"][r], rtol=2e-6)
        np.testing.assert_allclose(b["reward"].cpu().numpy(), g["reward"][r], rtol=1e-6, atol=1e-7)
        assert np.array_equal(b["discount"].cpu().numpy(), g["discount"][r])
        assert np.array_equal(b["is_state_terminal"].cpu().numpy(), g["terminal"][r])
        assert np.array_equal(b["action"].cpu().numpy(), g["action"][r])
        assert b["state"].shape == (batch, 4, 84, 84) and b["state"].dtype == torch.float32
        for key, c1, c2 in (("state", "s1", "s2"), ("next_state", "n1", "n2")):
            x = torch.round(b[key] * 255.0).to(torch.uint8)
            # f32 path: u8 * (1/255) is within 1 ulp of the reference's u8 / 255
            np.testing.assert_allclose(b[key][:4].cpu().numpy(),
                                       x[:4].cpu().numpy().astype(np.float32) / 255, rtol=2e-7)
            s1, s2 = _checksums_u8(x)
            assert np.array_equal(s1, g[c1][r]) and np.array_equal(s2, g[c2][r]), (r, key)
        if r == 0:
            # raw uint8 gather of the same sample (separate gather kernel)
            raw = buf._gather(exps, gamma, None, raw=True)
            for key, c1, c2 in (("state", "s1", "s2"), ("next_state", "n1", "n2")):
                assert raw[key].dtype == torch.uint8
                s1, s2 = _checksums_u8(raw[key])
                assert np.array_equal(s1, g[c1][r]) and np.array_equal(s2, g[c2][r])
        buf.update_errors([float(x) for x in g["errors"][r]])
        for j in range(between):
            a, rw, tm = int(g["between_actions"][bi]), float(g["between_rewards"][bi]), bool(
                g["between_terminals"][bi])
            bi += 1
            frames = frames[1:] + [host_frame(fid)]
            fid += 1
            nxt = LazyFrames(list(frames), stack_axis=0)
            buf.append(cur, a, rw, nxt, None, tm)
            if tm:
                frames = [host_frame(fid)] * STACK
                fid += 1
                cur = LazyFrames(list(frames), stack_axis=0)
            else:
                cur = nxt
    buf._flush()
    info = buf.store.info()
    assert len(buf) == int(g["final_len"])
    assert info["max_priority"] == float(g["final_max_priority"])
    assert info["total"] == float(g["final_total"])
    assert info["min"] == float(g["final_min"])
