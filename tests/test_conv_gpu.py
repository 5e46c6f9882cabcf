"""GPU: the hand-written fp32 Nature conv1 forward vs cuDNN (TF32 off)."""
import time

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 3, 148, 512])
def test_conv1_matches_cudnn_fp32(n):
    from pfrl_b200.nn.fast_conv import NatureConv1

    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(n)
    conv = NatureConv1().cuda()
    x = torch.rand(n, 4, 84, 84, device="cuda")
    ref = F.conv2d(x, conv.weight, conv.bias, stride=4)
    out = conv(x)
    assert out.shape == (n, 32, 20, 20)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    # weight / bias gradients (cuDNN backward) through the custom forward
    g = torch.randn_like(out)
    conv.zero_grad()
    out.backward(g)
    gw, gb = conv.weight.grad.clone(), conv.bias.grad.clone()
    conv.zero_grad()
    ref.backward(g)
    # both gradients come from cuDNN, possibly through different wgrad algorithms
    # (FFT / implicit GEMM): compare relative to the gradient's scale
    scale = conv.weight.grad.abs().max().item()
    torch.testing.assert_close(gw, conv.weight.grad, rtol=1e-3, atol=1e-3 * scale)
    torch.testing.assert_close(gb, conv.bias.grad, rtol=1e-3, atol=1e-3 * conv.bias.grad.abs().max().item())


def test_conv1_falls_back_when_input_needs_grad_or_other_shape():
    from pfrl_b200.nn.fast_conv import NatureConv1

    conv = NatureConv1().cuda()
    x = torch.rand(2, 4, 84, 84, device="cuda", requires_grad=True)
    conv(x).sum().backward()
    assert x.grad is not None
    y = conv(torch.rand(2, 4, 100, 100, device="cuda"))
    assert y.shape == (2, 32, 24, 24)


def test_conv1_speed_report():
    from pfrl_b200.nn.fast_conv import NatureConv1

    torch.backends.cudnn.allow_tf32 = False
    conv = NatureConv1().cuda()
    x = torch.rand(512, 4, 84, 84, device="cuda")

    def timeit(fn, n=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e6

    with torch.no_grad():
        ours = timeit(lambda: conv(x))
        cudnn = timeit(lambda: F.conv2d(x, conv.weight, conv.bias, stride=4))
    print("conv1 fwd B=512: ours %.1f us, cuDNN fp32 %.1f us" % (ours, cudnn))
    assert ours < cudnn


@pytest.mark.parametrize("n", [1, 5, 149, 512])
def test_conv1_on_uint8_images_is_bit_identical_to_the_f32_path(n):
    """phi = x / 255 folded into the layer (b2rl_conv_nature1_fwd_u8): same bits as the
    ScaleU8 gather output fed to b2rl_conv_nature1_fwd, same weight gradient."""
    from pfrl_b200.nn.fast_conv import NatureConv1
    from pfrl_b200.utils.phi import ScaleU8

    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(7 + n)
    conv = NatureConv1().cuda()
    x8 = torch.randint(0, 256, (n, 4, 84, 84), dtype=torch.uint8, device="cuda")
    scale = ScaleU8().b2rl_obs_scale
    assert scale == conv.input_scale
    xf = x8.to(torch.float32) * scale          # what the U8_TO_F32 gather writes
    out8, outf = conv(x8), conv(xf)
    assert torch.equal(out8, outf)
    g = torch.randn_like(out8)
    conv.zero_grad()
    out8.backward(g)
    gw8 = conv.weight.grad.clone()
    conv.zero_grad()
    outf.backward(g)
    # both weight gradients come from cuDNN (not bit-reproducible between calls)
    sc = conv.weight.grad.abs().max().item()
    torch.testing.assert_close(gw8, conv.weight.grad, rtol=1e-3, atol=1e-3 * sc)


def test_rainbow_update_with_byte_batches_equals_f32_batches():
    """RawU8 (bytes out of the replay gather, / 255 inside conv1) trains exactly like
    ScaleU8 (f32 batches): same parameters (to round-off) after a few prioritised updates."""
    import numpy as np

    from pfrl_b200 import agents, explorers, nn as pnn, q_functions
    from pfrl_b200.replay_buffers import PrioritizedReplayBuffer
    from pfrl_b200.utils.phi import RawU8, ScaleU8

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    params = []
    for phi in (ScaleU8(), RawU8()):
        torch.manual_seed(0)
        np.random.seed(0)
        q = q_functions.DistributionalDuelingDQN(6, 51, -10, 10)
        pnn.to_factorized_noisy(q, sigma_scale=0.5)
        buf = PrioritizedReplayBuffer(2000, alpha=0.5, beta0=0.4, betasteps=100, num_steps=3,
                                      normalize_by_max="memory")
        rng = np.random.RandomState(1)
        frames = rng.randint(0, 256, size=(604, 84, 84), dtype=np.uint8)
        term = rng.rand(600) < 0.02
        term[-1] = True
        buf.append_trajectory(frames, rng.randint(0, 6, size=600).astype(np.int64),
                              rng.randint(-1, 2, size=600).astype(np.float64), term)
        agent = agents.CategoricalDoubleDQN(
            q, torch.optim.Adam(q.parameters(), 6.25e-5, eps=1.5e-4), buf, gpu=0, gamma=0.99,
            explorer=explorers.Greedy(), minibatch_size=32, replay_start_size=32,
            target_update_interval=100, update_interval=1, batch_accumulator="mean", phi=phi)
        torch.manual_seed(5)
        for _ in range(6):
            agent.update(buf.sample(32))
        params.append([p.detach().clone() for p in agent.model.parameters()])
    # identical forward values; the cuDNN weight-gradient kernels are not bit-reproducible
    # between calls, so the parameters agree to Adam-amplified round-off, not bit for bit
    for a, b in zip(*params):
        torch.testing.assert_close(a, b, rtol=1e-3, atol=2e-5)
