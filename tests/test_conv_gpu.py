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
