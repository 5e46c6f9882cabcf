"""Dense layers on the tcgen05 tensor cores with fp32 results (csrc/gemm.cu).

``linear(x, weight, bias)`` is ``torch.nn.functional.linear`` for 2-D fp32 CUDA inputs: the
forward product and both backward products (dX = dY . W, dW = dY^T . X) run
``b2rl_gemm_tf32x3`` -- every fp32 operand split into two TF32 numbers, three tensor-core
products accumulated in fp32 -- instead of cuBLAS' CUDA-core SGEMM.  Operands are read in
place whatever their orientation (row stride = leading dimension), so the halves of a
``torch.chunk`` and the transposed backward products need no copies.

Replaces the F.linear calls of pfrl/q_functions/dueling_dqn.py:67-129,
pfrl/nn/noisy_linear.py:53-70 and pfrl/nn/atari_cnn.py:17-47.  Which products take the
tensor-core path is decided by ``worth_it`` (measured break-even against cuBLAS fp32);
``B2RL_LINEAR=cublas`` / ``=tcgen05`` in the environment force one side (A/B timing).
"""
import ctypes
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from pfrl_b200 import _lib

def worth_it(M, N, K):
    """Products that measured faster than cuBLAS' fp32 SGEMM on B200 (tools/bench_gemm.py):
    the 3136-deep layers at any batch, and everything from ~0.5 G multiply-adds up; the
    small heads (512 x 918 x 512 and the like) are launch- and epilogue-bound and stay on
    cuBLAS.  B2RL_LINEAR=tcgen05 forces the tensor-core path, =cublas forbids it."""
    forced = os.environ.get("B2RL_LINEAR", "auto")
    if forced == "tcgen05":
        return True
    macs = M * N * K
    return forced != "cublas" and ((K >= 1024 and macs >= 1 << 24) or macs >= 1 << 29)


def _rows(t):
    """A 2-D fp32 view whose columns are contiguous (row stride = leading dimension)."""
    if t.stride(1) != 1 or t.stride(0) < t.shape[1] or t.data_ptr() % 4:
        t = t.contiguous()
    return t


def gemm(a, b, a_mn_major=False, b_mn_major=False, bias=None, relu=False):
    """C[M,N] = A . B^T (+ bias) (relu) through ``b2rl_gemm_tf32x3``.

    a: [M, K] (or [K, M] when a_mn_major), b: [N, K] (or [K, N] when b_mn_major); fp32 CUDA,
    columns contiguous.  Returns a new contiguous [M, N] tensor."""
    L = _lib.load()
    a, b = _rows(a), _rows(b)
    K, M = (a.shape if a_mn_major else a.shape[::-1])
    Kb, N = (b.shape if b_mn_major else b.shape[::-1])
    assert K == Kb, (a.shape, b.shape, a_mn_major, b_mn_major)
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    need = L.b2rl_gemm_workspace_bytes(M, N, K)
    ws = torch.empty(need, dtype=torch.uint8, device=a.device) if need else None
    if bias is not None:
        bias = bias.detach().contiguous()
    _lib.check(L.b2rl_gemm_tf32x3(
        ctypes.c_void_p(a.data_ptr()), a.stride(0), int(a_mn_major),
        ctypes.c_void_p(b.data_ptr()), b.stride(0), int(b_mn_major),
        None if bias is None else ctypes.c_void_p(bias.data_ptr()), int(relu),
        ctypes.c_void_p(out.data_ptr()), N, M, N, K,
        None if ws is None else ctypes.c_void_p(ws.data_ptr()), need,
        ctypes.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)))
    return out


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return gemm(x.detach(), weight.detach(), bias=bias)

    @staticmethod
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        batch, n_out, n_in = x.shape[0], weight.shape[0], weight.shape[1]
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            # dX[batch, in] = dY[batch, out] . W[out, in]: W is read with `out` as the row index
            if worth_it(batch, n_in, n_out):
                gx = gemm(grad_out, weight.detach(), b_mn_major=True)
            else:
                gx = grad_out @ weight.detach()
        if ctx.needs_input_grad[1]:
            # dW[out, in] = dY^T . X: both read with the batch index as the row index
            if worth_it(n_out, n_in, batch):
                gw = gemm(grad_out, x.detach(), a_mn_major=True, b_mn_major=True)
            else:
                gw = grad_out.t() @ x.detach()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = grad_out.sum(0)
        return gx, gw, gb


def linear(x, weight, bias=None):
    """F.linear; 2-D fp32 CUDA products of a useful size run on the tensor cores."""
    if (x.is_cuda and x.ndim == 2 and x.dtype == torch.float32 and weight.dtype == torch.float32
            and weight.ndim == 2 and (bias is None or bias.dtype == torch.float32)
            and worth_it(x.shape[0], weight.shape[0], weight.shape[1])):
        return _LinearFn.apply(x, weight, bias)
    return F.linear(x, weight, bias)


class TCLinear(nn.Linear):
    """nn.Linear (same parameters, same state_dict keys) whose products go through
    ``linear`` above."""

    def forward(self, x):
        return linear(x, self.weight, self.bias)
