"""nn.Conv2d with a hand-written fp32 forward for the first Nature-DQN layer.

``NatureConv1`` IS an ``nn.Conv2d(4, 32, 8, stride=4)`` (same parameters,
same state_dict keys).  On CUDA, for contiguous fp32 [N, 4, 84, 84] inputs
that do not require grad (observations), the forward runs
``b2rl_conv_nature1_fwd`` (csrc/conv.cu: exact fp32 FFMA accumulation, several
times faster than cuDNN's TF32-off path); the weight gradient comes from
``aten::convolution_backward`` (cuDNN), or with B2RL_CONV=tcgen05 from the tensor-core
implicit GEMM of ops/conv.py (read straight from the bytes for uint8 inputs; slower today).  uint8 inputs (phi = utils.phi.RawU8: the replay gather
emits bytes) go through ``b2rl_conv_nature1_fwd_u8``, which applies ``x * input_scale``
while it stages the image, so the f32 batch is never written to HBM.  Everything else
falls through to cuDNN.
"""
import ctypes

import torch
import torch.nn as nn

from pfrl_b200 import _lib
from pfrl_b200.ops import conv as conv_ops


class _Conv1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        L = _lib.load()
        n = x.shape[0]
        out = torch.empty((n, 32, 20, 20), dtype=torch.float32, device=x.device)
        w = weight.detach().contiguous()
        b = None if bias is None else bias.detach().contiguous()
        _lib.check(L.b2rl_conv_nature1_fwd(
            ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(w.data_ptr()),
            None if b is None else ctypes.c_void_p(b.data_ptr()), n,
            ctypes.c_void_p(out.data_ptr()),
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        if conv_ops.mode() == "tcgen05" and x.shape[0] >= 8:
            gw = conv_ops.geometry(x.shape[0], 4, 84, 84, 32, 8, 8, 4, str(x.device)) \
                .wgrad(x, grad_out)
            return None, gw, grad_out.sum((0, 2, 3)) if ctx.has_bias else None
        _, gw, gb = torch.ops.aten.convolution_backward(
            grad_out.contiguous(), x, weight, [32] if ctx.has_bias else None, [4, 4], [0, 0],
            [1, 1], False, [0, 0], 1, [False, True, ctx.has_bias])
        return None, gw, gb if ctx.has_bias else None


class _Conv1U8Fn(torch.autograd.Function):
    """uint8 images in, float(x) * scale applied inside the kernel (phi = x / 255 folded
    into the layer); the weight gradient expands the bytes only when it is asked for."""

    @staticmethod
    def forward(ctx, x, weight, bias, scale):
        L = _lib.load()
        n = x.shape[0]
        out = torch.empty((n, 32, 20, 20), dtype=torch.float32, device=x.device)
        w = weight.detach().contiguous()
        b = None if bias is None else bias.detach().contiguous()
        _lib.check(L.b2rl_conv_nature1_fwd_u8(
            ctypes.c_void_p(x.data_ptr()), float(scale), ctypes.c_void_p(w.data_ptr()),
            None if b is None else ctypes.c_void_p(b.data_ptr()), n,
            ctypes.c_void_p(out.data_ptr()),
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        ctx.scale = float(scale)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        if conv_ops.mode() == "tcgen05" and x.shape[0] >= 8:
            # tensor-core weight gradient straight from the bytes (ops/conv.py)
            gw = conv_ops.geometry(x.shape[0], 4, 84, 84, 32, 8, 8, 4, str(x.device)) \
                .wgrad(x, grad_out, scale=ctx.scale)
            return None, gw, grad_out.sum((0, 2, 3)) if ctx.has_bias else None, None
        xf = x.to(torch.float32) * ctx.scale
        _, gw, gb = torch.ops.aten.convolution_backward(
            grad_out.contiguous(), xf, weight, [32] if ctx.has_bias else None, [4, 4], [0, 0],
            [1, 1], False, [0, 0], 1, [False, True, ctx.has_bias])
        return None, gw, gb if ctx.has_bias else None, None


class NatureConv1(nn.Conv2d):
    #: uint8 inputs are read as float(x) * input_scale (the Atari phi, x / 255)
    input_scale = float(torch.tensor(1.0 / 255.0, dtype=torch.float32))

    def __init__(self, n_input_channels=4):
        super().__init__(n_input_channels, 32, 8, stride=4)

    def forward(self, x):
        if x.dtype == torch.uint8:
            if (x.is_cuda and x.ndim == 4 and tuple(x.shape[1:]) == (4, 84, 84)
                    and self.in_channels == 4 and x.is_contiguous()
                    and self.weight.dtype == torch.float32):
                return _Conv1U8Fn.apply(x, self.weight, self.bias, self.input_scale)
            x = x.to(torch.float32) * self.input_scale
        if (x.is_cuda and x.dtype == torch.float32 and x.ndim == 4
                and tuple(x.shape[1:]) == (4, 84, 84) and self.in_channels == 4
                and not x.requires_grad and x.is_contiguous()
                and self.weight.dtype == torch.float32):
            return _Conv1Fn.apply(x, self.weight, self.bias)
        return super().forward(x)
