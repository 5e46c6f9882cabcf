"""Convolutions of the Atari trunks as implicit GEMMs on the tcgen05 tensor cores.

Forward, input gradient and weight gradient of an ``nn.Conv2d`` without padding
(pfrl/nn/atari_cnn.py:30-44, pfrl/q_functions/dueling_dqn.py:34-40,91-97 -- cuDNN fp32 in
the reference) are three products of ``b2rl_gemm_tf32x3_ex`` (csrc/gemm.cu, 3xTF32 split with
fp32 results) whose operands are read IN PLACE through index tables: no im2col buffer, no
layout change of activations or weights, NCHW in and out.

    forward   y[(b,oy,ox), oc]   = sum_(ic,ky,kx) x[b, ic, oy s + ky, ox s + kx] w[oc, (ic,ky,kx)]
    dgrad     gx[(b,y,x), ic]    = sum_(oc,ky,kx) gy[b, oc, (y - ky) / s, (x - kx) / s] w[oc, ic, ky, kx]
    wgrad     gw[(ic,ky,kx), oc] = sum_(b,oy,ox)  x[b, ic, oy s + ky, ox s + kx] gy[b, oc, oy, ox]

A ``ConvGeometry`` holds the tables of one (batch, layer) shape on one device; they are built
once (a few hundred KB of int32) and reused by every call.  uint8 inputs (the raw replay
minibatch, phi = x / 255 folded in as ``scale``) are read directly by forward and wgrad.
"""
import ctypes
import functools
import math
import os

import torch
import torch.nn as nn

from pfrl_b200 import _lib

K_MAJOR, MN_MAJOR, GATHER = 0, 1, 2


def _pack(lo, hi):
    """Two 16-bit fields in one int32 (lo: bits 0-15, hi: bits 16-31, both may be negative)."""
    v = (lo.to(torch.int64) & 0xFFFF) | ((hi.to(torch.int64) & 0xFFFF) << 16)
    return torch.where(v >= (1 << 31), v - (1 << 32), v).to(torch.int32)


def _table(offsets, lo=None, hi=None):
    offsets = offsets.reshape(-1).to(torch.int64)
    assert int(offsets.max()) < 2 ** 31
    zero = torch.zeros_like(offsets)
    code = _pack(zero if lo is None else lo.reshape(-1), zero if hi is None else hi.reshape(-1))
    return torch.stack([offsets.to(torch.int32), code], dim=1).contiguous()


class ConvGeometry:
    """Index tables of one convolution shape (see the module docstring)."""

    def __init__(self, batch, in_channels, height, width, out_channels, kh, kw, stride, device):
        s = stride
        assert s & (s - 1) == 0, "stride must be a power of two"
        self.B, self.IC, self.H, self.W = batch, in_channels, height, width
        self.OC, self.KH, self.KW, self.s = out_channels, kh, kw, s
        self.OH, self.OW = (height - kh) // s + 1, (width - kw) // s + 1
        self.device = torch.device(device)
        B, IC, H, W, OC, KH, KW, OH, OW = batch, in_channels, height, width, out_channels, \
            kh, kw, self.OH, self.OW
        ar = torch.arange
        # rows (b, oy, ox) of the output / k (ic, ky, kx) of the filter
        b, oy, ox = torch.meshgrid(ar(B), ar(OH), ar(OW), indexing="ij")
        ic, ky, kx = torch.meshgrid(ar(IC), ar(KH), ar(KW), indexing="ij")
        t = {}
        t["fwd_a_row"] = _table(b * (IC * H * W), oy * s, ox * s)
        t["fwd_a_k"] = _table(ic * (H * W), ky, kx)
        t["fwd_c_row"] = (b * (OC * OH * OW) + oy * OW + ox).reshape(-1).to(torch.int32)
        # dgrad: rows (b, y, x) of the input, k (oc, ky, kx)
        bi, y, x = torch.meshgrid(ar(B), ar(H), ar(W), indexing="ij")
        oc, ky2, kx2 = torch.meshgrid(ar(OC), ar(KH), ar(KW), indexing="ij")
        t["dg_a_row"] = _table(bi * (OC * OH * OW), y, x)
        t["dg_a_k"] = _table(oc * (OH * OW), -ky2, -kx2)
        t["dg_b_row"] = _table(ar(IC) * (KH * KW))
        t["dg_b_k"] = _table(oc * (IC * KH * KW) + ky2 * KW + kx2)
        t["dg_c_row"] = (bi * (IC * H * W) + y * W + x).reshape(-1).to(torch.int32)
        # wgrad: rows (ic, ky, kx), k (b, oy, ox)
        t["wg_a_row"] = _table(ic * (H * W), ky, kx)
        t["wg_a_k"] = _table(b * (IC * H * W), oy * s, ox * s)
        t["wg_b_row"] = _table(ar(OC) * (OH * OW))
        t["wg_b_k"] = _table(b * (OC * OH * OW) + oy * OW + ox)
        t["wg_c_row"] = ar(IC * KH * KW).to(torch.int32)
        self.t = {k: v.to(self.device) for k, v in t.items()}

    # -- operand / output descriptors ------------------------------------------------
    def _gather(self, data, row, k, y_limit=1, x_limit=1, shift=0, pitch=0, along_k=False,
                scale=None):
        o = _lib.GemmOperand()
        o.data, o.mode = data.data_ptr(), GATHER
        o.row_tab, o.k_tab = self.t[row].data_ptr(), self.t[k].data_ptr()
        o.y_limit, o.x_limit, o.shift, o.pitch = y_limit, x_limit, shift, pitch
        o.lanes_along_k = int(along_k)
        o.u8 = int(data.dtype == torch.uint8)
        o.scale = 1.0 if scale is None else float(scale)
        return o

    @staticmethod
    def _dense(data, ld):
        o = _lib.GemmOperand()
        o.data, o.mode, o.ld = data.data_ptr(), K_MAJOR, ld
        return o

    def _run(self, a, b, out, row_tab, col_stride, bias, relu, M, N, K):
        L = _lib.load()
        c = _lib.GemmOutput()
        c.data, c.ld = out.data_ptr(), N
        c.row_tab = self.t[row_tab].data_ptr()
        c.col_stride = col_stride
        c.bias = None if bias is None else bias.data_ptr()
        c.relu = int(relu)
        need = L.b2rl_gemm_workspace_bytes(M, N, K)
        ws = torch.empty(need, dtype=torch.uint8, device=out.device) if need else None
        _lib.check(L.b2rl_gemm_tf32x3_ex(
            ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), M, N, K,
            None if ws is None else ctypes.c_void_p(ws.data_ptr()), need,
            ctypes.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)))
        return out

    # -- the three products ---------------------------------------------------------------
    def forward(self, x, weight, bias=None, relu=False, scale=None):
        assert x.shape == (self.B, self.IC, self.H, self.W) and x.is_contiguous()
        w = weight.detach().contiguous()
        K = self.IC * self.KH * self.KW
        out = torch.empty((self.B, self.OC, self.OH, self.OW), dtype=torch.float32,
                          device=x.device)
        a = self._gather(x, "fwd_a_row", "fwd_a_k", self.H, self.W, 0, self.W, scale=scale)
        b = self._dense(w, K)
        bias = None if bias is None else bias.detach().contiguous()
        return self._run(a, b, out, "fwd_c_row", self.OH * self.OW, bias, relu,
                         self.B * self.OH * self.OW, self.OC, K)

    def dgrad(self, grad_out, weight):
        gy = grad_out.contiguous()
        w = weight.detach().contiguous()
        s = self.s
        out = torch.empty((self.B, self.IC, self.H, self.W), dtype=torch.float32,
                          device=gy.device)
        a = self._gather(gy, "dg_a_row", "dg_a_k", s * (self.OH - 1) + 1, s * (self.OW - 1) + 1,
                         int(math.log2(s)), self.OW)
        b = self._gather(w, "dg_b_row", "dg_b_k", along_k=True)
        return self._run(a, b, out, "dg_c_row", self.H * self.W, None, False,
                         self.B * self.H * self.W, self.IC, self.OC * self.KH * self.KW)

    def wgrad(self, x, grad_out, scale=None):
        gy = grad_out.contiguous()
        Kw = self.IC * self.KH * self.KW
        out = torch.empty((self.OC, self.IC, self.KH, self.KW), dtype=torch.float32,
                          device=gy.device)
        a = self._gather(x, "wg_a_row", "wg_a_k", self.H, self.W, 0, self.W, along_k=True,
                         scale=scale)
        b = self._gather(gy, "wg_b_row", "wg_b_k", along_k=True)
        return self._run(a, b, out, "wg_c_row", Kw, None, False, Kw, self.OC,
                         self.B * self.OH * self.OW)


@functools.lru_cache(maxsize=64)
def geometry(batch, in_channels, height, width, out_channels, kh, kw, stride, device):
    return ConvGeometry(batch, in_channels, height, width, out_channels, kh, kw, stride, device)


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, geo, scale):
        ctx.save_for_backward(x, weight)
        ctx.geo, ctx.scale, ctx.has_bias = geo, scale, bias is not None
        return geo.forward(x.detach(), weight, bias, scale=scale)

    @staticmethod
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        geo = ctx.geo
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = geo.dgrad(grad_out, weight)
        if ctx.needs_input_grad[1]:
            gw = geo.wgrad(x.detach(), grad_out, scale=ctx.scale)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = grad_out.sum((0, 2, 3))
        return gx, gw, gb, None, None


def enabled():
    return os.environ.get("B2RL_CONV", "tcgen05") != "cudnn"


def supported(module, x):
    return (x.is_cuda and x.ndim == 4 and x.is_contiguous() and enabled()
            and x.dtype in (torch.float32, torch.uint8)
            and module.weight.dtype == torch.float32 and module.padding == (0, 0)
            and module.dilation == (1, 1) and module.groups == 1
            and module.stride[0] == module.stride[1]
            and module.stride[0] & (module.stride[0] - 1) == 0
            and module.padding_mode == "zeros" and x.shape[0] >= 8
            and x.shape[1] == module.in_channels and max(x.shape[2:]) < 2 ** 15)


def conv2d(module, x, scale=None):
    """``module(x)`` for an nn.Conv2d without padding, on the tensor cores."""
    kh, kw = module.kernel_size
    geo = geometry(x.shape[0], x.shape[1], x.shape[2], x.shape[3], module.out_channels, kh, kw,
                   module.stride[0], str(x.device))
    return _ConvFn.apply(x, module.weight, module.bias, geo, scale)


class TCConv2d(nn.Conv2d):
    """nn.Conv2d (same parameters, same state_dict keys); CUDA fp32 batches run as implicit
    GEMMs on the tensor cores, everything else goes to cuDNN."""

    def forward(self, x):
        if supported(self, x) and x.dtype == torch.float32:
            return conv2d(self, x)
        return super().forward(x)
