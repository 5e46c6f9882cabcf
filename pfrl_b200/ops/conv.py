"""Convolutions of the Atari trunks as implicit GEMMs on the tcgen05 tensor cores.

Forward, input gradient and weight gradient of an ``nn.Conv2d`` without padding
(pfrl/nn/atari_cnn.py:30-44, pfrl/q_functions/dueling_dqn.py:34-40,91-97 -- cuDNN fp32 in
the reference) are three products of ``b2rl_gemm_tf32x3_ex`` (csrc/gemm.cu, 3xTF32 split with
fp32 results) whose operands are read IN PLACE through index tables: no im2col buffer, no
layout change of activations or weights, NCHW in and out.

    forward   y[(b,oy,ox), oc]   = sum_(ic,ky,kx) x[b, ic, oy s + ky, ox s + kx] w[oc, (ic,ky,kx)]
    dgrad     gx[(b,y,x), ic]    = sum_(oc,ky,kx) gy[b, oc, (y - ky) / s, (x - kx) / s] w[oc, ic, ky, kx]
              (one product per stride phase (y mod s, x mod s): only the taps of that phase)
    wgrad     gw[(ic,ky,kx), oc] = sum_(b,oy,ox)  x[b, ic, oy s + ky, ox s + kx] gy[b, oc, oy, ox]

A ``ConvGeometry`` holds the tables of one (batch, layer) shape on one device; they are built
once (a few hundred KB of int32) and reused by every call.  uint8 inputs (the raw replay
minibatch, phi = x / 255 folded in as ``scale``) are read directly by forward and wgrad.
"""
import ctypes
import functools
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from pfrl_b200 import _lib

K_MAJOR, MN_MAJOR, GATHER = 0, 1, 2


def _pack(lo, hi):
    """Two 16-bit fields in one int32 (lo: bits 0-15, hi: bits 16-31, both may be negative)."""
    v = (lo.to(torch.int64) & 0xFFFF) | ((hi.to(torch.int64) & 0xFFFF) << 16)
    return torch.where(v >= (1 << 31), v - (1 << 32), v).to(torch.int32)


def _i32(t, pad_to=1):
    """Flat int32 table, zero-padded to a multiple of ``pad_to`` entries."""
    t = t.reshape(-1).to(torch.int64)
    assert t.numel() == 0 or (int(t.max()) < 2 ** 31 and int(t.min()) >= -2 ** 31)
    n = -(-max(t.numel(), 1) // pad_to) * pad_to
    out = torch.zeros(n, dtype=torch.int32)
    out[:t.numel()] = t.to(torch.int32)
    return out


class _Product:
    """Tables of one product: gather operands A and B, scattered output."""

    def __init__(self, M, N, K, a, b, c_row, c_stride, device):
        self.M, self.N, self.K = M, N, K
        self.c_stride = c_stride
        self.a = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in a.items()}
        self.b = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in b.items()}
        self.c_row = c_row.to(device)


def _operand(row_off, k_off, row_yx=None, k_yx=None, limits=(0, 0), along_k=False):
    d = {"row_off": _i32(row_off), "k_off": _i32(k_off, 32), "along_k": along_k,
         "limits": limits, "row_yx": None, "k_yx": None}
    if row_yx is not None:
        d["row_yx"] = _pack(row_yx[0].reshape(-1), row_yx[1].reshape(-1))
        kyx = _pack(k_yx[0].reshape(-1), k_yx[1].reshape(-1))
        pad = torch.zeros(d["k_off"].numel(), dtype=torch.int32)
        pad[:kyx.numel()] = kyx
        d["k_yx"] = pad
    return d


class ConvGeometry:
    """Index tables of one convolution shape (see the module docstring)."""

    def __init__(self, batch, in_channels, height, width, out_channels, kh, kw, stride, device):
        s = stride
        self.B, self.IC, self.H, self.W = batch, in_channels, height, width
        self.OC, self.KH, self.KW, self.s = out_channels, kh, kw, s
        self.OH, self.OW = (height - kh) // s + 1, (width - kw) // s + 1
        self.device = torch.device(device)
        B, IC, H, W, OC, KH, KW, OH, OW = batch, in_channels, height, width, out_channels, \
            kh, kw, self.OH, self.OW
        HW, OHW, KHW = H * W, OH * OW, KH * KW
        ar = torch.arange
        dev = self.device
        b, oy, ox = torch.meshgrid(ar(B), ar(OH), ar(OW), indexing="ij")     # output pixels
        ic, ky, kx = torch.meshgrid(ar(IC), ar(KH), ar(KW), indexing="ij")   # filter taps
        x_at_out = b * (IC * HW) + (oy * s) * W + ox * s     # x[b, 0, oy s, ox s]
        tap = ic * HW + ky * W + kx                           # + x[0, ic, ky, kx]
        # forward: rows = output pixels, k = taps
        self.fwd = _Product(
            B * OHW, OC, IC * KHW,
            _operand(x_at_out, tap), {}, (b * (OC * OHW) + oy * OW + ox).reshape(-1)
            .to(torch.int32), OHW, dev)
        # weight gradient, transposed: rows = taps, k = output pixels, columns = oc
        self.wg = _Product(
            IC * KHW, OC, B * OHW,
            _operand(tap, x_at_out, along_k=True),
            _operand(ar(OC) * OHW, b * (OC * OHW) + oy * OW + ox, along_k=True),
            ar(IC * KHW).to(torch.int32), IC * KHW, dev)
        # input gradient: one product per stride phase (py, px): input pixels y = s a + py,
        # x = s c + px only meet the taps ky = py + s j, kx = px + s i, at output (a - j, c - i)
        self.dg = []
        self.dg_covers_input = True
        for py in range(s):
            for px in range(s):
                na, nc = len(range(py, H, s)), len(range(px, W, s))
                nj, ni = len(range(py, KH, s)), len(range(px, KW, s))
                if na == 0 or nc == 0:
                    continue
                if nj == 0 or ni == 0:
                    self.dg_covers_input = False  # these pixels get no gradient: zero-fill
                    continue
                bb, a, c = torch.meshgrid(ar(B), ar(na), ar(nc), indexing="ij")
                oc, j, i = torch.meshgrid(ar(OC), ar(nj), ar(ni), indexing="ij")
                self.dg.append(_Product(
                    B * na * nc, IC, OC * nj * ni,
                    _operand(bb * (OC * OHW) + a * OW + c, oc * OHW - j * OW - i,
                             row_yx=(a, c), k_yx=(-j, -i), limits=(OH, OW)),
                    _operand(ar(IC) * KHW, oc * (IC * KHW) + (py + s * j) * KW + (px + s * i),
                             along_k=True),
                    (bb * (IC * HW) + (s * a + py) * W + (s * c + px)).reshape(-1)
                    .to(torch.int32), HW, dev))

    # -- operand / output descriptors ------------------------------------------------
    @staticmethod
    def _gather(data, tab, scale=None):
        o = _lib.GemmOperand()
        o.data, o.mode = data.data_ptr(), GATHER
        o.row_off, o.k_off = tab["row_off"].data_ptr(), tab["k_off"].data_ptr()
        if tab["row_yx"] is not None:
            o.row_yx, o.k_yx = tab["row_yx"].data_ptr(), tab["k_yx"].data_ptr()
            o.y_limit, o.x_limit = tab["limits"]
        o.lanes_along_k = int(tab["along_k"])
        o.u8 = int(data.dtype == torch.uint8)
        o.scale = 1.0 if scale is None else float(scale)
        return o

    @staticmethod
    def _dense(data, ld):
        o = _lib.GemmOperand()
        o.data, o.mode, o.ld = data.data_ptr(), K_MAJOR, ld
        return o

    @staticmethod
    def _run(prod, a, b, out, bias=None, relu=False):
        L = _lib.load()
        c = _lib.GemmOutput()
        c.data, c.ld = out.data_ptr(), prod.N
        c.row_tab = prod.c_row.data_ptr()
        c.col_stride = prod.c_stride
        c.bias = None if bias is None else bias.data_ptr()
        c.relu = int(relu)
        need = L.b2rl_gemm_workspace_bytes(prod.M, prod.N, prod.K)
        ws = torch.empty(need, dtype=torch.uint8, device=out.device) if need else None
        _lib.check(L.b2rl_gemm_tf32x3_ex(
            ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), prod.M, prod.N, prod.K,
            None if ws is None else ctypes.c_void_p(ws.data_ptr()), need,
            ctypes.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)))
        return out

    # -- the three products ---------------------------------------------------------------
    def forward(self, x, weight, bias=None, relu=False, scale=None):
        assert x.shape == (self.B, self.IC, self.H, self.W) and x.is_contiguous()
        w = weight.detach().contiguous()
        out = torch.empty((self.B, self.OC, self.OH, self.OW), dtype=torch.float32,
                          device=x.device)
        bias = None if bias is None else bias.detach().contiguous()
        return self._run(self.fwd, self._gather(x, self.fwd.a, scale),
                         self._dense(w, self.fwd.K), out, bias, relu)

    def dgrad(self, grad_out, weight):
        gy = grad_out.contiguous()
        w = weight.detach().contiguous()
        alloc = torch.empty if self.dg_covers_input else torch.zeros
        out = alloc((self.B, self.IC, self.H, self.W), dtype=torch.float32, device=gy.device)
        for prod in self.dg:
            self._run(prod, self._gather(gy, prod.a), self._gather(w, prod.b), out)
        return out

    def wgrad(self, x, grad_out, scale=None):
        gy = grad_out.contiguous()
        out = torch.empty((self.OC, self.IC, self.KH, self.KW), dtype=torch.float32,
                          device=gy.device)
        return self._run(self.wg, self._gather(x, self.wg.a, scale),
                         self._gather(gy, self.wg.b), out)


@functools.lru_cache(maxsize=64)
def geometry(batch, in_channels, height, width, out_channels, kh, kw, stride, device):
    return ConvGeometry(batch, in_channels, height, width, out_channels, kh, kw, stride, device)


def mode():
    """B2RL_CONV: "auto" (default: each product goes where it measured faster on B200, see
    profiles/README.md), "tcgen05" (everything on the tensor-core path), "cudnn" (nothing)."""
    return os.environ.get("B2RL_CONV", "auto")


def enabled():
    return mode() != "cudnn"


def _choice(stride):
    """(forward, dgrad, wgrad) on the tensor-core path?  Measured at B = 512 on B200 against
    cuDNN fp32 (tools/bench_conv.py): the strided forward and the stride-1 input gradient win,
    the weight gradients (two gather operands, huge K) do not yet."""
    if mode() == "tcgen05":
        return True, True, True
    return stride >= 2, stride == 1, False


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, geo, scale):
        ctx.save_for_backward(x, weight)
        ctx.geo, ctx.scale, ctx.has_bias = geo, scale, bias is not None
        if _choice(geo.s)[0] or x.dtype == torch.uint8:
            return geo.forward(x.detach(), weight, bias, scale=scale)
        return F.conv2d(x, weight, bias, stride=geo.s)

    @staticmethod
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        geo = ctx.geo
        _, tc_dgrad, tc_wgrad = _choice(geo.s)
        need_x, need_w, need_b = ctx.needs_input_grad[:3]
        need_b = need_b and ctx.has_bias
        gx = gw = gb = None
        grad_out = grad_out.contiguous()
        if need_x and tc_dgrad:
            gx = geo.dgrad(grad_out, weight)
        if need_w and (tc_wgrad or x.dtype == torch.uint8):
            gw = geo.wgrad(x.detach(), grad_out, scale=ctx.scale)
        lib_x, lib_w = need_x and gx is None, need_w and gw is None
        if lib_x or lib_w:
            rx, rw, rb = torch.ops.aten.convolution_backward(
                grad_out, x, weight, [geo.OC] if need_b else None, [geo.s, geo.s], [0, 0], [1, 1],
                False, [0, 0], 1, [lib_x, lib_w, need_b])
            gx = rx if lib_x else gx
            gw = rw if lib_w else gw
            gb = rb if need_b else None
        elif need_b:
            gb = grad_out.sum((0, 2, 3))
        return gx, gw, gb, None, None


def supported(module, x):
    return (x.is_cuda and x.ndim == 4 and x.is_contiguous() and enabled()
            and x.dtype in (torch.float32, torch.uint8)
            and module.weight.dtype == torch.float32 and module.padding == (0, 0)
            and module.dilation == (1, 1) and module.groups == 1
            and module.stride[0] == module.stride[1]
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
