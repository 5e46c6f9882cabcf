"""K10 micro-benchmark: the Nature-trunk convolutions at B = 512 as tcgen05 implicit GEMMs
(ops/conv.py) against cuDNN fp32 (TF32 off) -- forward, input gradient, weight gradient.
CUDA events, L2 flushed between repetitions.  One JSON line per layer and direction."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfrl_b200.ops.conv import geometry  # noqa: E402
from tools.bench_gemm import time_it  # noqa: E402

LAYERS = [("conv1 8x8/4", 4, 84, 84, 32, 8, 4), ("conv2 4x4/2", 32, 20, 20, 64, 4, 2),
          ("conv3 3x3/1", 64, 9, 9, 64, 3, 1)]


def main(batch=512):
    torch.backends.cudnn.allow_tf32 = False
    flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
    aten = torch.ops.aten
    for name, IC, H, W, OC, K, s in LAYERS:
        x = torch.rand(batch, IC, H, W, device="cuda")
        w = torch.randn(OC, IC, K, K, device="cuda") * 0.05
        b = torch.zeros(OC, device="cuda")
        geo = geometry(batch, IC, H, W, OC, K, K, s, "cuda:0")
        y = geo.forward(x, w, b)
        gy = torch.randn_like(y)
        flops = 2.0 * y.numel() * IC * K * K

        def cudnn_bwd(mask):
            return aten.convolution_backward(gy, x, w, [OC], [s, s], [0, 0], [1, 1], False,
                                             [0, 0], 1, mask)

        rows = [
            ("fwd", lambda: geo.forward(x, w, b), lambda: F.conv2d(x, w, b, stride=s)),
            ("wgrad", lambda: geo.wgrad(x, gy), lambda: cudnn_bwd([False, True, False])),
        ]
        if name != "conv1 8x8/4":
            rows.insert(1, ("dgrad", lambda: geo.dgrad(gy, w),
                            lambda: cudnn_bwd([True, False, False])))
        for what, tc, ref in rows:
            t_tc, t_ref = time_it(tc, flush), time_it(ref, flush)
            print(json.dumps({"layer": name, "op": what, "batch": batch,
                              "tcgen05_3xtf32_us": round(t_tc, 1), "cudnn_fp32_us": round(t_ref, 1),
                              "speedup": round(t_ref / t_tc, 2),
                              "tflops_fp32_equiv": round(flops / t_tc * 1e-6, 1)}), flush=True)
        if name == "conv1 8x8/4":
            from pfrl_b200.nn.fast_conv import NatureConv1
            m = NatureConv1(4).cuda()
            xb = (x * 255).to(torch.uint8)
            with torch.no_grad():
                t_own = time_it(lambda: m(xb), flush)
                sc = m.input_scale
                t_tcu8 = time_it(lambda: geo.forward(xb, w, b, scale=sc), flush)
            print(json.dumps({"layer": name, "op": "fwd from uint8", "k_conv_nature1_us": round(t_own, 1),
                              "tcgen05_3xtf32_us": round(t_tcu8, 1)}), flush=True)


if __name__ == "__main__":
    main()
