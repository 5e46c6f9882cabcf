"""Where does a k_gemm_tf32x3 launch spend its time?  %globaltimer stamps per CTA
(b2rl_gemm_debug_times): entry, set-up done, first stage filled, last MMA issued, accumulators
complete, epilogue stores issued, CTA end.  Prints medians over the CTAs of one launch."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfrl_b200 import _lib  # noqa: E402
from pfrl_b200.ops.conv import geometry  # noqa: E402
from pfrl_b200.ops.linear import gemm  # noqa: E402

NAMES = ["setup", "first_fill", "mainloop", "drain", "epilogue", "exit"]


def report(name, fn, n_ctas_max=4096):
    L = _lib.load()
    buf = torch.zeros(n_ctas_max * 8, dtype=torch.int64, device="cuda")
    for _ in range(3):
        fn()
    flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
    flush.add_(1)
    L.b2rl_gemm_debug_times(ctypes.c_void_p(buf.data_ptr()))
    fn()
    torch.cuda.synchronize()
    L.b2rl_gemm_debug_times(None)
    t = buf.cpu().numpy().reshape(-1, 8).astype(np.int64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    d = np.diff(t[:, :7], axis=1)
    print(f"{name}: {len(t)} CTAs, launch span {(t[:, 6].max() - t0) / 1e3:.1f} us; "
          f"CTA start spread {(t[:, 0].max() - t0) / 1e3:.1f} us; median per-CTA phases (us): "
          + ", ".join(f"{n} {np.median(d[:, i]) / 1e3:.2f}" for i, n in enumerate(NAMES))
          + f"; CTA total {np.median(t[:, 6] - t[:, 0]) / 1e3:.1f}")


if __name__ == "__main__":
    a = torch.randn(512, 3136, device="cuda")
    w = torch.randn(1024, 3136, device="cuda")
    report("main fwd 512x1024x3136", lambda: gemm(a, w))
    a2 = torch.randn(512, 512, device="cuda")
    w2 = torch.randn(918, 512, device="cuda")
    report("adv fwd 512x918x512", lambda: gemm(a2, w2))
    a3 = torch.randn(4096, 4096, device="cuda")
    report("square 4096", lambda: gemm(a3, a3))
    x = torch.rand(512, 32, 20, 20, device="cuda")
    k = torch.randn(64, 32, 4, 4, device="cuda")
    geo = geometry(512, 32, 20, 20, 64, 4, 4, 2, "cuda:0")
    report("conv2 fwd", lambda: geo.forward(x, k))
    gy = torch.randn(512, 64, 9, 9, device="cuda")
    report("conv2 wgrad", lambda: geo.wgrad(x, gy))
    report("conv2 dgrad (last phase)", lambda: geo.dgrad(gy, k))
