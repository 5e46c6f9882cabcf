"""K10 micro-benchmark: b2rl_gemm_tf32x3 against cuBLAS (fp32 CUDA-core SGEMM, and TF32 for
orientation) on the Rainbow layer shapes at B = 512, forward / dX / dW.  CUDA events, L2
flushed between repetitions.  One JSON line per shape."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfrl_b200.ops.linear import gemm  # noqa: E402

SHAPES = [
    ("main fwd", 512, 1024, 3136, False, False),
    ("main dX", 512, 3136, 1024, False, True),
    ("main dW", 1024, 3136, 512, True, True),
    ("adv fwd", 512, 918, 512, False, False),
    ("adv dX", 512, 512, 918, False, True),
    ("adv dW", 918, 512, 512, True, True),
    ("val fwd", 512, 51, 512, False, False),
    ("dqn-head B=32", 32, 512, 3136, False, False),
    ("square 4096", 4096, 4096, 4096, False, False),
]


def time_it(fn, flush, reps=20):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(reps)]
    for a, b in ev:
        flush.add_(1)  # 256 MB > L2
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e3  # us, median


def main():
    flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except OSError:
        pass
    tf32_peak = peaks.get("bf16_tflops", 1668.1) / 2  # dense TF32 = half the bf16 rate
    for name, M, N, K, a_mn, b_mn in SHAPES:
        a = torch.randn((K, M) if a_mn else (M, K), device="cuda")
        b = torch.randn((K, N) if b_mn else (N, K), device="cuda")
        af = a.t() if a_mn else a
        bf = (b if b_mn else b.t())
        torch.backends.cuda.matmul.allow_tf32 = False
        t_fp32 = time_it(lambda: af @ bf, flush)
        torch.backends.cuda.matmul.allow_tf32 = True
        t_tf32 = time_it(lambda: af @ bf, flush)
        torch.backends.cuda.matmul.allow_tf32 = False
        t_tc = time_it(lambda: gemm(a, b, a_mn_major=a_mn, b_mn_major=b_mn), flush)
        flops = 2.0 * M * N * K
        print(json.dumps({
            "shape": name, "M": M, "N": N, "K": K,
            "tcgen05_3xtf32_us": round(t_tc, 2), "cublas_fp32_us": round(t_fp32, 2),
            "cublas_tf32_us": round(t_tf32, 2),
            "speedup_vs_cublas_fp32": round(t_fp32 / t_tc, 2),
            "tflops_fp32_equiv": round(flops / t_tc * 1e-6, 1),
            "tensor_frac_of_tf32_peak": round(3 * flops / t_tc * 1e-6 / tf32_peak, 3),
        }), flush=True)


if __name__ == "__main__":
    main()
