"""Layout probe for csrc/gemm.cu: products with index-coded operands, so that a wrong
shared-memory descriptor shows up as a readable permutation rather than as noise.
C[m, n] = sum_k A[m, k] B[n, k] with B = one-hot(k == n) gives C[m, n] = A[m, n]."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfrl_b200.ops.linear import gemm  # noqa: E402


def probe(M, N, K, a_mn, b_mn):
    dev = "cuda"
    m = torch.arange(M, device=dev, dtype=torch.float32)[:, None]
    k = torch.arange(K, device=dev, dtype=torch.float32)[None, :]
    A = m * 64 + k + 1          # codes (m, k); < 2^24: exact after the hi/lo split
    B = torch.zeros(N, K, device=dev)
    idx = torch.arange(min(N, K), device=dev)
    B[idx, idx] = 1.0
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    C = gemm(a, b, a_mn_major=a_mn, b_mn_major=b_mn)
    want = A @ B.t()
    bad = (C != want)
    print(f"probe A-coded M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: mismatches {int(bad.sum())}")
    if bad.any():
        ij = bad.nonzero()[:12]
        for i, j in ij.tolist():
            v = float(C[i, j]) - 1
            print(f"   C[{i},{j}] = {float(C[i, j])} -> (m={int(v // 64)}, k={int(v % 64)}) "
                  f"want {float(want[i, j])}")
    # and the mirror: B coded, A one-hot
    n = torch.arange(N, device=dev, dtype=torch.float32)[:, None]
    Bc = n * 64 + k + 1
    Ah = torch.zeros(M, K, device=dev)
    idx = torch.arange(min(M, K), device=dev)
    Ah[idx, idx] = 1.0
    a = Ah.t().contiguous() if a_mn else Ah
    b = Bc.t().contiguous() if b_mn else Bc
    C = gemm(a, b, a_mn_major=a_mn, b_mn_major=b_mn)
    want = Ah @ Bc.t()
    bad = (C != want)
    print(f"probe B-coded: mismatches {int(bad.sum())}")
    if bad.any():
        for i, j in bad.nonzero()[:12].tolist():
            v = float(C[i, j]) - 1
            print(f"   C[{i},{j}] = {float(C[i, j])} -> (n={int(v // 64)}, k={int(v % 64)}) "
                  f"want {float(want[i, j])}")


if __name__ == "__main__":
    for a_mn in (False, True):
        for b_mn in (False, True):
            probe(128, 128, 32, a_mn, b_mn)
    probe(128, 128, 64, False, False)
    probe(256, 256, 32, False, False)
