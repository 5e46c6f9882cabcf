"""CPU model of the v6 exact sampler's decision logic (tools, not product).

The main warp of k_sample_exact_v6 decides every draw on an APPROXIMATE copy A of the
top of the sum tree (updated by plain subtractions, all levels at once) and on the exact
bottom levels, and accepts a decision only when it is farther than EPS from every
boundary it crosses; otherwise the draw falls back to the reference's exact sequential
arithmetic.  This model checks, on adversarial trees, that the accepted decisions always
equal the exact chain's, and counts the fallbacks.
"""
import sys

import numpy as np


def build(leaves):
    n = len(leaves)
    h = np.zeros(2 * n)
    h[n:] = leaves
    for i in range(n - 1, 0, -1):
        h[i] = h[2 * i] + h[2 * i + 1]
    return h


def exact_draw(h, n, u, older=2):
    """reference arithmetic (collections/prioritized.py:245-258, 294-312)"""
    pos = h[1] * u
    node = older
    if not (pos < h[older]):
        pos = pos - h[older]
        node = older ^ 1
    while node < n:
        left = h[2 * node]
        if pos < left:
            node = 2 * node
        else:
            pos = pos - left
            node = 2 * node + 1
    return node


def zero_and_reduce(h, node):
    h[node] = 0.0
    p = node >> 1
    while p >= 1:
        h[p] = h[2 * p] + h[2 * p + 1]
        p >>= 1


def run(leaves, us, T, eps_rel, older=2, mispredict=0.0, rng=None):
    """Prefix-table formulation of the kernel: M / P_lo / P_hi over the level-(T-1)
    nodes in descent order (older half first), running sums Q of the leaves under the
    predicted node; a draw is accepted only when pos is farther than EPS from both ends
    of the node's interval and of the leaf's interval."""
    n = len(leaves)
    L = n.bit_length() - 1          # leaves at level L
    E = build(leaves)
    Ex = E.copy()                   # pure exact sampler for comparison
    ntop = 1 << (T - 1)             # nodes at level T-1
    flip = 0 if older == 2 else ntop // 2
    blk = max(1, int(np.sqrt(ntop)))
    while ntop % blk:
        blk -= 1
    M = np.array([E[ntop + (o ^ flip)] for o in range(ntop)])
    P_lo = np.zeros(ntop)
    P_hi = np.zeros(ntop // blk)
    run_hi = 0.0
    for b in range(ntop // blk):
        P_hi[b] = run_hi
        r = 0.0
        for i in range(blk):
            P_lo[b * blk + i] = r
            r += M[b * blk + i]
        run_hi += r
    rootA = E[1]
    eps = eps_rel * E[1]
    D = L - (T - 1)
    nleaf = 1 << D
    fallbacks = 0
    out = []
    for u in us:
        want = exact_draw(Ex, n, u, older)
        # ---- scout: predicted node from the tables (optionally perturbed)
        pos = rootA * u
        b = max(0, int(np.searchsorted(P_hi, pos, side="right")) - 1)
        w = max(0, int(np.searchsorted(P_lo[b * blk:(b + 1) * blk], pos - P_hi[b], side="right")) - 1)
        o = b * blk + w
        if rng is not None and rng.rand() < mispredict:
            o = min(ntop - 1, max(0, o + rng.randint(-1, 2)))
        node = ntop + (o ^ flip)
        # running sums of the leaves under the node (what the scout lays down)
        lv = E[node * nleaf:(node + 1) * nleaf]
        Q = np.concatenate([[0.0], np.cumsum(lv)])
        # ---- main
        pos12 = (rootA * u - P_hi[o // blk]) - P_lo[o]
        leaf = None
        if eps < pos12 < M[o] - eps:
            i = max(0, int(np.searchsorted(Q[:nleaf], pos12, side="right")) - 1)
            if pos12 - Q[i] > eps and Q[i + 1] - pos12 > eps:
                leaf = node * nleaf + i
        if leaf is None:
            fallbacks += 1
            leaf = exact_draw(E, n, u, older)
        assert leaf == want, ("decision differs from the exact chain", leaf, want, u)
        prio = E[leaf]
        out.append(leaf)
        zero_and_reduce(E, leaf)      # exact ascent (ascent warp)
        zero_and_reduce(Ex, leaf)
        # approximate update: everything after the drawn node
        od = ((leaf >> D) - ntop) ^ flip
        M[od] -= prio
        bb = od // blk
        P_lo[od + 1:(bb + 1) * blk] -= prio
        P_hi[bb + 1:] -= prio
        rootA = rootA - prio
    assert np.array_equal(E, Ex)
    return out, fallbacks


def main():
    rng = np.random.RandomState(0)
    total_fb = total = 0
    cases = []
    n = 1 << 12
    cases.append(("uniform random", rng.rand(n) + 0.01))
    cases.append(("all equal (ties on power-of-two sums)", np.ones(n)))
    cases.append(("huge dynamic range", np.exp(rng.randn(n) * 12)))
    sp = np.zeros(n)
    sp[rng.randint(0, n, 300)] = rng.rand(300)
    cases.append(("sparse (mostly empty slots)", sp))
    cases.append(("tiny among big", np.where(rng.rand(n) < 0.5, 1e-30, 1.0)))
    for name, leaves in cases:
        for T in (4, 7, 9):
            for eps_rel in (2.0 ** -36, 2.0 ** -20, 0.5):
                k = min(256, int((leaves > 0).sum()))
                us = rng.random_sample(k)
                us[::17] = 0.0
                us[5::31] = np.nextafter(1.0, 0.0)
                us[3::13] = np.round(us[3::13] * 64) / 64     # boundaries of equal trees
                _, fb = run(leaves, us, T, eps_rel, older=2 + (T & 1), mispredict=0.05, rng=rng)
                total_fb += fb if eps_rel < 1e-6 else 0
                total += k if eps_rel < 1e-6 else 0
                print("%-40s T=%d eps=2^%-4d draws=%d fallbacks=%d" % (
                    name, T, int(np.log2(eps_rel)), k, fb))
    print("fallback rate at the production EPS: %d / %d" % (total_fb, total))


if __name__ == "__main__":
    sys.exit(main())
