"""CPU prototype of FP64 GEMM emulation on 8-bit integer tensor cores (Ozaki scheme I), for the factorisation GEMMs.

B200's FP64 tensor pipe (DMMA) tops out at ~37 TFLOP/s; its int8 tcgen05 path is nominally 4.5 POP/s.  Splitting each
FP64 operand into s slices of w = 7 bits (row-wise exponent for A, column-wise for B), multiplying slice pairs exactly
in int32 and recombining in FP64 reproduces a DGEMM with s (s + 1) / 2 integer products.  This script measures, on the
matrices of this workload, how many slices are needed for (a) a product as accurate as a native FP64 GEMM and (b) the
same greedy selections and scores after a Cholesky + triangular inverse built on the emulated product.

    python tools/ozaki_prototype.py [n] [k]
"""
import json
import sys

import numpy as np

W = 7


def slices_rows(a, s):
    """a [m, k] -> (q [s, m, k] integer-valued float64 in [-127, 127], e [m]) with a = 2^e * sum_t q_t 2^(-W (t + 1))."""
    mx = np.max(np.abs(a), axis=1)
    e = np.where(mx > 0, np.floor(np.log2(np.where(mx > 0, mx, 1.0))) + 1, 0.0)
    r = a * np.exp2(-e)[:, None]                 # |r| < 1, exact scaling
    q = np.empty((s,) + a.shape)
    for t in range(s):
        r = r * (1 << W)
        q[t] = np.trunc(r)
        r = r - q[t]
    return q, e


def gemm_emulated(a, b, s):
    """a @ b through s (s + 1) / 2 exact integer products (float64 matmul of small integers is exact here)."""
    qa, ea = slices_rows(a, s)
    qb, eb = slices_rows(np.ascontiguousarray(b.T), s)
    assert a.shape[1] * 127 * 127 * s < 2 ** 31, "int32 accumulator would overflow: chunk k"
    c = np.zeros((a.shape[0], b.shape[1]))
    for g in range(s - 1, -1, -1):               # smallest terms first
        acc = np.zeros_like(c)
        for i in range(g + 1):
            acc += qa[i] @ qb[g - i].T           # one int8 GEMM, int32 accumulation shared by the group
        c += acc * 2.0 ** (-W * (g + 2))
    return c * np.exp2(ea)[:, None] * np.exp2(eb)[None, :]


def cloud_cov(n, seed=0, nugget=1e-2):
    x = np.random.default_rng(seed).uniform(-2, 2, (n, 3))
    ls = 0.5 * (1000.0 / max(n, 1000)) ** (1 / 3)
    d = x[:, None, :] - x[None, :, :]
    return np.exp(-np.einsum("ijk,ijk->ij", d, d) / (2 * ls * ls)) + nugget * np.eye(n)


def chol_inv_factor(a, gemm, nb=128):
    """Blocked right-looking Cholesky, then M = L^-1 by blocked forward substitution; all O(n^3) work through gemm."""
    n = a.shape[0]
    l = np.tril(a.copy())
    for j in range(0, n, nb):
        e = min(j + nb, n)
        l[j:e, j:e] = np.linalg.cholesky(l[j:e, j:e] + np.tril(l[j:e, j:e], -1).T)
        if e < n:
            inv = np.linalg.inv(l[j:e, j:e])
            l[e:, j:e] = gemm(l[e:, j:e], inv.T)
            l[e:, e:] -= np.tril(gemm(l[e:, j:e], l[e:, j:e].T))
    m = np.zeros_like(l)
    for j in range(0, n, nb):
        e = min(j + nb, n)
        m[j:e, j:e] = np.linalg.inv(l[j:e, j:e])
    for j in range(0, n, nb):                    # M[i, j] = -M[i, i] * sum_{j <= p < i} L[i, p] M[p, j]
        e = min(j + nb, n)
        for i in range(e, n, nb):
            ie = min(i + nb, n)
            m[i:ie, j:e] = -gemm(m[i:ie, i:ie], gemm(l[i:ie, j:i], m[j:i, j:e]))
    return l, m


def greedy_from_inverse_factor(cov, m, k):
    """Krause greedy on P0 = M^T M by the rank-1 downdate (the oracle's incremental form)."""
    p = m.T @ m
    n = cov.shape[0]
    num = np.diag(cov).copy()
    sel, scores = [], []
    w_hist = []
    taken = np.zeros(n, bool)
    for _ in range(k):
        den = 1.0 / np.diag(p)
        sc = np.where(taken, -np.inf, num / den)
        y = int(np.argmax(sc))
        sel.append(y)
        scores.append(sc[y])
        taken[y] = True
        col = cov[:, y].copy()
        for (u, piv) in w_hist:
            col -= u * (u[y] / piv)
        piv = col[y]
        w_hist.append((col, piv))
        num = num - col * col / piv
        py = p[:, y].copy()
        p -= np.outer(py, py) / py[y]
    return sel, np.array(scores)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    cov = cloud_cov(n)
    out = {"n": n, "k": k, "bits_per_slice": W, "gemm_error_vs_longdouble": {}, "placement": {}}
    # (a) one product of the factorisation's kind: L21 * L21^T with L from the workload
    l = np.linalg.cholesky(cov)
    a = l[n // 2:, :n // 2]
    exact = (a.astype(np.longdouble) @ a.T.astype(np.longdouble))
    scale = float(np.max(np.abs(exact)))
    out["gemm_error_vs_longdouble"]["native_fp64"] = float(np.max(np.abs(a @ a.T - exact)) / scale)
    for s in (5, 6, 7, 8, 9):
        out["gemm_error_vs_longdouble"]["s=%d (%d int8 GEMMs)" % (s, s * (s + 1) // 2)] = \
            float(np.max(np.abs(gemm_emulated(a, a.T, s) - exact)) / scale)
    # (b) placement on factors built with the emulated product
    l0, m0 = chol_inv_factor(cov, lambda x, y: x @ y)
    want_sel, want_sc = greedy_from_inverse_factor(cov, m0, k)
    for s in (6, 7, 8):
        _, ms = chol_inv_factor(cov, lambda x, y, s=s: gemm_emulated(x, y, s))
        sel, sc = greedy_from_inverse_factor(cov, ms, k)
        out["placement"]["s=%d" % s] = {
            "selection_identical": sel == want_sel,
            "max_rel_score_diff": float(np.max(np.abs(sc - want_sc) / np.abs(want_sc))),
            "max_rel_factor_diff": float(np.max(np.abs(ms - m0)) / np.max(np.abs(m0)))}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
