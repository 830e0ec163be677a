"""Where the factorisation time goes: a launch-by-launch model of potrf + trtri as csrc/dense.cu issues them.

Walks the same recursion (split(n) = 128 * (n / 128 / 2); 128-block leaves) and lists every GEMM launch with its shape;
times each with a two-parameter model fitted to the measured sweep (profiles/r01_gemm_sweep_tma.json): a launch runs
ceil(tiles / 296) waves of 128 x 64 tiles (two CTAs per SM), a wave costs t0 + k * t_k.  Reports the split of the time
by product size and, for G ranks, what the distributed inverse spends on products too small to distribute (run by every
rank) and on the two flag barriers around each distributed one.

    python tools/factor_schedule_model.py [n] [ranks ...]
"""
import json
import math
import os
import sys

NB = 128
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def split(n):
    n1 = (n // NB // 2) * NB
    return NB if n1 < NB else n1


class Trace:
    def __init__(self, slab_width=0, leaf=0):
        self.gemms, self.leaves, self.slabs, self.slab_width, self.leaf = [], 0, [], slab_width, leaf
        self.copies = 0                         # doubles copied back from scratch by the wide-leaf solves

    def wide(self, m, n, left=False):           # VGP_TRSM_LEAF: one product with the node's cached explicit inverse
        if self.leaf and NB < n <= self.leaf:
            self.gemm(n, m, n) if left else self.gemm(m, n, n)
            self.copies += m * n
            return True
        return False

    def slab(self, m, n):                       # one launch: m / 128 slabs, each walks an n-wide triangle (slab.cu)
        if self.slab_width and NB < n <= self.slab_width:
            self.slabs.append((m, n))
            return True
        return False

    def gemm(self, m, n, k, lower=False):
        self.gemms.append((m, n, k, lower))

    # -- the recursions of dense.cu (shapes only)
    def trsm_right(self, m, n):                 # both right-side forms: leaves m x 128 x 128, updates m x n2 x n1
        if self.slab(m, n) or self.wide(m, n):
            return
        if n == NB:
            return self.gemm(m, NB, NB)
        n1 = split(n)
        self.trsm_right(m, n1)
        self.gemm(m, n - n1, n1)
        self.trsm_right(m, n - n1)

    def trsm_left(self, n, nrhs):
        if self.slab(nrhs, n) or self.wide(nrhs, n, left=True):
            return
        if n == NB:
            return self.gemm(NB, nrhs, NB)
        n1 = split(n)
        self.trsm_left(n1, nrhs)
        self.gemm(n - n1, nrhs, n1)
        self.trsm_left(n - n1, nrhs)

    def potrf(self, n):
        if n == NB:
            self.leaves += 1
            return
        n1 = split(n)
        self.potrf(n1)
        self.trsm_right(n - n1, n1)
        self.gemm(n - n1, n - n1, n1, lower=True)
        self.potrf(n - n1)

    def trtri(self, n):
        if n == NB:
            self.leaves += 1
            return
        n1 = split(n)
        self.trsm_right(n - n1, n1)
        self.trsm_left(n - n1, n1)
        self.trtri(n1)
        self.trtri(n - n1)


def fit():
    """(t0, t_k) in microseconds from the square and panel shapes of the sweep."""
    rows = json.load(open(os.path.join(ROOT, "profiles", "r01_gemm_sweep_tma.json")))
    pts = []
    for r in rows:
        tiles = (r["m"] // 128) * (r["n"] // 64)
        waves = math.ceil(tiles / 296)
        pts.append((waves, waves * r["k"], r["ours_ms"] * 1e3))
    # least squares  t = t0 * waves + t_k * waves * k
    sxx = sum(a * a for a, _, _ in pts)
    sxy = sum(a * b for a, b, _ in pts)
    syy = sum(b * b for _, b, _ in pts)
    sxt = sum(a * t for a, _, t in pts)
    syt = sum(b * t for _, b, t in pts)
    det = sxx * syy - sxy * sxy
    t0 = (sxt * syy - syt * sxy) / det
    tk = (sxx * syt - sxy * sxt) / det
    return max(t0, 0.0), tk


def gemm_us(m, n, k, lower, t0, tk, share=1):
    tiles = (m // 128) * (n // 64)
    if lower:
        tiles = (m // 128) * (m // 128 + 1)            # lower 128 x 128 tiles, two 128 x 64 halves each
    tiles = math.ceil(tiles / share)
    return math.ceil(tiles / 296) * (t0 + tk * k)


def main():
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    opts = dict(a[2:].split("=") for a in sys.argv[1:] if a.startswith("--"))
    slab_width = int(opts.get("slab", 0))              # --slab=1024: solves of up to that width in one launch
    emulate = float(opts.get("emulate", 1.0))          # --emulate=2.5: products with m, n >= 2048, k >= 1024 that much faster
    leaf = int(opts.get("leaf", 0))                    # --leaf=512: solves against nodes of up to that size as one product
    n = int(argv[0]) if argv else 50000
    ranks = [int(v) for v in argv[1:]] or [1, 2, 4, 8]
    n_pad = (n + NB - 1) // NB * NB
    t0, tk = fit()
    tr = Trace(slab_width, leaf)
    tr.potrf(n_pad)
    tr.trtri(n_pad)
    launch_us, leaf_us, barrier_us = 3.0, 45.0, 12.0     # launch gap of a short kernel; potf2 + block inverse; flag barrier
    out = {"n": n, "n_pad": n_pad, "options": {"slab_width": slab_width, "emulated_speedup": emulate, "wide_leaf": leaf},
           "gemm_launches": len(tr.gemms), "slab_launches": len(tr.slabs), "diagonal_block_leaves": tr.leaves,
           "wave_model_us": {"t0": t0, "per_k": tk}, "ranks": {}}
    flops = sum((m * (m + 128) if lo else 2 * m * nn) * k for m, nn, k, lo in tr.gemms)
    out["gemm_flop"] = flops
    for g in ranks:
        classes = {"k<=128": [0, 0.0], "128<k<=1024": [0, 0.0], "k>1024": [0, 0.0]}
        dist_t = repl_t = 0.0
        ndist = 0
        for m, nn, k, lo in tr.gemms:
            tiles128 = (m // 128) * (m // 128 + 1) // 2 if lo else (m // 128) * (nn // 128)
            distributed = g > 1 and tiles128 >= 96 and k >= 256
            t = gemm_us(m, nn, k, lo, t0, tk, g if distributed else 1)
            if emulate != 1.0 and m >= 2048 and nn >= 2048 and 2 * k >= 2048:
                t /= emulate
            t += launch_us
            key = "k<=128" if k <= 128 else ("128<k<=1024" if k <= 1024 else "k>1024")
            classes[key][0] += 1
            classes[key][1] += t
            if distributed:
                ndist += 1
                dist_t += t
            else:
                repl_t += t
        # slab launches: one CTA per SM at the full per-SM DMMA rate (35.5 TF / 148), 128 (n^2 + 128 n) flop per slab,
        # slabs shared out over the ranks when there are at least two per rank
        slab_t, slab_barriers = 0.0, 0
        for m, nn in tr.slabs:
            slabs = m // 128
            share = g if (g > 1 and slabs >= 2 * g) else 1
            slab_barriers += 2 if share > 1 else 0
            per_cta_us = 128.0 * (nn * nn + 128.0 * nn) / (35.5e12 / 148) * 1e6
            slab_t += math.ceil(math.ceil(slabs / share) / 148) * (t0 + per_cta_us) + launch_us
        copy_t = tr.copies * 16 / 2.0e12 * 1e6            # pitched device copies: ~2 TB/s of read + write
        total = dist_t + repl_t + slab_t + copy_t + tr.leaves * leaf_us + (2 * ndist + slab_barriers) * barrier_us
        out["ranks"][str(g)] = {
            "modelled_seconds": total / 1e6,
            "distributed_products": ndist, "distributed_seconds": dist_t / 1e6,
            "replicated_products_seconds": repl_t / 1e6,
            "slab_solves_seconds": slab_t / 1e6,
            "diagonal_blocks_seconds": tr.leaves * leaf_us / 1e6,
            "barrier_seconds": 2 * ndist * barrier_us / 1e6,
            "by_k": {k: {"launches": v[0], "seconds": v[1] / 1e6} for k, v in classes.items()},
            "ideal_seconds_at_35.5_TF": flops / 35.5e12 / g}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
