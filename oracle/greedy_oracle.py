"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's greedy mutual-information placement.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / reference arm may import this
module; no product code does.  It restates `/root/reference/placement_algorithm2.py` in two forms:

* the **literal** form (``literal_*``) follows the reference line by line -- sub-matrix gathers, `pinv`,
  first-strict-maximum scans -- and is what the golden vectors in ``tests/golden/`` were checked
  against (the vectors themselves come from the reference's own functions, exec'd unmodified by
  ``oracle/ref_extract.py``).  It is O(n^4) per selection and only usable for n of a few hundred;
* the **incremental** form (``incremental_greedy``) is the algebraically identical O(n^2)-per-selection
  restatement that the CUDA path implements: sigma^2(y|A) by a growing Cholesky panel, and
  sigma^2(y | Abar \\ y) = 1 / (Sigma_AbarAbar^-1)_yy by rank-1 downdates of the precision.

Parity status: **pinned** -- both forms reproduce every golden vector in ``tests/golden/greedy_*.json``
(selection sequences bit-exactly, scores within 1e-10 relative).
"""
import numpy as np

GUARD_NUMPY = 1e-8      # placement_algorithm2.py:116,198
GUARD_TF_GRAPH = 1e-7   # snippets_a2.py:480
JITTER_TF_GRAPH = 1e-6  # snippets_a2.py:161-163


# --------------------------------------------------------------------------------------------------
# literal form
# --------------------------------------------------------------------------------------------------
def literal_make_slice(cov_vv, rows, cols):
    """Sub-matrix gather, placement_algorithm2.py:391-396 (the reference loops in Python)."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    out = np.zeros((len(rows), len(cols)))
    if len(rows) and len(cols):
        out[:, :] = cov_vv[np.ix_(rows, cols)]
    return out


def literal_call_pinv(a):
    """placement_algorithm2.py:399-405: reciprocal for 1x1, else SVD pseudo-inverse (rcond 1e-15)."""
    assert a.shape[0] == a.shape[1]
    if a.shape[0] == 1:
        return 1 / a
    return np.linalg.pinv(a)


def literal_nominator(y, A, cov_vv, jitter=0.0):
    """sigma^2(y | A), placement_algorithm2.py:371-388.  `jitter` is the TF-graph variant's diagonal
    shift of Sigma_AA (snippets_a2.py:161-163)."""
    A_ = [int(a) for a in A]
    sigm_yy = literal_make_slice(cov_vv, [y], [y])
    if len(A_) == 0:
        return sigm_yy
    cov_yA = literal_make_slice(cov_vv, [y], A_)
    cov_AA = literal_make_slice(cov_vv, A_, A_)
    cov_Ay = literal_make_slice(cov_vv, A_, [y])
    if jitter:
        cov_AA = cov_AA + jitter * np.eye(len(A_))
    return sigm_yy - np.dot(np.dot(cov_yA, literal_call_pinv(cov_AA)), cov_Ay)


def literal_denominator(y, A_hat, cov_vv, jitter=0.0):
    """sigma^2(y | Abar \\ y), placement_algorithm2.py:408-413."""
    rest = [int(a) for a in A_hat if int(a) != int(y)]
    return literal_nominator(y, rest, cov_vv, jitter)


def literal_delta(y, A, A_bar, cov_vv, small=GUARD_NUMPY, jitter=0.0):
    """Guarded ratio, placement_algorithm2.py:113-119 / :194-203."""
    nom = literal_nominator(y, A, cov_vv, jitter)
    den = literal_denominator(y, A_bar, cov_vv, jitter)
    if np.abs(den) < small or np.abs(nom) < small:
        return 0.0
    return float((nom / den).reshape(()))


def literal_placement_algorithm_1(cov_vv, k, small=GUARD_NUMPY, jitter=0.0, return_scores=False):
    """Naive greedy, placement_algorithm2.py:128-145 with argmax_ :105-125: first strict maximum,
    running best initialised to -1."""
    n = cov_vv.shape[0]
    A, A_bar = [], list(range(n))
    per_step = []
    while len(A) < k:
        y_st, delta_st = -1, -1
        scores = np.full(n, np.nan)
        for y in range(n):
            if y in A:
                continue
            d = literal_delta(y, A, A_bar, cov_vv, small, jitter)
            scores[y] = d
            if delta_st < d:
                delta_st, y_st = d, y
        A.append(y_st)
        A_bar.remove(y_st)          # raises ValueError for y_st == -1, as the reference does
        per_step.append(scores)
    return (A, per_step) if return_scores else A


def literal_placement_algorithm_2(cov_vv, k, small=GUARD_NUMPY, jitter=0.0):
    """Lazy greedy, placement_algorithm2.py:151-219 (cache starts at +inf, :164)."""
    n = cov_vv.shape[0]
    A, A_bar = [], list(range(n))
    cache = [np.inf] * n
    evaluations = []
    while len(A) < k:
        fresh = [False] * n
        while True:
            y_st, delta_st = -1, -1
            for y in range(n):          # argmax_cache_linear, :53-67
                if y in A:
                    continue
                if delta_st < cache[y]:
                    delta_st, y_st = cache[y], y
            if fresh[y_st]:
                break
            cache[y_st] = literal_delta(y_st, A, A_bar, cov_vv, small, jitter)
            fresh[y_st] = True
            evaluations.append((len(A), y_st, cache[y_st]))
        A.append(y_st)
        A_bar.remove(y_st)
    return A, evaluations


def _sparse_argmax(cache, A):
    """sparse_argmax_cache_linear, placement_algorithm2.py:24-50: max of the cache over V \\ A, first index achieving
    it (tf.where(...)[0, 0] over the index-ordered set difference)."""
    taken = np.zeros(len(cache), dtype=bool)
    taken[list(A)] = True
    s = np.where(taken, -np.inf, cache)
    return int(np.argmax(s))


def literal_sparse_placement_algorithm_2(cov_vv, k, small=GUARD_TF_GRAPH, jitter=JITTER_TF_GRAPH, inf=1e8):
    """The TF-graph lazy greedy, snippets_a2.py:679-822: cache starts at INF = 1e8 (:690), every entry stale at the
    start of a selection (:726), the arg-max entry is re-evaluated until an up-to-date one wins (while_true_outside),
    delta with the graph's jitter 1e-6 (:161-163) and guard 1e-7 (:480); the cache is stored in column len(A) - 1 of
    delta_cached_iters BEFORE the winner's entry is zeroed (:771-797).
    Returns (sorted A -- the graph's unordered sparse set --, len(A), delta_cached_iters [N, k],
    A_selection_and_delta [k, 2] = (index, delta) in selection order)."""
    n = cov_vv.shape[0]
    A, A_bar = [], list(range(n))
    cache = np.full(n, inf)
    dci = np.zeros((n, k))
    sel = np.zeros((k, 2))
    while len(A) < k:
        fresh = np.zeros(n, dtype=bool)
        while True:
            y = _sparse_argmax(cache, A)
            if fresh[y]:
                break
            cache[y] = literal_delta(y, A, A_bar, cov_vv, small, jitter)
            fresh[y] = True
        A.append(y)
        A_bar.remove(y)
        sel[len(A) - 1] = (y, cache[y])
        dci[:, len(A) - 1] = cache
        cache[y] = 0.0
    return sorted(A), len(A), dci, sel


def literal_sparse_placement_algorithm_3(cov_vv, k, cover_spatial, cutoff, small=GUARD_TF_GRAPH,
                                         jitter=JITTER_TF_GRAPH):
    """Local-kernel greedy, snippets_a3.py:43-364.  All deltas are evaluated once with A empty (whD, :71-120); then
    k - 1 times: arg-max of the cache over V \\ A, the winner joins A and its entry is zeroed, and ONLY the
    candidates inside the index box [i - cutoff, min(i + cutoff, I)) per axis around the winner (:231-262; index =
    I2 I1 i0 + I2 i1 + i2, :189-194) are re-evaluated against the new A (entries of selected points stay 0,
    :223-226); everything else keeps its stale value.  Column i + 1 of delta_cached_iters receives the cache after
    round i (:300-304); the k-th winner is taken from the last cache (:356-358).
    Returns (A in selection order, final cache [N], delta_cached_iters [N, k])."""
    i0n, i1n, i2n = (int(v) for v in cover_spatial)
    n = cov_vv.shape[0]
    assert n == i0n * i1n * i2n
    A, A_bar = [], list(range(n))
    cache = np.array([literal_delta(y, A, A_bar, cov_vv, small, jitter) for y in range(n)])
    dci = np.zeros((n, k))
    dci[:, 0] = cache
    for i in range(k - 1):
        y = _sparse_argmax(cache, A)
        A.append(y)
        A_bar.remove(y)
        cache[y] = 0.0
        c0 = y // (i1n * i2n)
        c1 = (y - c0 * i1n * i2n) // i2n
        c2 = y - c0 * i1n * i2n - c1 * i2n
        for j0 in range(max(c0 - cutoff, 0), min(c0 + cutoff, i0n)):
            for j1 in range(max(c1 - cutoff, 0), min(c1 + cutoff, i1n)):
                for j2 in range(max(c2 - cutoff, 0), min(c2 + cutoff, i2n)):
                    yj = i1n * i2n * j0 + i2n * j1 + j2
                    if yj in A:
                        cache[yj] = 0.0
                    else:
                        cache[yj] = literal_delta(yj, A, A_bar, cov_vv, small, jitter)
        snap = cache.copy()
        snap[y] = 0.0
        dci[:, i + 1] = snap
    y = _sparse_argmax(cache, A)
    A.append(y)
    return A, cache, dci


# --------------------------------------------------------------------------------------------------
# incremental form (what the CUDA kernels compute)
# --------------------------------------------------------------------------------------------------
def spd_inverse(cov):
    """Sigma^-1 through LAPACK potrf + potri (the same factor-then-invert route the GPU path takes)."""
    from scipy.linalg import lapack
    c, info = lapack.dpotrf(cov, lower=1, clean=1, overwrite_a=0)
    if info != 0:
        raise np.linalg.LinAlgError("covariance is not positive definite (potrf info=%d)" % info)
    inv, info = lapack.dpotri(c, lower=1, overwrite_c=1)
    if info != 0:
        raise np.linalg.LinAlgError("potri info=%d" % info)
    inv = np.tril(inv)
    return inv + np.tril(inv, -1).T


def guarded_scores(num, den, small):
    """delta = 0 where |den| < small or |num| < small, else num/den (placement_algorithm2.py:116-119)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = num / den
    bad = (np.abs(den) < small) | (np.abs(num) < small)
    return np.where(bad, 0.0, ratio)


def first_argmax(scores, taken):
    """First index of the strict maximum over candidates not yet taken, running best starting at -1
    (placement_algorithm2.py:106-123).  Returns -1 when no score exceeds -1."""
    s = np.where(taken, -np.inf, scores)
    y = int(np.argmax(s))               # np.argmax returns the first maximal index
    return y if s[y] > -1 else -1


class IncrementalState:
    """State of the incremental greedy over a contiguous block of candidate columns [c0, c1).

    One instance over [0, n) is the single-device algorithm; G instances over a partition of [0, n)
    are the shards of the multi-GPU layout.  All updates are element-wise in
    the column index, so the results do not depend on the partition.
    """

    def __init__(self, cov_cols, prec_cols, c0, small=GUARD_NUMPY, jitter=0.0):
        self.c0 = int(c0)
        self.n = cov_cols.shape[0]
        self.nloc = cov_cols.shape[1]
        self.cov = cov_cols                       # Sigma[:, c0:c1]   (row y = Sigma[y, c0:c1])
        self.P = np.array(prec_cols, dtype=np.float64, order="C")   # precision panel, updated in place
        self.small = small
        self.jitter = jitter
        idx = np.arange(self.nloc)
        self.num = np.array(cov_cols[self.c0 + idx, idx], dtype=np.float64)
        if jitter:
            self.num = self.num + jitter          # run on Sigma + jitter*I, subtract it at scoring
        self.W = np.zeros((0, self.nloc))         # growing conditioning panel, rows = selections
        self.taken = np.zeros(self.nloc, dtype=bool)

    def local_scores(self):
        idx = np.arange(self.nloc)
        pdiag = self.P[self.c0 + idx, idx]
        with np.errstate(divide="ignore"):
            den = 1.0 / pdiag
        return guarded_scores(self.num - self.jitter, den - self.jitter, self.small)

    def local_best(self):
        """(score, global index, numerator) of this block's first strict maximum; index -1 if none."""
        s = self.local_scores()
        j = first_argmax(s, self.taken)
        if j < 0:
            return -np.inf, -1, 0.0
        return float(s[j]), self.c0 + j, float(self.num[j])

    def segments(self, y, num_y, w_hist_y):
        """Row segments for the exchange: w = (Sigma[y, J] - W^T W[:, y]) / sqrt(num_y), p = P[y, J].
        `w_hist_y` holds W[0..t-1, y] (column y of the full panel)."""
        row = self.cov[y, :].astype(np.float64, copy=True)
        if self.jitter and self.c0 <= y < self.c0 + self.nloc:
            row[y - self.c0] += self.jitter
        for s in range(self.W.shape[0]):          # fixed summation order over earlier selections
            row -= self.W[s, :] * w_hist_y[s]
        return row / np.sqrt(num_y), self.P[y, :].copy()

    def apply(self, y, w_seg, p_full):
        """Condition the numerators on y and drop y from the precision of the unselected set."""
        self.W = np.vstack([self.W, w_seg[None, :]])
        self.num = self.num - w_seg * w_seg
        inv_pivot = 1.0 / p_full[y]
        p_loc = p_full[self.c0:self.c0 + self.nloc]
        self.P -= (p_full[:, None] * p_loc[None, :]) * inv_pivot
        self.P[y, :] = 0.0                         # exact zeros instead of rounding residue
        if self.c0 <= y < self.c0 + self.nloc:
            self.P[:, y - self.c0] = 0.0
            self.taken[y - self.c0] = True


def pick_winner(candidates):
    """Global winner among per-block (score, index, num) triples: largest score, lowest index on exact
    ties -- what a single ascending scan with strict '<' returns (placement_algorithm2.py:121)."""
    best = (-np.inf, -1, 0.0)
    for c in candidates:
        if c[1] < 0:
            continue
        if c[0] > best[0] or (c[0] == best[0] and c[1] < best[1]) or best[1] < 0:
            best = c
    return best


def incremental_greedy(cov_vv, k, small=GUARD_NUMPY, jitter=0.0, shards=1, prec=None,
                       return_all_scores=False):
    """O(n^2)-per-selection greedy.  Returns (selection list, winning scores [k], per-step score
    vectors [k, n] or None, min relative top-2 gap per step [k])."""
    cov_vv = np.ascontiguousarray(cov_vv, dtype=np.float64)
    n = cov_vv.shape[0]
    if prec is None:
        prec = spd_inverse(cov_vv + jitter * np.eye(n) if jitter else cov_vv)
    bounds = [(n * g) // shards for g in range(shards + 1)]
    blocks = [IncrementalState(cov_vv[:, a:b], prec[:, a:b], a, small, jitter)
              for a, b in zip(bounds[:-1], bounds[1:])]
    selection, win_scores, gaps = [], [], []
    all_scores = [] if return_all_scores else None
    w_rows = np.zeros((0, n))
    for _ in range(k):
        full = np.concatenate([b.local_scores() for b in blocks])
        taken = np.concatenate([b.taken for b in blocks])
        score, y, num_y = pick_winner([b.local_best() for b in blocks])
        if y < 0:
            raise ValueError("list.remove(x): x not in list")   # the reference's failure mode (:144)
        masked = np.where(taken, -np.inf, full)
        if all_scores is not None:
            all_scores.append(np.where(taken, np.nan, full))
        top2 = np.partition(masked, -2)[-2:] if n - len(selection) >= 2 else None
        gaps.append(float((top2[1] - top2[0]) / abs(top2[1])) if top2 is not None and top2[1] != 0
                    else np.inf)
        segs = [b.segments(y, num_y, w_rows[:, y]) for b in blocks]
        w_full = np.concatenate([s[0] for s in segs])
        p_full = np.concatenate([s[1] for s in segs])
        for b, s in zip(blocks, segs):
            b.apply(y, s[0], p_full)
        w_rows = np.vstack([w_rows, w_full[None, :]])
        selection.append(y)
        win_scores.append(score)
    return selection, np.array(win_scores), (np.array(all_scores) if all_scores is not None else None), \
        np.array(gaps)


def replay_lazy_evaluations(step_scores, selection):
    """Replay alg. 2's lazy cache (placement_algorithm2.py:157-214) on dense per-step score vectors
    [k, n]; returns the (step, y, delta) evaluation trace it would print."""
    k, n = step_scores.shape
    cache = np.full(n, np.inf)
    taken = np.zeros(n, dtype=bool)
    trace, picks = [], []
    for t in range(k):
        fresh = np.zeros(n, dtype=bool)
        while True:
            y = first_argmax(cache, taken)
            if fresh[y]:
                break
            cache[y] = step_scores[t, y]
            fresh[y] = True
            trace.append((t, y, float(cache[y])))
        picks.append(y)
        taken[y] = True
    return picks, trace


# --------------------------------------------------------------------------------------------------
# C/OpenMP form of the same step (oracle/greedy_step.c) -- CPU baseline for bench.py
# --------------------------------------------------------------------------------------------------
def _load_c():
    import ctypes
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libgreedy_oracle.so")
    if not os.path.exists(path):
        try:
            from . import build_oracle
            build_oracle.build()
        except Exception:
            return None
    try:
        lib = ctypes.CDLL(path)
    except OSError:
        return None
    i64, dbl, vp = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
    lib.oracle_scores.restype = i64
    lib.oracle_scores.argtypes = [vp, i64, i64, i64, vp, vp, dbl, dbl, vp, ctypes.POINTER(dbl)]
    lib.oracle_segments.restype = None
    lib.oracle_segments.argtypes = [vp, vp, i64, i64, i64, vp, i64, i64, dbl, i64, dbl, vp, vp]
    lib.oracle_apply.restype = None
    lib.oracle_apply.argtypes = [vp, i64, i64, i64, i64, vp, vp, i64, vp, vp]
    lib.oracle_set_threads.restype = ctypes.c_int
    lib.oracle_set_threads.argtypes = [ctypes.c_int]
    return lib


def set_threads(n):
    """Thread count of the C/OpenMP step (omp_set_num_threads: overrides an OMP_NUM_THREADS=1 exported by a launcher
    such as torch.distributed.run).  Returns the count in effect, or None without the C library."""
    lib = _load_c()
    if lib is None:
        return None
    return int(lib.oracle_set_threads(int(n)))


def incremental_greedy_c(cov_vv, k, small=GUARD_NUMPY, jitter=0.0, prec=None, timings=None, all_scores=None):
    """Single-block incremental greedy with the per-step work in C/OpenMP.  Returns (selection, scores).
    `timings`, if a dict, receives 'setup_s' (inverse) and 'steps_s' (list of per-selection seconds); `all_scores`, if
    a list, receives every step's full score vector (NaN where taken), as incremental_greedy(return_all_scores=True)."""
    import ctypes
    import time
    lib = _load_c()
    if lib is None:
        sel, sc, _, _ = incremental_greedy(cov_vv, k, small, jitter, prec=prec)
        return sel, sc
    cov = np.ascontiguousarray(cov_vv, dtype=np.float64)
    n = cov.shape[0]
    t0 = time.perf_counter()
    if prec is None:
        prec = spd_inverse(cov + jitter * np.eye(n) if jitter else cov)
    P = np.ascontiguousarray(prec, dtype=np.float64).copy()
    if timings is not None:
        timings["setup_s"] = time.perf_counter() - t0
        timings["steps_s"] = []
    num = np.ascontiguousarray(np.diag(cov) + jitter)
    taken = np.zeros(n, dtype=np.uint8)
    wfull = np.zeros((k, n))
    scores = np.empty(n)
    w_seg, p_seg = np.empty(n), np.empty(n)
    selection, win = [], []
    ptr = lambda a: a.ctypes.data  # noqa: E731
    for t in range(k):
        t1 = time.perf_counter()
        bs = ctypes.c_double(0.0)
        y = lib.oracle_scores(ptr(P), n, 0, n, ptr(num), ptr(taken), small, jitter, ptr(scores), ctypes.byref(bs))
        if y < 0:
            raise ValueError("list.remove(x): x not in list")
        if all_scores is not None:
            all_scores.append(scores.copy())
        lib.oracle_segments(ptr(cov), ptr(P), n, 0, n, ptr(wfull), n, t, jitter, y, float(num[y]), ptr(w_seg),
                            ptr(p_seg))
        wfull[t] = w_seg
        lib.oracle_apply(ptr(P), n, n, 0, n, ptr(w_seg), ptr(p_seg), y, ptr(num), ptr(taken))
        selection.append(int(y))
        win.append(bs.value)
        if timings is not None:
            timings["steps_s"].append(time.perf_counter() - t1)
    return selection, np.array(win)


# --------------------------------------------------------------------------------------------------
# rank-deficient covariances: the reference's pinv semantics in closed form
# --------------------------------------------------------------------------------------------------
def pinv_step_scores(cov_vv, A, small=GUARD_NUMPY, rank_tol=1e-12, leverage_tol=1e-6):
    """delta_y for every y not in A on a positive SEMI-definite cov_vv, as the reference computes them with
    np.linalg.pinv (placement_algorithm2.py:371-413), from two symmetric eigendecompositions instead of one SVD per
    candidate:

      nominator    sigma^2(y | A)        = Sigma_yy - Sigma_yA pinv(Sigma_AA) Sigma_Ay
      denominator  sigma^2(y | Abar \\ y) = 0 when e_y is not in the range of Sigma_AbarAbar (y's factor-space vector lies
                   in the span of the others: leverage (Sigma Sigma^+)_yy < 1), else 1 / (Sigma_AbarAbar^+)_yy

    (block-inverse identity on the range of Sigma_AbarAbar).  Eigenvalues below rank_tol x the largest are treated as
    zero -- on exactly rank-deficient inputs they are rounding noise ~1e-17, far below either this cut or pinv's 1e-15."""
    cov = np.asarray(cov_vv, dtype=np.float64)
    n = cov.shape[0]
    A = [int(a) for a in A]
    rest = np.array([v for v in range(n) if v not in A], dtype=np.int64)
    scale = float(np.max(np.diag(cov)))
    nom = np.diag(cov)[rest].copy()
    if A:
        w, v = np.linalg.eigh(cov[np.ix_(A, A)])
        keep = w > rank_tol * max(w.max(), 0.0) if w.max() > 0 else np.zeros_like(w, dtype=bool)
        proj = (v[:, keep].T @ cov[np.ix_(A, rest)]) / np.sqrt(w[keep])[:, None]
        nom = nom - np.sum(proj * proj, axis=0)
    w, v = np.linalg.eigh(cov[np.ix_(rest, rest)])
    keep = w > rank_tol * scale
    lev = np.sum(v[:, keep] ** 2, axis=1)
    pinv_diag = np.sum(v[:, keep] ** 2 / w[keep][None, :], axis=1)
    den = np.where((lev >= 1.0 - leverage_tol) & (pinv_diag > 0), 1.0 / np.where(pinv_diag > 0, pinv_diag, 1.0), 0.0)
    out = np.full(n, np.nan)
    out[rest] = guarded_scores(nom, den, small)
    return out


def pinv_greedy(cov_vv, k, algorithm=1, small=GUARD_NUMPY):
    """placement_algorithm_1 (algorithm=1: first strict maximum of the fresh scores, :128-145) or placement_algorithm_2
    (algorithm=2: the lazy cache, :151-219 -- NOT equivalent to alg. 1 here, because on rank-deficient inputs a delta
    can rise from 0 to a positive value between steps, which the stale cache never sees) on a rank-deficient PSD
    matrix.  Returns (selection, winning deltas, per-step score vectors [k, n])."""
    n = np.asarray(cov_vv).shape[0]
    A, win, steps = [], [], []
    cache = np.full(n, np.inf)
    taken = np.zeros(n, dtype=bool)
    for _ in range(k):
        scores = pinv_step_scores(cov_vv, A, small)
        steps.append(scores)
        if algorithm == 1:
            y = first_argmax(np.nan_to_num(scores, nan=-np.inf), taken)
        else:
            fresh = np.zeros(n, dtype=bool)
            while True:
                y = first_argmax(cache, taken)
                if y < 0 or fresh[y]:
                    break
                cache[y] = scores[y]
                fresh[y] = True
        if y < 0:
            raise ValueError("list.remove(x): x not in list")
        win.append(float(scores[y]))
        A.append(int(y))
        taken[y] = True
    return A, np.array(win), np.array(steps)


def pinv_well_posed(cov_vv, selection, lo=1e-15, hi=1e-9):
    """Is the pinv greedy numerically well-posed along `selection`?  Returns (True, -1), or (False, t) for the first
    selection t at which Sigma_AbarAbar has an eigenvalue between `lo` and `hi` of the largest variance: whether such a
    direction counts as part of the range decides denominators by orders of magnitude, and np.linalg.pinv's own answer
    (rcond 1e-15 on singular values it computes to ~1e-16 absolute) is rounding noise there.  Typical place: the one
    selection at which card Abar equals the rank.  Parity tests assert this on their inputs, like the top-2 gap."""
    cov = np.asarray(cov_vv, dtype=np.float64)
    n = cov.shape[0]
    scale = float(np.max(np.diag(cov)))
    for t in range(len(selection)):
        taken = set(int(v) for v in selection[:t])
        rest = [v for v in range(n) if v not in taken]
        w = np.linalg.eigvalsh(cov[np.ix_(rest, rest)]) / scale
        if np.any((w > lo) & (w < hi)):
            return False, t
    return True, -1
