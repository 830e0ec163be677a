"""Drop-in for the reference's `placement_algorithm2.py` (same names, argument meaning, return types and
error behaviour), computing on a B200 through libvgposp.so.

    import vgposp_b200.placement_algorithm2 as alg2          # instead of: import placement_algorithm2 as alg2
    A = alg2.placement_algorithm_2(cov_vv, k)                # main.py:605, snippets_a2.py:964

What runs where: the O(n^3) setup (Cholesky + inverse of cov_vv) and every selection step run in CUDA
kernels; this module only moves arrays and mirrors the reference's host-visible behaviour (list of
np.int64 out, the per-evaluation prints of alg. 2, the ValueError when candidates run out).  There is no CPU
fallback: without a GPU every function that needs arithmetic raises.

Semantics, documented in DESIGN.md:
  * the reference takes `np.linalg.pinv` of Sigma_AA and of Sigma_{Abar\\y}.  A symmetric positive definite cov_vv
    (pinv == inv) runs on the Cholesky-based kernels; when that factorisation meets a pivot below 1e-12 of the
    largest variance -- a rank-deficient covariance such as the empirical M M^T / S the reference really feeds
    (main_architecture_2.py:391-444) -- the call continues on the pseudo-inverse path (csrc/pinv.cu), which
    reproduces the reference's results there (golden vectors: tests/golden/greedy_lowrank_golden.json).  A matrix
    that is not even positive semi-definite raises `NotPositiveDefiniteError`;
  * on positive definite inputs alg. 1 and alg. 2 return the same selection (the lazy evaluation of alg. 2 is an
    optimisation of the same arg-max while deltas only fall), so both are served by the dense per-step scores and
    alg. 2's print trace is replayed from them on the host; on rank-deficient inputs the two DIFFER in the reference
    (a delta can rise from the guarded 0), and the pseudo-inverse path follows each algorithm's own rule;
  * `argmax_` needs A_bar == V \\ A (every caller in the reference passes that).
"""
import time

import numpy as np

from . import _ffi, greedy as _greedy
from ._ffi import NotPositiveDefiniteError, VgpError  # noqa: F401  (re-exported)

PRINTS = True          # the reference prints per evaluation (placement_algorithm2.py:188,205)
DEVICE = 0


# --------------------------------------------------------------------------------------------------
# fixtures / generators (host-side data, no arithmetic of the path)
# --------------------------------------------------------------------------------------------------
def cov_vv_4x4():
    """The reference's 4x4 fixture (placement_algorithm2.py:473-479)."""
    rows = ((1.10, 0.31, 0.33, 0.27),
            (0.31, 1.01, 0.30, 0.27),
            (0.33, 0.30, 0.97, 0.33),
            (0.27, 0.27, 0.33, 1.2))
    return np.array(rows, dtype=np.float64)


def dg_create_random_cov(n):
    """U U^T with U ~ Uniform(0,1)^{n x n} from NumPy's global RNG, as placement_algorithm2.py:441-444 (the
    product runs through vgp_dgemm)."""
    u = np.random.uniform(0, 1, n ** 2).reshape(-1, n)
    ud = _ffi.DeviceArray.from_host(u, DEVICE)
    out = _ffi.DeviceArray((n, n), np.float64, DEVICE)
    _ffi.call("vgp_dgemm", DEVICE, 0, 1, n, n, n, 1.0, ud.ptr, n, ud.ptr, n, 0.0, out.ptr, n, None)
    return out.to_host()


# --------------------------------------------------------------------------------------------------
# single-candidate pieces (placement_algorithm2.py:371-413)
# --------------------------------------------------------------------------------------------------
def make_slice(cov_vv, y, A):
    """Sub-matrix cov_vv[y, A] (placement_algorithm2.py:391-396) -- a gather, no arithmetic."""
    y = [int(v) for v in y]
    A = [int(v) for v in A]
    out = np.zeros(shape=[len(y), len(A)])
    if y and A:
        out[:, :] = np.asarray(cov_vv)[np.ix_(y, A)]
    return out


def call_pinv(a):
    """placement_algorithm2.py:399-405.  1x1 -> reciprocal; else the inverse of an SPD matrix on device."""
    a = np.asarray(a, dtype=np.float64)
    assert a.shape[0] == a.shape[1]
    if a.shape[0] == 1:
        return 1 / a
    d = _ffi.DeviceArray.from_host(a, DEVICE)
    info = _ffi.c_int(0)
    _ffi.call("vgp_spd_inverse", DEVICE, d.ptr, a.shape[0], a.shape[0], _ffi.ctypes.byref(info), None)
    return d.to_host()


def _conditional_variance(y, cond, cov_vv):
    """sigma^2(y | cond) = 1 / (Sigma_SS^-1)_yy with S = cond + [y]  -> [1, 1] array."""
    cov = np.asarray(cov_vv, dtype=np.float64)
    y = int(y)
    cond = [int(c) for c in cond if int(c) != y]
    if not cond:
        return cov[y:y + 1, y:y + 1].copy()
    s = cond + [y]
    sub = np.ascontiguousarray(cov[np.ix_(s, s)])
    d = _ffi.DeviceArray.from_host(sub, DEVICE)
    info = _ffi.c_int(0)
    _ffi.call("vgp_spd_inverse", DEVICE, d.ptr, len(s), len(s), _ffi.ctypes.byref(info), None)
    last = np.empty(1)
    _ffi.call("vgp_memcpy_d2h", DEVICE, last.ctypes.data, d.ptr + (len(s) * len(s) - 1) * 8, 8, None)
    _ffi.call("vgp_stream_sync", DEVICE, None)
    return np.array([[1.0 / last[0]]])


def nominator(y, A, cov_vv):
    """sigma^2(y | A) (placement_algorithm2.py:371-388)."""
    return _conditional_variance(y, list(A), cov_vv)


def denominator(y, A_hat, cov_vv):
    """sigma^2(y | A_hat \\ y) (placement_algorithm2.py:408-413)."""
    return _conditional_variance(y, [a for a in A_hat if int(a) != int(y)], cov_vv)


# --------------------------------------------------------------------------------------------------
# arg-max helpers
# --------------------------------------------------------------------------------------------------
def argmax_cache_linear(cache, A, V):
    """First strict maximum of `cache` over V \\ A, running best starting at -1
    (placement_algorithm2.py:53-67).  Pure index scan on the host."""
    y_st, delta_st = -1, -1
    taken = set(int(a) for a in A)
    for y in V:
        if int(y) in taken:
            continue
        if delta_st < cache[y]:
            delta_st, y_st = cache[y], y
    return y_st


def _indices(s):
    """Index array of a set given as a tf.SparseTensor-like (`.values`), an array or a list."""
    v = getattr(s, "values", s)
    v = v.numpy() if hasattr(v, "numpy") else v
    return np.asarray(v, dtype=np.int64).reshape(-1)


def sparse_argmax_cache_linear(cache, A, V):
    """Index of the maximum of `cache` over V \\ A, first index on ties (placement_algorithm2.py:24-50).
    A and V may be tf.SparseTensor sets (their `.values` are used) or plain index arrays."""
    a, v = _indices(A), _indices(V)
    rest = v[~np.isin(v, a)]
    c = np.asarray(cache.numpy() if hasattr(cache, "numpy") else cache, dtype=np.float64).reshape(-1)
    vals = c[rest]
    return np.int64(rest[int(np.argmax(vals))])


def argmax_(A, A_bar, V, cov_vv):
    """(y*, delta*) of one naive step (placement_algorithm2.py:105-125)."""
    n = np.asarray(cov_vv).shape[0]
    A = [int(a) for a in A]
    if sorted(int(a) for a in A_bar) != sorted(set(range(n)) - set(A)):
        raise ValueError("argmax_ on the CUDA path needs A_bar == V \\ A")
    scores = step_scores_given(A, cov_vv)
    y_st, delta_st = -1, -1
    for y in V:
        y = int(y)
        if y in A:
            continue
        if delta_st < scores[y]:
            delta_st, y_st = scores[y], y
    return y_st, delta_st


def step_scores_given(A, cov_vv, small=_greedy.GUARD_NUMPY, jitter=0.0):
    """delta_y for all y given the selected list A, through the one-call C-ABI: the selection is replayed
    as a prefix, which is valid whenever A is itself a greedy prefix; for arbitrary A the sub-problem
    is solved per candidate with nominator/denominator."""
    cov = np.asarray(cov_vv, dtype=np.float64)
    n = cov.shape[0]
    A = [int(a) for a in A]
    if len(A) < n:
        k = len(A) + 1
        sel, _, steps, _ = _greedy.place_single(cov, k, DEVICE, small, jitter, want_step_scores=True)
        if [int(s) for s in sel[:len(A)]] == A:
            return steps[len(A)]
    out = np.full(n, np.nan)
    A_bar = [v for v in range(n) if v not in A]
    for y in A_bar:
        nom = nominator(y, A, cov)
        den = denominator(y, A_bar, cov)
        out[y] = 0.0 if (np.abs(den) < small or np.abs(nom) < small) else float((nom / den).reshape(()))
    return out


# --------------------------------------------------------------------------------------------------
# the two placement algorithms
# --------------------------------------------------------------------------------------------------
def placement_algorithm_1(cov_vv, k):
    """Naive greedy MI placement (placement_algorithm2.py:128-145): list of k np.int64, selection order."""
    sel, _, _, _ = _greedy.place_single(cov_vv, k, DEVICE, algorithm=1)
    return [np.int64(s) for s in sel]


def placement_algorithm_2(cov_vv, k):
    """Lazy greedy MI placement (placement_algorithm2.py:151-219).  Same selection as alg. 1; when PRINTS is
    set the reference's evaluation trace ('delta_y= .. y_st= ..' / 'y*= ..', :188,205) is replayed from the
    per-step scores computed on the device."""
    if not PRINTS:
        sel, _, _, _ = _greedy.place_single(cov_vv, k, DEVICE, algorithm=2)
        return [np.int64(s) for s in sel]
    sel, _, steps, _ = _greedy.place_single(cov_vv, k, DEVICE, want_step_scores=True, algorithm=2)
    for line in lazy_trace_lines(steps, sel):
        print(line)
    return [np.int64(s) for s in sel]


def lazy_trace_lines(step_scores, selection):
    """Lines alg. 2 prints, from dense per-step scores [k, n] (cache starts at +inf, all stale each round:
    placement_algorithm2.py:157-208)."""
    k, n = step_scores.shape
    cache = np.full(n, np.inf)
    taken = np.zeros(n, dtype=bool)
    lines = []
    for t in range(k):
        fresh = np.zeros(n, dtype=bool)
        while True:
            masked = np.where(taken, -np.inf, cache)
            y = int(np.argmax(masked))
            if fresh[y]:
                lines.append("y*= %d" % y)
                break
            d = step_scores[t, y]
            cache[y] = d
            fresh[y] = True
            shown = "0" if d == 0 else str(np.array([[d]]))
            lines.append("delta_y= %s y_st= %d" % (shown, y))
        if y != int(selection[t]):
            raise RuntimeError("lazy replay picked %d, device picked %d (near-tie below 1e-12?)" % (y, selection[t]))
        taken[y] = True
    return lines


# --------------------------------------------------------------------------------------------------
# timing harness (placement_algorithm2.py:416-429)
# --------------------------------------------------------------------------------------------------
def calculate_running_time_algorithm_1(cov_vv, k):
    strt = time.time()
    pl = placement_algorithm_1(cov_vv, k)
    print("placement_array : ", pl)
    return time.time() - strt


def calculate_running_time_algorithm_2(cov_vv, k):
    strt = time.time()
    pl = placement_algorithm_2(cov_vv, k)
    print("placement_array : ", pl)
    return time.time() - strt


# --------------------------------------------------------------------------------------------------
# TF-graph named variants (placement_algorithm2.py:336-368): same arithmetic, array-like or DLPack inputs
# --------------------------------------------------------------------------------------------------
def tf_nominator(y, Ain, cov_vv):
    """sigma^2(y | Ain U y \\ y); y and Ain may be tf.SparseTensor sets, index arrays or ints."""
    yy = int(_indices(y)[0])
    return nominator(yy, [a for a in _indices(Ain) if a != yy], _host_matrix(cov_vv))


def tf_denominator(y, A_hat, cov_vv):
    yy = int(_indices(y)[0])
    return denominator(yy, list(_indices(A_hat)), _host_matrix(cov_vv))


def _host_matrix(m):
    if hasattr(m, "numpy"):
        return m.numpy()
    return np.asarray(m)
