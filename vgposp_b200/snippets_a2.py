"""Drop-in for the TF-graph lazy greedy of the reference's `snippets_a2.py`.

    A, len_A, delta_cached_iters, A_selection_and_delta = sparse_placement_algorithm_2(cov_vv, k, COVER_spatial)

The reference builds this out of nested `tf.while_loop`s over SparseTensor sets (snippets_a2.py:679-822) with the
graph's numerics: +1e-6 on diag(Sigma_AA) (:161-163), guard 1e-7 (:480), cache initialised to INF = 1e8 (:690), the
winner's cache entry zeroed after it is stored (:796).  Here the per-step deltas of every candidate come from the
device in one call (`vgp_placement_host_ex` with those two constants); the lazy cache bookkeeping -- which entries the
graph would have re-evaluated and which stay stale -- is an index scan over those rows on the host.
"""
import numpy as np

from . import greedy as _greedy

DEVICE = 0
INF = 1e8                                   # snippets_a2.py:690
SMALL = _greedy.GUARD_TF_GRAPH              # :480
JITTER = _greedy.JITTER_TF_GRAPH            # :161-163


def lazy_cache_replay(step_scores, k, inf=INF):
    """The cache the graph carries, replayed from dense fresh deltas `step_scores` [k, n] (row t = deltas given the
    first t winners): returns (selection order, delta_cached_iters [n, k], A_selection_and_delta [k, 2])."""
    n = step_scores.shape[1]
    cache = np.full(n, inf)
    taken = np.zeros(n, dtype=bool)
    dci = np.zeros((n, k))
    sel = np.zeros((k, 2))
    order = []
    for t in range(k):
        fresh = np.zeros(n, dtype=bool)
        while True:
            y = int(np.argmax(np.where(taken, -np.inf, cache)))      # sparse_argmax_cache_linear: first maximum
            if fresh[y]:
                break
            cache[y] = step_scores[t, y]
            fresh[y] = True
        order.append(y)
        taken[y] = True
        sel[t] = (y, cache[y])
        dci[:, t] = cache
        cache[y] = 0.0
    return order, dci, sel


def sparse_placement_algorithm_2(cov_vv, k, COVER_spatial=None):
    """snippets_a2.py:679-822.  Returns (A, len(A), delta_cached_iters [N, k], A_selection_and_delta [k, 2]); `A` is the
    sorted index array (the graph returns an unordered sparse set -- the order of selection is in
    A_selection_and_delta[:, 0], :822)."""
    cov = np.asarray(cov_vv.numpy() if hasattr(cov_vv, "numpy") else cov_vv, dtype=np.float64)
    n = cov.shape[0]
    if COVER_spatial is not None:
        assert n == int(np.prod(COVER_spatial)), "N must equal COVER_spatial[0] * [1] * [2]"       # :692
    sel, _, steps, _ = _greedy.place_single(cov, k, DEVICE, SMALL, JITTER, want_step_scores=True)
    order, dci, sel_delta = lazy_cache_replay(steps, int(k))
    assert order == [int(v) for v in sel], "lazy replay and device arg-max disagree"
    return np.sort(np.asarray(order, dtype=np.int64)), len(order), dci, sel_delta
