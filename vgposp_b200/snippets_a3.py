"""Drop-in for the local-kernel greedy of the reference's `snippets_a3.py`.

    A, delta_cached, delta_cached_iters = sparse_placement_algorithm_3(cov_vv, k, COVER_spatial, cutoff)

Algorithm 3 (snippets_a3.py:43-364) evaluates every delta once and afterwards re-evaluates only the candidates inside
an index box of half-width `cutoff` around each new winner; with a tapered ("local kernel") covariance the stale rest
is the approximation the reference accepts.  On the device this is the lazy-column formulation with one extra array:
the step kernel refreshes the cache entries inside the box and takes the arg-max over the cache (csrc/lazy.cu,
`vgp_lazy_set_local`).  Numerics of the TF graph: jitter 1e-6, guard 1e-7.
"""
import numpy as np

from . import greedy as _greedy

DEVICE = 0


def sparse_placement_algorithm_3(cov_vv, k, COVER_spatial, cutoff, small=_greedy.GUARD_TF_GRAPH,
                                 jitter=_greedy.JITTER_TF_GRAPH):
    """Returns (A in selection order [k] int64, final cache [N], delta_cached_iters [N, k]).  (The graph returns A as
    an unordered sparse set; sort it for that view.)"""
    cov = np.asarray(cov_vv.numpy() if hasattr(cov_vv, "numpy") else cov_vv, dtype=np.float64)
    n = cov.shape[0]
    assert n == int(np.prod(COVER_spatial)), "N must equal COVER_spatial[0] * [1] * [2]"           # :51
    k, cutoff = int(k), int(cutoff)
    assert cutoff >= 1, "cutoff is an index distance >= 1"
    h = _greedy.LazyGreedy(n, k, DEVICE, small=small, jitter=jitter, mode=0)
    try:
        h.load_cov_host(cov)
        h.factor()
        h.set_local(COVER_spatial, cutoff)
        h.record_scores(True)
        h.run(k)
        sel, _ = h.results()
        dci = h.step_scores().T.copy()
    finally:
        h.close()
    _greedy.check_selection(sel)
    cache = dci[:, k - 1].copy()
    return np.asarray(sel, dtype=np.int64), cache, dci
