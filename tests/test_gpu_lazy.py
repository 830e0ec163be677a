"""GPU parity: the lazy-column formulation of the greedy placement (csrc/lazy.cu) against the reference's golden
vectors, the CPU oracle and the dense precision-downdate formulation (placement_algorithm2.py:105-145, :371-413)."""
import numpy as np
import pytest

from oracle import greedy_oracle as go
from vgposp_b200 import _ffi, greedy

pytestmark = pytest.mark.gpu
D = 0
LAZY = ["lazy_precision", "lazy_factor"]


def cloud_cov(n, seed, nugget=1e-2):
    x = np.random.default_rng(seed).uniform(-2, 2, (n, 3))
    ls = 0.5 * (1000.0 / max(n, 1000)) ** (1 / 3)
    d = x[:, None, :] - x[None, :, :]
    return np.exp(-np.einsum("ijk,ijk->ij", d, d) / (2 * ls * ls)) + nugget * np.eye(n)


@pytest.mark.parametrize("formulation", LAZY)
def test_golden_selections_bit_exact(golden, golden_name, formulation):
    cov, k = golden.cov(golden_name), golden.cases[golden_name]["k"]
    sel, scores, steps, _ = greedy.place_single(cov, k, D, want_step_scores=True, formulation=formulation)
    assert [int(s) for s in sel] == golden.selection(golden_name)
    ref = golden.step_scores(golden_name)
    if ref is not None:
        tol = 1e-10 if golden_name.startswith(("fixture", "expquad")) else 1e-7
        np.testing.assert_allclose(steps, ref, rtol=tol, equal_nan=True)
        np.testing.assert_allclose(scores, [np.nanmax(r) for r in ref], rtol=tol)


@pytest.mark.parametrize("n,k,seed", [(1, 1, 0), (5, 5, 1), (130, 130, 2), (700, 20, 3), (1500, 30, 4)])
def test_precision_mode_is_bitwise_the_dense_downdate(n, k, seed):
    """Mode 0 replays the rank-1 history with the same fused multiply-adds in the same order as downdate_kernel."""
    cov = cloud_cov(n, seed)
    dense = greedy.place_single(cov, k, D, want_step_scores=True, formulation="dense")
    lazy = greedy.place_single(cov, k, D, want_step_scores=True, formulation="lazy_precision")
    np.testing.assert_array_equal(lazy[0], dense[0])
    np.testing.assert_array_equal(lazy[1], dense[1])
    np.testing.assert_array_equal(lazy[2], dense[2])             # every score of every step, NaNs included


@pytest.mark.parametrize("formulation", LAZY)
@pytest.mark.parametrize("n,k,seed", [(257, 12, 1), (1000, 10, 2), (2000, 25, 3), (3000, 6, 4), (1111, 40, 5)])
def test_matches_cpu_oracle_on_seeded_clouds(n, k, seed, formulation):
    cov = cloud_cov(n, seed)
    want_sel, want_scores, want_steps, gaps = go.incremental_greedy(cov, k, return_all_scores=True)
    assert gaps.min() > 1e-12, "near-tie in the test input"
    sel, scores, steps, _ = greedy.place_single(cov, k, D, want_step_scores=True, formulation=formulation)
    assert [int(s) for s in sel] == want_sel
    np.testing.assert_allclose(scores, want_scores, rtol=1e-9)
    np.testing.assert_allclose(steps, want_steps, rtol=1e-9, equal_nan=True)


@pytest.mark.parametrize("formulation", LAZY)
def test_history_longer_than_one_shared_memory_chunk(formulation):
    """k > 512: the replay of the rank-1 history is staged through shared memory in chunks."""
    n, k = 900, 600
    cov = cloud_cov(n, 21)
    dense = greedy.place_single(cov, k, D, formulation="dense")
    lazy = greedy.place_single(cov, k, D, formulation=formulation)
    np.testing.assert_array_equal(lazy[0], dense[0])
    np.testing.assert_allclose(lazy[1], dense[1], rtol=1e-9)


@pytest.mark.parametrize("formulation", LAZY)
def test_tf_graph_compat_mode(formulation):
    cov = cloud_cov(300, 9)
    want = go.incremental_greedy(cov, 6, small=go.GUARD_TF_GRAPH, jitter=go.JITTER_TF_GRAPH, return_all_scores=True)
    sel, scores, steps, _ = greedy.place_single(cov, 6, D, small=greedy.GUARD_TF_GRAPH, jitter=greedy.JITTER_TF_GRAPH,
                                                want_step_scores=True, formulation=formulation)
    assert [int(s) for s in sel] == want[0]
    np.testing.assert_allclose(steps, want[2], rtol=1e-9, equal_nan=True)


@pytest.mark.parametrize("formulation", LAZY)
def test_errors_like_the_dense_path(formulation):
    with pytest.raises(np.linalg.LinAlgError):
        greedy.place_single(np.ones((6, 6)), 2, D, formulation=formulation, pinv_fallback=False)
    sel, *_ = greedy.place_single(np.ones((6, 6)), 2, D, formulation=formulation)      # pseudo-inverse path takes over
    assert [int(v) for v in sel] == [0, 1]
    with pytest.raises(ValueError, match="not in list"):
        greedy.place_single(cloud_cov(4, 0), 5, D, formulation=formulation)


@pytest.mark.parametrize("mode", [0, 1])
def test_handle_continues_and_replays(mode):
    """run(a) + run(b) == run(a + b); reset replays identically; launches: 1 (mode 0) or 2 (mode 1) per step."""
    n, k = 1300, 14
    cov = cloud_cov(n, 8)
    h = greedy.LazyGreedy(n, k, D, mode=mode)
    h.load_cov_host(cov)
    h.factor()
    base = h.launch_count()
    h.run(5)
    h.run(k - 5)
    first = h.results()
    per_step = (h.launch_count() - base) / k
    assert per_step <= (1 if mode == 0 else 2)
    h.reset()
    h.run(k)
    second = h.results()
    np.testing.assert_array_equal(first[0], second[0])
    np.testing.assert_array_equal(first[1], second[1])
    want_sel, want_scores = go.incremental_greedy_c(cov, k)
    assert [int(s) for s in first[0]] == want_sel
    np.testing.assert_allclose(first[1], want_scores, rtol=1e-9)
    h.close()


def test_properties_at_size():
    """n = 6000 (no CPU oracle): both lazy modes against the dense formulation on a device-built covariance."""
    n, k = 6000, 24
    x = np.random.default_rng(42).uniform(-2, 2, (n, 3))
    ls = 0.5 * (1000.0 / n) ** (1 / 3)
    xd = _ffi.DeviceArray.from_host(x, D)
    shard = greedy.GreedyShard(n, 0, n, k, D)
    shard.build_cov_expquad(xd.ptr, 3, 1.0, ls, 1e-2)
    shard.factor()
    shard.run(k)
    dsel, dscores = shard.results()
    shard.close()
    for mode in (0, 1):
        h = greedy.LazyGreedy(n, k, D, mode=mode)
        h.build_cov_expquad(xd.ptr, 3, 1.0, ls, 1e-2)
        h.factor()
        h.run(k)
        sel, scores = h.results()
        h.close()
        np.testing.assert_array_equal(sel, dsel)
        if mode == 0:
            np.testing.assert_array_equal(scores, dscores)
        else:
            np.testing.assert_allclose(scores, dscores, rtol=1e-10)
        assert np.all(np.diff(scores) <= 1e-12 * scores[:-1])


def _jittered_grid_cov(cover, ls, seed):
    idx = np.indices(cover).reshape(3, -1).T.astype(np.float64)
    pts = idx + np.random.default_rng(seed).uniform(-0.3, 0.3, idx.shape)       # breaks the lattice's exact ties
    d = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
    return np.exp(-d / (2 * ls * ls)) + 1e-2 * np.eye(len(pts))


@pytest.mark.parametrize("cover,cutoff,k", [((4, 4, 2), 2, 6), ((5, 3, 3), 1, 5), ((6, 6, 1), 3, 7)])
def test_algorithm_3_local_kernel_greedy(cover, cutoff, k):
    """snippets_a3.sparse_placement_algorithm_3 against its literal CPU restatement: selection order, the final cache
    and every column of delta_cached_iters (stale entries included)."""
    from vgposp_b200 import snippets_a3
    cov = _jittered_grid_cov(cover, 1.2, sum(cover))
    want_A, want_cache, want_dci = go.literal_sparse_placement_algorithm_3(cov, k, cover, cutoff)
    A, cache, dci = snippets_a3.sparse_placement_algorithm_3(cov, k, cover, cutoff)
    assert [int(v) for v in A] == want_A
    np.testing.assert_allclose(dci, want_dci, rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(cache, want_cache, rtol=1e-7, atol=1e-12)
    # a box that covers the whole grid makes every entry fresh: algorithm 3 == exact greedy (graph numerics)
    A_full, _, _ = snippets_a3.sparse_placement_algorithm_3(cov, k, cover, max(cover))
    assert [int(v) for v in A_full] == go.literal_placement_algorithm_1(cov, k, small=1e-7, jitter=1e-6)


def test_tf_graph_algorithm_2_outputs():
    """snippets_a2.sparse_placement_algorithm_2: the four outputs of the graph (:822) against the literal restatement."""
    from vgposp_b200 import snippets_a2
    cover = (4, 3, 3)
    cov = _jittered_grid_cov(cover, 1.0, 7)
    k = 6
    want_A, want_len, want_dci, want_sel = go.literal_sparse_placement_algorithm_2(cov, k)
    A, len_A, dci, sel = snippets_a2.sparse_placement_algorithm_2(cov, k, cover)
    assert [int(v) for v in A] == want_A and len_A == want_len == k
    assert [int(v) for v in sel[:, 0]] == [int(v) for v in want_sel[:, 0]]
    np.testing.assert_allclose(sel[:, 1], want_sel[:, 1], rtol=1e-7)
    np.testing.assert_allclose(dci, want_dci, rtol=1e-7, atol=1e-12)


@pytest.mark.parametrize("formulation", ["lazy_factor", "lazy_precision"])
def test_pinned_and_pageable_sources_give_the_same_bits(formulation, vgp_options):
    """The one-call path copies the lower triangle in row chunks underneath the factorisation: from pinned memory this
    thread issues every copy up front; from pageable memory (a NumPy array) a helper thread issues them while the
    factorisation is enqueued and the row gate waits on the host for each chunk's event.  Same results, also with the
    overlap switched off, and with several chunks (n > 2048)."""
    import ctypes
    from vgposp_b200 import _ffi
    n, k = 5000, 9
    cov = cloud_cov(n, 12)
    want_sel, want_scores = go.incremental_greedy_c(cov, k)
    host = ctypes.c_void_p()
    _ffi.call("vgp_host_alloc", cov.nbytes, ctypes.byref(host))
    try:
        pinned = np.ctypeslib.as_array(ctypes.cast(host, ctypes.POINTER(ctypes.c_double)), shape=(n, n))
        pinned[:] = cov
        results = []
        for overlap in (1, 0):
            vgp_options(h2d_overlap=overlap)
            for src in (cov, pinned):
                sel, scores, steps, secs = greedy.place_single(src, k, D, want_step_scores=True, formulation=formulation)
                results.append((sel, scores, steps))
        for sel, scores, steps in results:
            assert [int(v) for v in sel] == want_sel
            np.testing.assert_array_equal(scores, results[0][1])
            np.testing.assert_array_equal(steps, results[0][2])
        np.testing.assert_allclose(results[0][1], want_scores, rtol=1e-9)
    finally:
        _ffi.call("vgp_host_free", host)
