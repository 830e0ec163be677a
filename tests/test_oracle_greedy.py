"""CPU: the oracle (oracle/greedy_oracle.py) against the golden vectors produced by the reference's own
NumPy code (tests/golden/make_golden.py), plus the host-side logic that needs no GPU."""
import numpy as np
import pytest

from oracle import greedy_oracle as go
from oracle import ref_extract


def test_incremental_oracle_matches_reference_selection(golden, golden_name):
    cov = golden.cov(golden_name)
    k = golden.cases[golden_name]["k"]
    sel, scores, steps, gaps = go.incremental_greedy(cov, k, return_all_scores=True)
    assert sel == golden.selection(golden_name)               # bit-exact index parity
    assert gaps.min() > 1e-12                                 # no near-tie in the fixtures
    ref = golden.step_scores(golden_name)
    if ref is not None:
        # score agreement degrades with cond(Sigma) (the reference goes through SVD pinv)
        tol = 1e-10 if golden_name.startswith(("fixture", "expquad")) else 1e-7
        np.testing.assert_allclose(steps, ref, rtol=tol, equal_nan=True)


@pytest.mark.parametrize("shards", [2, 3, 5])
def test_sharding_does_not_change_results(golden, shards):
    for name in ("fixture4x4", "legacy_random_n40", "expquad_n100"):
        cov, k = golden.cov(name), golden.cases[name]["k"]
        if shards > cov.shape[0]:
            continue
        one = go.incremental_greedy(cov, k, return_all_scores=True)
        many = go.incremental_greedy(cov, k, shards=shards, return_all_scores=True)
        assert one[0] == many[0]
        np.testing.assert_array_equal(one[1], many[1])        # element-wise updates: bitwise equal
        np.testing.assert_array_equal(one[2], many[2])


def test_literal_oracle_matches_reference_small(golden):
    for name in golden.names(max_n=50):
        cov, k = golden.cov(name), golden.cases[name]["k"]
        assert go.literal_placement_algorithm_1(cov, k) == golden.selection(name)
        sel2, evaluations = go.literal_placement_algorithm_2(cov, k)
        assert sel2 == golden.cases[name]["alg2_selection"]


def test_first_step_scores_of_fixture(golden):
    # step-1 scores of the 4x4 fixture
    _, _, steps, _ = go.incremental_greedy(golden.cov("fixture4x4"), 4, return_all_scores=True)
    np.testing.assert_allclose(steps[0], [1.1865662348005577, 1.179414784756388, 1.2342192027371512,
                                          1.1518962290768235], rtol=1e-13)
    np.testing.assert_allclose(steps[3][0], 0.8427679556953547, rtol=1e-13)


def test_lazy_replay_reproduces_alg2_print_trace(golden):
    from vgposp_b200.placement_algorithm2 import lazy_trace_lines
    for name in golden.names(max_n=200):
        c = golden.cases[name]
        if "alg2_stdout" not in c:
            continue
        sel, _, steps, _ = go.incremental_greedy(golden.cov(name), c["k"], return_all_scores=True)
        assert_same_trace(lazy_trace_lines(steps, sel), c["alg2_stdout"],
                          exact=name.startswith(("fixture", "expquad")))
        picks, trace = go.replay_lazy_evaluations(steps, sel)
        assert picks == sel


def parse_trace(lines):
    out = []
    for ln in lines:
        if ln.startswith("y*="):
            out.append(("pick", int(ln.split("=")[1]), None))
        else:
            val, y = ln[len("delta_y="):].split("y_st=")
            out.append(("eval", int(y), float(val.strip().strip("[]"))))
    return out


def assert_same_trace(got, want, exact):
    """Evaluation order and picks must be identical; the printed 8-decimal scores are compared as text on
    well-conditioned inputs and to 1e-7 relative on the U U^T inputs (cond up to 1e6, where the reference's
    own SVD pinv carries ~1e-9 relative noise)."""
    if exact:
        assert got == want
        return
    g, w = parse_trace(got), parse_trace(want)
    assert [(a, b) for a, b, _ in g] == [(a, b) for a, b, _ in w]
    np.testing.assert_allclose([v for _, _, v in g if v is not None], [v for _, _, v in w if v is not None],
                               rtol=1e-7)


def test_tf_graph_compat_mode_matches_literal():
    # jitter 1e-6 on diag(Sigma_AA) and guard 1e-7 (snippets_a2.py:161-163,480) against the literal form
    rng = np.random.default_rng(3)
    x = rng.uniform(-2, 2, (30, 3))
    d = x[:, None] - x[None]
    cov = np.exp(-np.sum(d * d, -1) / (2 * 0.7 ** 2)) + 1e-2 * np.eye(30)
    lit, per_step = go.literal_placement_algorithm_1(cov, 4, small=go.GUARD_TF_GRAPH, jitter=go.JITTER_TF_GRAPH,
                                                     return_scores=True)
    inc = go.incremental_greedy(cov, 4, small=go.GUARD_TF_GRAPH, jitter=go.JITTER_TF_GRAPH, return_all_scores=True)
    assert inc[0] == lit
    np.testing.assert_allclose(inc[2], np.array(per_step), rtol=1e-9, equal_nan=True)


def test_guard_zeroes_scores():
    num = np.array([1.0, 1e-9, 2.0, -3e-9])
    den = np.array([1e-9, 1.0, 4.0, 1.0])
    np.testing.assert_array_equal(go.guarded_scores(num, den, 1e-8), [0.0, 0.0, 0.5, 0.0])


def test_first_argmax_tie_rule_and_exhaustion():
    s = np.array([0.5, 0.7, 0.7, 0.1])
    assert go.first_argmax(s, np.zeros(4, bool)) == 1                     # lowest index wins an exact tie
    assert go.first_argmax(s, np.array([0, 1, 0, 0], bool)) == 2
    assert go.first_argmax(np.full(3, -2.0), np.zeros(3, bool)) == -1     # nothing beats the initial -1
    assert go.pick_winner([(0.7, 5, 1.0), (0.7, 2, 1.0), (-np.inf, -1, 0.0)])[1] == 2


def test_not_positive_definite_is_an_error():
    cov = np.ones((4, 4))
    with pytest.raises(np.linalg.LinAlgError):
        go.incremental_greedy(cov, 2)


@pytest.mark.skipif(not ref_extract.available(), reason="reference tree only exists in the build container")
def test_golden_still_matches_reference_code(golden):
    ref = ref_extract.load()
    for name in ("fixture4x4", "legacy_random_n10", "expquad_n50"):
        cov, k = golden.cov(name), golden.cases[name]["k"]
        sel, _ = ref_extract.run_quiet(ref["placement_algorithm_1"], cov, k)
        assert [int(s) for s in sel] == golden.selection(name)
        sel2, out = ref_extract.run_quiet(ref["placement_algorithm_2"], cov, k)
        assert out.splitlines() == golden.cases[name]["alg2_stdout"]


def _grid_cov(cover, ls, seed):
    idx = np.indices(cover).reshape(3, -1).T.astype(np.float64)
    pts = idx + np.random.default_rng(seed).uniform(-0.3, 0.3, idx.shape)
    d = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
    return np.exp(-d / (2 * ls * ls)) + 1e-2 * np.eye(len(pts))


def test_tf_graph_algorithm_2_restatement_is_consistent():
    """snippets_a2.sparse_placement_algorithm_2 restated: same winners as the pinned literal lazy greedy run with the
    graph's constants; delta_cached_iters column t holds the winner's delta at its row and zeros at earlier winners."""
    cov = _grid_cov((3, 3, 2), 1.0, 3)
    k = 5
    A, len_A, dci, sel = go.literal_sparse_placement_algorithm_2(cov, k)
    want, _ = go.literal_placement_algorithm_2(cov, k, small=go.GUARD_TF_GRAPH, jitter=go.JITTER_TF_GRAPH)
    assert [int(v) for v in sel[:, 0]] == want and A == sorted(want) and len_A == k
    for t in range(k):
        assert dci[want[t], t] == sel[t, 1]
        assert np.all(dci[want[:t], t] == 0.0)
        assert dci[:, t].max() <= 1e8


def test_algorithm_3_restatement_limits():
    """snippets_a3 restated: a box covering the grid re-evaluates everything (== exact greedy with the graph's
    constants); with a small box the first two winners still agree (the first cache is complete, the second arg-max
    only sees refreshed neighbours and still-valid upper bounds), and column 0 of delta_cached_iters is the A = {}
    delta of every point."""
    cover = (4, 3, 2)
    cov = _grid_cov(cover, 1.1, 5)
    k = 5
    exact = go.literal_placement_algorithm_1(cov, k, small=go.GUARD_TF_GRAPH, jitter=go.JITTER_TF_GRAPH)
    A_full, cache, dci = go.literal_sparse_placement_algorithm_3(cov, k, cover, max(cover))
    assert A_full == exact
    A_loc, _, dci_loc = go.literal_sparse_placement_algorithm_3(cov, k, cover, 1)
    assert A_loc[0] == exact[0] and len(set(A_loc)) == k
    n = cov.shape[0]
    first = [go.literal_delta(y, [], list(range(n)), cov, go.GUARD_TF_GRAPH, go.JITTER_TF_GRAPH) for y in range(n)]
    np.testing.assert_allclose(dci_loc[:, 0], first, rtol=1e-14)
    assert np.all(dci_loc[A_loc[:-1], -1] == 0.0)


def test_lazy_cache_replay_host_logic():
    """vgposp_b200.snippets_a2.lazy_cache_replay (host index scan over device step scores) against the literal graph
    restatement, fed with the literal per-step deltas."""
    from vgposp_b200 import snippets_a2
    cov = _grid_cov((3, 2, 2), 0.9, 8)
    k = 4
    A, _, want_dci, want_sel = go.literal_sparse_placement_algorithm_2(cov, k)
    order, per_step = go.literal_placement_algorithm_1(cov, k, small=go.GUARD_TF_GRAPH, jitter=go.JITTER_TF_GRAPH,
                                                       return_scores=True)
    got_order, dci, sel = snippets_a2.lazy_cache_replay(np.array(per_step), k)
    assert got_order == order == [int(v) for v in want_sel[:, 0]]
    np.testing.assert_allclose(dci, want_dci, rtol=1e-14)
    np.testing.assert_allclose(sel, want_sel, rtol=1e-14)
