"""GPU parity on RANK-DEFICIENT covariances (the reference's pinv semantics, csrc/pinv.cu) through the drop-in calls:
golden vectors from the reference's unmodified NumPy code (tests/golden/make_golden_lowrank.py), the closed-form CPU
oracle at the reference's own sizes (625 locations, 144 samples: main_architecture_2.py:391-444), and the producer ->
placement chain without any nugget."""
import contextlib
import io
import json
import os

import numpy as np
import pytest

from oracle import greedy_oracle as go
from vgposp_b200 import greedy
import vgposp_b200.placement_algorithm2 as alg2

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
D = 0

with open(os.path.join(HERE, "golden", "greedy_lowrank_golden.json")) as _fh:
    CASES = json.load(_fh)["cases"]
INPUTS = np.load(os.path.join(HERE, "golden", "greedy_lowrank_inputs.npz"))


def golden_steps(case):
    return np.array([[np.nan if v is None else v for v in row] for row in case["step_scores"]])


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_goldens_through_the_dropin(name, quiet_alg2):
    case, cov = CASES[name], INPUTS[name]
    assert alg2.placement_algorithm_1(cov, case["k"]) == case["alg1_selection"]
    assert alg2.placement_algorithm_2(cov, case["k"]) == case["alg2_selection"]
    sel, scores, steps, _ = greedy.place_single(cov, case["k"], D, want_step_scores=True)      # falls through to pinv
    assert greedy.place_single_pinv.last_rank == case["rank"]
    assert [int(v) for v in sel] == case["alg1_selection"]
    ref = golden_steps(case)
    # the reference's scores carry the noise of one SVD per candidate; exact zeros (guarded) must be exact zeros
    np.testing.assert_allclose(steps, ref, rtol=1e-6, atol=1e-12, equal_nan=True)
    assert np.array_equal(steps == 0.0, ref == 0.0)
    np.testing.assert_allclose(scores, [row[int(y)] for row, y in zip(ref, sel)], rtol=1e-6, atol=1e-12)


def test_alg2_print_trace_on_a_rank_deficient_input():
    """placement_algorithm_2 prints per evaluation (placement_algorithm2.py:188,205); compare with the oracle's cache
    walk on the device's step scores."""
    case, cov = CASES["midrank_n30_s22"], INPUTS["midrank_n30_s22"]
    saved, alg2.PRINTS = alg2.PRINTS, True
    try:
        with contextlib.redirect_stdout(io.StringIO()) as out:
            sel = alg2.placement_algorithm_2(cov, case["k"])
    finally:
        alg2.PRINTS = saved
    assert sel == case["alg2_selection"]
    lines = out.getvalue().splitlines()
    assert [ln for ln in lines if ln.startswith("y*= ")] == ["y*= %d" % y for y in case["alg2_selection"]]


@pytest.mark.parametrize("n,s,k", [(625, 144, 20), (625, 250, 12), (2000, 300, 10)])
def test_reference_regime_matches_the_oracle(n, s, k):
    """The sizes of the reference's own runs: rank s - 1 << n.  Every delta is 0, the selection is [0 .. k-1]."""
    m = np.random.default_rng(n + s).standard_normal((n, s))
    cov = np.cov(m, bias=True)
    sel, scores, steps, _ = greedy.place_single(cov, k, D, want_step_scores=True)
    assert greedy.place_single_pinv.last_rank == s - 1
    assert [int(v) for v in sel] == list(range(k))
    assert np.all(scores == 0.0) and np.nanmax(np.abs(steps)) == 0.0
    if n <= 625:
        want = go.pinv_greedy(cov, 3)
        assert want[0] == [0, 1, 2]
        np.testing.assert_array_equal(np.nan_to_num(want[2], nan=-1.0), np.nan_to_num(steps[:3], nan=-1.0))


@pytest.mark.parametrize("n,s,k,alg,seed", [(120, 90, 100, 1, 930), (120, 90, 100, 2, 930), (200, 150, 70, 1, 350),
                                             (200, 150, 70, 2, 2350)])
def test_mid_rank_matches_the_oracle(n, s, k, alg, seed):
    """rank > n / 2: between selection n - rank and selection rank both conditional variances are non-zero and the
    scores are real numbers; the factor is rebuilt on the remaining candidates once they stop spanning it.  Inputs are
    checked to be numerically well-posed (no eigenvalue of any Sigma_AbarAbar in the window where pinv's own rank
    decision is rounding noise): about half of all random draws are not, at the one selection where card Abar = rank."""
    m = np.random.default_rng(seed).standard_normal((n, s))
    cov = np.cov(m, bias=True)
    want_sel, want_scores, want_steps = go.pinv_greedy(cov, k, algorithm=alg)
    assert go.pinv_well_posed(cov, want_sel)[0], "ill-posed test input"
    sel, scores, steps, _ = greedy.place_single_pinv(cov, k, D, want_step_scores=True, algorithm=alg)
    live = want_steps[np.isfinite(want_steps) & (want_steps != 0)]
    assert live.size > 100, "test input does not reach the non-degenerate steps"
    assert [int(v) for v in sel] == want_sel
    np.testing.assert_allclose(steps, want_steps, rtol=1e-7, atol=1e-12, equal_nan=True)
    np.testing.assert_allclose(scores, want_scores, rtol=1e-7, atol=1e-12)


def test_producer_feeds_the_placement_without_a_nugget(quiet_alg2):
    """f-1 -> a8/a9: create_cov_matrix over more locations than samples (rank-deficient by construction) straight into
    placement_algorithm_2 -- the chain of main_architecture_2.py:391-444 -> :722."""
    from vgposp_b200 import cov_producer
    rng = np.random.default_rng(5)
    fields = rng.standard_normal((4, 5, 6, 30))                 # 120 locations, 30 samples each
    cov = cov_producer.empirical_cov(fields.reshape(120, 30))
    np.testing.assert_allclose(cov, np.cov(fields.reshape(120, 30), bias=True), rtol=1e-10, atol=1e-14)
    sel = alg2.placement_algorithm_2(cov, 7)
    assert sel == go.pinv_greedy(cov, 7, algorithm=2)[0] == list(range(7))


def test_ill_posed_transition_does_not_fail():
    """An input whose Sigma_AbarAbar is numerically singular exactly where card Abar = rank (seed found with
    go.pinv_well_posed): scores there are rounding noise in the reference too, but the call must complete."""
    m = np.random.default_rng(7 * 300 + 260).standard_normal((300, 260))
    cov = np.cov(m, bias=True)
    sel, scores, _, _ = greedy.place_single_pinv(cov, 60, D)
    assert [int(v) for v in sel[:40]] == list(range(40)) and len(set(int(v) for v in sel)) == 60
    assert np.all(np.isfinite(scores))


def test_indefinite_and_wellconditioned_inputs_take_the_right_path():
    with pytest.raises(np.linalg.LinAlgError):
        greedy.place_single(np.array([[1.0, 2.0, 0.0], [2.0, 1.0, 0.0], [0.0, 0.0, 1.0]]), 2, D)
    x = np.random.default_rng(0).uniform(-2, 2, (300, 3))
    d = x[:, None, :] - x[None, :, :]
    cov = np.exp(-np.einsum("ijk,ijk->ij", d, d) / (2 * 0.5 ** 2)) + 1e-2 * np.eye(300)
    greedy.place_single_pinv.last_rank = None
    sel, *_ = greedy.place_single(cov, 5, D)
    assert greedy.place_single_pinv.last_rank is None            # SPD: the Cholesky path served it
    # the pseudo-inverse path on a full-rank matrix is the same mathematics: same selection, same scores
    sel_p, sc_p, _, _ = greedy.place_single_pinv(cov, 5, D)
    assert [int(v) for v in sel_p] == [int(v) for v in sel] and greedy.place_single_pinv.last_rank == 300
    np.testing.assert_allclose(sc_p, go.incremental_greedy_c(cov, 5)[1], rtol=1e-8)
