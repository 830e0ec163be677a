"""CPU: the closed-form restatement of the reference's pinv semantics on rank-deficient covariances
(oracle/greedy_oracle.pinv_greedy) against golden vectors produced by the reference's own, unmodified NumPy code
(tests/golden/make_golden_lowrank.py) -- and, when the reference tree is present, against the reference run live."""
import json
import os

import numpy as np
import pytest

from oracle import greedy_oracle as go
from oracle import ref_extract

HERE = os.path.dirname(os.path.abspath(__file__))


def lowrank_golden():
    with open(os.path.join(HERE, "golden", "greedy_lowrank_golden.json")) as fh:
        cases = json.load(fh)["cases"]
    inputs = np.load(os.path.join(HERE, "golden", "greedy_lowrank_inputs.npz"))
    return cases, inputs


CASES, INPUTS = lowrank_golden()


def golden_steps(case):
    return np.array([[np.nan if v is None else v for v in row] for row in case["step_scores"]])


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_reference_on_rank_deficient_inputs(name):
    case, cov = CASES[name], INPUTS[name]
    assert np.linalg.matrix_rank(cov) == case["rank"] == case["s"] - 1 < case["n"]
    for alg in (1, 2):
        sel, win, steps = go.pinv_greedy(cov, case["k"], algorithm=alg)
        assert sel == case["alg%d_selection" % alg], alg
        if alg == 1:
            # the reference's own scores carry the noise of one SVD per candidate (cond up to 1e6 on these inputs)
            np.testing.assert_allclose(steps, golden_steps(case), rtol=1e-6, atol=1e-12, equal_nan=True)


def test_low_rank_regime_degenerates_to_the_first_indices():
    """rank < n / 2 (the reference's own runs: 144 or 250 samples over >= 625 locations): every delta is 0 and the
    first strict maximum above -1 is the lowest free index."""
    for name in ("lowrank_n30_s8", "lowrank_n60_s20", "lowrank_n100_s40"):
        case = CASES[name]
        assert case["alg1_selection"] == case["alg2_selection"] == list(range(case["k"]))
        assert np.nanmax(np.abs(golden_steps(case))) == 0.0


def test_alg1_and_alg2_differ_when_a_delta_rises_from_zero():
    case = CASES["midrank_n30_s22"]
    assert case["alg1_selection"] != case["alg2_selection"]
    assert case["alg1_selection"][:9] == case["alg2_selection"][:9] == list(range(9))


@pytest.mark.skipif(not ref_extract.available(), reason="reference tree not present (GPU box)")
def test_oracle_matches_live_reference_on_a_fresh_input():
    ref = ref_extract.load()
    m = np.random.default_rng(11).standard_normal((26, 17))
    cov = np.cov(m, bias=True)
    for alg in (1, 2):
        want, _ = ref_extract.run_quiet(ref["placement_algorithm_%d" % alg], cov, 22)
        got, _, _ = go.pinv_greedy(cov, 22, algorithm=alg)
        assert got == [int(v) for v in want]
