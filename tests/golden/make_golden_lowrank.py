"""Golden vectors for RANK-DEFICIENT covariances from the reference's own NumPy code (np.linalg.pinv inside every
nominator / denominator, placement_algorithm2.py:371-413), exec'd unmodified through oracle/ref_extract.py.

    python tests/golden/make_golden_lowrank.py        # build container only (needs /root/reference)

Inputs are empirical covariances of S samples over n > S locations -- np.cov(M, bias=True), what
gp_functions.py:1019-1057 / main_architecture_2.py:391-444 feed the placement -- of rank S - 1.  Writes
tests/golden/greedy_lowrank_golden.json (selections, per-step score vectors) and greedy_lowrank_inputs.npz."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_extract  # noqa: E402
from make_golden import step_scores  # noqa: E402

CASES = [
    # rank < n / 2: the regime of the reference's own runs (S = 144 or 250 samples, n >= 625 locations)
    dict(name="lowrank_n30_s8", n=30, s=8, seed=0, k=10),
    dict(name="lowrank_n60_s20", n=60, s=20, seed=1, k=12),
    dict(name="lowrank_n100_s40", n=100, s=40, seed=2, k=8),
    # rank > n / 2: between step n - rank and step rank both conditional variances are non-zero
    dict(name="midrank_n30_s22", n=30, s=22, seed=3, k=26),
    dict(name="midrank_n40_s33", n=40, s=33, seed=4, k=36),
]


def empirical_cov(n, s, seed):
    m = np.random.default_rng(seed).standard_normal((n, s))
    return np.cov(m, bias=True)


def main():
    ref = ref_extract.load()
    golden, arrays = {}, {}
    for case in CASES:
        cov = empirical_cov(case["n"], case["s"], case["seed"])
        arrays[case["name"]] = cov
        sel1, _ = ref_extract.run_quiet(ref["placement_algorithm_1"], cov, case["k"])
        sel2, out = ref_extract.run_quiet(ref["placement_algorithm_2"], cov, case["k"])
        rec = dict(case, rank=int(np.linalg.matrix_rank(cov)), alg1_selection=[int(v) for v in sel1],
                   alg2_selection=[int(v) for v in sel2], step_scores=step_scores(ref, cov, sel1))
        golden[case["name"]] = rec
        nz = [sum(1 for v in row if v) for row in rec["step_scores"]]
        print(case["name"], "rank", rec["rank"], sel1, "alg2 equal", sel1 == sel2, "non-zero deltas per step", nz, flush=True)
    meta = {"numpy": np.__version__, "generator": "tests/golden/make_golden_lowrank.py",
            "source": "/root/reference/placement_algorithm2.py (functions exec'd unmodified)"}
    with open(os.path.join(HERE, "greedy_lowrank_golden.json"), "w") as fh:
        json.dump({"meta": meta, "cases": golden}, fh, indent=1)
    np.savez_compressed(os.path.join(HERE, "greedy_lowrank_inputs.npz"), **arrays)


if __name__ == "__main__":
    main()
