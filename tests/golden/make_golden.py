"""Generate the greedy-placement golden vectors from the reference's own NumPy code.

Run once in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

It exec's the unmodified functions of /root/reference/placement_algorithm2.py (via
oracle/ref_extract.py) on seeded inputs and writes

    tests/golden/greedy_golden.json   selections, per-step score vectors, alg. 2 print trace
    tests/golden/greedy_inputs.npz    the input matrices for n <= 100 (larger ones are rebuilt from the
                                      recipe and checked against a sha256)

The GPU box has no /root/reference; tests only read the two files above.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_extract  # noqa: E402


def expquad_cloud(n, seed=0, amplitude=1.0, length_scale=0.5, nugget=1e-2):
    """SURVEY.md section 8c recipe: uniform cloud in [-2,2]^3, direct squared distances."""
    x = np.random.default_rng(seed).uniform(-2, 2, (n, 3))
    d = x[:, None, :] - x[None, :, :]
    k = amplitude ** 2 * np.exp(-np.sum(d * d, axis=-1) / (2 * length_scale ** 2))
    return k + nugget * np.eye(n)


def legacy_random_cov(ref, n):
    np.random.seed(0)
    return ref["dg_create_random_cov"](n)      # placement_algorithm2.py:441-444


def build_inputs(ref, case):
    kind = case["kind"]
    if kind == "fixture4x4":
        return ref["cov_vv_4x4"]()              # placement_algorithm2.py:473-479
    if kind == "legacy_random":
        return legacy_random_cov(ref, case["n"])
    if kind == "expquad_cloud":
        return expquad_cloud(case["n"], case["seed"], 1.0, case["length_scale"], case["nugget"])
    raise ValueError(kind)


CASES = [
    dict(name="fixture4x4", kind="fixture4x4", n=4, k=4, algs=[1, 2]),
    dict(name="legacy_random_n10", kind="legacy_random", n=10, k=5, algs=[1, 2]),
    dict(name="legacy_random_n20", kind="legacy_random", n=20, k=5, algs=[1, 2]),
    dict(name="legacy_random_n40", kind="legacy_random", n=40, k=5, algs=[1, 2]),
    dict(name="expquad_n50", kind="expquad_cloud", n=50, seed=0, length_scale=0.5, nugget=1e-2, k=5, algs=[1, 2]),
    dict(name="expquad_n100", kind="expquad_cloud", n=100, seed=0, length_scale=0.5, nugget=1e-2, k=8, algs=[1, 2]),
    dict(name="expquad_n200", kind="expquad_cloud", n=200, seed=0, length_scale=0.5, nugget=1e-2, k=6, algs=[2]),
    dict(name="expquad_n200_nugget1e-6", kind="expquad_cloud", n=200, seed=0, length_scale=0.5, nugget=1e-6, k=6, algs=[2]),
    dict(name="expquad_n400", kind="expquad_cloud", n=400, seed=0, length_scale=0.5, nugget=1e-2, k=3, algs=[2]),
]


def step_scores(ref, cov, selection):
    """Per-step score of every remaining candidate, from the reference's nominator/denominator."""
    n = cov.shape[0]
    out = []
    A, A_bar = [], list(range(n))
    for y_sel in selection:
        row = [None] * n
        for y in range(n):
            if y in A:
                continue
            nom = ref["nominator"](y, A, cov)
            den = ref["denominator"](y, A_bar, cov)
            if np.abs(den) < 1e-8 or np.abs(nom) < 1e-8:    # placement_algorithm2.py:116-119
                row[y] = 0.0
            else:
                row[y] = float((nom / den).reshape(()))
        out.append(row)
        A.append(int(y_sel))
        A_bar.remove(int(y_sel))
    return out


def main():
    if not ref_extract.available():
        raise SystemExit("reference tree not found at %s" % ref_extract.REFERENCE_ROOT)
    ref = ref_extract.load()
    golden, arrays = {}, {}
    for case in CASES:
        cov = build_inputs(ref, case)
        rec = {k: v for k, v in case.items() if k != "algs"}
        rec["sha256"] = hashlib.sha256(np.ascontiguousarray(cov).tobytes()).hexdigest()
        if case["n"] <= 100:
            arrays[case["name"]] = cov
        for alg in case["algs"]:
            t0 = time.time()
            sel, out = ref_extract.run_quiet(ref["placement_algorithm_%d" % alg], cov, case["k"])
            rec["alg%d_selection" % alg] = [int(s) for s in sel]
            rec["alg%d_seconds_8core_container" % alg] = round(time.time() - t0, 3)
            if alg == 2:
                rec["alg2_stdout"] = out.splitlines()
        sel = rec.get("alg1_selection", rec.get("alg2_selection"))
        if case["n"] <= 100:
            rec["step_scores"] = step_scores(ref, cov, sel)
        golden[case["name"]] = rec
        print(case["name"], sel, {k: v for k, v in rec.items() if k.endswith("container")}, flush=True)
    meta = {"numpy": np.__version__, "generator": "tests/golden/make_golden.py",
            "source": "/root/reference/placement_algorithm2.py (functions exec'd unmodified)"}
    with open(os.path.join(HERE, "greedy_golden.json"), "w") as fh:
        json.dump({"meta": meta, "cases": golden}, fh, indent=1)
    np.savez_compressed(os.path.join(HERE, "greedy_inputs.npz"), **arrays)


if __name__ == "__main__":
    main()
