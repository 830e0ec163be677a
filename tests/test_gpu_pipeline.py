"""GPU: the reference's main_architecture_2 flow through the drop-in modules, end to end at toy size
(examples/placement_pipeline.py): VGP training -> predicted tracers -> empirical covariance -> placement -> CSVs."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import greedy_oracle as go
from vgposp_b200 import cov_producer as cp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pipeline_end_to_end(tmp_path):
    spec = importlib.util.spec_from_file_location("placement_pipeline", os.path.join(ROOT, "examples", "placement_pipeline.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    res = mod.main(cover=4, samples=9, k=4, steps=25, n_obs=1500, m=32, batch=128, out=str(tmp_path), quiet=True,
                   nugget=1e-6)
    assert res["losses"][-1] < res["losses"][0]                          # the ELBO improves
    cov = res["cov_vv"]
    n = cov.shape[0]
    assert cov.shape == (64, 64) and np.array_equal(cov, cov.T) and np.all(np.diag(cov) >= 0)
    # the selection is what the CPU oracle picks on the same (nugget-regularised) matrix
    cov_spd = cov + (1e-6 * np.trace(cov) / n) * np.eye(n)
    want, _ = go.incremental_greedy_c(cov_spd, 4)
    assert res["selection"] == want
    # hand-off files round-trip
    assert np.array_equal(cp.load_cov_vv(str(tmp_path / "cov_vv_small.csv")), cov)
    sel = cp.read_indexed_csv(str(tmp_path / "placement_algorithm_selection_idxs.csv"), np.int64)[:, 0]
    assert [int(v) for v in sel] == res["selection"]


def test_pipeline_without_nugget_takes_the_pinv_path(tmp_path):
    """More locations than samples: the empirical covariance is rank-deficient and the placement runs in the
    reference's pinv semantics (every delta 0 for rank < n / 2: the first k indices)."""
    spec = importlib.util.spec_from_file_location("placement_pipeline", os.path.join(ROOT, "examples", "placement_pipeline.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    res = mod.main(cover=5, samples=5, k=4, steps=10, n_obs=1500, m=32, batch=128, out=None, quiet=True)
    cov = res["cov_vv"]
    assert cov.shape == (125, 125)
    assert np.linalg.matrix_rank(cov) <= 25
    assert res["selection"] == [0, 1, 2, 3]
