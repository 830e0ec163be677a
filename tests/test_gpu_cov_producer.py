"""GPU parity for the rows either side of the placement path: the empirical covariance
producer and the decay filter, against NumPy statements of the reference's per-pair loops."""
import numpy as np
import pytest

from vgposp_b200 import cov_producer as cp
import vgposp_b200.gp_functions as gpf
import vgposp_b200.placement_algorithm2 as alg2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,s", [(1, 5), (7, 3), (64, 300), (130, 257), (500, 1000)])
def test_empirical_cov_matches_np_cov(n, s):
    m = np.random.default_rng(n * s).standard_normal((n, s)) * 3.0 + 5.0
    got = cp.empirical_cov(m)
    want = np.atleast_2d(np.cov(m, bias=True))                     # np.cov(a, b, bias=True)[0, 1] for every pair
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)
    assert np.array_equal(got, got.T)
    # the reference's literal form for a few pairs (gp_functions.py:1054)
    for i, j in [(0, 0), (0, n - 1), (n // 2, n // 3)]:
        assert got[i, j] == pytest.approx(np.cov(m[i], m[j], bias=True)[0, 1], rel=1e-10, abs=1e-12)


def test_cov_taper_matches_decay_fn():
    idx = cp.gen_idxs([4, 5, 3])
    n = idx.shape[0]
    c = np.random.default_rng(3).standard_normal((n, n))
    c = c @ c.T
    beta = 1.3
    got = cp.cov_taper(c, idx, beta)
    d = np.sqrt(((idx[:, None, :] - idx[None, :, :]).astype(np.float64) ** 2).sum(-1))
    decay = np.exp(-(beta * d) ** 2 / (2 * np.pi))                 # ..._sampledistribution.py:388-391
    decay[decay < 0.01] = 0.0
    np.testing.assert_allclose(got, c * decay, rtol=1e-13, atol=0)
    assert (got == 0).sum() == (decay == 0).sum() > 0


def test_create_cov_matrix_feeds_the_greedy():
    """gp_functions.create_cov_matrix (:1019-1057) with a VGP predictor as the encoder, then placement on the result:
    the reference's main_architecture_2 pipeline end to end at toy size (27 locations, 6 x 6 (p, t) samples)."""
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 2, (300, 5))
    y = np.sin(x[:, 0] + 0.3 * x[:, 3]) * np.cos(x[:, 1]) + 0.2 * x[:, 2] * x[:, 4] + 0.05 * rng.standard_normal(300)
    z = rng.uniform(0, 2, (40, 5))
    k = gpf.MaternFiveHalves(1.0, 1.5)                              # main_architecture_2.py:184
    loc, scale = gpf.VariationalGaussianProcess.optimal_variational_posterior(k, z, x, y, 0.05)

    def encoder(points):
        return gpf.VariationalGaussianProcess(k, points, z, loc, scale, 0.05, predictive_noise_variance=0.0).mean()

    cov = gpf.create_cov_matrix([0, 2], [0, 2], [0, 2], [0.0, 2.0], [0.0, 2.0], 3, 6, encoder, None)
    assert cov.shape == (27, 27) and np.array_equal(cov, cov.T)
    # literal statement of the reference's loop for a few pairs
    grid = np.array(np.meshgrid(np.linspace(0, 2, 6), np.linspace(0, 2, 6))).reshape(2, -1).T

    def tracers(i0, i1, i2):
        pts = np.column_stack([np.full(36, i0), np.full(36, i1), np.full(36, i2), grid])
        return encoder(pts)

    for (a, b) in [((0, 0, 0), (2, 1, 0)), ((1, 2, 2), (1, 2, 2)), ((2, 2, 1), (0, 1, 2))]:
        i = a[0] + a[1] * 3 + a[2] * 9
        j = b[0] + b[1] * 3 + b[2] * 9
        assert cov[i, j] == pytest.approx(np.cov(tracers(*a), tracers(*b), bias=True)[0, 1], rel=1e-8, abs=1e-14)
    # 27 locations from 36 samples: full rank; a nugget makes it safely SPD for the inverse-based path
    sel = alg2.placement_algorithm_2(cov + 1e-6 * np.trace(cov) / 27 * np.eye(27), 5)
    assert len(set(int(v) for v in sel)) == 5
