"""External anchor for the exact-GP rows of the oracle: scikit-learn's GaussianProcessRegressor is an independent,
widely used implementation of the same arithmetic (kernel matrix -> Cholesky -> log marginal likelihood, posterior mean
and standard deviation) with the same kernel families (RBF == ExponentiatedQuadratic, Matern nu = 1/2, 3/2, 5/2).
TensorFlow-Probability -- what the reference calls -- is not installable here, so this is the
third-party implementation the restatement in oracle/gp_oracle.py is pinned to for a2 / a3 and the kernel formulas."""
import numpy as np
import pytest

from oracle import gp_oracle as gpo

sk = pytest.importorskip("sklearn.gaussian_process")
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern  # noqa: E402

KERNELS = {"expquad": lambda l: RBF(l), "matern12": lambda l: Matern(l, nu=0.5), "matern32": lambda l: Matern(l, nu=1.5),
           "matern52": lambda l: Matern(l, nu=2.5)}


def data(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-2, 2, (n, d))
    y = np.sin(x[:, 0]) * np.sin(x[:, 1 % d]) + 0.1 + 0.1 * rng.standard_normal(n)
    return x, y


@pytest.mark.parametrize("kind", sorted(KERNELS))
def test_kernel_matrix_matches_sklearn(kind):
    x1, _ = data(40, 3, 1)
    x2, _ = data(55, 3, 2)
    amp, ls = 1.3, 0.7
    want = (ConstantKernel(amp ** 2) * KERNELS[kind](ls))(x1, x2)
    np.testing.assert_allclose(gpo.kernel_matrix(kind, x1, x2, amp, ls), want, rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("kind", sorted(KERNELS))
@pytest.mark.parametrize("n,d", [(25, 2), (120, 3), (60, 5)])
def test_log_marginal_likelihood_and_posterior_match_sklearn(kind, n, d):
    x, y = data(n, d, n + d)
    xt, _ = data(33, d, 99)
    amp, ls, noise, jitter = 0.9, 0.8, 0.05, 1e-6
    kernel = ConstantKernel(amp ** 2, "fixed") * KERNELS[kind](ls).clone_with_theta(np.log([ls]))
    gpr = sk.GaussianProcessRegressor(kernel=kernel, alpha=noise + jitter, optimizer=None, normalize_y=False).fit(x, y)
    want_ll = gpr.log_marginal_likelihood(gpr.kernel_.theta)
    assert gpo.gp_log_prob(x, y, amp, ls, noise, jitter, kind=kind) == pytest.approx(want_ll, rel=1e-10)
    mean, std = gpr.predict(xt, return_std=True)
    got_mean, got_var = gpo.gp_regression(x, y, xt, amp, ls, noise, 0.0, divisor_jitter=jitter, kind=kind)
    np.testing.assert_allclose(got_mean, mean, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(np.sqrt(got_var), std, rtol=1e-6, atol=1e-9)


def test_log_marginal_likelihood_gradient_matches_sklearn():
    """sklearn differentiates with respect to log-parameters: d/d log(theta) = theta d/d theta."""
    x, y = data(50, 3, 7)
    amp, ls, noise = 1.1, 0.6, 0.1
    from sklearn.gaussian_process.kernels import WhiteKernel
    kernel = ConstantKernel(amp ** 2) * RBF(ls) + WhiteKernel(noise)
    gpr = sk.GaussianProcessRegressor(kernel=kernel, alpha=0.0, optimizer=None).fit(x, y)
    want_ll, want_g = gpr.log_marginal_likelihood(gpr.kernel_.theta, eval_gradient=True)      # theta = log(a^2, l, s2)
    ll, g = gpo.gp_log_prob_grad(x, y, amp, ls, noise, jitter=0.0, kind="expquad")
    assert ll == pytest.approx(want_ll, rel=1e-10)
    np.testing.assert_allclose([0.5 * amp * g[0], ls * g[1], noise * g[2]], want_g, rtol=1e-7)
