"""CPU: internal consistency of the GP oracle (oracle/gp_oracle.py) and its agreement with the reference's only
first-party NumPy statement of kernel -> Cholesky -> solve -> posterior (plot_confidence_interval.py:17-51)."""
import numpy as np
import pytest

from oracle import gp_oracle as gpo


def _restated_kernel(a, b, param):
    # plot_confidence_interval.py:17-19: exp(-.5 / param * sqdist) with the expanded squared distance
    sqdist = np.sum(a ** 2, 1).reshape(-1, 1) + np.sum(b ** 2, 1) - 2 * np.dot(a, b.T)
    return np.exp(-.5 * (1 / param) * sqdist)


def _reference_kernel():
    """The reference's own `kernel` (plot_confidence_interval.py:17-19), cut out of the file with `ast` and exec'd
    unmodified (the module top imports matplotlib and plots) -- only in the build container; on the GPU box the tree is
    absent and the restatement above stands in (the test below checks the two against each other here)."""
    import ast
    import os
    path = os.path.join(os.environ.get("VGPOSP_REFERENCE_ROOT", "/root/reference"), "plot_confidence_interval.py")
    if not os.path.isfile(path):
        return None
    tree = ast.parse(open(path).read(), filename=path)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "kernel"]
    ns = {"np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns["kernel"]


reference_style_kernel = _reference_kernel() or _restated_kernel


def test_restated_kernel_is_the_references_own():
    ref = _reference_kernel()
    if ref is None:
        pytest.skip("reference tree not present (GPU box)")
    rng = np.random.default_rng(3)
    a, b = rng.uniform(-3, 3, (17, 2)), rng.uniform(-3, 3, (23, 2))
    np.testing.assert_array_equal(ref(a, b, 0.3), _restated_kernel(a, b, 0.3))


def test_kernel_matches_reference_numpy_statement():
    rng = np.random.default_rng(0)
    a, b = rng.uniform(-3, 3, (40, 1)), rng.uniform(-3, 3, (55, 1))
    param = 0.3                                       # length_scale^2
    np.testing.assert_allclose(gpo.expquad_matrix(a, b, 1.0, np.sqrt(param)), reference_style_kernel(a, b, param),
                               rtol=1e-12, atol=1e-15)


def test_regression_matches_reference_numpy_chain():
    # plot_confidence_interval.py:38-51: K + 5e-5 I, Lk = L^-1 K_s, mu = Lk^T L^-1 y, s2 = diag(K_ss) - sum Lk^2
    xtrain = np.array([-4, -3, -2, -1, 1.0]).reshape(5, 1)
    ytrain = np.sin(xtrain).reshape(-1)
    xtest = np.linspace(-5, 5, 50).reshape(-1, 1)
    param = 0.3
    K = reference_style_kernel(xtrain, xtrain, param)
    L = np.linalg.cholesky(K + 0.00005 * np.eye(5))
    Lk = np.linalg.solve(L, reference_style_kernel(xtrain, xtest, param))
    mu = np.dot(Lk.T, np.linalg.solve(L, ytrain))
    s2 = np.diag(reference_style_kernel(xtest, xtest, param)) - np.sum(Lk ** 2, axis=0)
    mean, var = gpo.gp_regression(xtrain, ytrain, xtest, 1.0, np.sqrt(param), 0.00005)
    np.testing.assert_allclose(mean, mu, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(var, s2, rtol=1e-8, atol=1e-10)


def test_log_prob_against_scipy():
    from scipy.stats import multivariate_normal
    rng = np.random.default_rng(1)
    x = rng.uniform(-2, 2, (30, 3))
    y = rng.standard_normal(30)
    cov = gpo.expquad_matrix(x, x, 1.2, 0.8, diag_add=0.1 + 1e-6)
    assert gpo.gp_log_prob(x, y, 1.2, 0.8, 0.1) == pytest.approx(multivariate_normal(np.zeros(30), cov).logpdf(y),
                                                                   rel=1e-10)


def test_elbo_at_optimum_equals_collapsed_bound():
    """With q at the Titsias optimum, full batch and kl_weight 1, the negative ELBO equals the collapsed bound
    -log N(y | 0, Q + s2 I) + tr(K - Q) / (2 s2), Q = K_xz K_zz^-1 K_zx (up to the jitter)."""
    rng = np.random.default_rng(2)
    n, m = 200, 15
    x = rng.uniform(-2, 2, (n, 2))
    y = np.sin(x[:, 0]) + 0.1 * rng.standard_normal(n)
    z = rng.uniform(-2, 2, (m, 2))
    amp, ls, s2, jit = 1.0, 0.9, 0.05, 1e-10
    loc, scale = gpo.optimal_variational_posterior(z, x, y, amp, ls, s2, jitter=jit)
    loss = gpo.vgp_loss(z, loc, scale, x, y, amp, ls, s2, 1.0, jitter=jit)
    kzz = gpo.expquad_matrix(z, z, amp, ls, diag_add=jit)
    kzx = gpo.expquad_matrix(z, x, amp, ls)
    q = kzx.T @ np.linalg.solve(kzz, kzx)
    from scipy.stats import multivariate_normal
    collapsed = -multivariate_normal(np.zeros(n), q + s2 * np.eye(n)).logpdf(y) + (n * amp ** 2 - np.trace(q)) / (2 * s2)
    assert loss == pytest.approx(collapsed, rel=1e-6)
    # and any other q is worse
    worse = gpo.vgp_loss(z, loc * 1.05, scale, x, y, amp, ls, s2, 1.0, jitter=jit)
    assert worse > loss


def test_predictive_mean_interpolates_training_signal():
    rng = np.random.default_rng(3)
    x = rng.uniform(-2, 2, (400, 1))
    y = np.sin(2 * x[:, 0]) + 0.05 * rng.standard_normal(400)
    z = np.linspace(-2, 2, 25)[:, None]
    loc, scale = gpo.optimal_variational_posterior(z, x, y, 1.0, 0.5, 0.05 ** 2)
    xt = np.linspace(-1.8, 1.8, 50)[:, None]
    mean, var = gpo.vgp_predict(z, loc, scale, xt, 1.0, 0.5)
    assert np.max(np.abs(mean - np.sin(2 * xt[:, 0]))) < 0.05
    assert np.all(var > 0) and np.all(var < 0.05)


def test_softplus_inverse_and_adam():
    v = np.array([-3.0, 0.0, 0.54, 5.0])
    np.testing.assert_allclose(gpo.softplus_inverse(gpo.softplus(v)), v, rtol=1e-12, atol=1e-12)
    opt = gpo.TfAdam((2,), lr=0.01)
    th = opt.step(np.array([1.0, -1.0]), np.array([0.5, -2.0]))
    # first Adam step moves every coordinate by ~lr against the gradient sign
    np.testing.assert_allclose(th, [1.0 - 0.01, -1.0 + 0.01], atol=1e-6)
