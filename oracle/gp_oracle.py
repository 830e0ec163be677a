"""TEST INFRASTRUCTURE ONLY -- float64 NumPy/SciPy restatement of the GP arithmetic behind the path.

**Parity unpinned.**  The reference delegates this arithmetic to TensorFlow-Probability
(`tfkern.ExponentiatedQuadratic`, `tfd.GaussianProcess`, `tfd.GaussianProcessRegressionModel`,
`tfd.VariationalGaussianProcess`), which is neither vendored under /root/reference nor installed
here, and the reference pins no TFP version and asserts no numeric result for these calls (SURVEY.md
section 8c).  The functions below restate TFP's published algorithm as recorded in SURVEY.md
Appendix A.2-A.5 and are anchored on the reference's call sites:

    kernel                variational_Gaussian_process_example.py:55-57, 3D_sin_wave.py:158-159
    exact GP log_prob     gp_functions.py:166-172 (fit_gp), 3D_sin_wave.py:161-172
    GP regression model   gp_functions.py:283-297, 3D_sin_wave.py:262-268
    optimal posterior     variational_Gaussian_process_example.py:68-74, main_architecture_2.py:200-206
    VGP loss / mean       variational_Gaussian_process_example.py:83-99,141-142
    Adam                  gp_functions.py:179-182 (tf.train.AdamOptimizer defaults)
    softplus params       gp_functions.py:124-135, variational_Gaussian_process_example.py:47-61

The exact-GP functions and all four kernel families are additionally pinned to an independent third-party
implementation available in this image, scikit-learn's GaussianProcessRegressor (tests/test_oracle_gp_sklearn.py:
kernel matrices 1e-12, log marginal likelihood 1e-10, posterior mean/std, LML gradient).

The one first-party NumPy statement of kernel -> Cholesky -> solve -> posterior variance in the
reference, `plot_confidence_interval.py:17-19,26,43-51`, is reproduced by `expquad_matrix`,
`gp_regression` (see tests/test_oracle_gp.py).

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this module.
"""
import numpy as np
from scipy.linalg import cholesky, solve_triangular

LOG_2PI = float(np.log(2.0 * np.pi))
DEFAULT_JITTER = 1e-6        # TFP class default; passed explicitly at main_architecture_2.py:225


def softplus(v):
    return np.logaddexp(0.0, v)


def softplus_inverse(p):
    """gp_functions.py:106-109 invert_softplus: v = log(exp(p) - 1)."""
    return np.log(np.expm1(p))


def sqdist(x1, x2):
    """Direct sum of squared coordinate differences (no |x|^2+|y|^2-2xy expansion)."""
    d = x1[:, None, :] - x2[None, :, :]
    return np.einsum("ijk,ijk->ij", d, d)


def expquad_matrix(x1, x2, amplitude, length_scale, diag_add=0.0):
    """k(x,y) = a^2 exp(-|x-y|^2 / (2 l^2)); `diag_add` is added where row index == column index."""
    k = amplitude ** 2 * np.exp(sqdist(x1, x2) * (-0.5 / length_scale ** 2))
    if diag_add:
        m = min(k.shape)
        k[np.arange(m), np.arange(m)] += diag_add
    return k


KERNEL_KINDS = {"expquad": 0, "matern12": 1, "matern32": 2, "matern52": 3}


def kernel_matrix(kind, x1, x2, amplitude, length_scale, diag_add=0.0):
    """The stationary kernels the reference instantiates, in TFP's published form (r = |x - y|):
    ExponentiatedQuadratic (variational_Gaussian_process_example.py:55-57); MaternOneHalf a^2 exp(-r/l)
    (gp_functions.py:160-163); MaternThreeHalves a^2 (1 + z) exp(-z), z = sqrt(3) r / l; MaternFiveHalves
    a^2 (1 + z + z^2/3) exp(-z), z = sqrt(5) r / l (main_architecture_2.py:184)."""
    if kind == "expquad":
        return expquad_matrix(x1, x2, amplitude, length_scale, diag_add)
    r = np.sqrt(sqdist(x1, x2))
    if kind == "matern12":
        k = np.exp(-r / length_scale)
    elif kind == "matern32":
        z = np.sqrt(3.0) * r / length_scale
        k = (1.0 + z) * np.exp(-z)
    elif kind == "matern52":
        z = np.sqrt(5.0) * r / length_scale
        k = (1.0 + z + z * z / 3.0) * np.exp(-z)
    else:
        raise ValueError(kind)
    k *= amplitude ** 2
    if diag_add:
        m = min(k.shape)
        k[np.arange(m), np.arange(m)] += diag_add
    return k


def _chol(a):
    return cholesky(a, lower=True, check_finite=False)


def _fwd(l, b):
    return solve_triangular(l, b, lower=True, check_finite=False)


def _bwd(l, b):
    return solve_triangular(l, b, lower=True, trans="T", check_finite=False)


def _kern(kind):
    return lambda x1, x2, a, l, diag_add=0.0: kernel_matrix(kind, x1, x2, a, l, diag_add)


def gp_log_prob(x, y, amplitude, length_scale, noise_variance, jitter=DEFAULT_JITTER, mean=0.0, kind="expquad"):
    """Exact-GP marginal log-likelihood: L = chol(K + (s2 + jitter) I)."""
    expquad_matrix = _kern(kind)
    n = x.shape[0]
    l = _chol(expquad_matrix(x, x, amplitude, length_scale, noise_variance + jitter))
    z = _fwd(l, y - mean)
    return -0.5 * float(z @ z) - float(np.sum(np.log(np.diag(l)))) - 0.5 * n * LOG_2PI


def gp_regression(x_obs, y_obs, x_pred, amplitude, length_scale, noise_variance,
                  predictive_noise_variance=0.0, divisor_jitter=0.0, mean=0.0, full_cov=False, kind="expquad"):
    """Posterior mean / (co)variance of GaussianProcessRegressionModel."""
    expquad_matrix = _kern(kind)
    l = _chol(expquad_matrix(x_obs, x_obs, amplitude, length_scale, noise_variance + divisor_jitter))
    k_xt = expquad_matrix(x_obs, x_pred, amplitude, length_scale)
    c = _fwd(l, k_xt)
    post_mean = mean + c.T @ _fwd(l, y_obs - mean)
    if full_cov:
        cov = expquad_matrix(x_pred, x_pred, amplitude, length_scale, predictive_noise_variance) - c.T @ c
        return post_mean, cov
    var = amplitude ** 2 - np.sum(c * c, axis=0) + predictive_noise_variance
    return post_mean, var


def optimal_variational_posterior(z, x, y, amplitude, length_scale, noise_variance,
                                  jitter=DEFAULT_JITTER, legacy_scale_orientation=False, kind="expquad"):
    """Titsias optimum.  Returns (loc [m], scale [m,m]) with S = scale @ scale.T =
    K_zz Sigma K_zz; `legacy_scale_orientation` returns L_Sigma^-1 K_zz (S = scale.T @ scale)."""
    expquad_matrix = _kern(kind)
    k_zz = expquad_matrix(z, z, amplitude, length_scale)
    k_zx = expquad_matrix(z, x, amplitude, length_scale)
    sigma_inv = k_zz + (k_zx @ k_zx.T) / noise_variance
    sigma_inv[np.diag_indices_from(sigma_inv)] += jitter
    l_s = _chol(sigma_inv)
    loc = (k_zz @ _bwd(l_s, _fwd(l_s, k_zx @ y))) / noise_variance
    scale = _fwd(l_s, k_zz)
    return loc, (scale if legacy_scale_orientation else scale.T)


def vgp_terms(z, q_loc, q_scale, x_b, y_b, amplitude, length_scale, noise_variance, kl_weight,
              jitter=DEFAULT_JITTER, kind="expquad"):
    """All pieces of the negative ELBO of SURVEY.md A.2 (dict), S = q_scale @ q_scale.T."""
    expquad_matrix = _kern(kind)
    m, b = z.shape[0], x_b.shape[0]
    l = _chol(expquad_matrix(z, z, amplitude, length_scale, jitter))
    k_zb = expquad_matrix(z, x_b, amplitude, length_scale)
    alpha = _bwd(l, _fwd(l, q_loc))
    mu_b = k_zb.T @ alpha
    r = y_b - mu_b
    ll = -0.5 * float(r @ r) / noise_variance - 0.5 * b * (LOG_2PI + np.log(noise_variance))
    c = _fwd(l, k_zb)
    d = _bwd(l, c)
    tr1 = b * amplitude ** 2 - float(np.sum(c * c))
    e = q_scale.T @ d
    tr2 = float(np.sum(e * e))
    li_a = _fwd(l, q_scale)                         # tr(Kzz^-1 S) = |L^-1 A|_F^2
    li_mu = _fwd(l, q_loc)
    sign, logdet_s = np.linalg.slogdet(q_scale @ q_scale.T)
    logdet_k = 2.0 * float(np.sum(np.log(np.diag(l))))
    kl = 0.5 * (float(np.sum(li_a * li_a)) + float(li_mu @ li_mu) - m + logdet_k - logdet_s)
    loss = -(ll - 0.5 * (tr1 + tr2) / noise_variance - kl_weight * kl)
    return dict(loss=loss, ll=ll, tr1=tr1, tr2=tr2, kl=kl, mu_b=mu_b, chol_kzz=l, alpha=alpha)


def vgp_loss(*args, **kwargs):
    return vgp_terms(*args, **kwargs)["loss"]


def vgp_predict(z, q_loc, q_scale, x_t, amplitude, length_scale, predictive_noise_variance=0.0,
                jitter=DEFAULT_JITTER, kind="expquad"):
    """Predictive mean and marginal variance of the VGP at x_t (last paragraph)."""
    expquad_matrix = _kern(kind)
    l = _chol(expquad_matrix(z, z, amplitude, length_scale, jitter))
    k_zt = expquad_matrix(z, x_t, amplitude, length_scale)
    c = _fwd(l, k_zt)
    d = _bwd(l, c)
    mean = d.T @ q_loc
    e = q_scale.T @ d
    var = amplitude ** 2 - np.sum(c * c, axis=0) + np.sum(e * e, axis=0) + predictive_noise_variance
    return mean, var


def gp_log_prob_grad(x, y, amplitude, length_scale, noise_variance, jitter=DEFAULT_JITTER, kind="expquad"):
    """(log_prob, d log_prob / d (amplitude, length_scale, noise_variance)): dL/dtheta = tr(W dC/dtheta) with
    W = (alpha alpha^T - C^-1) / 2 -- what TF autodiff gives tf_train_gp_adam (gp_functions.py:179-182)."""
    c = kernel_matrix(kind, x, x, amplitude, length_scale, noise_variance + jitter)
    cinv = np.linalg.inv(c)
    alpha = cinv @ y
    w = 0.5 * (np.outer(alpha, alpha) - cinv)
    k = kernel_matrix(kind, x, x, amplitude, length_scale)
    r2 = sqdist(x, x)
    r = np.sqrt(r2)
    if kind == "expquad":
        dk = k * r2 / length_scale ** 3
    elif kind == "matern12":
        dk = k * r / length_scale ** 2
    elif kind == "matern32":
        z = np.sqrt(3.0) * r / length_scale
        dk = amplitude ** 2 * z * z * np.exp(-z) / length_scale
    else:
        z = np.sqrt(5.0) * r / length_scale
        dk = amplitude ** 2 * (z * z / 3.0) * (1.0 + z) * np.exp(-z) / length_scale
    grads = np.array([2.0 * np.sum(w * k) / amplitude, np.sum(w * dk), np.trace(w)])
    return gp_log_prob(x, y, amplitude, length_scale, noise_variance, jitter, kind=kind), grads


def gp_train_adam(x, y, v_init, lr, iters, jitter=DEFAULT_JITTER, kind="matern12", tiny=np.finfo(np.float64).tiny):
    """The exact-GP training loop of gpf.tf_optimize_model_params (gp_functions.py:228-259) on the unconstrained
    variables v = (v_amplitude, v_length_scale, v_noise), theta = tiny + softplus(v) (:124-135), Adam on -log_prob
    (:179-182).  One initial run, then iters + 1 recorded steps; returns (lls [iters + 1], final v)."""
    v = np.array(v_init, dtype=np.float64)
    opt = TfAdam(3, lr)
    lls = np.zeros(iters + 1)
    for it in range(iters + 2):
        theta = tiny + softplus(v)
        ll, g = gp_log_prob_grad(x, y, theta[0], theta[1], theta[2], jitter, kind)
        if it > 0:
            lls[it - 1] = ll
        v = opt.step(v, -g / (1.0 + np.exp(-v)))
    return lls, v


class TfAdam:
    """tf.train.AdamOptimizer update (beta1 .9, beta2 .999, eps 1e-8; 'epsilon hat' form):
    lr_t = lr sqrt(1-b2^t)/(1-b1^t);  theta -= lr_t m / (sqrt(v) + eps)."""

    def __init__(self, shape, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self.m = np.zeros(shape)
        self.v = np.zeros(shape)
        self.t = 0
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps

    def step(self, theta, grad):
        self.t += 1
        self.m = self.b1 * self.m + (1 - self.b1) * grad
        self.v = self.b2 * self.v + (1 - self.b2) * grad * grad
        lr_t = self.lr * np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        return theta - lr_t * self.m / (np.sqrt(self.v) + self.eps)
