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
    gp = gpf.fit_gp(gpf.ExponentiatedQuadratic(amp, ls), x, noise)
    want = gpo.gp_log_prob(x, y, amp, ls, noise)
    assert gp.log_prob(y) == pytest.approx(want, rel=RTOL)
    gp = gpf.fit_gp(gpf.create_cov_kernel(amp, ls), x, noise)              # the reference's choice: MaternOneHalf
    assert gp.log_prob(y) == pytest.approx(gpo.gp_log_prob(x, y, amp, ls, noise, kind="matern12"), rel=RTOL)


def test_gp_regression_model():
    x, y = data(400, 3)
    xt = np.random.default_rng(9).uniform(-2, 2, (333, 3))
    amp, ls, noise, pnoise = 0.9, 0.8, 0.02, 0.01
    gprm = gpf.tf_gp_regression_model(gpf.ExponentiatedQuadratic(amp, ls), xt, x, y, noise, pnoise)
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
            want = gpo.gp_log_prob(x, y, 40 * (1 + j) / 2, 40 * (1 + i) / 3, float(noise), kind="matern12")
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
    assert gp.log_prob(yt) == pytest.approx(gpo.gp_log_prob(x, y, 1.0, 0.5, 0.1, kind="matern12"), rel=RTOL)
    k = gpf.ExponentiatedQuadratic(1.0, 0.5).matrix(xt, xt)
    np.testing.assert_allclose(k, gpo.expquad_matrix(x, x, 1.0, 0.5), rtol=1e-13)
    with pytest.raises(TypeError):
        _ffi.as_device_f64(xt.float(), 0)


KINDS = {"expquad": gpf.ExponentiatedQuadratic, "matern12": gpf.MaternOneHalf, "matern32": gpf.MaternThreeHalves,
         "matern52": gpf.MaternFiveHalves}


@pytest.mark.parametrize("kind", ["matern12", "matern32", "matern52"])
def test_matern_kernels_through_gp_and_vgp(kind):
    """The kernels the reference actually instantiates (MaternOneHalf main.py:94, MaternFiveHalves over 5-D xyztp
    inputs main_architecture_2.py:184) through every GP / VGP entry point."""
    x, y = data(500, 11, d=5)
    xt = np.random.default_rng(12).uniform(-2, 2, (200, 5))
    amp, ls, noise = 1.1, 1.4, 0.04
    k = KINDS[kind](amp, ls)
    assert gpf.fit_gp(k, x, noise).log_prob(y) == pytest.approx(gpo.gp_log_prob(x, y, amp, ls, noise, kind=kind), rel=RTOL)
    gprm = gpf.tf_gp_regression_model(k, xt, x, y, noise, 0.0)
    mean, var = gpo.gp_regression(x, y, xt, amp, ls, noise, 0.0, kind=kind)
    np.testing.assert_allclose(gprm.mean(), mean, rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(gprm.variance(), var, rtol=1e-7, atol=1e-12)
    z = np.random.default_rng(13).uniform(-2, 2, (64, 5))
    loc, scale = gpf.VariationalGaussianProcess.optimal_variational_posterior(k, z, x, y, noise)
    want_loc, want_scale = gpo.optimal_variational_posterior(z, x, y, amp, ls, noise, kind=kind)
    np.testing.assert_allclose(loc, want_loc, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(scale @ scale.T, want_scale @ want_scale.T, rtol=1e-7, atol=1e-10)
    vgp = gpf.VariationalGaussianProcess(k, xt, z, want_loc, want_scale, noise)
    idx = np.random.default_rng(14).integers(500, size=128)
    got = vgp.variational_loss(y[idx], x[idx], kl_weight=128 / 500, return_terms=True)
    want = gpo.vgp_terms(z, want_loc, want_scale, x[idx], y[idx], amp, ls, noise, 128 / 500, kind=kind)
    for key in ("loss", "ll", "kl"):
        assert got[key] == pytest.approx(want[key], rel=1e-8)
    wm, wv = gpo.vgp_predict(z, want_loc, want_scale, xt, amp, ls, kind=kind)
    np.testing.assert_allclose(vgp.mean(), wm, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(vgp.variance(), wv, rtol=1e-6, atol=1e-10)


@pytest.mark.parametrize("kind", ["expquad", "matern12", "matern32", "matern52"])
@pytest.mark.parametrize("n,d", [(25, 2), (300, 3), (700, 5)])
def test_gp_log_prob_gradient(kind, n, d):
    """vgp_gp_logprob_grad_k against the closed form of the oracle (itself checked by finite differences)."""
    x, y = data(n, 3 * n, d=max(d, 2))
    x = x[:, :d]
    amp, ls, noise = 0.9, 0.7, 0.05
    gp = gpf.fit_gp(KINDS[kind](amp, ls), x, noise)
    ll, grads = gp.log_prob_and_grad(y)
    want_ll, want_g = gpo.gp_log_prob_grad(x, y, amp, ls, noise, kind=kind)
    assert ll[0] == pytest.approx(want_ll, rel=RTOL)
    np.testing.assert_allclose(grads[0], want_g, rtol=1e-8, atol=1e-9 * np.abs(want_g).max())


def test_tf_train_gp_adam_training_loop():
    """main.py:80-110 + gpf.tf_optimize_model_params: softplus-constrained (amplitude, length_scale, noise), Adam(0.1)
    on -log_likelihood, MaternOneHalf kernel.  The whole trajectory must follow the CPU restatement."""
    x, y = data(25, 21, d=2)                                            # main.py:418-419 uses 25 points
    amp, amp_assign, amp_p, lensc, lensc_assign, lensc_p, emb, emb_assign, emb_p, noise = \
        gpf.tf_Placeholder_assign_test(np.array([0.54]), np.array([0.54]), np.array([0.54]))
    sess = gpf.reset_session()
    obs = gpf.placeholder(np.float64, (1, 25), "obs")
    gp = gpf.fit_gp(gpf.create_cov_kernel(amp, lensc), x, noise)
    log_likelihood = gp.log_prob(obs)
    train_op = gpf.tf_train_gp_adam(log_likelihood, 0.1)
    lls = gpf.tf_optimize_model_params(sess, 30, train_op, log_likelihood, None, None, None, None, None, y, obs)
    want_lls, want_v = gpo.gp_train_adam(x, y, [0.54, 0.54, 0.54], 0.1, 30, kind="matern12")
    assert lls.shape == (31, 1)
    np.testing.assert_allclose(lls[:, 0], want_lls, rtol=1e-8)
    got_v = [amp.variable.value[0], lensc.variable.value[0], noise.variable.value[0]]
    np.testing.assert_allclose(got_v, want_v, rtol=1e-7)
    assert lls[-1, 0] > lls[0, 0]
    # sess.run([train_op, log_likelihood]) returns the value the update was computed from
    before = log_likelihood(y)
    _, ll_run = sess.run([train_op, log_likelihood], feed_dict={obs: y.reshape(1, 25)})
    assert ll_run[0] == pytest.approx(before[0], rel=1e-12)
    assert gpf.do_assign(sess, amp, amp_assign, amp_p, [1.7])[0] == pytest.approx(1.7, rel=1e-14)


def test_batch_of_two_gps():
    """INIT arrays of length 2 (main.py:80-88): two independent GPs over the same index points, observations [2, n],
    log_likelihood[0] / [1] (main.py:105-107), both trained by one train op."""
    x, y0 = data(40, 31, d=2)
    y = np.stack([y0, np.cos(x[:, 0]) + 0.05 * np.random.default_rng(1).standard_normal(40)])
    amp, _, _, lensc, _, _, _, _, _, noise = gpf.tf_Placeholder_assign_test(np.array([0.54, 0.9]), np.array([0.54, 0.3]),
                                                                            np.array([0.54, 0.2]))
    obs = gpf.placeholder(np.float64, (2, 40))
    gp = gpf.fit_gp(gpf.create_cov_kernel(amp, lensc), x, noise)
    ll = gp.log_prob(y)
    a, l, s = amp.numpy(), lensc.numpy(), noise.numpy()
    for i in range(2):
        assert ll[i] == pytest.approx(gpo.gp_log_prob(x, y[i], a[i], l[i], s[i], kind="matern12"), rel=RTOL)
    node = gp.log_prob(obs)
    assert node[1](y) == pytest.approx(ll[1], rel=1e-14)
    train_op = gpf.tf_train_gp_adam(node, 0.1)
    lls = gpf.tf_optimize_model_params(None, 5, train_op, node, None, None, None, None, None, y, obs)
    assert lls.shape == (6, 2)
    for i, v0 in enumerate([[0.54, 0.54, 0.54], [0.9, 0.3, 0.2]]):
        want, _ = gpo.gp_train_adam(x, y[i], v0, 0.1, 5, kind="matern12")
        np.testing.assert_allclose(lls[:, i], want, rtol=1e-8)


@pytest.mark.parametrize("kind", ["expquad", "matern12", "matern52"])
@pytest.mark.parametrize("n,d", [(1, 1), (25, 2), (100, 5), (127, 3), (140, 3)])
def test_batched_log_prob_sweep(kind, n, d):
    """vgp_gp_logprob_batch_k (one CTA per hyper-parameter triple for n <= 127, per-triple calls above) against the
    single-evaluation oracle."""
    import ctypes
    x, y = data(n, 5 * n + d, d=max(d, 2))
    x = x[:, :d]
    rng = np.random.default_rng(n)
    params = np.column_stack([rng.uniform(0.3, 3.0, 40), rng.uniform(0.2, 4.0, 40), rng.uniform(0.01, 0.5, 40)])
    out = np.empty(40)
    xd, yd = gpf._points(x), gpf._vector(y)
    gpf.call("vgp_gp_logprob_batch_k", 0, gpo.KERNEL_KINDS[kind], xd.ptr, n, d, yd.ptr, params.ctypes.data, 40, 1e-6,
             out.ctypes.data, None)
    want = [gpo.gp_log_prob(x, y, a, l, s, kind=kind) for a, l, s in params]
    np.testing.assert_allclose(out, want, rtol=RTOL)


def test_calc_H_is_one_launch_and_matches_the_loop():
    """gpf.calc_H over a 12 x 9 grid on 25 points (main.py:400-419 uses 160 x 160 on 25): batched path == the
    sequential statement of gp_functions.py:864-876, and the variables end where the loop leaves them."""
    x, y = data(25, 77, d=2)
    amp, amp_assign, amp_p, lensc, lensc_assign, lensc_p, _, _, _, noise = \
        gpf.tf_Placeholder_assign_test(np.array([0.54]), np.array([0.54]), np.array([0.3]))
    gp = gpf.fit_gp(gpf.create_cov_kernel(amp, lensc), x, noise)
    H = gpf.calc_H(12, 9, lensc, lensc_assign, lensc_p, amp, amp_assign, amp_p, gp.log_prob, None, None, y)
    assert float(lensc) == pytest.approx(40.0) and float(amp) == pytest.approx(40.0)
    s2 = float(noise)
    for i in range(12):
        for j in range(9):
            want = gpo.gp_log_prob(x, y, 40 * (1 + j) / 9, 40 * (1 + i) / 12, s2, kind="matern12")
            assert H[i, j] == pytest.approx(want, rel=RTOL)
    obs = gpf.placeholder(np.float64, (1, 25))
    H2 = gpf.calc_H(12, 9, lensc, lensc_assign, lensc_p, amp, amp_assign, amp_p, gp.log_prob(obs), None, obs, y)
    assert np.array_equal(H, H2)
