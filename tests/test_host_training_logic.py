"""CPU tests of the host-side training logic of the drop-in gp_functions module (no GPU: the GP is a stub with an
analytic log-likelihood): softplus parameterisation, TF-Adam update, Session.run fetch semantics, lls shapes."""
import numpy as np

from oracle import gp_oracle as gpo
import vgposp_b200.gp_functions as gpf


class _Kernel:
    KIND = 0

    def __init__(self, amp, ls):
        self.amplitude, self.length_scale = amp, ls


class _StubGP:
    """log-likelihood -sum_b |theta_b - target|^2 / 2 of the constrained parameters theta = (amp, ls, noise)."""

    def __init__(self, amp, ls, noise, target):
        self.kernel, self.observation_noise_variance, self.target = _Kernel(amp, ls), noise, np.asarray(target)

    def theta(self):
        return np.stack([np.atleast_1d(p.numpy()) for p in
                         (self.kernel.amplitude, self.kernel.length_scale, self.observation_noise_variance)], axis=1)

    def log_prob(self, observations):
        return -0.5 * np.sum((self.theta() - self.target) ** 2, axis=1)

    def log_prob_and_grad(self, observations):
        th = self.theta()
        return -0.5 * np.sum((th - self.target) ** 2, axis=1), -(th - self.target)


def test_adam_train_op_follows_tf_adam_in_softplus_space():
    amp, _, _, lensc, _, _, _, _, _, noise = gpf.tf_Placeholder_assign_test(np.array([0.54]), np.array([0.3]), np.array([1.2]))
    target = np.array([2.0, 0.7, 0.05])
    gp = _StubGP(amp, lensc, noise, target)
    obs = gpf.placeholder(np.float64, (1, 4))
    node = gpf.LogProb(gp, obs)
    op = gpf.tf_train_gp_adam(node, 0.1)
    lls = gpf.tf_optimize_model_params(gpf.reset_session(), 25, op, node, None, None, None, None, None,
                                       np.zeros(4), obs)
    # the same loop with the oracle's Adam on the unconstrained variables
    v = np.array([0.54, 0.3, 1.2])
    adam = gpo.TfAdam(3, 0.1)
    want = []
    tiny = np.finfo(np.float64).tiny
    for it in range(27):
        th = tiny + gpo.softplus(v)
        if it > 0:
            want.append(-0.5 * np.sum((th - target) ** 2))
        v = adam.step(v, (th - target) / (1.0 + np.exp(-v)))
    assert lls.shape == (26, 1)
    np.testing.assert_allclose(lls[:, 0], want, rtol=1e-12)
    got_v = [amp.variable.value[0], lensc.variable.value[0], noise.variable.value[0]]
    np.testing.assert_allclose(got_v, v, rtol=1e-12)
    assert lls[-1, 0] > lls[0, 0]


def test_session_run_fetches_and_batch_of_two():
    amp, amp_assign, amp_p, lensc, _, _, _, _, _, noise = gpf.tf_Placeholder_assign_test(
        np.array([0.54, 0.9]), np.array([0.3, 0.4]), np.array([1.2, 0.8]))
    gp = _StubGP(amp, lensc, noise, np.array([1.0, 1.0, 1.0]))
    obs = gpf.placeholder(np.float64, (2, 3))
    node = gpf.LogProb(gp, obs)
    op = gpf.tf_train_gp_adam(node, 0.05)
    sess = gpf.reset_session()
    before = node(np.zeros((2, 3)))
    _, ll, ll1 = sess.run([op, node, node[1]], feed_dict={obs: np.zeros((2, 3))})
    np.testing.assert_allclose(ll, before)                      # fetched with the train op: the pre-update value
    assert ll1 == before[1]
    assert not np.allclose(node(), before)                      # parameters moved
    lls = gpf.tf_optimize_model_params(sess, 3, op, node, None, None, None, None, None, np.zeros((2, 3)), obs)
    assert lls.shape == (4, 2)
    amp_assign([2.0, 3.0])
    np.testing.assert_allclose(amp.numpy(), [2.0, 3.0], rtol=1e-14)
    assert np.allclose(gpf.do_assign(sess, amp, amp_assign, amp_p, [0.5, 0.6]), [0.5, 0.6])


def test_likelihood_surfaces_walk_the_reference_grids(monkeypatch):
    """calc_H / calc_H_1d (gp_functions.py:864-889): (length_scale, amplitude) = span (1+i)/X, span (1+j)/Y with span 40
    and 2; calc_H feeds the observations, calc_H_1d evaluates the node as it stands."""
    class FakeGP:
        def __init__(self):
            self.calls = []

        def log_prob(self, obs):
            self.calls.append((float(lensc), float(amp), None if obs is None else np.asarray(obs).copy()))
            return np.array([float(lensc) * 100 + float(amp)])

    amp, amp_assign, amp_p, lensc, lensc_assign, lensc_p, _, _, _, _ = \
        gpf.tf_Placeholder_assign_test(np.array([0.54]), np.array([0.54]), np.array([0.1]))
    gp = FakeGP()
    obs = gpf.placeholder(np.float64, (1, 3))
    node = gpf.LogProb(gp, obs)
    y = np.array([1.0, 2.0, 3.0])
    H = gpf.calc_H(3, 2, lensc, lensc_assign, lensc_p, amp, amp_assign, amp_p, node, None, obs, y)
    want = np.array([[40 * (1 + i) / 3 * 100 + 40 * (1 + j) / 2 for j in range(2)] for i in range(3)])
    np.testing.assert_allclose(H, want, rtol=1e-12)
    assert all(np.array_equal(c[2], y) for c in gp.calls)
    gp.calls.clear()
    obs.value = y
    H1 = gpf.calc_H_1d(4, 3, lensc, lensc_assign, lensc_p, amp, amp_assign, amp_p, node, None)
    want1 = np.array([[2 * (1 + i) / 4 * 100 + 2 * (1 + j) / 3 for j in range(3)] for i in range(4)])
    np.testing.assert_allclose(H1, want1, rtol=1e-12)
    assert len(gp.calls) == 12 and all(np.array_equal(c[2], y) for c in gp.calls)


def test_vgp_input_rows_for_one_location():
    """gp_functions.py:995-1016: columns (x, y, z, t, p) with t = vec_pt[:, 1], p = vec_pt[:, 0]."""
    vec_pt = np.array([[10.0, 1.0], [20.0, 2.0], [30.0, 3.0]])
    got = gpf.graph_get_vgp_input_xyztp([0.5, -1.0, 2.0], vec_pt)
    want = np.array([[0.5, -1.0, 2.0, 1.0, 10.0], [0.5, -1.0, 2.0, 2.0, 20.0], [0.5, -1.0, 2.0, 3.0, 30.0]])
    assert got.dtype == np.float64
    np.testing.assert_array_equal(got, want)
