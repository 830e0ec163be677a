"""Drop-in for the hot-path part of the reference's `gp_functions.py`, computing on a B200 through libvgposp.so.

The reference builds TF1 graph nodes out of TensorFlow-Probability objects and runs them in a session; here
the same names return small eager objects whose numeric methods call the CUDA library:

    reference (gp_functions.py)                          here
    ---------------------------------------------------  -------------------------------------------------
    create_cov_kernel(amp, lensc)            :160-163    MaternOneHalf(amp, lensc) (also ExponentiatedQuadratic,
                                                         MaternThreeHalves, MaternFiveHalves)
    fit_gp(kernel, idx_pts, noise_var)       :166-172    GaussianProcess(...).log_prob(y | placeholder)
    tf_train_gp_adam(log_likelihood, lr)     :179-182    AdamTrainOp: device log-prob gradient + TF-Adam in softplus space
    tf_optimize_model_params(...)            :228-259    the training loop, returns lls
    reset_session / do_assign                :112-121,222-225   Session stand-in
    create_meshgrid, slice_grid_xyz, py_get_coord_idxs, denormalize_coord, create_cov_matrix  :262,1215-1247,1019
    tf_gp_regression_model(...)              :283-297    GaussianProcessRegressionModel(...).mean()/.variance()
    tf_Variable / invert_softplus / tf_Placeholder_assign_test :48-65,106-109,124-149   Positive parameters
    calc_H(...)                              :864-876    likelihood surface by re-evaluating log_prob
    placement_algorithm_1/2, nominator, ...  :576-685    forwarded to vgposp_b200.placement_algorithm2

and, for the variational path (variational_Gaussian_process_example.py:68-99),
`VariationalGaussianProcess.optimal_variational_posterior / .variational_loss / .mean / .variance`.

Notes
  * `create_cov_kernel` returns MaternOneHalf like the reference (gp_functions.py:162); the ExpQuad kernel named by
    the north star is `ExponentiatedQuadratic`.  The hand-derived ELBO training step (VgpTrainer) is ExpQuad-only.
  * gp_functions.py carries an older copy of the greedy whose `nominator` appends y to A (:653-661) and therefore
    degenerates to [0, 1, 2, ...]; the names here forward to the corrected semantics of placement_algorithm2.py.
  * TF1 graph tensors cannot be exported through DLPack; eager tensors (TF >= 2.2: tf.experimental.dlpack.to_dlpack)
    or anything else speaking `__dlpack__` are consumed zero-copy when they already live on the device.
  * No CPU fallback: numeric methods raise without a GPU.
"""
import ctypes

import numpy as np

from . import _ffi
from ._ffi import call
from .placement_algorithm2 import (argmax_, call_pinv, calculate_running_time_algorithm_1,  # noqa: F401
                                   calculate_running_time_algorithm_2, denominator, make_slice, nominator,
                                   placement_algorithm_1, placement_algorithm_2)

DEVICE = 0
DEFAULT_JITTER = 1e-6
TINY = float(np.finfo(np.float64).tiny)


# --------------------------------------------------------------------------------------------------
# positive parameters (gp_functions.py:106-109,124-149)
# --------------------------------------------------------------------------------------------------
def softplus(v):
    return np.logaddexp(0.0, np.asarray(v, dtype=np.float64))


def softplus_inverse(p):
    return np.log(np.expm1(np.asarray(p, dtype=np.float64)))


class Variable:
    """Unconstrained trainable value (the reference's tf.Variable)."""

    def __init__(self, initial_value, name=None):
        self.value = np.array(initial_value, dtype=np.float64)
        self.name = name

    def assign(self, value):
        self.value = np.array(value, dtype=np.float64)
        return self.value


class Positive:
    """`offset + softplus(variable)` -- the constrained view the kernel consumes (gp_functions.py:134)."""

    def __init__(self, variable, offset=TINY):
        self.variable, self.offset = variable, offset

    def numpy(self):
        return self.offset + softplus(self.variable.value)

    def __float__(self):
        return float(np.asarray(self.numpy()).reshape(-1)[0])


def tf_Variable(FEATURE_s, FEATURE_n, INIT):
    """(variable, tiny + softplus(variable)) (gp_functions.py:124-135)."""
    var = Variable(INIT, FEATURE_n)
    return var, Positive(var)


def invert_softplus(place_holder, variable, name="assign_op"):
    """Assign `variable` so that softplus(variable) == place_holder (gp_functions.py:106-109)."""
    return variable.assign(softplus_inverse(place_holder))


class Placeholder:
    def __init__(self, shape, name=None, dtype=np.float64):
        self.shape, self.name, self.value = tuple(shape), name, None


def placeholder(dtype=np.float64, shape=(), name=None):
    """tf.placeholder's role for observation feeds (main.py: obs_value_placeholder)."""
    return Placeholder(shape, name)


class AssignOp:
    """Callable standing for the reference's `feature_assign` graph op: run it with the placeholder's value."""

    def __init__(self, variable, placeholder):
        self.variable, self.placeholder = variable, placeholder
        self.shape = placeholder.shape

    def __call__(self, value):
        return invert_softplus(value, self.variable)


def tf_Placeholder_assignments(feature_var, FEATUREPLH_s, FEATUREPLH_n, INIT):
    plh = Placeholder(np.shape(INIT), FEATUREPLH_n)
    return plh, AssignOp(feature_var, plh)


def tf_Placeholder_assign_test(AMPLITUDE_INIT, LENGTHSCALE_INIT, INIT_OBSNOISEVAR):
    """The 10-tuple of gp_functions.py:48-65."""
    AMPLITUDE_INIT, LENGTHSCALE_INIT = np.asarray(AMPLITUDE_INIT), np.asarray(LENGTHSCALE_INIT)
    amp_var, amp = tf_Variable("amplitude", "amplitude", AMPLITUDE_INIT)
    amp_plh, amp_assign = tf_Placeholder_assignments(amp_var, "amplitude_assign", "amplitude_assign", AMPLITUDE_INIT)
    lensc_var, lensc = tf_Variable("lengthscale", "lengthscale", LENGTHSCALE_INIT)
    lensc_plh, lensc_assign = tf_Placeholder_assignments(lensc_var, "lengthscale_assign", "lengthscale_assign",
                                                         LENGTHSCALE_INIT)
    _, obs_noise_var = tf_Variable("observation_noise_variance", "observation_noise_variance", INIT_OBSNOISEVAR)
    emb_var, emb = tf_Variable("log_probability_embedding", "log_probability_embedding", LENGTHSCALE_INIT)
    emb_plh, emb_assign = tf_Placeholder_assignments(emb_var, "log_probability_embedding_assign",
                                                     "log_probability_embedding_assign", LENGTHSCALE_INIT)
    assert amp_assign.shape == AMPLITUDE_INIT.shape
    assert lensc_assign.shape == LENGTHSCALE_INIT.shape
    assert amp_assign.shape == lensc_assign.shape
    return amp, amp_assign, amp_plh, lensc, lensc_assign, lensc_plh, emb, emb_assign, emb_plh, obs_noise_var


def _values(v):
    return np.asarray(v.numpy() if hasattr(v, "numpy") else v, dtype=np.float64).reshape(-1)


def _size(v):
    return int(_values(v).size)


def _scalar(v, index=0):
    """Element `index` of a parameter (length-1 parameters broadcast over a batch of GPs, Appendix A.4)."""
    a = _values(v)
    return float(a[index if a.size > 1 else 0])


def _points(x):
    """Index points as a device float64 [n, d] array (1-D input becomes [n, 1])."""
    if isinstance(x, _ffi.DeviceArray):
        return x
    if hasattr(x, "__dlpack__") and not isinstance(x, np.ndarray):
        d = _ffi.as_device_f64(x, DEVICE)
        if len(d.shape) == 1:
            d.shape = (d.shape[0], 1)
        return d
    a = np.asarray(x, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    return _ffi.DeviceArray.from_host(a, DEVICE)


def _vector(y):
    if isinstance(y, _ffi.DeviceArray):
        return y
    if hasattr(y, "__dlpack__") and not isinstance(y, np.ndarray):
        return _ffi.as_device_f64(y, DEVICE)
    return _ffi.DeviceArray.from_host(np.asarray(y, dtype=np.float64).reshape(-1), DEVICE)


# --------------------------------------------------------------------------------------------------
# kernel (a1)
# --------------------------------------------------------------------------------------------------
class _StationaryKernel:
    """Common part of the stationary kernels; `KIND` is the VGP_KERNEL_* code of include/vgposp.h."""
    KIND = 0

    def __init__(self, amplitude, length_scale, feature_ndims=1):
        assert feature_ndims == 1
        self.amplitude, self.length_scale = amplitude, length_scale

    def params(self, index=0):
        return _scalar(self.amplitude, index), _scalar(self.length_scale, index)

    def batch_size(self):
        return max(_size(self.amplitude), _size(self.length_scale))

    def matrix_device(self, x1, x2, diag_add=0.0):
        a, l = self.params()
        d1, d2 = _points(x1), (None if x2 is x1 else _points(x2))
        d2 = d1 if d2 is None else d2
        n1, n2, d = d1.shape[0], d2.shape[0], d1.shape[1]
        assert d2.shape[1] == d, "feature dimensions differ"
        ld = n2 + (n2 % 2)
        out = _ffi.DeviceArray((n1, ld), np.float64, DEVICE)
        call("vgp_kernel_matrix", DEVICE, self.KIND, d1.ptr, n1, d2.ptr, n2, d, a, l, float(diag_add), 0, out.ptr, ld,
             None)
        return out, n2

    def matrix(self, x1, x2):
        out, n2 = self.matrix_device(x1, x2)
        return out.to_host()[:, :n2]

    def apply(self, x1, x2):
        """k of two single points."""
        return self.matrix(np.asarray(x1, dtype=np.float64).reshape(1, -1),
                           np.asarray(x2, dtype=np.float64).reshape(1, -1))[0, 0]

    _apply = apply


class ExponentiatedQuadratic(_StationaryKernel):
    """k(x, y) = amplitude^2 exp(-|x - y|^2 / (2 length_scale^2)) -- tfkern.ExponentiatedQuadratic's role at
    variational_Gaussian_process_example.py:55-57, 3D_sin_wave.py:158-159."""
    KIND = 0


class MaternOneHalf(_StationaryKernel):
    """amplitude^2 exp(-|x - y| / length_scale) -- tfkern.MaternOneHalf (main.py:94, gp_functions.py:162)."""
    KIND = 1


class MaternThreeHalves(_StationaryKernel):
    """amplitude^2 (1 + z) exp(-z), z = sqrt(3) |x - y| / length_scale -- tfkern.MaternThreeHalves."""
    KIND = 2


class MaternFiveHalves(_StationaryKernel):
    """amplitude^2 (1 + z + z^2/3) exp(-z), z = sqrt(5) |x - y| / length_scale -- tfkern.MaternFiveHalves
    (main_architecture_2.py:184)."""
    KIND = 3


def create_cov_kernel(amp, lensc):
    """gp_functions.py:160-163: the reference returns MaternOneHalf (ExponentiatedQuadratic is left in its comment)."""
    return MaternOneHalf(amp, lensc)


# --------------------------------------------------------------------------------------------------
# exact GP (a2, a3)
# --------------------------------------------------------------------------------------------------
class GaussianProcess:
    """tfd.GaussianProcess's role in fit_gp (gp_functions.py:166-172): zero mean, stationary kernel.

    Parameters of length 2 (the reference's INIT arrays of shape [2], main.py:80-88) make a batch of two independent
    GPs over the same index points; `log_prob` of observations [2, n] then returns both values (main.py:105-107)."""

    def __init__(self, kernel, index_points, observation_noise_variance=0.0, jitter=DEFAULT_JITTER,
                 validate_args=False):
        self.kernel, self.jitter = kernel, jitter
        self.observation_noise_variance = observation_noise_variance
        self._x = _points(index_points)

    def batch_size(self):
        return max(self.kernel.batch_size(), _size(self.observation_noise_variance))

    def _rows(self, observations):
        n = self._x.shape[0]
        y = np.asarray(observations.to_host() if isinstance(observations, _ffi.DeviceArray) else observations,
                       dtype=np.float64)
        assert y.size % n == 0, "observations must have one value per index point"
        return y.reshape(-1, n)

    def log_prob(self, observations):
        """float for one GP, ndarray [b] for a batch; a `Placeholder` gives a lazy `LogProb` node to run with a feed
        (the reference's graph tensor)."""
        if isinstance(observations, Placeholder):
            return LogProb(self, observations)
        on_device = isinstance(observations, _ffi.DeviceArray) or \
            (hasattr(observations, "__dlpack__") and not isinstance(observations, np.ndarray))
        if on_device and self.batch_size() == 1:
            rows = [observations]                            # consumed where it lies (DLPack / DeviceArray)
        else:
            rows = self._rows(observations)
        b = max(self.batch_size(), len(rows))
        n, d = self._x.shape
        out = np.zeros(b)
        for i in range(b):
            a, l = self.kernel.params(i)
            y = _vector(rows[i if len(rows) > 1 else 0])
            assert y.size == n, "observations must have one value per index point"
            v = ctypes.c_double()
            call("vgp_gp_logprob_k", DEVICE, self.kernel.KIND, self._x.ptr, n, d, y.ptr, a, l,
                 _scalar(self.observation_noise_variance, i), float(self.jitter), ctypes.byref(v), None)
            out[i] = v.value
        return float(out[0]) if b == 1 else out

    def log_prob_and_grad(self, observations):
        """(log_prob [b], gradient [b, 3] with respect to the constrained (amplitude, length_scale, noise variance))."""
        rows = self._rows(observations)
        b = max(self.batch_size(), len(rows))
        n, d = self._x.shape
        out, grads = np.zeros(b), np.zeros((b, 3))
        for i in range(b):
            a, l = self.kernel.params(i)
            y = _vector(rows[i if len(rows) > 1 else 0])
            v = ctypes.c_double()
            call("vgp_gp_logprob_grad_k", DEVICE, self.kernel.KIND, self._x.ptr, n, d, y.ptr, a, l,
                 _scalar(self.observation_noise_variance, i), float(self.jitter), ctypes.byref(v),
                 grads[i].ctypes.data, None)
            out[i] = v.value
        return out, grads


class LogProb:
    """`gp.log_prob(placeholder)`: the reference's `log_likelihood` graph tensor.  Call it with the observations
    (or `.run({placeholder: value})`); index it (`log_likelihood[0]`) for one element of a batch."""

    def __init__(self, gp, placeholder, index=None):
        self.gp, self.placeholder, self.index = gp, placeholder, index

    def __call__(self, observations=None):
        if observations is None:
            observations = self.placeholder.value
        v = np.atleast_1d(self.gp.log_prob(observations))
        return v if self.index is None else v[self.index]

    def run(self, feed_dict=None):
        return self(None if not feed_dict else feed_dict.get(self.placeholder))

    def __getitem__(self, index):
        return LogProb(self.gp, self.placeholder, index)


class AdamTrainOp:
    """`tf.train.AdamOptimizer(lr).minimize(-log_likelihood)` (gp_functions.py:179-182) on the softplus-unconstrained
    variables behind the GP's amplitude, length scale and noise variance (TF defaults beta1 .9, beta2 .999, eps 1e-8;
    TF's update lr_t = lr sqrt(1 - b2^t) / (1 - b1^t), v -= lr_t m / (sqrt(v) + eps)).  `run(observations)` evaluates
    log-likelihood and gradient on the device at the current parameters, applies one update, and returns the
    log-likelihood(s) of before the update -- what `sess.run([train_op, log_likelihood], feed)` yields."""

    def __init__(self, node, learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8):
        assert isinstance(node, LogProb), "tf_train_gp_adam expects gp.log_prob(placeholder)"
        self.node, self.lr, self.b1, self.b2, self.eps = node, float(learning_rate), beta1, beta2, epsilon
        gp = node.gp
        self.params = [gp.kernel.amplitude, gp.kernel.length_scale, gp.observation_noise_variance]
        self.m = [np.zeros_like(_unconstrained(p)) if isinstance(p, Positive) else None for p in self.params]
        self.v = [np.zeros_like(_unconstrained(p)) if isinstance(p, Positive) else None for p in self.params]
        self.t = 0

    def run(self, observations=None):
        if observations is None:
            observations = self.node.placeholder.value
        ll, grads = self.node.gp.log_prob_and_grad(observations)
        self.t += 1
        lr_t = self.lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for q, p in enumerate(self.params):
            if not isinstance(p, Positive):
                continue                                     # a constant: not trainable
            u = _unconstrained(p)
            g = grads[:, q] if u.size > 1 else np.array([grads[:, q].sum()])
            g = -g.reshape(u.shape) * _sigmoid(u)            # d(-ll)/du = -(dll/dtheta) softplus'(u)
            self.m[q] = self.b1 * self.m[q] + (1.0 - self.b1) * g
            self.v[q] = self.b2 * self.v[q] + (1.0 - self.b2) * g * g
            p.variable.assign(u - lr_t * self.m[q] / (np.sqrt(self.v[q]) + self.eps))
        return ll


def _unconstrained(p):
    return np.atleast_1d(np.asarray(p.variable.value, dtype=np.float64))


def _sigmoid(u):
    return 1.0 / (1.0 + np.exp(-u))


def fit_gp(kernel, obs_idx_pts, obs_noise_var):
    return GaussianProcess(kernel=kernel, index_points=obs_idx_pts, observation_noise_variance=obs_noise_var,
                           validate_args=True)


def tf_train_gp_adam(feature, LEARNING_RATE):
    """gp_functions.py:179-182: Adam on -feature, feature = gp.log_prob(placeholder)."""
    return AdamTrainOp(feature, LEARNING_RATE)


def do_assign(sess, feature, feature_assign, feature_plh, feature_arr):
    """gp_functions.py:222-225: run the assign op with the placeholder's value, return the constrained feature."""
    feature_assign(feature_arr)
    return np.asarray(feature.numpy())


class Session:
    """Stand-in for the reference's module-global tf.InteractiveSession (gp_functions.py:112-121): `run` evaluates
    the eager nodes of this module (LogProb, AdamTrainOp, AssignOp, Positive) with a feed dict."""

    def run(self, fetches, feed_dict=None):
        feed_dict = feed_dict or {}
        for plh, value in feed_dict.items():
            plh.value = value
        single = not isinstance(fetches, (list, tuple))
        out = []
        for f in ([fetches] if single else fetches):
            if isinstance(f, AdamTrainOp):
                f.last = f.run()
                out.append(None)
            elif isinstance(f, LogProb):
                out.append(f())
            elif isinstance(f, AssignOp):
                out.append(f(f.placeholder.value))
            elif hasattr(f, "numpy"):
                out.append(np.asarray(f.numpy()))
            else:
                out.append(f)
        # a train op fetched together with its log-likelihood: both come from the same evaluation (pre-update)
        ops = [f for f in ([fetches] if single else fetches) if isinstance(f, AdamTrainOp)]
        if ops and not single:
            for i, f in enumerate(fetches):
                if isinstance(f, LogProb) and f.gp is ops[0].node.gp:
                    v = np.atleast_1d(ops[0].last)
                    out[i] = v if f.index is None else v[f.index]
        return out[0] if single else out

    def close(self):
        pass


sess = None


def reset_session():
    """gp_functions.py:112-121: a fresh global session."""
    global sess
    sess = Session()
    return sess


def tf_optimize_model_params(sess, num_iters, train_op, log_likelihood, summ=None, writer=None, saver=None,
                             LOGDIR=None, LOGCHECKPT=None, obs_train_dataset=None, obs_value_placeholder=None):
    """gp_functions.py:228-259: one initial run, then num_iters + 1 training steps; returns the log-likelihoods
    `lls` [num_iters + 1, 1 or 2] (of before each update).  Summaries / checkpoints (`summ`, `writer`, `saver`) are
    TensorBoard plumbing outside the hot path: when given they are called the way the reference calls them."""
    data = np.asarray(obs_train_dataset, dtype=np.float64)
    cols = 2 if data.ndim == 2 else 1
    lls = np.zeros([num_iters + 1, cols], np.float64)
    feed = data.reshape(obs_value_placeholder.shape) if obs_value_placeholder is not None else data
    train_op.run(feed)                                               # :245-247, the initial run
    for i in range(num_iters + 1):
        lls[i] = np.atleast_1d(train_op.run(feed))[:cols]
        if writer is not None and summ is not None:
            writer.add_summary(summ, i)
        if saver is not None and i % 200 == 0:
            import os
            saver.save(sess, os.path.join(LOGDIR, LOGCHECKPT), i)
    return lls


class GaussianProcessRegressionModel:
    """tfd.GaussianProcessRegressionModel's role in tf_gp_regression_model (gp_functions.py:283-297)."""

    def __init__(self, kernel, index_points, observation_index_points, observations,
                 observation_noise_variance=0.0, predictive_noise_variance=0.0, divisor_jitter=0.0):
        self.kernel = kernel
        self._xt, self._x = _points(index_points), _points(observation_index_points)
        self._y = _vector(observations)
        self.observation_noise_variance = observation_noise_variance
        self.predictive_noise_variance = predictive_noise_variance
        self.divisor_jitter = divisor_jitter
        self._cache = None

    def _run(self):
        a, l = self.kernel.params()
        n, d = self._x.shape
        t = self._xt.shape[0]
        mean = _ffi.DeviceArray((t,), np.float64, DEVICE)
        var = _ffi.DeviceArray((t,), np.float64, DEVICE)
        call("vgp_gp_regression_k", DEVICE, self.kernel.KIND, self._x.ptr, n, d, self._y.ptr, self._xt.ptr, t, a, l,
             _scalar(self.observation_noise_variance), _scalar(self.predictive_noise_variance),
             float(self.divisor_jitter), mean.ptr, var.ptr, None)
        self._cache = (mean.to_host(), var.to_host())
        return self._cache

    def mean(self):
        return self._run()[0]

    def variance(self):
        return self._run()[1]

    def stddev(self):
        return np.sqrt(self.variance())


def tf_gp_regression_model(kernel, pred_idx_pts, obs_idx_pts, obs, obs_noise_var, pred_noise_var):
    obs = np.asarray(obs.numpy() if hasattr(obs, "numpy") else obs, dtype=np.float64).reshape(-1)
    return GaussianProcessRegressionModel(kernel=kernel, index_points=pred_idx_pts,
                                          observation_index_points=obs_idx_pts, observations=obs,
                                          observation_noise_variance=obs_noise_var,
                                          predictive_noise_variance=pred_noise_var)


def create_meshgrid(pred_x, pred_y):
    """gp_functions.py:262-280: predictive index points [len(pred_x) * len(pred_y), 2]."""
    h = np.array(np.meshgrid(pred_x, pred_y, sparse=False))
    return h.swapaxes(0, -1).reshape(-1, 2)


def slice_grid_xyz(i0, i1, i2, linsp_x, linsp_y, linsp_z):
    """gp_functions.py:1215-1223: the coordinate triple at grid indices (i0, i1, i2)."""
    return np.array([np.asarray(linsp_x)[int(i0)], np.asarray(linsp_y)[int(i1)], np.asarray(linsp_z)[int(i2)]])


def py_get_coord_idxs(sel_idx, xyz_idxs):
    """gp_functions.py:1226-1237: grid-index rows of the first 7 selected locations, in selection order."""
    xyz_idxs = np.asarray(xyz_idxs)
    return np.vstack([xyz_idxs[sel_idx[i], :].reshape(1, 3) for i in range(7)])


def denormalize_coord(sel_norm_coord):
    """gp_functions.py:1239-1247 (in place, like the reference)."""
    stdev_var = 0.0007434639347162126 * 3000000
    mean_var = 0.0018159087825037148
    sel_norm_coord[:, :3] = sel_norm_coord[:, :3] * stdev_var + mean_var
    return sel_norm_coord


def graph_get_vgp_input_xyztp(loc_xyz, vec_pt):
    """[S, 5] VGP inputs (x, y, z, t, p) for one location and S (pressure, temperature) samples: the location repeated,
    then column 1 of `vec_pt`, then column 0 (gp_functions.py:995-1016; call site main_architecture_2.py:341)."""
    vec_pt = np.asarray(vec_pt, dtype=np.float64)
    out = np.empty((vec_pt.shape[0], 5))
    out[:, 0], out[:, 1], out[:, 2] = float(loc_xyz[0]), float(loc_xyz[1]), float(loc_xyz[2])
    out[:, 3] = vec_pt[:, 1]
    out[:, 4] = vec_pt[:, 0]
    return out


def create_cov_matrix(minmax_x, minmax_y, minmax_z, minmax_pressure, minmax_temperature, SPATIAL_COVER,
                      SPATIAL_COV_PR_TEMP, encoder, sess=None):
    """gp_functions.py:1019-1057 (see vgposp_b200.cov_producer.create_cov_matrix)."""
    from .cov_producer import create_cov_matrix as impl
    return impl(minmax_x, minmax_y, minmax_z, minmax_pressure, minmax_temperature, SPATIAL_COVER,
                SPATIAL_COV_PR_TEMP, encoder, sess)


def create_cov_matrix_while_loops(minmax_x, minmax_y, minmax_z, minmax_pressure, minmax_temperature, SPATIAL_COVER,
                                  SPATIAL_COV_PR_TEMP, encoder, sess=None):
    """gp_functions.py:1060-1117: the same matrix as create_cov_matrix written with while loops."""
    return create_cov_matrix(minmax_x, minmax_y, minmax_z, minmax_pressure, minmax_temperature, SPATIAL_COVER,
                             SPATIAL_COV_PR_TEMP, encoder, sess)


def tf_create_cov_matrix(mm_x, mm_y, mm_z, mm_p, mm_t, sptl_const, sptl_pt_const, encoder, sess=None):
    """gp_functions.py:1120-1179: again the same matrix (the reference's third copy of the loop nest)."""
    return create_cov_matrix(mm_x, mm_y, mm_z, mm_p, mm_t, sptl_const, sptl_pt_const, encoder, sess)


def calc_H(XEDGES, YEDGES, lensc, lensc_assign, lensc_p, amp, amp_assign, amp_p, log_likelihood, sess=None,
           obs_values_placeholder=None, obs_train_dataset=None):
    """Likelihood surface over (length_scale, amplitude) = 40 (1+i)/X, 40 (1+j)/Y (gp_functions.py:864-876):
    H[i, j] = log_likelihood[0] with both parameters assigned.  `log_likelihood` is `gp.log_prob` (bound method) or
    the `gp.log_prob(placeholder)` node; `sess` is accepted and ignored.

    The reference issues X * Y session runs, each a kernel build + Cholesky (25 600 of them at main.py:400-401).  When
    the GP is known and has at most 127 observations the whole sweep is ONE launch -- one CTA per grid point, kernel
    matrix, Cholesky and solve in registers (vgp_gp_logprob_batch_k); otherwise the evaluations run one by one."""
    return _likelihood_surface(40.0, XEDGES, YEDGES, lensc, lensc_assign, amp, amp_assign, log_likelihood,
                               obs_train_dataset)


def calc_H_1d(XEDGES, YEDGES, lensc, lensc_assign, lensc_p, amp, amp_assign, amp_p, log_likelihood, sess=None):
    """The same sweep on (length_scale, amplitude) = 2 (1+i)/X, 2 (1+j)/Y with the observations already bound to the
    `log_likelihood` node (gp_functions.py:879-889: nothing but the two parameters is fed)."""
    y = log_likelihood.placeholder.value if isinstance(log_likelihood, LogProb) else None
    return _likelihood_surface(2.0, XEDGES, YEDGES, lensc, lensc_assign, amp, amp_assign, log_likelihood, y)


def _likelihood_surface(span, XEDGES, YEDGES, lensc, lensc_assign, amp, amp_assign, log_likelihood, obs_train_dataset):
    H = np.zeros([XEDGES, YEDGES])
    y = None if obs_train_dataset is None else np.asarray(obs_train_dataset, dtype=np.float64)
    gp = log_likelihood.gp if isinstance(log_likelihood, LogProb) else getattr(log_likelihood, "__self__", None)
    if isinstance(gp, GaussianProcess) and y is not None and gp._x.shape[0] <= 127 and \
            gp.kernel.amplitude is amp and gp.kernel.length_scale is lensc:
        n, d = gp._x.shape
        y0 = _vector(gp._rows(y)[0])
        ls_grid = span * (1.0 + np.arange(XEDGES)) / XEDGES
        amp_grid = span * (1.0 + np.arange(YEDGES)) / YEDGES
        params = np.empty((XEDGES, YEDGES, 3))
        params[:, :, 0] = amp_grid[None, :]
        params[:, :, 1] = ls_grid[:, None]
        params[:, :, 2] = _scalar(gp.observation_noise_variance)
        out = np.empty(XEDGES * YEDGES)
        call("vgp_gp_logprob_batch_k", DEVICE, gp.kernel.KIND, gp._x.ptr, n, d, y0.ptr, params.ctypes.data,
             XEDGES * YEDGES, float(gp.jitter), out.ctypes.data, None)
        lensc_assign(np.full(np.shape(_values(lensc)), ls_grid[-1]))         # the state the reference's loop leaves
        amp_assign(np.full(np.shape(_values(amp)), amp_grid[-1]))
        return out.reshape(XEDGES, YEDGES)
    for i in range(XEDGES):
        for j in range(YEDGES):
            lensc_assign([span * np.double((1 + i) / XEDGES)])
            amp_assign([span * np.double((1 + j) / YEDGES)])
            H[i, j] = np.atleast_1d(log_likelihood(y) if y is not None else log_likelihood())[0]      # :875, L[0]
    return H


# --------------------------------------------------------------------------------------------------
# variational GP (a4, a5)
# --------------------------------------------------------------------------------------------------
class VariationalGaussianProcess:
    """tfd.VariationalGaussianProcess's role at variational_Gaussian_process_example.py:83-99."""

    def __init__(self, kernel, index_points, inducing_index_points, variational_inducing_observations_loc,
                 variational_inducing_observations_scale, observation_noise_variance=0.0,
                 predictive_noise_variance=0.0, jitter=DEFAULT_JITTER):
        self.kernel, self.jitter = kernel, jitter
        self._xt = None if index_points is None else _points(index_points)
        self._z = _points(inducing_index_points)
        self._loc = _vector(variational_inducing_observations_loc)
        m = self._z.shape[0]
        sc = variational_inducing_observations_scale
        self._scale = sc if isinstance(sc, _ffi.DeviceArray) else \
            _ffi.DeviceArray.from_host(np.asarray(sc, dtype=np.float64).reshape(m, m), DEVICE)
        self.observation_noise_variance = observation_noise_variance
        self.predictive_noise_variance = predictive_noise_variance

    @staticmethod
    def optimal_variational_posterior(kernel, inducing_index_points, observation_index_points, observations,
                                      observation_noise_variance, jitter=DEFAULT_JITTER,
                                      legacy_scale_orientation=False, as_device=False):
        """(loc [m], scale [m, m]) of the Titsias optimum (variational_Gaussian_process_example.py:68-74)."""
        a, l = kernel.params()
        z, x, y = _points(inducing_index_points), _points(observation_index_points), _vector(observations)
        m, d = z.shape
        loc = _ffi.DeviceArray((m,), np.float64, DEVICE)
        scale = _ffi.DeviceArray((m, m), np.float64, DEVICE)
        call("vgp_vgp_optimal_posterior_k", DEVICE, kernel.KIND, z.ptr, m, x.ptr, x.shape[0], d, y.ptr, a, l,
             _scalar(observation_noise_variance), float(jitter), 1 if legacy_scale_orientation else 0, loc.ptr,
             scale.ptr, None)
        return (loc, scale) if as_device else (loc.to_host(), scale.to_host())

    def variational_loss(self, observations, observation_index_points=None, kl_weight=1.0, return_terms=False):
        a, l = self.kernel.params()
        xb = self._xt if observation_index_points is None else _points(observation_index_points)
        yb = _vector(observations)
        m, d = self._z.shape
        terms = _ffi.VgpTerms()
        call("vgp_vgp_loss_k", DEVICE, self.kernel.KIND, self._z.ptr, m, d, self._loc.ptr, self._scale.ptr, xb.ptr,
             yb.ptr, xb.shape[0],
             a, l, _scalar(self.observation_noise_variance), float(kl_weight), float(self.jitter),
             ctypes.byref(terms), None)
        if return_terms:
            return {k: getattr(terms, k) for k, _ in terms._fields_}
        return terms.loss

    def _predict(self, want_var):
        a, l = self.kernel.params()
        m, d = self._z.shape
        t = self._xt.shape[0]
        mean = _ffi.DeviceArray((t,), np.float64, DEVICE)
        var = _ffi.DeviceArray((t,), np.float64, DEVICE) if want_var else None
        call("vgp_vgp_predict_k", DEVICE, self.kernel.KIND, self._z.ptr, m, d, self._loc.ptr, self._scale.ptr,
             self._xt.ptr, t, a, l,
             _scalar(self.predictive_noise_variance), float(self.jitter), mean.ptr, var.ptr if want_var else None,
             None)
        return mean.to_host(), (var.to_host() if want_var else None)

    def mean(self):
        return self._predict(False)[0]

    def variance(self):
        return self._predict(True)[1]

    def stddev(self):
        return np.sqrt(self.variance())


# --------------------------------------------------------------------------------------------------
# VGP training (a6, a7 for the variational path)
# --------------------------------------------------------------------------------------------------
class VgpTrainer:
    """The training graph of variational_Gaussian_process_example.py:46-104 as one object:

        trainer = VgpTrainer(x_train, y_train, inducing_init, batch_size)      # graph(...)
        loss = trainer.step(x_batch, y_batch)                                  # sess.run([train_op, loss], feed)

    amplitude = softplus(v_a), length_scale = 1e-5 + softplus(v_l), noise = softplus(v_s), all initialised at
    v = 0.54 (:47-61); (loc, scale) follow the parameters through `optimal_variational_posterior` over the full
    training set (:68-74); Adam(0.01) with TF's update rule (:101-102).  Loss, gradient and update run on device.
    """

    def __init__(self, x_train, y_train, inducing_index_points, batch_size, v_amplitude=0.54, v_length_scale=0.54,
                 v_noise=0.54, length_scale_offset=1e-5, jitter=DEFAULT_JITTER, learning_rate=0.01, allreduce=None,
                 n_total=None, kernel=None):
        """`kernel`: the kernel CLASS of the model -- ExponentiatedQuadratic (default, the example script),
        MaternFiveHalves (main_architecture_2.py:184: the VGP over 5-D (x, y, z, t, p) inputs) or MaternThreeHalves.
        `allreduce(device_view)`: sum-all-reduce over the ranks of a 1-D float64 device buffer, in place and in stream
        order (e.g. `lambda t: torch.distributed.all_reduce(t)`); then (x_train, y_train) is THIS rank's slice of the
        observations, `n_total` their number over all ranks, and every rank must feed the same minibatches."""
        self._x, self._y = _points(x_train), _vector(y_train)          # kept alive: the handle borrows them
        z = np.ascontiguousarray(np.asarray(inducing_index_points, dtype=np.float64))
        if z.ndim == 1:
            z = z[:, None]
        self.n, self.d = self._x.shape
        self.m, self.batch = z.shape[0], int(batch_size)
        assert z.shape[1] == self.d
        h = ctypes.c_void_p()
        call("vgp_elbo_create", ctypes.byref(h), DEVICE, self._x.ptr, self._y.ptr, self.n, self.d, z.ctypes.data,
             self.m, self.batch, float(v_amplitude), float(v_length_scale), float(v_noise),
             float(length_scale_offset), float(jitter), float(learning_rate))
        self.handle = h.value
        self.kernel_class = ExponentiatedQuadratic if kernel is None else kernel
        if self.kernel_class.KIND != 0:
            call("vgp_elbo_set_kernel", self.handle, self.kernel_class.KIND)
        self.length_scale_offset, self.jitter = length_scale_offset, jitter
        self._cb = None
        if allreduce is not None:
            views = {}

            def _cb(ctx, ptr, count, stream):
                try:
                    key = (ptr, count)
                    if key not in views:
                        views[key] = _ffi.torch_view_f64(ptr, count, DEVICE)
                    allreduce(views[key])
                    return 0
                except Exception as e:       # noqa: BLE001 -- must not unwind through the C frame
                    self._cb_error = e
                    return 1

            self._cb = _ffi.ALLREDUCE_FN(_cb)                      # kept alive with the trainer
            call("vgp_elbo_set_exchange", self.handle, int(n_total if n_total is not None else self.n), self._cb, None)

    def _batch(self, xb, yb):
        xb, yb = _points(xb), _vector(yb)
        assert xb.shape == (self.batch, self.d) and yb.size == self.batch, "minibatch shape is fixed at construction"
        return xb, yb

    def loss_and_grad(self, x_batch, y_batch):
        """(loss, d loss / d (v_a, v_l, v_s), d loss / d Z [m, d], terms dict) at the current parameters."""
        xb, yb = self._batch(x_batch, y_batch)
        loss, g = ctypes.c_double(), np.zeros(3)
        gz = np.zeros((self.m, self.d))
        terms = _ffi.VgpTerms()
        call("vgp_elbo_loss_grad", self.handle, xb.ptr, yb.ptr, ctypes.byref(loss), g.ctypes.data, gz.ctypes.data,
             ctypes.byref(terms), None)
        return loss.value, g, gz, {k: getattr(terms, k) for k, _ in terms._fields_}

    def step(self, x_batch, y_batch):
        xb, yb = self._batch(x_batch, y_batch)
        loss = ctypes.c_double()
        call("vgp_elbo_step", self.handle, xb.ptr, yb.ptr, ctypes.byref(loss), None)
        return loss.value

    def step_device(self, xb_ptr, yb_ptr):
        loss = ctypes.c_double()
        call("vgp_elbo_step", self.handle, xb_ptr, yb_ptr, ctypes.byref(loss), None)
        return loss.value

    def variables(self):
        v, z = np.zeros(3), np.zeros((self.m, self.d))
        call("vgp_elbo_get_params", self.handle, v.ctypes.data, z.ctypes.data, None)
        return v, z

    def assign(self, v=None, z=None):
        vp = None if v is None else np.ascontiguousarray(v, dtype=np.float64)
        zp = None if z is None else np.ascontiguousarray(z, dtype=np.float64)
        call("vgp_elbo_set_params", self.handle, None if vp is None else vp.ctypes.data,
             None if zp is None else zp.ctypes.data, None)

    def parameters(self):
        """Constrained (amplitude, length_scale, observation_noise_variance) and Z."""
        v, z = self.variables()
        return float(softplus(v[0])), float(self.length_scale_offset + softplus(v[1])), float(softplus(v[2])), z

    def vgp(self, index_points=None):
        """The VariationalGaussianProcess at the current parameters (for `.mean()` etc., :141-142)."""
        amp, ls, noise, z = self.parameters()
        k = self.kernel_class(amp, ls)
        loc, scale = VariationalGaussianProcess.optimal_variational_posterior(k, z, self._x, self._y, noise,
                                                                              jitter=self.jitter, as_device=True)
        return VariationalGaussianProcess(k, index_points, z, loc, scale, noise, jitter=self.jitter)

    def launch_count(self):
        c = ctypes.c_int64(0)
        call("vgp_elbo_launch_count", self.handle, ctypes.byref(c))
        return c.value

    def close(self):
        if getattr(self, "handle", None):
            call("vgp_elbo_destroy", self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
