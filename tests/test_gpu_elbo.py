"""GPU parity: the VGP training step (loss, hand-derived gradient, Adam update) against torch autograd of the CPU
restatement (oracle/gp_oracle_torch.py; TFP semantics, parity unpinned).  Tolerances: loss 1e-9 relative;
gradients 1e-6 relative to the gradient's scale (they pass through three m x m inverses with cond up to 1e8)."""
import numpy as np
import pytest

from oracle import gp_oracle as gpo
from oracle import gp_oracle_torch as gt
import vgposp_b200.gp_functions as gpf

pytestmark = pytest.mark.gpu


def problem(n, m, b, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-2, 2, (n, d))
    y = np.sin(x[:, 0]) * np.sin(x[:, 1 % d]) + 0.1 + 0.1 * rng.standard_normal(n)
    z = rng.uniform(-2, 2, (m, d))
    idx = rng.integers(n, size=b)
    return x, y, z, idx


# v_ls: unconstrained length-scale variable.  The loss contains logdet K_zz of the UN-jittered kernel matrix (through
# logdet S); with many inducing points and a long length scale K_zz is singular to working precision and
# that term is noise in ANY implementation (cond(K_zz) > 1e12 for m = 100 points at l = 0.97 in 2-D), so the larger
# cases are checked at a shorter length scale.
@pytest.mark.parametrize("n,m,b,d,v_ls", [(600, 24, 64, 3, 0.54), (1000, 32, 64, 3, 0.54), (3000, 100, 300, 2, -1.2),
                                          (5000, 130, 256, 3, -0.5)])
def test_loss_and_gradient_match_autograd(n, m, b, d, v_ls):
    x, y, z, idx = problem(n, m, b, d, n + m)
    tr = gpf.VgpTrainer(x, y, z, b, v_length_scale=v_ls)
    loss, g, gz, terms = tr.loss_and_grad(x[idx], y[idx])
    want_loss, want = gt.loss_and_grads(0.54, v_ls, 0.54, z, x, y, x[idx], y[idx])
    assert loss == pytest.approx(want_loss, rel=1e-9)
    for i in range(3):
        assert g[i] == pytest.approx(float(want[i]), rel=1e-6, abs=1e-8 * abs(want_loss))
    np.testing.assert_allclose(gz, want[3], rtol=1e-5, atol=1e-6 * np.abs(want[3]).max())
    # the pieces agree with the forward-only entry point and with the NumPy oracle
    amp, ls, noise = gpo.softplus(0.54), 1e-5 + gpo.softplus(v_ls), gpo.softplus(0.54)
    loc, scale = gpo.optimal_variational_posterior(z, x, y, amp, ls, noise)
    ref = gpo.vgp_terms(z, loc, scale, x[idx], y[idx], amp, ls, noise, b / n)
    for key in ("ll", "tr1", "tr2", "kl"):
        assert terms[key] == pytest.approx(ref[key], rel=1e-8, abs=1e-9 * abs(want_loss)), key
    tr.close()


@pytest.mark.parametrize("kernel,name,n,m,b,d,v_ls", [
    (gpf.MaternFiveHalves, "matern52", 1500, 40, 128, 5, 0.54),       # the reference's VGP: 5-D (x, y, z, t, p) inputs
    (gpf.MaternFiveHalves, "matern52", 4000, 128, 256, 5, 0.0),
    (gpf.MaternFiveHalves, "matern52", 2000, 64, 200, 3, -0.5),
    (gpf.MaternThreeHalves, "matern32", 1500, 40, 128, 5, 0.54),
])
def test_matern_training_step_matches_autograd(kernel, name, n, m, b, d, v_ls):
    """main_architecture_2.py:184-249 trains tfkern.MaternFiveHalves over 5-D inputs with jitter 1e-6: hand-derived
    gradient (csrc/elbo.cu, kernel_q) against torch autograd of the same loss with TFP's kernel form."""
    x, y, z, idx = problem(n, m, b, d, n + m + d)
    tr = gpf.VgpTrainer(x, y, z, b, v_length_scale=v_ls, kernel=kernel, jitter=1e-6)
    loss, g, gz, _ = tr.loss_and_grad(x[idx], y[idx])
    want_loss, want = gt.loss_and_grads(0.54, v_ls, 0.54, z, x, y, x[idx], y[idx], kernel=name)
    assert loss == pytest.approx(want_loss, rel=1e-9)
    for i in range(3):
        assert g[i] == pytest.approx(float(want[i]), rel=1e-6, abs=1e-8 * abs(want_loss))
    np.testing.assert_allclose(gz, want[3], rtol=1e-5, atol=1e-6 * np.abs(want[3]).max())
    # two Adam steps follow the oracle optimiser
    params = [np.array(0.54), np.array(v_ls), np.array(0.54), z.copy()]
    opt = gt.TfAdamTorch([p.shape for p in params], lr=0.01)
    for it in range(2):
        got = tr.step(x[idx], y[idx])
        wl, grads = gt.loss_and_grads(params[0], params[1], params[2], params[3], x, y, x[idx], y[idx], kernel=name)
        assert got == pytest.approx(wl, rel=1e-7), it
        params = opt.step(params, grads)
    v, zz = tr.variables()
    np.testing.assert_allclose(v, [float(p) for p in params[:3]], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(zz, params[3], rtol=1e-6, atol=1e-8)
    mean = tr.vgp(x[:50]).mean()
    assert np.all(np.isfinite(mean))
    tr.close()


def test_matern12_training_is_refused():
    x, y, z, _ = problem(300, 8, 32, 3, 1)
    with pytest.raises(Exception, match="MATERN"):
        gpf.VgpTrainer(x, y, z, 32, kernel=gpf.MaternOneHalf)


def test_training_steps_follow_tf_adam():
    n, m, b, d = 800, 20, 64, 3
    x, y, z, _ = problem(n, m, b, d, 5)
    tr = gpf.VgpTrainer(x, y, z, b, learning_rate=0.01)
    params = [np.array(0.54), np.array(0.54), np.array(0.54), z.copy()]
    opt = gt.TfAdamTorch([p.shape for p in params], lr=0.01)
    rng = np.random.default_rng(0)
    for it in range(5):
        idx = rng.integers(n, size=b)                        # variational_Gaussian_process_example.py:119
        loss = tr.step(x[idx], y[idx])
        want_loss, grads = gt.loss_and_grads(params[0], params[1], params[2], params[3], x, y, x[idx], y[idx])
        assert loss == pytest.approx(want_loss, rel=1e-7), it
        params = opt.step(params, grads)
        v, zz = tr.variables()
        np.testing.assert_allclose(v, [float(p) for p in params[:3]], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(zz, params[3], rtol=1e-6, atol=1e-8)
    assert tr.launch_count() > 100
    tr.close()


def test_loss_decreases_and_prediction_tracks_the_field():
    n, m, b = 4000, 64, 256
    rng = np.random.default_rng(1)
    x = rng.uniform(-2, 2, (n, 2))
    f = lambda p: np.sin(p[:, 0]) * np.sin(p[:, 1]) + 0.1     # noqa: E731  (3D_sin_wave.py:96-103)
    y = f(x) + 0.05 * rng.standard_normal(n)
    z = rng.uniform(-2, 2, (m, 2))
    tr = gpf.VgpTrainer(x, y, z, b, learning_rate=0.05)
    losses = []
    for it in range(60):
        idx = rng.integers(n, size=b)
        losses.append(tr.step(x[idx], y[idx]))
    assert np.mean(losses[-10:]) < np.mean(losses[:10]) - 50
    xt = rng.uniform(-1.8, 1.8, (500, 2))
    mean = tr.vgp(xt).mean()
    assert np.sqrt(np.mean((mean - f(xt)) ** 2)) < 0.05
    amp, ls, noise, _ = tr.parameters()
    assert noise < 0.2 and 0.3 < ls < 3.0
    tr.close()


def test_config3_shape_runs_and_is_finite():
    """BASELINE configs[2] shape at reduced N (m = 512, B = 4096): finite loss, gradient matches a central
    finite difference of the device loss in v_length_scale."""
    n, m, b = 20000, 512, 4096
    rng = np.random.default_rng(2)
    x = rng.uniform(-2, 2, (n, 3))
    y = np.sum(np.sin(2 * np.pi * x), axis=1) + 0.1 * rng.standard_normal(n)      # gp_functions.py:78-95
    z = rng.uniform(-2, 2, (m, 3))
    idx = rng.integers(n, size=b)
    v_ls = -1.5                                              # l = 0.2: the scale of the sin(2 pi x) field
    tr = gpf.VgpTrainer(x, y, z, b, v_length_scale=v_ls)
    loss, g, gz, _ = tr.loss_and_grad(x[idx], y[idx])
    assert np.isfinite(loss) and np.all(np.isfinite(g)) and np.all(np.isfinite(gz))
    h = 1e-5
    tr.assign(v=[0.54, v_ls + h, 0.54])
    lp = tr.loss_and_grad(x[idx], y[idx])[0]
    tr.assign(v=[0.54, v_ls - h, 0.54])
    lm = tr.loss_and_grad(x[idx], y[idx])[0]
    assert g[1] == pytest.approx((lp - lm) / (2 * h), rel=1e-5)
    tr.close()


def test_n_axis_sharding_two_ranks_as_threads():
    """the training step sharded over the observations.  Two trainers, each on half of the data,
    exchange the three all-reduced buffers through a callback (here: two threads on one device, summed with torch);
    losses and parameters must follow the single-trainer run on all the data, and the two replicas must stay bitwise
    identical."""
    import threading
    torch = pytest.importorskip("torch")
    n, m, b, d = 6000, 64, 256, 3
    rng = np.random.default_rng(4)
    x = rng.uniform(-2, 2, (n, d))
    y = np.sum(np.sin(2 * np.pi * x), axis=1) + 0.1 * rng.standard_normal(n)
    z = rng.uniform(-2, 2, (m, d))
    single = gpf.VgpTrainer(x, y, z, b)
    barrier = threading.Barrier(2)
    pending = [None, None]

    def make_allreduce(rank):
        def allreduce(t):
            torch.cuda.synchronize()
            pending[rank] = t
            barrier.wait()
            total = pending[0] + pending[1]
            torch.cuda.synchronize()
            barrier.wait()
            t.copy_(total)
            torch.cuda.synchronize()
            barrier.wait()
        return allreduce

    halves = [gpf.VgpTrainer(x[r::2], y[r::2], z, b, allreduce=make_allreduce(r), n_total=n) for r in range(2)]
    batches = [rng.integers(n, size=b) for _ in range(6)]
    want = [single.step(x[i], y[i]) for i in batches]
    got = [[], []]
    errors = []

    def work(rank):
        try:
            for i in batches:
                got[rank].append(halves[rank].step(x[i], y[i]))
        except Exception as e:       # noqa: BLE001
            errors.append(e)
            barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors, errors
    np.testing.assert_allclose(got[0], want, rtol=1e-9)
    assert got[0] == got[1]                                   # replicas: identical sums in, identical arithmetic
    v0, z0 = halves[0].variables()
    v1, z1 = halves[1].variables()
    vs, zs = single.variables()
    assert np.array_equal(v0, v1) and np.array_equal(z0, z1)
    np.testing.assert_allclose(v0, vs, rtol=1e-8)
    np.testing.assert_allclose(z0, zs, rtol=1e-7, atol=1e-9)
    for t in halves + [single]:
        t.close()
