"""GPU parity AT the BASELINE.json configurations: the same synthetic inputs bench.py times
(tools/workloads.py), compared with the CPU oracle where the oracle finishes in seconds.

  cfg1  3D_sin_wave field, n = 1000: VGP fit with 32 inducing points + greedy placement k = 10
  cfg2  mesh-shaped cloud n = 10 768, ExpQuad l = 0.226, k = 50: all three formulations vs the C/OpenMP oracle
  cfg3  VGP ELBO at m = 512, B = 4096 (N reduced to 20 000 so the torch-CPU oracle runs in seconds): loss + gradients
  cfg4  n = 50 000 is checked on the n = 8192 prefix of the very same cloud (what bench.py's CPU arm computes) -- the
        full size runs in bench.py, where the `parity` object of the JSON line carries the same comparison.

Selections must be bit-exact; scores / losses within 1e-9 relative (float64); the minimum relative top-2 gap of every
greedy input is printed and must stay above the 1e-12 near-tie threshold."""
import numpy as np
import pytest

from oracle import gp_oracle as gpo
from oracle import gp_oracle_torch as gt
from oracle import greedy_oracle as go
from tools import workloads
from vgposp_b200 import greedy
import vgposp_b200.gp_functions as gpf
import vgposp_b200.placement_algorithm2 as alg2

pytestmark = pytest.mark.gpu
D = 0
FORMULATIONS = ("dense", "lazy_precision", "lazy_factor")


def top2_gaps(step_scores):
    """Minimum over the steps of (best - second) / |best| on dense per-step scores [k, n] (NaN = taken)."""
    s = np.where(np.isnan(step_scores), -np.inf, step_scores)
    part = np.partition(s, -2, axis=1)[:, -2:]
    return (part[:, 1] - part[:, 0]) / np.abs(part[:, 1])


def check_against_c_oracle(cov, k, label):
    want_steps = []
    want_sel, want_scores = go.incremental_greedy_c(cov, k, all_scores=want_steps)
    want_steps = np.array(want_steps)
    gaps = top2_gaps(want_steps)
    print("%s: min relative top-2 gap %.3e (step %d)" % (label, gaps.min(), int(gaps.argmin())))
    assert gaps.min() > 1e-12, "near-tie in the input: index parity would depend on summation order"
    for form in FORMULATIONS:
        sel, scores, steps, _ = greedy.place_single(cov, k, D, want_step_scores=True, formulation=form)
        assert [int(s) for s in sel] == want_sel, form                      # bit-exact index parity
        np.testing.assert_allclose(scores, want_scores, rtol=1e-9, err_msg=form)
        np.testing.assert_allclose(steps, want_steps, rtol=1e-9, equal_nan=True, err_msg=form)
    return want_sel


def test_cfg2_mesh_cloud_n10768_k50_all_formulations(quiet_alg2):
    x, amp, ls, nugget = workloads.mesh_cloud()
    assert x.shape == (10768, 3) and abs(ls - 0.226) < 1e-3
    cov = gpf.ExponentiatedQuadratic(amp, ls).matrix(x, x)                  # the device builder (row a1)
    cov[np.diag_indices(len(cov))] += nugget
    host_rows = workloads.expquad_cov_host(x[:512], amp, ls, 0.0)
    np.testing.assert_allclose(cov[:512, :512] - nugget * np.eye(512), host_rows, rtol=1e-13)
    sel = check_against_c_oracle(cov, 50, "cfg2 n=10768 k=50")
    assert quiet_alg2.placement_algorithm_2(cov, 50) == sel                 # the drop-in call itself


def test_cfg4_cloud_prefix_n8192_matches_bench_cpu_arm():
    """The first 8192 points of the n = 50 000 cloud at the neighbour density of the full workload: exactly the sample
    bench.py's cpu_baseline / reference arm computes."""
    x, amp, _, nugget = workloads.cloud(50000)
    xs = x[:8192]
    cov = workloads.expquad_cov_host(xs, amp, workloads.length_scale_for(8192), nugget)
    check_against_c_oracle(cov, 25, "cfg4 prefix n=8192 k=25")


def test_cfg1_sin_wave_n1000_vgp_m32_then_placement_k10():
    """configs[0] end to end: VGP fit (m = 32, B = 64, Adam 0.01, kl_weight B/N; variational_Gaussian_process_example.py
    :47-125) followed by greedy placement (k = 10) on K(X; a = 1, l = 0.5) + 1e-2 I."""
    x, y, z = workloads.sin_field(1000)
    n, b, iters = 1000, 64, 12
    tr = gpf.VgpTrainer(x, y, z, b, learning_rate=0.01)
    params = [np.array(0.54), np.array(0.54), np.array(0.54), z.copy()]
    opt = gt.TfAdamTorch([p.shape for p in params], lr=0.01)
    rng = np.random.default_rng(0)
    for it in range(iters):
        idx = rng.integers(n, size=b)                                       # :119
        loss = tr.step(x[idx], y[idx])
        want_loss, grads = gt.loss_and_grads(params[0], params[1], params[2], params[3], x, y, x[idx], y[idx])
        assert loss == pytest.approx(want_loss, rel=1e-8), it
        params = opt.step(params, grads)
    v, zz = tr.variables()
    np.testing.assert_allclose(v, [float(p) for p in params[:3]], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(zz, params[3], rtol=1e-6, atol=1e-8)
    # posterior mean of the fitted model against the NumPy oracle at the fitted parameters
    amp, ls, noise, zfit = tr.parameters()
    xt = np.random.default_rng(1).uniform(-2, 2, (200, 3))
    loc, scale = gpo.optimal_variational_posterior(zfit, x, y, amp, ls, noise)
    want_mean = gpo.vgp_predict(zfit, loc, scale, xt, amp, ls)[0]
    np.testing.assert_allclose(tr.vgp(xt).mean(), want_mean, rtol=1e-8, atol=1e-10)
    tr.close()
    # placement on the ExpQuad covariance of the same points
    cov = gpf.ExponentiatedQuadratic(1.0, 0.5).matrix(x, x) + 1e-2 * np.eye(n)
    check_against_c_oracle(cov, 10, "cfg1 n=1000 k=10")


def test_cfg3_shape_m512_b4096_loss_and_gradients_match_torch_oracle():
    """BASELINE configs[2] shapes (m = 512, B = 4096, d = 3) with N = 20 000 observations: loss to 1e-9, gradients to
    1e-6 of their scale against torch autograd of the CPU restatement (oracle/gp_oracle_torch.py)."""
    n, m, b = 20000, 512, 4096
    x, y, z = workloads.elbo_problem(n, m)
    idx = np.random.default_rng(2).integers(n, size=b)
    v_ls = -1.5                                              # l = 0.2: the scale of the sin(2 pi x) field
    tr = gpf.VgpTrainer(x, y, z, b, v_length_scale=v_ls)
    loss, g, gz, _ = tr.loss_and_grad(x[idx], y[idx])
    want_loss, want = gt.loss_and_grads(0.54, v_ls, 0.54, z, x, y, x[idx], y[idx])
    assert loss == pytest.approx(want_loss, rel=1e-9)
    for i in range(3):
        assert g[i] == pytest.approx(float(want[i]), rel=1e-6, abs=1e-8 * abs(want_loss))
    np.testing.assert_allclose(gz, want[3], rtol=1e-5, atol=1e-6 * np.abs(want[3]).max())
    tr.close()
