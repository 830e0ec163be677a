"""GPU check of the experimental FP64-on-int8 GEMM (csrc/emulated.cu, vgp_gemm_emulated): the digit-plane product must
agree with a float64 matmul to the accuracy tools/ozaki_prototype.py predicts for the slice count.

The kernel is not on any default path and has not been validated on hardware yet, so these tests only run with
VGP_TEST_EMULATED=1 (first thing to do with a GPU at hand: VGP_TEST_EMULATED=1 pytest tests/test_gpu_emulated_gemm.py)."""
import os

import numpy as np
import pytest

from vgposp_b200 import _ffi

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("VGP_TEST_EMULATED") != "1", reason="experimental kernel: opt in")]
D = 0


@pytest.fixture(params=[1, 2], ids=["by_group", "planes_resident"], autouse=True)
def variant(request, monkeypatch):
    """Both kernels: group by group (128 x 128 tiles) and all planes of a k block resident (128 x 64 tiles)."""
    monkeypatch.setenv("VGP_GEMM_EMULATE_VARIANT", str(request.param))
    return request.param


def emulated(a, b, trans_a, trans_b, m, n, k, alpha=1.0, beta=0.0, c=None, slices=8, lower=0):
    da, db = _ffi.DeviceArray.from_host(a, D), _ffi.DeviceArray.from_host(b, D)
    c = np.zeros((m, n + (n % 2))) if c is None else c
    dc = _ffi.DeviceArray.from_host(c, D)
    _ffi.call("vgp_gemm_emulated", D, trans_a, trans_b, m, n, k, alpha, da.ptr, a.shape[1], db.ptr, b.shape[1], beta,
              dc.ptr, c.shape[1], slices, lower, None)
    out = dc.to_host()
    for x in (da, db, dc):
        x.free()
    return out[:, :n]


@pytest.mark.parametrize("trans_a,trans_b", [(0, 1), (0, 0), (1, 0), (1, 1)])
@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (256, 384, 512), (200, 130, 77), (1000, 1024, 3000)])
def test_matches_float64_matmul(m, n, k, trans_a, trans_b):
    rng = np.random.default_rng(m + n + k)
    a = rng.standard_normal((k, m) if trans_a else (m, k)) * np.exp(rng.uniform(-8, 8, (1, m) if trans_a else (m, 1)))
    b = rng.standard_normal((n, k) if trans_b else (k, n))
    want = (a.T if trans_a else a) @ (b.T if trans_b else b)
    got = emulated(np.ascontiguousarray(a), np.ascontiguousarray(b), trans_a, trans_b, m, n, k)
    scale = np.abs(a.T if trans_a else a) @ np.abs(b.T if trans_b else b)
    assert np.max(np.abs(got - want) / scale) < 1e-13


@pytest.mark.parametrize("slices,tol", [(6, 1e-9), (7, 1e-11), (8, 1e-13), (9, 1e-15)])
def test_accuracy_follows_the_slice_count(slices, tol, variant):
    if slices == 9 and variant == 2:
        pytest.skip("nine group accumulators do not fit the 512 TMEM columns: falls back to the by-group kernel")
    rng = np.random.default_rng(slices)
    a, b = rng.standard_normal((256, 1024)), rng.standard_normal((256, 1024))
    want = (a.astype(np.longdouble) @ b.T.astype(np.longdouble)).astype(np.float64)
    got = emulated(a, b, 0, 1, 256, 256, 1024, slices=slices)
    assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < 30 * tol


def test_alpha_beta_lower_and_k_chunks():
    rng = np.random.default_rng(5)
    m, k = 384, 20000                                            # three k chunks of 8192
    a = rng.standard_normal((m, k))
    c0 = rng.standard_normal((m, m))
    want = -1.0 * (a @ a.T) + 1.0 * c0
    got = emulated(a, a, 0, 1, m, m, k, alpha=-1.0, beta=1.0, c=c0.copy(), lower=1)
    tiles = np.kron(np.tril(np.ones((3, 3))), np.ones((128, 128))).astype(bool)
    np.testing.assert_allclose(got[tiles], want[tiles], rtol=0, atol=1e-12 * k)
    np.testing.assert_array_equal(got[~tiles], c0[~tiles])      # tiles above the diagonal untouched


def test_workload_product_keeps_the_selection():
    """A factorisation-shaped product on this workload's matrices: L21 L21^T from a cloud covariance."""
    x = np.random.default_rng(0).uniform(-2, 2, (1024, 3))
    d = x[:, None, :] - x[None, :, :]
    cov = np.exp(-np.einsum("ijk,ijk->ij", d, d) / (2 * 0.5 ** 2)) + 1e-2 * np.eye(1024)
    l = np.linalg.cholesky(cov)
    a = np.ascontiguousarray(l[512:, :512])
    want = (a.astype(np.longdouble) @ a.T.astype(np.longdouble)).astype(np.float64)
    got = emulated(a, a, 0, 1, 512, 512, 512)
    assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < 1e-13


def test_both_kernels_give_the_same_bits(monkeypatch):
    rng = np.random.default_rng(11)
    a, b = rng.standard_normal((384, 640)), rng.standard_normal((320, 640))
    monkeypatch.setenv("VGP_GEMM_EMULATE_VARIANT", "1")
    one = emulated(a, b, 0, 1, 384, 320, 640)
    monkeypatch.setenv("VGP_GEMM_EMULATE_VARIANT", "2")
    two = emulated(a, b, 0, 1, 384, 320, 640)
    np.testing.assert_array_equal(one, two)          # same integer sums, same FP64 recombination order


def test_placement_on_emulated_factorisation_matches_the_oracle():
    """The whole one-call placement with the large products of potrf + trtri routed through the int8 kernels
    (VGP_GEMM_EMULATE is read once per process, hence the subprocess)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import numpy as np, json, sys\n"
        "sys.path.insert(0, %r)\n"
        "from vgposp_b200 import greedy\n"
        "x = np.random.default_rng(7).uniform(-2, 2, (3000, 3)); ls = 0.5 * (1000.0 / 3000) ** (1 / 3)\n"
        "d = x[:, None, :] - x[None, :, :]\n"
        "cov = np.exp(-np.einsum('ijk,ijk->ij', d, d) / (2 * ls * ls)) + 1e-2 * np.eye(3000)\n"
        "sel, sc, _, _ = greedy.place_single(cov, 12, 0, formulation='lazy_factor')\n"
        "print(json.dumps({'sel': [int(v) for v in sel], 'scores': [float(v) for v in sc]}))\n" % root)
    outs = []
    for emulate in ("0", "8"):
        env = dict(os.environ, VGP_GEMM_EMULATE=emulate, VGP_GEMM_EMULATE_MIN="512")
        res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env, cwd=root)
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
        import json
        outs.append(json.loads(res.stdout.strip().splitlines()[-1]))
    assert outs[0]["sel"] == outs[1]["sel"]
    np.testing.assert_allclose(outs[1]["scores"], outs[0]["scores"], rtol=1e-10)
