"""GPU check of the FP64-class GEMM on the int8 tensor cores (csrc/emulated.cu, vgp_gemm_emulated; the default route of the
large products of potrf / trtri / lauum): the digit-plane product must agree with a float64 matmul to the accuracy
tools/ozaki_prototype.py predicts for the slice count, and a factorisation built on it must leave the placement
unchanged.  First validated on a B200 in round 2 (profiles/r02_first_contact_experimental_paths.md)."""
import os

import numpy as np
import pytest

from vgposp_b200 import _ffi

pytestmark = pytest.mark.gpu
D = 0


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


@pytest.mark.parametrize("slices,tol", [(4, 1e-5), (6, 1e-9), (7, 1e-11), (8, 1e-13)])
def test_accuracy_follows_the_slice_count(slices, tol):
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


def test_result_is_independent_of_the_k_split():
    """Integer partial sums: splitting k (here: two calls with beta = 1 against one call) changes nothing but the
    order of the final FP64 additions of two exact group sums -- agreement to the last few ulps, and the single call
    is run-to-run bitwise reproducible."""
    rng = np.random.default_rng(11)
    a, b = rng.standard_normal((384, 640)), rng.standard_normal((320, 640))
    one = emulated(a, b, 0, 1, 384, 320, 640)
    again = emulated(a, b, 0, 1, 384, 320, 640)
    np.testing.assert_array_equal(one, again)
    first = emulated(np.ascontiguousarray(a[:, :256]), np.ascontiguousarray(b[:, :256]), 0, 1, 384, 320, 256)
    c = np.zeros((384, 320))
    c[:, :] = first
    two = emulated(np.ascontiguousarray(a[:, 256:]), np.ascontiguousarray(b[:, 256:]), 0, 1, 384, 320, 384, beta=1.0, c=c)
    np.testing.assert_allclose(two, one, rtol=0, atol=1e-13)


def test_rejects_bad_arguments():
    a = _ffi.DeviceArray.from_host(np.zeros((128, 128)), D)
    c = _ffi.DeviceArray.from_host(np.zeros((128, 128)), D)
    with pytest.raises(_ffi.VgpError, match="slices"):           # more planes than tensor memory holds
        _ffi.call("vgp_gemm_emulated", D, 0, 1, 128, 128, 128, 1.0, a.ptr, 128, a.ptr, 128, 0.0, c.ptr, 128, 9, 0, None)
    with pytest.raises(_ffi.VgpError, match="aliases"):          # C aliases A
        _ffi.call("vgp_gemm_emulated", D, 0, 1, 128, 128, 128, 1.0, a.ptr, 128, a.ptr, 128, 0.0, a.ptr, 128, 8, 0, None)
    a.free()
    c.free()


@pytest.mark.parametrize("formulation", ["lazy_factor", "lazy_precision", "dense"])
def test_placement_on_emulated_factorisation_matches_the_fp64_pipe_and_the_oracle(formulation, vgp_options):
    """The whole one-call placement with the large products of potrf + trtri (+ lauum) on the int8 tensor cores
    (threshold lowered so that n = 3000 has such products) against the same call on the FP64 pipe and the CPU oracle."""
    from oracle import greedy_oracle as go
    from vgposp_b200 import greedy
    n, k = 3000, 12
    x = np.random.default_rng(7).uniform(-2, 2, (n, 3))
    ls = 0.5 * (1000.0 / n) ** (1 / 3)
    d = x[:, None, :] - x[None, :, :]
    cov = np.exp(-np.einsum("ijk,ijk->ij", d, d) / (2 * ls * ls)) + 1e-2 * np.eye(n)
    vgp_options(gemm_emulate_slices=0)
    sel0, sc0, _, _ = greedy.place_single(cov, k, D, formulation=formulation)
    vgp_options(gemm_emulate_slices=8, gemm_emulate_min=512)
    sel1, sc1, _, _ = greedy.place_single(cov, k, D, formulation=formulation)
    want_sel, want_scores = go.incremental_greedy_c(cov, k)
    assert [int(v) for v in sel0] == [int(v) for v in sel1] == want_sel
    np.testing.assert_allclose(sc1, sc0, rtol=1e-11)
    np.testing.assert_allclose(sc1, want_scores, rtol=1e-9)
    assert not np.array_equal(sc1, sc0), "the int8 route was not taken"


def test_spd_inverse_on_emulated_products(vgp_options):
    import ctypes
    n = 4096
    x = np.random.default_rng(3).uniform(-2, 2, (n, 3))
    ls = 0.5 * (1000.0 / n) ** (1 / 3)
    cov = np.empty((n, n))
    for i in range(0, n, 512):
        dd = x[i:i + 512, None, :] - x[None, :, :]
        cov[i:i + 512] = np.exp(-np.einsum("ijk,ijk->ij", dd, dd) / (2 * ls * ls))
    cov[np.diag_indices(n)] += 1e-2
    outs = []
    for slices in (0, 8):
        vgp_options(gemm_emulate_slices=slices, gemm_emulate_min=1024)
        dev = _ffi.DeviceArray.from_host(cov, D)
        info = ctypes.c_int(0)
        _ffi.call("vgp_spd_inverse", D, dev.ptr, n, n, ctypes.byref(info), None)
        outs.append(dev.to_host())
        dev.free()
    np.testing.assert_allclose(outs[1], outs[0], rtol=0, atol=1e-11 * np.abs(outs[0]).max())
    np.testing.assert_allclose(outs[1] @ cov, np.eye(n), atol=1e-9)
