"""GPU parity: ExpQuad builder and the dense float64 layer, through the C-ABI, against NumPy/SciPy."""
import ctypes

import numpy as np
import pytest
import scipy.linalg as sla

from oracle import gp_oracle as gpo
from vgposp_b200 import _ffi

pytestmark = pytest.mark.gpu
D = 0


def dev(a):
    return _ffi.DeviceArray.from_host(np.ascontiguousarray(a, dtype=np.float64), D)


def expquad(x1, x2, amp, ls, diag_add=0.0, diag_col0=0, ld=None, same=False):
    n1, n2 = x1.shape[0], x2.shape[0]
    ld = ld or n2
    d1 = dev(x1)
    d2 = d1 if same else dev(x2)
    out = _ffi.DeviceArray((n1, ld), np.float64, D).zero_()
    _ffi.call("vgp_expquad_matrix", D, d1.ptr, n1, d2.ptr, n2, x1.shape[1], amp, ls, diag_add, diag_col0, out.ptr, ld,
              None)
    return out.to_host()[:, :n2]


@pytest.mark.parametrize("n1,n2,d", [(1, 1, 3), (5, 7, 1), (33, 1030, 3), (513, 77, 5), (300, 300, 8), (129, 640, 2)])
def test_expquad_rectangular(n1, n2, d):
    rng = np.random.default_rng(n1 * 1000 + n2)
    x1, x2 = rng.uniform(-2, 2, (n1, d)), rng.uniform(-2, 2, (n2, d))
    got = expquad(x1, x2, 1.3, 0.45)
    want = gpo.expquad_matrix(x1, x2, 1.3, 0.45)
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-300)      # float64, tolerance 1e-9 in north_star


@pytest.mark.parametrize("n", [128, 130, 191, 1000, 2049])
def test_expquad_symmetric_path(n):
    x = np.random.default_rng(n).uniform(-2, 2, (n, 3))
    got = expquad(x, x, 1.0, 0.5, diag_add=1e-2, same=True, ld=n + (n % 2))
    want = gpo.expquad_matrix(x, x, 1.0, 0.5, diag_add=1e-2)
    np.testing.assert_allclose(got, want, rtol=1e-13)
    assert np.array_equal(got, got.T)                                   # mirrored tiles: bitwise symmetric


def test_expquad_odd_ld_and_panel_diagonal():
    x = np.random.default_rng(5).uniform(-2, 2, (200, 3))
    got = expquad(x, x[60:131], 1.1, 0.3, diag_add=0.25, diag_col0=60, ld=73)
    want = gpo.expquad_matrix(x, x, 1.1, 0.3, diag_add=0.25)[:, 60:131]
    np.testing.assert_allclose(got, want, rtol=1e-13)


@pytest.mark.parametrize("kind", ["expquad", "matern12", "matern32", "matern52"])
@pytest.mark.parametrize("n1,n2,d,same", [(33, 1030, 3, False), (513, 77, 5, False), (700, 700, 5, True), (130, 130, 3, True)])
def test_kernel_matrix_kinds(kind, n1, n2, d, same):
    """vgp_kernel_matrix, rectangular and symmetric paths, against the float64 NumPy statement of TFP's formulas.
    Tolerance: 2 ulp of the fast exp + a few ulp of the Matern prefactor and sqrt -> 1e-13 relative."""
    rng = np.random.default_rng(n1 + 31 * n2 + d)
    x1 = rng.uniform(-2, 2, (n1, d))
    x2 = x1 if same else rng.uniform(-2, 2, (n2, d))
    d1 = dev(x1)
    d2 = d1 if same else dev(x2)
    out = _ffi.DeviceArray((n1, n2), np.float64, D).zero_()
    _ffi.call("vgp_kernel_matrix", D, gpo.KERNEL_KINDS[kind], d1.ptr, n1, d2.ptr, n2, d, 0.8, 0.9, 0.03 if same else 0.0, 0,
              out.ptr, n2, None)
    got = out.to_host()
    want = gpo.kernel_matrix(kind, x1, x2, 0.8, 0.9, diag_add=0.03 if same else 0.0)
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-300)
    if same:
        assert np.array_equal(got, got.T)


def test_expquad_fast_exp_edges():
    """Arguments at the edges of the fast exp: exact zero distance (value a^2 exactly), arguments below -650 (the fast
    path returns exactly 0 where libm still has 1e-283 a^2 -- hence atol), huge amplitude (libm kernels, no flush),
    and a dense sweep of exponents over the whole fast range."""
    x1 = np.zeros((3, 1))
    x1[:, 0] = [0.0, 1.0, 40.0]
    x2 = np.array([[0.0], [1.0 + 1e-9], [3.0], [12.0], [26.0], [27.5], [39.0]])
    for amp, atol in ((1.0, 1e-282), (0.37, 1e-282), (1e25, 1e-310)):
        got = expquad(x1, x2, amp, 0.7)
        want = gpo.expquad_matrix(x1, x2, amp, 0.7)
        np.testing.assert_allclose(got, want, rtol=2e-13, atol=atol)
        assert got[0, 0] == amp * amp
    t = -np.linspace(0.0, 700.0, 4001)[:, None]               # exp(t) via l = 1/sqrt(2), x = sqrt(-t)
    xs = np.sqrt(-t)
    got = expquad(xs, np.zeros((1, 1)), 1.0, np.sqrt(0.5))
    want = gpo.expquad_matrix(xs, np.zeros((1, 1)), 1.0, np.sqrt(0.5))
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-282)
    assert np.all(got[t[:, 0] > -649.0] > 0.0)


def test_expquad_empty_and_bad_arguments():
    out = _ffi.DeviceArray((4,), np.float64, D)
    _ffi.call("vgp_expquad_matrix", D, out.ptr, 0, out.ptr, 4, 3, 1.0, 1.0, 0.0, 0, out.ptr, 4, None)   # n1 == 0: no-op
    with pytest.raises(_ffi.VgpError):
        _ffi.call("vgp_expquad_matrix", D, out.ptr, 1, out.ptr, 1, 9, 1.0, 1.0, 0.0, 0, out.ptr, 1, None)


def gemm(ta, tb, a, b, c=None, alpha=1.0, beta=0.0):
    m = a.shape[1] if ta else a.shape[0]
    k = a.shape[0] if ta else a.shape[1]
    n = b.shape[0] if tb else b.shape[1]
    da, db = dev(a), dev(b)
    dc = dev(c) if c is not None else _ffi.DeviceArray((m, n), np.float64, D).zero_()
    _ffi.call("vgp_dgemm", D, ta, tb, m, n, k, alpha, da.ptr, a.shape[1], db.ptr, b.shape[1], beta, dc.ptr, n, None)
    return dc.to_host()


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (256, 384, 512), (100, 37, 59), (1, 1, 1), (300, 129, 1000)])
def test_dgemm_all_layouts(ta, tb, m, n, k):
    rng = np.random.default_rng(m + 7 * n + 13 * k)
    a = rng.standard_normal((k, m) if ta else (m, k))
    b = rng.standard_normal((n, k) if tb else (k, n))
    c = rng.standard_normal((m, n))
    got = gemm(ta, tb, a, b, c, alpha=-0.7, beta=1.5)
    want = -0.7 * ((a.T if ta else a) @ (b.T if tb else b)) + 1.5 * c
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12 * np.sqrt(k))


def spd(n, seed):
    x = np.random.default_rng(seed).uniform(-2, 2, (n, 3))
    return gpo.expquad_matrix(x, x, 1.0, 0.5 * (1000.0 / max(n, 1000)) ** (1 / 3), diag_add=1e-2)


@pytest.mark.parametrize("n", [1, 2, 127, 128, 129, 256, 640, 1000, 2500])
def test_potrf_matches_lapack(n):
    a = spd(n, n)
    marker = np.triu(np.full((n, n), 7.0), 1)
    d = dev(np.tril(a) + marker)                      # strict upper triangle must come back untouched
    info = ctypes.c_int(-1)
    _ffi.call("vgp_potrf", D, d.ptr, n, n, ctypes.byref(info), None)
    got = d.to_host()
    want = np.linalg.cholesky(a)
    assert info.value == 0
    np.testing.assert_allclose(np.tril(got), want, rtol=1e-9, atol=1e-13)      # factors within 1e-9 relative
    np.testing.assert_array_equal(np.triu(got, 1), marker)
    np.testing.assert_allclose(np.tril(got) @ np.tril(got).T, a, rtol=1e-12, atol=1e-14)


def test_potrf_reports_not_positive_definite():
    a = spd(300, 1)
    a[200, 200] = -1.0
    d = dev(a)
    info = ctypes.c_int(0)
    with pytest.raises(_ffi.NotPositiveDefiniteError):
        _ffi.call("vgp_potrf", D, d.ptr, 300, 300, ctypes.byref(info), None)
    assert info.value == 201                          # 1 + first failing row


@pytest.mark.parametrize("n", [1, 3, 128, 200, 384, 1000, 2304])
def test_spd_inverse(n):
    a = spd(n, 100 + n)
    d = dev(a)
    info = ctypes.c_int(0)
    _ffi.call("vgp_spd_inverse", D, d.ptr, n, n, ctypes.byref(info), None)
    got = d.to_host()
    want = np.linalg.inv(a)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
    assert np.array_equal(got, got.T)
    np.testing.assert_allclose(got @ a, np.eye(n), atol=1e-10)


@pytest.mark.parametrize("side,trans", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("n,nrhs", [(128, 128), (300, 5), (640, 257), (1, 3)])
def test_trsm(side, trans, n, nrhs):
    l = np.linalg.cholesky(spd(n, 7 * n))
    rng = np.random.default_rng(n + nrhs)
    b = rng.standard_normal((n, nrhs) if side == 0 else (nrhs, n))
    dl, db = dev(l + np.triu(np.full((n, n), 9.0), 1)), dev(b)      # garbage above the diagonal must be ignored
    _ffi.call("vgp_trsm", D, side, trans, n, nrhs, dl.ptr, n, db.ptr, b.shape[1], None)
    got = db.to_host()
    op = l.T if trans else l
    want = sla.solve_triangular(op, b, lower=not trans) if side == 0 else \
        sla.solve_triangular(op.T, b.T, lower=bool(trans)).T
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-11 * np.abs(want).max())


@pytest.mark.parametrize("trans", [0, 1])
def test_trsm_right_in_place_many_rows_is_deterministic(trans):
    """The right-side solves multiply B by inverted diagonal blocks IN PLACE (C aliases A).  With tiles narrower than
    the 128-column block two CTAs would read what the other overwrites; the product must run on the 128x128
    configuration (dense.cu, `in_place`).  Many row tiles and repeated runs make the race visible if it comes back."""
    n, rows = 512, 16384
    l = np.linalg.cholesky(spd(n, 3))
    b = np.random.default_rng(9).standard_normal((rows, n))
    dl = dev(l)
    want = sla.solve_triangular(l if trans else l.T, b.T, lower=bool(trans)).T
    first = None
    for _ in range(4):
        db = dev(b)
        _ffi.call("vgp_trsm", D, 1, trans, n, rows, dl.ptr, n, db.ptr, n, None)
        got = db.to_host()
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-11 * np.abs(want).max())
        if first is None:
            first = got
        assert np.array_equal(got, first)


@pytest.mark.parametrize("trans", [0, 1])
def test_trsm_left_in_place_is_deterministic(trans):
    """The left-side solves multiply B by inverted diagonal blocks in place with C aliasing B (m = 128 rows, many
    columns): one CTA must own all 128 rows of its column block, so the 64-row tile shape must not be chosen.
    Few columns (small tile count -> the small-tile heuristic would apply) and repeated runs expose the race."""
    n, cols = 512, 384
    l = np.linalg.cholesky(spd(n, 5))
    b = np.random.default_rng(11).standard_normal((n, cols))
    dl = dev(l)
    want = sla.solve_triangular(l.T if trans else l, b, lower=not trans)
    first = None
    for _ in range(20):
        db = dev(b)
        _ffi.call("vgp_trsm", D, 0, trans, n, cols, dl.ptr, n, db.ptr, cols, None)
        got = db.to_host()
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-11 * np.abs(want).max())
        if first is None:
            first = got
        assert np.array_equal(got, first)


def test_spd_inverse_is_run_to_run_deterministic():
    a = spd(512, 9)
    outs = []
    for _ in range(10):
        d = dev(a)
        info = ctypes.c_int(0)
        _ffi.call("vgp_spd_inverse", D, d.ptr, 512, 512, ctypes.byref(info), None)
        outs.append(d.to_host())
    for o in outs[1:]:
        assert np.array_equal(o, outs[0])
