"""GPU check of the two experimental forms of the triangular solves: slab kernels (csrc/slab.cu, VGP_TRSM_SLAB=<width>: one
launch per solve of up to <width> columns instead of the recursion down to 128 x 128 leaves) and wide leaves
(csrc/dense.cu, VGP_TRSM_LEAF=<width>: one product with the cached explicit inverse of a diagonal node).  Same mathematics in a different
summation order, so results agree to rounding with the recursive path, and the distributed factorisation must still
be bitwise the single-device one when both use it.

Not validated on hardware yet and off by default: these tests run only with VGP_TEST_SLAB=1."""
import ctypes
import os
import threading

import numpy as np
import pytest

from vgposp_b200 import _ffi, greedy
from vgposp_b200.dist_inverse import DistInverse

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("VGP_TEST_SLAB") != "1", reason="experimental kernels: opt in")]
D = 0


def spd(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-2, 2, (n, 3))
    d = x[:, None, :] - x[None, :, :]
    ls = 0.5 * (1000.0 / n) ** (1 / 3)
    return np.exp(-np.einsum("ijk,ijk->ij", d, d) / (2 * ls * ls)) + 1e-2 * np.eye(n)


def inverse(a):
    n = a.shape[0]
    d = _ffi.DeviceArray.from_host(a, D)
    info = ctypes.c_int(0)
    _ffi.call("vgp_spd_inverse", D, d.ptr, n, n, ctypes.byref(info), None)
    out = d.to_host()
    d.free()
    return out


@pytest.mark.parametrize("width", [256, 512, 1024, 4096])
@pytest.mark.parametrize("n", [384, 1000, 2050, 3333])
def test_inverse_with_slab_solves_matches_the_recursive_path(n, width, monkeypatch):
    a = spd(n, n)
    monkeypatch.delenv("VGP_TRSM_SLAB", raising=False)
    want = inverse(a)
    monkeypatch.setenv("VGP_TRSM_SLAB", str(width))
    got = inverse(a)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-11 * np.abs(want).max())
    np.testing.assert_allclose(got @ a, np.eye(n), atol=1e-9)
    np.testing.assert_array_equal(got, inverse(a))              # run-to-run deterministic


def test_placement_with_slab_solves_matches_the_oracle(monkeypatch):
    from oracle import greedy_oracle as go
    monkeypatch.setenv("VGP_TRSM_SLAB", "1024")
    cov = spd(3000, 5)
    want_sel, want_scores = go.incremental_greedy(cov, 15)[:2]
    for form in ("dense", "lazy_factor", "lazy_precision"):
        sel, scores, _, _ = greedy.place_single(cov, 15, D, formulation=form)
        assert [int(v) for v in sel] == want_sel
        np.testing.assert_allclose(scores, want_scores, rtol=1e-9)


@pytest.mark.timeout(240)
@pytest.mark.parametrize("n,world", [(2050, 2), (3333, 4)])
def test_distributed_inverse_with_slab_solves_is_bitwise_the_single_device_one(n, world, monkeypatch):
    monkeypatch.setenv("VGP_TRSM_SLAB", "1024")
    monkeypatch.setenv("VGP_DIST_MIN_TILES", "2")
    monkeypatch.setenv("VGP_DIST_MIN_K", "256")
    a = spd(n, n + 2)
    want = inverse(a)
    streams = []
    for _ in range(world):
        s = ctypes.c_void_p()
        _ffi.call("vgp_stream_create", D, ctypes.byref(s))
        streams.append(s)
    ranks = [DistInverse(n, r, world, D, stream=streams[r]) for r in range(world)]
    for r in ranks:
        r.connect_pointers([q.pointers for q in ranks])
        r.fill_padding()
        r.load_host(a)
    errors = []

    def work(r):
        try:
            r.invert()
        except Exception as e:       # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(r,)) for r in ranks]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=200)
    assert not errors, errors
    for r in ranks:
        np.testing.assert_array_equal(r.to_host(), want)
        r.close()
    for s in streams:
        _ffi.call("vgp_stream_destroy", D, s)


@pytest.mark.parametrize("width", [256, 512, 1024])
@pytest.mark.parametrize("n", [384, 1000, 2050, 3333])
def test_inverse_with_wide_leaves_matches_the_recursive_path(n, width, monkeypatch):
    a = spd(n, n + 7)
    monkeypatch.delenv("VGP_TRSM_LEAF", raising=False)
    want = inverse(a)
    monkeypatch.setenv("VGP_TRSM_LEAF", str(width))
    got = inverse(a)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-11 * np.abs(want).max())
    np.testing.assert_allclose(got @ a, np.eye(n), atol=1e-9)
    np.testing.assert_array_equal(got, inverse(a))


def test_placement_with_wide_leaves_matches_the_oracle(monkeypatch):
    from oracle import greedy_oracle as go
    monkeypatch.setenv("VGP_TRSM_LEAF", "512")
    cov = spd(3000, 6)
    want_sel, want_scores = go.incremental_greedy(cov, 15)[:2]
    for form in ("dense", "lazy_factor", "lazy_precision"):
        sel, scores, _, _ = greedy.place_single(cov, 15, D, formulation=form)
        assert [int(v) for v in sel] == want_sel
        np.testing.assert_allclose(scores, want_scores, rtol=1e-9)


@pytest.mark.timeout(240)
@pytest.mark.parametrize("n,world", [(2050, 2), (3333, 4)])
def test_distributed_inverse_with_wide_leaves_is_bitwise_the_single_device_one(n, world, monkeypatch):
    monkeypatch.setenv("VGP_TRSM_LEAF", "512")
    monkeypatch.setenv("VGP_DIST_MIN_TILES", "2")
    monkeypatch.setenv("VGP_DIST_MIN_K", "256")
    a = spd(n, n + 3)
    want = inverse(a)
    streams = []
    for _ in range(world):
        s = ctypes.c_void_p()
        _ffi.call("vgp_stream_create", D, ctypes.byref(s))
        streams.append(s)
    ranks = [DistInverse(n, r, world, D, stream=streams[r]) for r in range(world)]
    for r in ranks:
        r.connect_pointers([q.pointers for q in ranks])
        r.fill_padding()
        r.load_host(a)
    errors = []

    def work(r):
        try:
            r.invert()
        except Exception as e:       # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(r,)) for r in ranks]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=200)
    assert not errors, errors
    for r in ranks:
        np.testing.assert_array_equal(r.to_host(), want)
        r.close()
    for s in streams:
        _ffi.call("vgp_stream_destroy", D, s)
