"""GPU: the process-wide pieces of the library -- workspace cache, option table, option fingerprint across ranks."""
import ctypes

import numpy as np
import pytest

from vgposp_b200 import _ffi, greedy
from vgposp_b200.dist_inverse import DistInverse

pytestmark = pytest.mark.gpu
D = 0


def cloud_cov(n, seed):
    x = np.random.default_rng(seed).uniform(-2, 2, (n, 3))
    ls = 0.5 * (1000.0 / n) ** (1 / 3)
    d = x[:, None, :] - x[None, :, :]
    return np.exp(-np.einsum("ijk,ijk->ij", d, d) / (2 * ls * ls)) + 1e-2 * np.eye(n)


def free_bytes():
    total, free = ctypes.c_size_t(0), ctypes.c_size_t(0)
    _ffi.call("vgp_device_info", D, None, 0, None, ctypes.byref(total), ctypes.byref(free))
    return free.value


def test_one_call_path_reuses_its_device_memory_and_trim_returns_it():
    cov = cloud_cov(3000, 1)
    _ffi.workspace_trim(D)
    before = free_bytes()
    sel1, sc1, _, _ = greedy.place_single(cov, 5, D)
    held = before - free_bytes()
    assert held >= 2 * 8 * 3072 * 3072                      # the two padded matrices stay cached
    wall = np.zeros(4)
    sel2, sc2, _, _ = greedy.place_single(cov, 5, D)
    _ffi.call("vgp_placement_host_wall", wall.ctypes.data)
    assert before - free_bytes() == held                    # the second call allocated nothing new
    assert np.array_equal(sel1, sel2) and np.array_equal(sc1, sc2)
    assert wall[0] < 0.05 and wall[3] >= wall[1] > 0        # allocation phase of a cached call: milliseconds
    released = _ffi.workspace_trim(D)
    assert released >= held and free_bytes() >= before - (1 << 20)
    assert _ffi.workspace_trim(D) == 0


def test_cache_limit_zero_disables_caching(vgp_options):
    cov = cloud_cov(1500, 2)
    _ffi.workspace_trim(D)
    vgp_options(workspace_cache_bytes=0)
    before = free_bytes()
    greedy.place_single(cov, 3, D)
    assert _ffi.workspace_trim(D) == 0 and free_bytes() >= before - (1 << 20)


def test_ranks_with_different_options_cannot_connect(vgp_options):
    """vgp_dist_connect compares a fingerprint of the options that change numerics or buffer sizes (a rank
    configured differently would store into scratch of another size, or produce other bits in its tiles)."""
    n = 640
    a = DistInverse(n, 0, 2, D)
    vgp_options(gemm_emulate_min=512)
    b = DistInverse(n, 1, 2, D)
    try:
        with pytest.raises(_ffi.VgpError, match="different options"):
            a.connect_pointers([a.pointers, b.pointers])
    finally:
        a.close()
        b.close()


def test_options_changed_after_create_are_refused(vgp_options):
    n = 640
    a = DistInverse(n, 0, 1, D)
    try:
        a.fill_padding()
        a.load_host(cloud_cov(n, 3))
        vgp_options(dist_min_k=512)
        with pytest.raises(_ffi.VgpError, match="options changed"):
            a.invert()
    finally:
        a.close()
