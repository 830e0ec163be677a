"""CPU: the C-ABI library builds, loads, exports every symbol include/vgposp.h declares, the ctypes binding
covers all of them, and compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from vgposp_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vgposp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vgp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = _ffi.load()
    names = declared_symbols()
    assert len(names) >= 45
    out = subprocess.run(["nm", "-D", "--defined-only", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (vgp_\w+)", out))
    for name in names:
        assert name in exported, "%s declared in vgposp.h but not exported" % name
        assert name in _ffi.SIGNATURES, "%s has no ctypes signature" % name
        assert hasattr(lib, name)
    assert set(_ffi.SIGNATURES) <= exported
    # nothing but the C-ABI leaks out of the library
    assert not [s for s in re.findall(r" T (\w+)", out) if not s.startswith("vgp_")]


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_abi_version_and_error_text():
    lib = _ffi.load()
    assert lib.vgp_abi_version() == 1
    with pytest.raises(_ffi.VgpError) as e:
        _ffi.call("vgp_expquad_matrix", 0, None, 4, None, 4, 3, 1.0, -1.0, 0.0, 0, None, 4, None)
    assert e.value.status == _ffi.VGP_ERR_INVALID and "length_scale" in str(e.value)


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_ffi.Candidate) == 32
    assert ctypes.sizeof(_ffi.VgpTerms) == 40
    assert ctypes.sizeof(_ffi.TensorView) == 8 + 6 * 4 + 2 * 64


def test_dlpack_view_of_host_tensors():
    torch = pytest.importorskip("torch")
    t = torch.arange(24, dtype=torch.float64).reshape(4, 6)
    ft = _ffi.ForeignTensor(t)
    assert ft.shape == (4, 6) and ft.is_f64 and ft.contiguous and not ft.on_device
    assert ft.ptr == t.data_ptr()
    tt = t.t()
    ft2 = _ffi.ForeignTensor(tt)
    assert ft2.shape == (6, 4) and not ft2.contiguous
    a = np.arange(6, dtype=np.float64)
    ft3 = _ffi.ForeignTensor(a)            # numpy >= 1.23 speaks __dlpack__
    assert ft3.ptr == a.ctypes.data and ft3.shape == (6,)
    f32 = _ffi.ForeignTensor(torch.zeros(3, dtype=torch.float32))
    assert not f32.is_f64


@pytest.mark.skipif(_ffi.device_count() > 0, reason="checks the no-GPU failure mode")
def test_compute_calls_fail_loudly_without_a_gpu():
    import vgposp_b200.placement_algorithm2 as alg2
    with pytest.raises(_ffi.VgpError, match="no CPU fallback"):
        alg2.placement_algorithm_1(alg2.cov_vv_4x4(), 2)
    with pytest.raises(_ffi.VgpError):
        alg2.nominator(0, [1, 2], alg2.cov_vv_4x4())
    with pytest.raises(_ffi.VgpError):
        _ffi.DeviceArray((4,), np.float64)


def test_host_only_helpers_work_without_gpu():
    import vgposp_b200.placement_algorithm2 as alg2
    c = alg2.cov_vv_4x4()
    assert c.shape == (4, 4) and c[3, 3] == 1.2 and np.array_equal(c, c.T)
    np.testing.assert_array_equal(alg2.make_slice(c, [0, 2], [1, 3]), c[np.ix_([0, 2], [1, 3])])
    assert alg2.make_slice(c, [1], []).shape == (1, 0)
    assert alg2.argmax_cache_linear([0.1, 0.9, 0.9, 0.3], [1], range(4)) == 2
    assert alg2.argmax_cache_linear([-2.0, -3.0], [], range(2)) == -1
    assert alg2.sparse_argmax_cache_linear(np.array([[.1], [.9], [.9], [.3]]), [1], np.arange(4)) == 2
    assert alg2.call_pinv(np.array([[4.0]]))[0, 0] == 0.25


def test_shard_bounds_cover_everything():
    from vgposp_b200.greedy import shard_bounds
    for n, g in ((10, 3), (50000, 8), (7, 7), (100000, 8), (5, 1)):
        b = shard_bounds(n, g)
        assert b[0] == 0 and b[-1] == n and len(b) == g + 1
        assert all(y >= x for x, y in zip(b[:-1], b[1:]))
        assert max(y - x for x, y in zip(b[:-1], b[1:])) - min(y - x for x, y in zip(b[:-1], b[1:])) <= 1


def test_triangle_bounds_balance_the_lower_triangle():
    from vgposp_b200.greedy import triangle_bounds
    for n, g in ((10, 3), (50000, 8), (700, 2), (100000, 8), (5, 1), (1200, 4)):
        b = triangle_bounds(n, g)
        assert b[0] == 0 and b[-1] == n and len(b) == g + 1
        assert all(y >= x for x, y in zip(b[:-1], b[1:]))
    b = triangle_bounds(50000, 8)
    share = [(y * (y + 1) - x * (x + 1)) / 2 for x, y in zip(b[:-1], b[1:])]       # lower-triangle entries per slab
    assert max(share) / (50000 * 50001 / 2 / 8) < 1.03


def test_options_table_roundtrip_and_validation():
    """vgp_set_option / vgp_get_option (no device needed): defaults as documented in the header, bad values refused,
    and the library does not read the environment."""
    assert _ffi.get_option("gemm_emulate_slices") == 8
    assert _ffi.get_option("gemm_emulate_min") == 1024
    assert _ffi.get_option("dist_min_tiles") == 96 and _ffi.get_option("dist_min_k") == 256
    assert _ffi.get_option("dist_emulate_min") == -1
    with pytest.raises(_ffi.VgpError):
        _ffi.set_option("dist_emulate_min", 64)
    old = _ffi.set_option("gemm_emulate_slices", 0)
    assert old == 8 and _ffi.get_option("gemm_emulate_slices") == 0
    _ffi.set_option("gemm_emulate_slices", old)
    for bad in (1, 9, -3):
        with pytest.raises(_ffi.VgpError):
            _ffi.set_option("gemm_emulate_slices", bad)
    with pytest.raises(_ffi.VgpError):
        _ffi.call("vgp_set_option", 99, 1)
    assert sorted(_ffi.OPTIONS.values()) == list(range(len(_ffi.OPTIONS)))
    src = "".join(open(os.path.join(ROOT, "vgposp_b200", "csrc", f)).read()
                  for f in os.listdir(os.path.join(ROOT, "vgposp_b200", "csrc")))
    assert "getenv" not in src
