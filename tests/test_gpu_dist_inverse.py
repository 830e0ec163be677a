"""GPU parity: the distributed SPD inverse (csrc/dist.cu) -- GEMM tiles split over ranks, every finished tile stored
into all replicas by the GEMM epilogue -- must equal the single-device inverse bit for bit and NumPy's inverse to
1e-9 (the pseudo-inverses of placement_algorithm2.py:399-413 on an SPD input)."""
import ctypes
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from vgposp_b200 import _ffi
from vgposp_b200.dist_inverse import UPLOAD_LOWER, DistInverse

pytestmark = pytest.mark.gpu
D = 0


def spd(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-2, 2, (n, 3))
    d = x[:, None, :] - x[None, :, :]
    ls = 0.5 * (1000.0 / n) ** (1 / 3)
    return np.exp(-np.einsum("ijk,ijk->ij", d, d) / (2 * ls * ls)) + 1e-2 * np.eye(n)


def single_inverse(a):
    n = a.shape[0]
    d = _ffi.DeviceArray.from_host(a, D)
    info = ctypes.c_int(0)
    _ffi.call("vgp_spd_inverse", D, d.ptr, n, n, ctypes.byref(info), None)
    out = d.to_host()
    d.free()
    return out


@pytest.mark.timeout(240)
@pytest.mark.parametrize("emulate_min", [4096, 512], ids=["fp64_products", "int8_products"])
@pytest.mark.parametrize("n,world", [(1000, 2), (1664, 3), (2050, 4)])
def test_ranks_as_threads_one_device(n, world, emulate_min, vgp_options):
    """G ranks as threads of this process on one device (each with its own stream and replica)."""
    vgp_options(dist_min_tiles=2, dist_min_k=256, gemm_emulate_min=emulate_min, dist_emulate_min=emulate_min)
    a = spd(n, n)
    want = single_inverse(a)
    streams = []
    for _ in range(world):
        s = ctypes.c_void_p()
        _ffi.call("vgp_stream_create", D, ctypes.byref(s))
        streams.append(s)
    ranks = [DistInverse(n, r, world, D, stream=streams[r]) for r in range(world)]
    for r in ranks:
        r.connect_pointers([q.pointers for q in ranks])
        r.fill_padding()
    # rank r uploads only its row slab and pushes it to the other replicas
    bounds = [(n * g) // world for g in range(world + 1)]
    for r in ranks:
        r.load_host(a, bounds[r.rank], bounds[r.rank + 1])
        r.push_rows(bounds[r.rank], bounds[r.rank + 1])
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
    assert ranks[0].stats()["distributed_gemms"] > 0
    for r in ranks:
        got = r.to_host()
        np.testing.assert_array_equal(got, want)            # same kernels, same summation order -> same bits
    np.testing.assert_allclose(want @ a, np.eye(n), atol=1e-9)
    for r in ranks:
        r.close()
    for s in streams:
        _ffi.call("vgp_stream_destroy", D, s)


@pytest.mark.timeout(300)
def test_two_processes_ipc():
    cnt = _ffi.c_int(0)
    _ffi.call("vgp_device_count", ctypes.byref(cnt))
    if cnt.value < 2:
        pytest.skip("needs two CUDA devices")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(root, "tests", "dist_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=280, cwd=root)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "DIST_WORKER_OK" in res.stdout


@pytest.mark.timeout(240)
@pytest.mark.parametrize("emulate_min", [4096, 512], ids=["fp64_products", "int8_products"])
@pytest.mark.parametrize("n,world,k", [(1000, 2, 12), (1664, 3, 20), (2050, 4, 9), (700, 2, 600)])
def test_sharded_lazy_factor_greedy_threads(n, world, k, emulate_min, vgp_options):
    """The one-call path on G ranks: replicas factorised to L^-1 by the distributed potrf + trtri, the triangular
    matrix-vector product of every selection split over the ranks (csrc/lazy.cu, vgp_lazy_create_dist).  Every rank
    must return the single-device lazy-factor result bit for bit, and the CPU oracle's selection."""
    from oracle import greedy_oracle as go
    from vgposp_b200 import greedy
    vgp_options(dist_min_tiles=2, dist_min_k=256, gemm_emulate_min=emulate_min, dist_emulate_min=emulate_min)
    a = spd(n, n + 1)
    want = greedy.place_single(a, k, D, want_step_scores=True, formulation="lazy_factor")
    streams = []
    for _ in range(world):
        s = ctypes.c_void_p()
        _ffi.call("vgp_stream_create", D, ctypes.byref(s))
        streams.append(s)
    ranks = [DistInverse(n, r, world, D, stream=streams[r]) for r in range(world)]
    for r in ranks:
        r.connect_pointers([q.pointers for q in ranks])
        r.fill_padding()
        r.load_host(np.full((n, n), np.nan))        # poison: nothing may read the strict upper triangle
    lazies = [greedy.LazyGreedy.from_dist(r, k) for r in ranks]
    for lz in lazies:
        # allocates the per-step score buffer: here, not inside the rank threads -- a cudaMalloc (workspace-cache miss)
        # may wait for the device, on which another rank's barrier kernel is spinning in this single-device emulation
        lz.record_scores(True)
    bounds = [(n * g) // world for g in range(world + 1)]
    for r in ranks:     # each rank uploads the lower-triangle share of its row slab; the peer copies ride along
        r0, r1 = bounds[r.rank], bounds[r.rank + 1]
        r.upload_rows(a[r0:r1], r0, r1, ncols=UPLOAD_LOWER)
    errors, out = [], [None] * world

    def work(i):
        try:
            r, lz = ranks[i], lazies[i]
            r.barrier()
            lz.load_cov_device(r.ptr, r.ld)
            r.factor_inverse()
            lz.adopt_factor()
            lz.run(k)
            sel, sc = lz.results()
            out[i] = (sel[:k], sc[:k], lz.step_scores()[:k])
        except Exception as e:       # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=200)
    assert not errors, errors
    for sel, sc, steps in out:
        np.testing.assert_array_equal(sel, want[0])
        np.testing.assert_array_equal(sc, want[1])
        np.testing.assert_array_equal(steps, want[2])
    if k <= 50:
        oracle_sel, oracle_scores = go.incremental_greedy(a, k)[:2]
        assert [int(s) for s in out[0][0]] == oracle_sel
        np.testing.assert_allclose(out[0][1], oracle_scores, rtol=1e-9)
    for lz in lazies:
        lz.close()
    for r in ranks:
        r.close()
    for s in streams:
        _ffi.call("vgp_stream_destroy", D, s)
