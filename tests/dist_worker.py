"""Launched by tests/test_gpu_dist_inverse.py under torch.distributed.run with two ranks (one per GPU): the
distributed inverse across processes (replicas mapped through CUDA IPC), then the sharded greedy seeded from it."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oracle import greedy_oracle as go
    from vgposp_b200 import _ffi, greedy
    from vgposp_b200.dist_inverse import DistInverse
    _ffi.set_option("dist_min_tiles", 2)
    _ffi.set_option("dist_min_k", 256)
    _ffi.set_option("gemm_emulate_min", 512)       # small n: still exercise the int8 products, distributed
    _ffi.set_option("dist_emulate_min", 512)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, k = 2500, 10
    x = np.random.default_rng(9).uniform(-2, 2, (n, 3))
    ls = 0.5 * (1000.0 / n) ** (1 / 3)
    d = x[:, None, :] - x[None, :, :]
    cov = np.exp(-np.einsum("ijk,ijk->ij", d, d) / (2 * ls * ls)) + 1e-2 * np.eye(n)
    stream = torch.cuda.current_stream().cuda_stream
    inv = DistInverse(n, rank, world, local, stream=stream)
    inv.connect_torch(dist, "cuda:%d" % local)
    inv.fill_padding()
    bounds = greedy.shard_bounds(n, world)
    inv.load_host(cov, bounds[rank], bounds[rank + 1])          # each rank uploads its row slab only
    inv.push_rows(bounds[rank], bounds[rank + 1])
    inv.invert()
    assert inv.stats()["distributed_gemms"] > 0
    got = inv.to_host()
    # single-device inverse on this rank for the bitwise comparison
    full = _ffi.DeviceArray.from_host(cov, local)
    info = ctypes.c_int(0)
    _ffi.call("vgp_spd_inverse", local, full.ptr, n, n, ctypes.byref(info), None)
    want = full.to_host()
    np.testing.assert_array_equal(got, want)
    np.testing.assert_allclose(got @ cov, np.eye(n), atol=1e-9)
    # seed the sharded greedy from the replica
    shard = greedy.GreedyShard(n, bounds[rank], bounds[rank + 1], k, local, stream=stream)
    shard.load_cov_host(cov)
    shard.load_prec_device(inv.ptr, inv.ld)
    shard.reset()
    shard.sync()
    greedy.connect_peers_torch(shard, rank, world, dist, "cuda:%d" % local)
    shard.run_peer(k)
    shard.comm_status()
    sel, scores = shard.results()
    want_sel, want_scores = go.incremental_greedy_c(cov, k)
    assert [int(s) for s in sel] == want_sel, (rank, sel, want_sel)
    np.testing.assert_allclose(scores, want_scores, rtol=1e-9)
    dist.barrier()
    shard.close()
    inv.close()
    # the persistent one-call form: connect once, place twice (a second covariance through the same mappings)
    cov2 = cov + 0.05 * np.eye(n)
    want2, want2_scores = go.incremental_greedy_c(cov2, k - 3)
    for form in ("dense", "lazy", "auto"):
        placer = greedy.ShardedPlacer(n, k, rank, world, dist, local, stream=stream, formulation=form)
        sel1, sc1, secs = placer.place(cov[placer.r0:placer.r1], k)
        assert secs["formulation"] == ("dense" if form == "dense" else "lazy")
        assert [int(s) for s in sel1] == want_sel, (form, rank, sel1, want_sel)
        np.testing.assert_allclose(sc1, want_scores, rtol=1e-9)
        sel2, sc2, _ = placer.place(cov2, k - 3)                       # full matrix accepted too
        assert [int(s) for s in sel2] == want2, (form, rank, sel2, want2)
        np.testing.assert_allclose(sc2, want2_scores, rtol=1e-9)
        dist.barrier()
        placer.close()
    # the ELBO training step sharded over the observations: NCCL all-reduces on the library's own buffers
    import vgposp_b200.gp_functions as gpf
    gpf.DEVICE = local
    nobs, m, b = 4000, 64, 256
    rng = np.random.default_rng(3)
    xo = rng.uniform(-2, 2, (nobs, 3))
    yo = np.sum(np.sin(2 * np.pi * xo), axis=1) + 0.1 * rng.standard_normal(nobs)
    zo = rng.uniform(-2, 2, (m, 3))
    batches = [rng.integers(nobs, size=b) for _ in range(4)]
    single = gpf.VgpTrainer(xo, yo, zo, b)
    want_losses = [single.step(xo[i], yo[i]) for i in batches]
    sharded = gpf.VgpTrainer(xo[rank::world], yo[rank::world], zo, b, allreduce=lambda t: dist.all_reduce(t), n_total=nobs)
    got_losses = [sharded.step(xo[i], yo[i]) for i in batches]
    np.testing.assert_allclose(got_losses, want_losses, rtol=1e-9)
    v_mine, z_mine = sharded.variables()
    gathered = [None] * world
    dist.all_gather_object(gathered, (v_mine.tolist(), z_mine.tolist()))
    assert all(g == gathered[0] for g in gathered), "replicas diverged"
    single.close()
    sharded.close()
    dist.barrier()
    if rank == 0:
        print("DIST_WORKER_OK", [int(s) for s in sel], flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
