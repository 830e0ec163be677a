"""Launched by tests/test_gpu_greedy.py under torch.distributed.run with two ranks (one per GPU): the sharded
greedy with the peer-memory exchange across processes (CUDA IPC mailboxes) against the CPU oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oracle import greedy_oracle as go
    from vgposp_b200 import greedy
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, k = 1500, 12
    x = np.random.default_rng(5).uniform(-2, 2, (n, 3))
    ls = 0.5 * (1000.0 / n) ** (1 / 3)
    d = x[:, None, :] - x[None, :, :]
    cov = np.exp(-np.einsum("ijk,ijk->ij", d, d) / (2 * ls * ls)) + 1e-2 * np.eye(n)
    want_sel, want_scores = go.incremental_greedy_c(cov, k)
    bounds = greedy.shard_bounds(n, world)
    shard = greedy.GreedyShard(n, bounds[rank], bounds[rank + 1], k, local,
                               stream=torch.cuda.current_stream().cuda_stream)
    shard.load_cov_host(cov)
    shard.load_prec_host(go.spd_inverse(cov))
    shard.reset()
    shard.sync()
    greedy.connect_peers_torch(shard, rank, world, dist, "cuda:%d" % local)
    shard.run_peer(5)
    shard.run_peer(k - 5)
    shard.comm_status()
    sel, scores = shard.results()
    assert [int(s) for s in sel] == want_sel, (rank, sel, want_sel)
    np.testing.assert_allclose(scores, want_scores, rtol=1e-9)
    dist.barrier()
    shard.close()
    if rank == 0:
        print("PEER_WORKER_OK", [int(s) for s in sel], flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
