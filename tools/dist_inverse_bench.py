"""The distributed factorisation alone (csrc/dist.cu), under torch.distributed.run, for a list of option settings in ONE
launch: for each setting a fresh DistInverse (the options are part of its connection fingerprint), the n x n ExpQuad
covariance of the bench workload built on every rank, then `factor_inverse` (potrf + trtri, what the lazy
formulations need) and `invert` (potrf + trtri + lauum) timed by host wall clock between barriers, max over ranks.

    python -m torch.distributed.run --nproc-per-node 8 ... tools/dist_inverse_bench.py 50000 \
        dist_emulate_min=-1 dist_emulate_min=2048 gemm_emulate_slices=0

A setting is a comma-separated list of name=value; every line printed by rank 0 is one JSON object."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from vgposp_b200 import _ffi
    from vgposp_b200.dist_inverse import DistInverse
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(sys.argv[1])
    settings = sys.argv[2:] or ["dist_emulate_min=-1"]
    stream = torch.cuda.current_stream().cuda_stream
    x, amp, ls, nugget = bench.workload(n)
    xd = _ffi.DeviceArray.from_host(x, local)
    defaults = {name: _ffi.get_option(name) for name in _ffi.OPTIONS}

    def timed(fn):
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda:%d" % local)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for setting in settings:
        for name, value in defaults.items():
            _ffi.set_option(name, value)
        applied = {}
        for item in setting.split(","):
            name, _, value = item.partition("=")
            _ffi.set_option(name, int(value))
            applied[name] = int(value)
        inv = DistInverse(n, rank, world, local, stream=stream)
        inv.connect_torch(dist, "cuda:%d" % local)
        inv.fill_padding()
        out = {"n": n, "ranks": world, "options": applied}
        for rep in range(2):
            inv.build_expquad(xd.ptr, 3, amp, ls, nugget)
            out["factor_inverse_s" + ("_first" if rep == 0 else "")] = round(timed(inv.factor_inverse), 4)
        inv.build_expquad(xd.ptr, 3, amp, ls, nugget)
        out["invert_s"] = round(timed(inv.invert), 4)
        out["stats"] = inv.stats()
        inv.close()
        _ffi.workspace_trim(local)
        if rank == 0:
            print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
