#!/usr/bin/env python
"""bench.py -- greedy mutual-information placement throughput (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[3], the one the metric is quoted on; fits one GPU):
    synthetic 3-D point cloud, n = 50 000 candidates in [-2,2]^3, ExpQuad covariance a = 1, l = 0.136,
    nugget 1e-2, k = 100 selections.
A "step" is one greedy selection: score + arg-max, numerator update, rank-1 precision downdate of the
panel this rank owns.  `value` = selections/s with Sigma and P resident in HBM (the O(n^3) inverse that seeds
P is setup and is reported separately as `setup_s`); `e2e` = the same selections/s through the one-call C-ABI
`vgp_placement_host` with the covariance in pinned HOST memory (H2D + inverse + k selections + D2H inside
the timed region).

One JSON line on stdout (rank 0).  See DESIGN.md section "Measurement" for every field.
"""
import argparse
import ctypes
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FULL = 50000
K_FULL = 100
SEED = 20261018 + 3


def workload(n):
    rng = np.random.default_rng(SEED)
    x = rng.uniform(-2.0, 2.0, (n, 3))
    length_scale = 0.5 * (1000.0 / n) ** (1.0 / 3.0)
    return x, 1.0, length_scale, 1e-2


# ------------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Polls NVML (SM clock, throttle reasons) every 20 ms on a side thread while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, device):
        self.samples, self.reasons, self.max_mhz, self.err = [], set(), None, None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = device
            if vis:
                try:
                    idx = int(vis.split(",")[device])
                except Exception:
                    idx = device
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:        # noqa: BLE001
            self.nv, self.err = None, repr(e)
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) \
                    if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception as e:    # noqa: BLE001
                self.err = repr(e)
                return
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv:
            self.thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": self.err or "no samples"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port (C/OpenMP + LAPACK) on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_greedy(n_full, n_sample, steps, warmup):
    """The incremental oracle run for real on the first n_sample points of the same cloud; per-selection
    cost is O(n^2) HBM/DRAM streaming, so selections/s at n_full = measured * (n_sample / n_full)^2."""
    from oracle import greedy_oracle as go
    from oracle import gp_oracle as gpo
    x, amp, _, nugget = workload(n_full)
    xs = x[:n_sample]
    ls = 0.5 * (1000.0 / n_sample) ** (1.0 / 3.0)          # same neighbour density as the full workload
    t0 = time.perf_counter()
    cov = np.empty((n_sample, n_sample))
    for i in range(0, n_sample, 1024):                      # blocked: bounded temporary
        cov[i:i + 1024] = gpo.expquad_matrix(xs[i:i + 1024], xs, amp, ls)
    cov[np.diag_indices(n_sample)] += nugget
    build_s = time.perf_counter() - t0
    tm = {}
    sel, _ = go.incremental_greedy_c(cov, warmup + steps, timings=tm)
    per_step = float(np.mean(tm["steps_s"][warmup:]))
    scale = (n_sample / n_full) ** 2
    cores = os.cpu_count() or 1
    threads = int(os.environ.get("OMP_NUM_THREADS", cores))
    k_call = K_FULL
    setup_scaled = tm["setup_s"] * (n_full / n_sample) ** 3          # potrf + potri are O(n^3)
    whole_call = k_call / (setup_scaled + k_call * per_step / scale)
    return {
        "value": scale / per_step, "unit": "selections/s", "cores": threads, "kind": "port",
        "whole_call_value": whole_call,
        "whole_call_note": "k=%d selections / (inverse %.1f s scaled by (n/%d)^3 = %.0f s + k scaled steps): the CPU "
                           "counterpart of `e2e` (setup inside the timed region)" % (k_call, tm["setup_s"], n_sample,
                                                                                    setup_scaled),
        "sample": "incremental oracle (C/OpenMP step + LAPACK inverse) run for real at n=%d of the same cloud, "
                  "%d selections after %d warm-up; per-selection time scaled by (n/%d)^2 to n=%d; setup "
                  "(inverse %.1f s, kernel build %.1f s at the sample size) excluded like `value`"
                  % (n_sample, steps, warmup, n_sample, n_full, tm["setup_s"], build_s),
        "ms_per_step_sample": per_step * 1e3, "ms_per_step_scaled": per_step * 1e3 / scale,
        "host": {"cpu_count": cores, "omp_threads": threads, "blas": _blas_vendor(),
                 "env": {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS")}},
    }


def _blas_vendor():
    """Which BLAS / LAPACK numpy links (the inverse of the CPU arm runs there)."""
    try:
        cfg = np.show_config(mode="dicts")
        b = cfg.get("Build Dependencies", {}).get("blas", {})
        return "%s %s" % (b.get("name"), b.get("version"))
    except Exception:      # noqa: BLE001
        return None


def run_reference(args, rank, world):
    if rank != 0:
        return
    n_sample = int(os.environ.get("VGP_BENCH_CPU_N", 8192))
    base = cpu_greedy(args.n, n_sample, args.steps, args.warmup)
    # The reference's public call is placement_algorithm_2(cov_vv, k) on a host matrix: setup + k selections.
    # That whole call is what our `e2e` times, so it is this arm's value; the steps-only rate (the counterpart
    # of our `value`, P already resident) is reported beside it.
    whole = base["whole_call_value"]
    base = dict(base, steps_only_value=base["value"], value=whole)
    line = {
        "impl": "reference", "metric": "greedy_mi_selections_per_s", "value": whole,
        "unit": "selections/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / whole, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args),
        "cpu_baseline": base,
        "steps_only": {"value": base["steps_only_value"], "unit": "selections/s",
                       "ms_per_step": base["ms_per_step_scaled"]},
        "e2e": {"value": whole, "unit": "selections/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference arm = CPU restatement of the reference's greedy (oracle/, pinned to the reference's "
                "own golden vectors); the literal reference is O(n^4)/selection and Python-only "
                "(0.063 selections/s at n=400, BASELINE.md) and cannot run at this size",
    }
    print(json.dumps(line), flush=True)


def ncu_traffic(n, nloc):
    """Per-launch DRAM bytes of the downdate kernel from the committed `ncu --set full` capture of this same
    configuration (profiles/*_traffic.json); None when no capture matches (n, nloc)."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), reverse=True):
        try:
            d = json.load(open(path))
        except Exception:      # noqa: BLE001
            continue
        if d.get("n") == n and d.get("nloc") == nloc:
            return d["traffic_bytes_per_launch"], os.path.basename(path)
    return None, None


def config_dict(args):
    _, amp, ls, nugget = workload(args.n)
    return {"workload": "greedy_mi_placement_n%d_k%d_expquad_cloud" % (args.n, args.k), "n": args.n, "k": args.k,
            "amplitude": amp, "length_scale": round(ls, 6), "nugget": nugget, "seed": SEED,
            "parallelism": "column-panel shards x%d" % args.gpus,
            "exchange": ("none" if args.gpus == 1 else
                         ("peer-memory mailboxes over NVLink (in-kernel)" if args.exchange == "peer"
                          else "2 NCCL all-gathers per selection")),
            "l2": "inputs larger than L2: every step streams the %.1f GB precision panel"
                  % (8.0 * args.n * args.n / args.gpus / 1e9)}


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    from vgposp_b200 import _ffi, greedy
    from vgposp_b200._ffi import call

    dev = local_rank
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    n, k = args.n, args.k
    x, amp, ls, nugget = workload(n)
    stream = torch.cuda.current_stream().cuda_stream
    bounds = greedy.shard_bounds(n, world)
    c0, c1 = bounds[rank], bounds[rank + 1]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if \
        os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm_peak, peak_kind = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")

    def ev():
        e = ctypes.c_void_p()
        call("vgp_event_record", dev, stream, ctypes.byref(e))
        return e

    def elapsed(a, b):
        ms = ctypes.c_float()
        call("vgp_event_elapsed_ms", dev, a, b, ctypes.byref(ms))
        return ms.value

    # ---- setup: Sigma panel from coordinates, P = Sigma^-1 (untimed, reported) -----------------------
    xd = _ffi.DeviceArray.from_host(x, dev)
    shard = greedy.GreedyShard(n, c0, c1, max(k, args.steps + args.warmup), dev, stream=stream)
    shard.build_cov_expquad(xd.ptr, 3, amp, ls, nugget)       # first launch pays CUDA's lazy module load
    e0 = ev()
    shard.build_cov_expquad(xd.ptr, 3, amp, ls, nugget)
    e1 = ev()
    build_ms = elapsed(e0, e1)
    dist_stats = None
    t0 = time.perf_counter()
    if world == 1:
        shard.factor()
    else:
        # distributed setup: every rank holds a replica, GEMM tiles are split over the ranks and stored into all
        # replicas by the GEMM epilogues over NVLink (csrc/dist.cu); each rank then keeps its column panel
        from vgposp_b200.dist_inverse import DistInverse
        inv = DistInverse(n, rank, world, dev, stream=stream)
        inv.connect_torch(dist, "cuda:%d" % dev)
        inv.fill_padding()
        inv.build_expquad(xd.ptr, 3, amp, ls, nugget)
        shard.sync()
        dist.barrier()
        t0 = time.perf_counter()
        inv.invert()
        dist_stats = inv.stats()
        factor_only_s = time.perf_counter() - t0
        shard.load_prec_device(inv.ptr, inv.ld)
        shard.reset()
        shard.sync()
        inv.close()
    shard.sync()
    factor_s = time.perf_counter() - t0
    shard.save_precision()

    if world > 1 and args.exchange == "peer":
        # the two per-selection exchanges are stores into the peers' mailboxes over NVLink (CUDA IPC), signalled
        # with release/acquire flags inside the kernels: no collective launch, k selections enqueued at once
        greedy.connect_peers_torch(shard, rank, world, dist, "cuda:%d" % dev)
        run_steps = shard.run_peer
    elif world > 1:
        def make_buffer(m):
            return torch.zeros(m, dtype=torch.float64, device="cuda:%d" % dev)

        def all_gather(dst, src):
            dist.all_gather_into_tensor(dst, src)
        driver = greedy.ShardedGreedy(shard, n, rank, world, make_buffer, all_gather)
        run_steps = driver.run
    else:
        run_steps = shard.run

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then restore the precision so that the timed region is selections 1..K -------------
    run_steps(args.warmup)
    shard.sync()
    shard.restore_precision()
    call("vgp_greedy_profile", shard.handle, 1)
    launches0 = shard.launch_count()
    barrier()
    with ClockSampler(dev) as clocks:
        a = ev()
        run_steps(args.steps)
        b = ev()
        ms = elapsed(a, b)
        barrier()
    launches = shard.launch_count() - launches0
    if world > 1 and args.exchange == "peer":
        shard.comm_status()                 # raises if any wait on a peer's flag timed out
    kms, kcount = ctypes.c_double(), ctypes.c_int64()
    call("vgp_greedy_profile_read", shard.handle, ctypes.byref(kms), ctypes.byref(kcount))
    call("vgp_greedy_profile", shard.handle, 0)
    if dist:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda:%d" % dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    sel, scores = shard.results()
    gaps_ok = bool(np.all(np.diff(scores) <= 1e-12 * np.abs(scores[:-1]))) if len(scores) > 1 else True

    # ---- e2e: host covariance in pinned memory through the one-call C-ABI -----------------------------
    e2e = None
    if world == 1 and not args.no_e2e:
        e2e = measure_e2e(args, shard, dev, sel)
    elif world > 1 and not args.no_e2e:
        e2e = measure_e2e_sharded(args, shard, rank, world, dev, dist, sel, xd, (amp, ls, nugget))
    shard.close()
    elbo_sharded = None
    if world > 1 and not args.no_elbo:
        try:
            elbo_sharded = measure_elbo(dev, False, dist=dist, rank=rank, world=world)
        except Exception as e:      # noqa: BLE001
            elbo_sharded = {"error": repr(e)}

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    nloc = c1 - c0
    algo_bytes = 16.0 * n * nloc
    kernel_ms = kms.value / max(kcount.value, 1)
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else None
    traffic, traffic_src = ncu_traffic(n, nloc)
    line = {
        "metric": "greedy_mi_selections_per_s", "value": args.steps / (ms * 1e-3), "unit": "selections/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args),
        "roofline": {"bound": "hbm", "kernel": "downdate_kernel", "achieved": achieved, "peak": hbm_peak,
                     "unit": "GB/s", "frac": (achieved / hbm_peak) if achieved else None,
                     "peak_kind": "%s copy bandwidth (MEASURED_PEAKS.json)" % peak_kind, "traffic": traffic,
                     "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms_avg": kernel_ms,
                     "kernel_launches_timed": kcount.value,
                     "kernel_share_of_step": kms.value / ms if ms > 0 else None},
        "clocks": clocks.summary(),
        "e2e": e2e,
        "gpu_launches": int(launches),
        "setup_s": {"expquad_panel_build": build_ms * 1e-3, "inverse_potrf_potri": factor_s,
                    "expquad_GBps": 8.0 * n * nloc / (build_ms * 1e-3) / 1e9 if build_ms > 0 else None,
                    "inverse_tflops": (float(n) ** 3) / factor_s / 1e12 if factor_s > 0 else None,
                    "inverse_distribution": dist_stats},
        "selection_head": [int(s) for s in sel[:8]], "scores_non_increasing": gaps_ok,
    }
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_greedy(n, int(os.environ.get("VGP_BENCH_CPU_N", 8192)), 20, 2)
    if world == 1 and not args.no_lazy:
        try:
            line["lazy_column"] = measure_lazy(args, dev, hbm_peak)
            line["lazy_column"]["selections_equal_dense"] = all(
                v["selection_head"] == line["selection_head"] for v in line["lazy_column"].values()
                if isinstance(v, dict) and "selection_head" in v)
        except Exception as e:      # noqa: BLE001
            line["lazy_column"] = {"error": repr(e)}
    if world == 1 and not args.no_elbo:
        try:
            line["elbo"] = measure_elbo(dev, not args.no_cpu)
        except Exception as e:      # noqa: BLE001 -- secondary metric: report, do not lose the headline line
            line["elbo"] = {"error": repr(e)}
    if world > 1 and not args.no_elbo and elbo_sharded is not None:
        line["elbo"] = elbo_sharded
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def measure_elbo(dev, with_cpu, n=200000, m=512, b=4096, steps=10, warmup=3, dist=None, rank=0, world=1):
    """Second metric of BASELINE.json: VGP ELBO training steps/s at configs[2] (N = 200k observations, m = 512
    inducing points, minibatch 4096, float64), reference-faithful mode: the optimal variational posterior over all
    N observations is re-derived every step (variational_Gaussian_process_example.py:68-74), so each step is a
    512 x 512 x 200k SYRK forward plus the same GEMM shape backward."""
    import torch
    import vgposp_b200.gp_functions as gpf
    from vgposp_b200._ffi import call
    rng = np.random.default_rng(SEED + 1)
    x = rng.uniform(-2.0, 2.0, (n, 3))
    y = np.sum(np.sin(2 * np.pi * x), axis=1) + 0.1 * rng.standard_normal(n)        # gp_functions.py:78-95
    z = rng.uniform(-2.0, 2.0, (m, 3))
    gpf.DEVICE = dev
    if world > 1:
        # N-axis sharding: every rank owns a slice of the observations; three sum-all-reduces per
        # step (G, v, push-through sums: ~2 MB) through NCCL on the library's own buffers; same minibatch on all ranks
        tr = gpf.VgpTrainer(x[rank::world], y[rank::world], z, b, allreduce=lambda t: dist.all_reduce(t), n_total=n)
    else:
        tr = gpf.VgpTrainer(x, y, z, b)
    xb = torch.empty((b, 3), dtype=torch.float64, device="cuda:%d" % dev)
    yb = torch.empty((b,), dtype=torch.float64, device="cuda:%d" % dev)
    xt = torch.as_tensor(x, device=xb.device)
    yt = torch.as_tensor(y, device=xb.device)
    losses, dev_ms = [], []

    def one():
        idx = torch.as_tensor(rng.integers(n, size=b), device=xb.device)              # :119
        xb.copy_(xt[idx])
        yb.copy_(yt[idx])
        torch.cuda.synchronize()
        e0, e1, ms = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_float()
        call("vgp_event_record", dev, None, ctypes.byref(e0))
        losses.append(tr.step_device(xb.data_ptr(), yb.data_ptr()))
        call("vgp_event_record", dev, None, ctypes.byref(e1))
        call("vgp_event_elapsed_ms", dev, e0, e1, ctypes.byref(ms))
        dev_ms.append(ms.value)

    for _ in range(warmup):
        one()
    l0 = tr.launch_count()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    launches = (tr.launch_count() - l0) / steps
    mp = 512
    flop = 2.0 * mp * mp * n * 2 + 2.0 * mp * mp * b * 2 + 26 * 2.0 * mp ** 3
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=xb.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    out = {"metric": "vgp_elbo_steps_per_s", "value": 1.0 / dt, "unit": "steps/s", "ms_per_step": dt * 1e3,
           "n_gpus": world, "sharding": "observations split over the ranks, 3 NCCL all-reduces per step" if world > 1 else "none",
           "config": {"workload": "vgp_elbo_train_N%d_m%d_B%d_f64_reference_faithful" % (n, m, b), "d": 3},
           "device_ms_per_step": float(np.mean(dev_ms[-steps:])),
           "flop_per_step": flop, "tflops": flop / dt / 1e12, "launches_per_step": launches,
           "loss_first_last": [losses[0], losses[-1]], "roofline_bound": "fp64 tensor pipe (DMMA)",
           "fp64_peak_tflops_cublas_dgemm_measured": 35.5}
    tr.close()
    if with_cpu:
        out["cpu_baseline"] = cpu_elbo(x, y, z, b, n_sample=20000)
    return out


def cpu_elbo(x, y, z, b, n_sample):
    """torch-CPU float64 autograd of the same step (oracle/gp_oracle_torch.py) on the first n_sample observations;
    the N-dependent part (kernel block + SYRK and their backward) dominates and is linear in N."""
    import torch
    from oracle import gp_oracle_torch as gt
    xs, ys = x[:n_sample], y[:n_sample]
    idx = np.random.default_rng(0).integers(n_sample, size=b)
    gt.loss_and_grads(0.54, 0.54, 0.54, z, xs, ys, xs[idx], ys[idx])
    t0 = time.perf_counter()
    reps = 2
    for _ in range(reps):
        gt.loss_and_grads(0.54, 0.54, 0.54, z, xs, ys, xs[idx], ys[idx])
    dt = (time.perf_counter() - t0) / reps
    scale = x.shape[0] / n_sample
    return {"value": 1.0 / (dt * scale), "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "torch CPU float64 autograd of the same step at N=%d (m=%d, B=%d), %d repetitions; time scaled "
                      "by N/%d (the step is linear in N)" % (n_sample, z.shape[0], b, reps, n_sample),
            "ms_per_step_sample": dt * 1e3}


def measure_e2e(args, shard, dev, expect_sel):
    """k selections through vgp_placement_host with Sigma in pinned host memory."""
    from vgposp_b200._ffi import call
    import psutil
    n, k = args.n, args.k
    nbytes = 8 * n * n
    if psutil.virtual_memory().available < nbytes * 1.3 + (8 << 30):
        return {"value": None, "unit": "selections/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                "note": "host has less than %.0f GB available for the pinned covariance" % (nbytes * 1.3 / 1e9)}
    host = ctypes.c_void_p()
    call("vgp_host_alloc", nbytes, ctypes.byref(host))
    try:
        # fill the host matrix from the device panel (outside the timed region)
        call("vgp_memcpy2d_d2h", dev, host, n * 8, shard.cov_ptr, shard.ld * 8, n * 8, n, shard.stream)
        shard.sync()
        shard.close()                       # free the 3 panels before the one-call path allocates its own
        sel = np.full(k, -1, dtype=np.int64)
        sc = np.zeros(k)
        secs = np.zeros(4)
        t0 = time.perf_counter()
        from vgposp_b200.greedy import FORMULATIONS
        call("vgp_placement_host_ex", dev, host, n, n, k, 1e-8, 0.0, FORMULATIONS[args.e2e_formulation],
             sel.ctypes.data, sc.ctypes.data, None, secs.ctypes.data)
        wall = time.perf_counter() - t0
    finally:
        call("vgp_host_free", host)
    same = bool(np.array_equal(sel[:len(expect_sel)], expect_sel[:k]))
    if args.e2e_formulation == "dense":
        copied = float(nbytes)
    else:       # lazy formulations copy the lower triangle in 2048-row chunks (rows [r0, r1) x columns [0, r1))
        copied = float(sum((min(r0 + 2048, n) - r0) * min(r0 + 2048, n) * 8 for r0 in range(0, n, 2048)))
    return {"value": k / secs[3], "unit": "selections/s", "h2d_bytes_per_step": copied / k, "d2h_bytes_per_step": 16,
            "seconds": {"h2d": secs[0], "factorisation": secs[1], "selections_and_d2h": secs[2],
                        "total_events": secs[3], "total_wall": wall},
            "formulation": args.e2e_formulation if args.e2e_formulation != "auto" else
            ("auto -> lazy_factor (potrf + trtri, trigemv per selection)" if 35 * k < n else
             "auto -> lazy_precision (potrf + trtri + lauum)"),
            "k": k, "selection_equals_resident_run": same,
            "api": "vgp_placement_host_ex == vgposp_b200.placement_algorithm2.placement_algorithm_1(cov_vv, k)"}


def measure_lazy(args, dev, hbm_peak):
    """Resident-state selections/s of the lazy-column formulation (csrc/lazy.cu), both modes, same workload.
    Not the headline `value` (that is the north-star dense downdate); this is what `e2e` runs on."""
    import torch
    from vgposp_b200 import _ffi, greedy
    from vgposp_b200._ffi import call
    n = args.n
    x, amp, ls, nugget = workload(n)
    xd = _ffi.DeviceArray.from_host(x, dev)
    steps, warm = args.steps, args.warmup
    out = {}
    for mode, name in ((1, "lazy_factor"), (0, "lazy_precision")):
        h = greedy.LazyGreedy(n, steps + warm, dev, mode=mode)
        h.build_cov_expquad(xd.ptr, 3, amp, ls, nugget)
        h.sync()
        t0 = time.perf_counter()
        h.factor()
        h.sync()
        factor_s = time.perf_counter() - t0
        h.run(warm)
        h.sync()
        h.reset()
        h.sync()
        l0 = h.launch_count()
        e0, e1, ms = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_float()
        call("vgp_event_record", dev, None, ctypes.byref(e0))
        h.run(steps)
        call("vgp_event_record", dev, None, ctypes.byref(e1))
        h.sync()
        call("vgp_event_elapsed_ms", dev, e0, e1, ctypes.byref(ms))
        sel, scores = h.results()
        launches = h.launch_count() - l0
        n_pad = h.n_pad
        rec = {"value": steps / (ms.value * 1e-3), "unit": "selections/s", "ms_per_step": ms.value / steps,
               "gpu_launches": int(launches), "setup_s": factor_s,
               "setup_flop": (2.0 if mode == 1 else 3.0) / 3.0 * float(n) ** 3,
               "setup_tflops": (2.0 if mode == 1 else 3.0) / 3.0 * float(n) ** 3 / factor_s / 1e12,
               "selection_head": [int(v) for v in sel[:8]]}
        if mode == 1:
            # trigemv_kernel: rows i >= y of the lower triangle of M, once: 8 * sum_{i >= y} (i + 1) bytes
            h.reset()
            h.profile(True)
            h.run(steps)
            h.sync()
            tms, cnt = h.profile(False)
            ys = np.asarray(sel[:steps - 1], dtype=np.float64)       # launch t streams the column of winner t - 1
            algo = float(np.sum(4.0 * (float(n_pad) * (n_pad + 1) - ys * (ys + 1))))
            rec["roofline"] = {"bound": "hbm", "kernel": "trigemv_kernel", "achieved": algo / (tms * 1e-3) / 1e9,
                               "peak": hbm_peak, "unit": "GB/s", "frac": algo / (tms * 1e-3) / 1e9 / hbm_peak,
                               "algorithmic_bytes_per_launch": algo / max(cnt, 1), "kernel_ms_avg": tms / max(cnt, 1),
                               "kernel_launches_timed": int(cnt),
                               "kernel_share_of_step": tms / ms.value if ms.value > 0 else None, "traffic": None}
        else:
            t = np.arange(steps, dtype=np.float64)
            algo = float(np.sum(8.0 * n * (2.0 * np.maximum(t - 1, 0) + 6.0)))
            rec["roofline"] = {"bound": "hbm (latency-limited at these sizes)", "kernel": "lazy_step_kernel",
                               "achieved": algo / (ms.value * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                               "frac": algo / (ms.value * 1e-3) / 1e9 / hbm_peak,
                               "algorithmic_bytes_per_launch": algo / steps, "kernel_ms_avg": ms.value / steps,
                               "kernel_launches_timed": steps, "traffic": None}
        h.close()
        out[name] = rec
    xd.free()
    torch.cuda.synchronize()
    return out


def measure_e2e_sharded(args, shard, rank, world, dev, dist, expect_sel, xd, kernel_params):
    """N > 1: every rank holds its ROW slab of Sigma in pinned host memory (placer.r0:r1 x n); one call to
    vgposp_b200.greedy.ShardedPlacer.place per rank does H2D, NVLink push, distributed factorisation, k selections,
    D2H."""
    import torch
    from vgposp_b200 import _ffi, greedy
    from vgposp_b200._ffi import call
    n, k = args.n, args.k
    amp, ls, nugget = kernel_params
    shard.close()
    torch.cuda.synchronize()
    dist.barrier()
    tc = time.perf_counter()
    form = {"auto": "auto", "dense": "dense"}.get(args.e2e_formulation, "lazy")
    placer = greedy.ShardedPlacer(n, k, rank, world, dist, dev, stream=shard.stream, formulation=form)
    torch.cuda.synchronize()
    dist.barrier()
    connect = time.perf_counter() - tc
    r0, r1 = placer.r0, placer.r1
    rows = r1 - r0
    host = ctypes.c_void_p()
    call("vgp_host_alloc", max(rows, 1) * n * 8, ctypes.byref(host))
    try:
        # this rank's row slab of Sigma, built on the device in chunks and read back (outside the timed region)
        slab = np.ctypeslib.as_array(ctypes.cast(host, ctypes.POINTER(ctypes.c_double)), shape=(rows, n))
        tmp = _ffi.DeviceArray((min(4096, max(rows, 1)), n), np.float64, dev)
        for a in range(r0, r1, 4096):
            h = min(4096, r1 - a)
            call("vgp_expquad_matrix", dev, xd.ptr + a * 3 * 8, h, xd.ptr, n, 3, float(amp), float(ls), float(nugget),
                 -a, tmp.ptr, n, shard.stream)
            call("vgp_memcpy2d_d2h", dev, slab.ctypes.data + (a - r0) * n * 8, n * 8, tmp.ptr, n * 8, n * 8, h,
                 shard.stream)
            call("vgp_stream_sync", dev, shard.stream)
        tmp.free()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        sel, sc, secs = placer.place(slab, k)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        # a second call through the same mappings (warm modules, allocations in place)
        dist.barrier()
        t1 = time.perf_counter()
        sel_b, _, secs_b = placer.place(slab, k)
        torch.cuda.synchronize()
        wall_b = time.perf_counter() - t1
        placer_bounds = placer.bounds
        placer.close()
    finally:
        call("vgp_host_free", host)
    t = torch.tensor([wall, connect, wall_b], dtype=torch.float64, device="cuda:%d" % dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, connect, wall_b = float(t[0].item()), float(t[1].item()), float(t[2].item())
    same = bool(np.array_equal(sel[:len(expect_sel)], expect_sel[:k])) and bool(np.array_equal(sel, sel_b))
    lazy = secs["formulation"] == "lazy"
    return {"value": k / wall, "unit": "selections/s",
            "h2d_bytes_per_step": (4.0 * n * (n + 1) if lazy else 8.0 * n * n) / k, "d2h_bytes_per_step": 16,
            "formulation": secs["formulation"], "row_slab_bounds": [int(b) for b in placer_bounds],
            "seconds": dict(secs, total_wall_max_over_ranks=wall, connect_once_max_over_ranks=connect),
            "second_call": {"value": k / wall_b, "seconds": dict(secs_b, total_wall_max_over_ranks=wall_b)},
            "cold_call_value": k / (wall + connect), "k": k, "selection_equals_resident_run": same,
            "api": "vgposp_b200.greedy.ShardedPlacer(n, k, rank, world, torch.distributed, device).place(row_slab): one "
                   "process per GPU; H2D of the row slabs, NVLink push, distributed factorisation (lazy: potrf + trtri, "
                   "the triangular matrix-vector product of every selection split over the ranks; dense: full inverse + "
                   "precision downdate on column panels), selections, D2H inside the timed region (host wall clock, max "
                   "over ranks).  The constructor (allocation, CUDA IPC mapping of the peers' replicas, peer-access "
                   "enable) is the once-per-process connection, reported as connect_once; cold_call_value = k / "
                   "(connect + place); second_call = the same call again on the connected placer"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=K_FULL)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("VGP_BENCH_N", N_FULL)),
                    help="candidate count (default: the BASELINE workload; VGP_BENCH_N overrides it under torchrun, "
                         "whose own parser claims a bare --n)")
    ap.add_argument("--k", type=int, default=None, help="selections of the e2e call (default: --steps)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: per-selection exchange through peer-memory mailboxes (default) or two NCCL all-gathers")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-elbo", action="store_true")
    ap.add_argument("--no-lazy", action="store_true")
    ap.add_argument("--e2e-formulation", default="auto", choices=["auto", "dense", "lazy_precision", "lazy_factor"])
    args = ap.parse_args()
    if args.k is None:
        args.k = args.steps
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch N>1 with torch.distributed.run (see the module docstring)")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
