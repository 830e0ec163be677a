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

from tools import workloads  # noqa: E402

N_FULL = 50000
K_FULL = 100
SEED = workloads.SEED0 + 3


def workload(n):
    """BASELINE configs[3] / [4]: uniform cloud in [-2, 2]^3, a = 1, l = 0.5 (1000 / n)^(1/3), nugget 1e-2."""
    return workloads.cloud(n, SEED)


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
def host_threads():
    """Cores this process may use (affinity-aware) -- the thread count the CPU arm is given EXPLICITLY: a launcher's
    OMP_NUM_THREADS=1 (torch.distributed.run exports it) must not turn the CPU arm single-threaded."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample(n_full, n_sample, steps, warmup, keep=False):
    """The incremental oracle (C/OpenMP step + LAPACK potrf/potri) run for real on the first n_sample points of the
    same cloud, at the neighbour density of the full workload: `warmup` + `steps` selections, everything measured."""
    from oracle import greedy_oracle as go
    from threadpoolctl import threadpool_info, threadpool_limits
    threads = host_threads()
    omp = go.set_threads(threads)
    x, amp, _, nugget = workload(n_full)
    xs = x[:n_sample]
    ls = workloads.length_scale_for(n_sample)
    t0 = time.perf_counter()
    cov = workloads.expquad_cov_host(xs, amp, ls, nugget)
    build_s = time.perf_counter() - t0
    tm, all_scores = {}, []
    with threadpool_limits(limits=threads):
        blas = sorted({(p.get("internal_api"), p.get("num_threads")) for p in threadpool_info()})
        sel, scores = go.incremental_greedy_c(cov, warmup + steps, timings=tm, all_scores=all_scores)
    per_step = float(np.mean(tm["steps_s"][warmup:]))
    st = np.where(np.isnan(np.array(all_scores)), -np.inf, np.array(all_scores))
    top2 = np.partition(st, -2, axis=1)[:, -2:]
    out = {"n": n_sample, "length_scale": round(ls, 6), "selections": warmup + steps, "timed_selections": steps,
           "setup_inverse_s": tm["setup_s"], "kernel_build_s": build_s, "ms_per_selection": per_step * 1e3,
           "selections_per_s_steps_only": 1.0 / per_step,
           "selections_per_s_whole_call": steps / (tm["setup_s"] + steps * per_step),
           "min_rel_top2_gap": float(np.min((top2[:, 1] - top2[:, 0]) / np.abs(top2[:, 1]))),
           "threads": {"used": threads, "openmp_step": omp, "blas_pools": blas}}
    if keep:
        out["_cov"], out["_sel"], out["_scores"] = cov, sel, scores
    return out


def host_dgemm_gflops(m=4096):
    """DGEMM rate of the host BLAS with all threads (the CPU arm's inverse cannot run faster than n^3 flop at it)."""
    from threadpoolctl import threadpool_limits
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((m, m)), rng.standard_normal((m, m))
    best = None
    with threadpool_limits(limits=host_threads()):
        for _ in range(3):
            t0 = time.perf_counter()
            a @ b
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return 2.0 * m ** 3 / best / 1e9


def cpu_extrapolate(samples, n_full, k, dgemm_gflops=None):
    """Selections/s at n_full from measured samples.  One sample: the O(n^2) per-selection and O(n^3) setup laws.  Two
    samples: also the exponents fitted between them (LAPACK is still gaining efficiency at these sizes, so the fitted
    setup exponent is below 3); the prediction kept is the one MORE favourable to the CPU, the setup floored by
    n^3 flop at the host's measured DGEMM rate (no inverse runs faster than that)."""
    big = max(samples, key=lambda r: r["n"])
    r = n_full / big["n"]
    step_s = {"theory_n2": big["ms_per_selection"] * 1e-3 * r ** 2}
    setup_s = {"theory_n3": big["setup_inverse_s"] * r ** 3}
    fitted = None
    if len(samples) > 1:
        small = min(samples, key=lambda q: q["n"])
        lr = math.log(big["n"] / small["n"])
        e_step = math.log(big["ms_per_selection"] / small["ms_per_selection"]) / lr
        e_setup = math.log(big["setup_inverse_s"] / small["setup_inverse_s"]) / lr
        # a two-point fit is noisy: used only near the algorithmic law (step: within half a power of n^2; setup:
        # between n^1.5 and n^3.5 -- the host LAPACK was measured at 200 -> 415 GFLOP/s from n = 8192 to 16 384, an
        # apparent exponent of 1.94, and the DGEMM-rate floor below bounds what a low exponent can claim); outside that
        # window the law alone is kept
        ok_step, ok_setup = abs(e_step - 2.0) <= 0.5, 1.5 <= e_setup <= 3.5
        fitted = {"step_exponent": e_step, "setup_exponent": e_setup, "between_n": [small["n"], big["n"]],
                  "step_fit_used": ok_step, "setup_fit_used": ok_setup}
        if ok_step:
            step_s["fitted"] = big["ms_per_selection"] * 1e-3 * r ** e_step
        if ok_setup:
            setup_s["fitted"] = big["setup_inverse_s"] * r ** e_setup
    step, setup = min(step_s.values()), min(setup_s.values())
    floor = None
    if dgemm_gflops:
        floor = float(n_full) ** 3 / (dgemm_gflops * 1e9)          # potrf + potri = n^3 flop
        setup_s["floor_n3_flop_at_host_dgemm_rate"] = floor
        setup = max(setup, floor)
    return {"n": n_full, "k": k, "seconds_per_selection": step, "setup_seconds": setup,
            "host_dgemm_gflops": dgemm_gflops,
            "steps_only_selections_per_s": 1.0 / step, "whole_call_selections_per_s": k / (setup + k * step),
            "candidates": {"seconds_per_selection": step_s, "setup_seconds": setup_s}, "fitted_exponents": fitted,
            "rule": "smallest predicted CPU time among the candidates (most favourable to the CPU arm)"}


def cpu_greedy(n_full, n_sample, steps, warmup, keep=False):
    """`cpu_baseline` of our arm: ONE measured sample, scaled by the n^2 law, flagged as extrapolated."""
    m = cpu_sample(n_full, n_sample, steps, warmup, keep=keep)
    ex = cpu_extrapolate([m], n_full, steps)
    out = {"value": ex["steps_only_selections_per_s"], "unit": "selections/s", "cores": m["threads"]["used"],
           "kind": "port", "extrapolated": True,
           "sample": "incremental oracle (oracle/: C/OpenMP step + LAPACK inverse, pinned to the reference's golden "
                     "vectors) run for real on the first %d points of the same cloud, %d selections after %d warm-up: "
                     "%.2f ms per selection measured; `value` = that rate scaled by (%d/%d)^2 (O(n^2) streaming per "
                     "selection); setup (inverse %.1f s at the sample size) excluded like our `value`"
                     % (n_sample, steps, warmup, m["ms_per_selection"], n_sample, n_full, m["setup_inverse_s"]),
           "measured": {k: v for k, v in m.items() if not k.startswith("_")},
           "extrapolation": ex,
           "host": {"cpu_count": os.cpu_count(), "affinity": host_threads(), "blas": _blas_vendor(),
                    "env": {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS",
                                                           "OPENBLAS_NUM_THREADS")}}}
    if keep:
        out["_sample"] = m
    return out


def _blas_vendor():
    """Which BLAS / LAPACK numpy links (the inverse of the CPU arm runs there)."""
    try:
        cfg = np.show_config(mode="dicts")
        b = cfg.get("Build Dependencies", {}).get("blas", {})
        return "%s %s" % (b.get("name"), b.get("version"))
    except Exception:      # noqa: BLE001
        return None


def run_reference(args, rank, world):
    """Reference arm: the reference's algorithm on the host cores.  The literal reference (NumPy pinv per candidate,
    O(n^4) per selection, 0.063 selections/s at n = 400: BASELINE.md) cannot run at n = 50 000, and neither can its
    O(n^2)-per-selection restatement inside a few minutes (the inverse alone is n^3 = 1.25e14 flop on the CPU): the
    restatement is therefore MEASURED at two sizes of the same cloud and the n = 50 000 figures are extrapolated from
    them, flagged as such, with the measured numbers beside them."""
    if rank != 0:
        return
    sizes = [int(v) for v in os.environ.get("VGP_BENCH_CPU_N", "8192,16384").split(",")]
    k = args.k
    cpu_sample(args.n, 1024, 2, 1)                       # thread pools up, libraries paged in: not a sample
    samples = [cpu_sample(args.n, ns, args.steps, args.warmup) for ns in sizes]
    ex = cpu_extrapolate(samples, args.n, k, dgemm_gflops=host_dgemm_gflops())
    steps_only, whole = ex["steps_only_selections_per_s"], ex["whole_call_selections_per_s"]
    line = {
        "impl": "reference", "metric": "greedy_mi_selections_per_s", "value": steps_only,
        "unit": "selections/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / steps_only, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "extrapolated": True,
        "config": config_dict(args),
        "cpu_baseline": {"value": steps_only, "unit": "selections/s", "cores": samples[0]["threads"]["used"],
                         "kind": "port", "extrapolated": True,
                         "sample": "incremental oracle (C/OpenMP step + LAPACK inverse) run for real at n = %s of the "
                                   "same cloud, %d selections after %d warm-up each; n = %d figures extrapolated "
                                   "(see `extrapolation`)" % (sizes, args.steps, args.warmup, args.n),
                         "measured": samples, "extrapolation": ex,
                         "host": {"cpu_count": os.cpu_count(), "affinity": host_threads(), "blas": _blas_vendor(),
                                  "env": {v: os.environ.get(v) for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS",
                                                                         "OPENBLAS_NUM_THREADS")}}},
        # value = steps-only rate (P resident): the counterpart of our `value`.  e2e = the whole public call
        # placement_algorithm_2(cov_vv, k): setup (inverse) + k selections, same k as our `e2e`: its counterpart.
        "e2e": {"value": whole, "unit": "selections/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "k": k,
                "extrapolated": True,
                "seconds": {"setup_inverse": ex["setup_seconds"], "selections": k * ex["seconds_per_selection"]}},
        "measured_at": [{"n": m["n"], "steps_only_selections_per_s": m["selections_per_s_steps_only"],
                         "whole_call_selections_per_s": m["selections_per_s_whole_call"],
                         "ms_per_selection": m["ms_per_selection"], "setup_inverse_s": m["setup_inverse_s"]}
                        for m in samples],
        "gpu_launches": 0,
        "note": "reference arm = CPU restatement of the reference's greedy (oracle/, pinned to the reference's own "
                "golden vectors), all host threads set explicitly.  `value` and `e2e.value` are EXTRAPOLATED to n = %d "
                "from the two measured sizes in `measured_at` (what was actually timed in this run); `value` is the "
                "steps-only rate, `e2e.value` the whole call with k = %d, like the two fields of our arm" % (args.n, k),
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
    cfg = {"workload": "greedy_mi_placement_n%d_k%d_expquad_cloud" % (args.n, args.k), "n": args.n, "k": args.k,
            "amplitude": amp, "length_scale": round(ls, 6), "nugget": nugget, "seed": SEED,
            "parallelism": "column-panel shards x%d" % args.gpus,
            "exchange": ("none" if args.gpus == 1 else
                         ("peer-memory mailboxes over NVLink (in-kernel)" if args.exchange == "peer"
                          else "2 NCCL all-gathers per selection")),
            "l2": "inputs larger than L2: every step streams the %.1f GB precision panel"
                  % (8.0 * args.n * args.n / args.gpus / 1e9)}
    if getattr(args, "option", None):
        cfg["library_options"] = list(args.option)           # non-default vgp_set_option values of this run
    return cfg


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def resident_run(args, n, rank, world, dev, dist, stream, kmax):
    """Setup (Sigma panel from coordinates, P = Sigma^-1) and the timed selections with Sigma and P resident in HBM,
    for one problem size.  Returns a dict of raw measurements plus the live GreedyShard (caller closes it)."""
    import torch
    from vgposp_b200 import _ffi, greedy
    from vgposp_b200._ffi import call
    x, amp, ls, nugget = workload(n)
    bounds = greedy.shard_bounds(n, world)
    c0, c1 = bounds[rank], bounds[rank + 1]

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
    shard = greedy.GreedyShard(n, c0, c1, kmax, dev, stream=stream)
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
        shard.load_prec_device(inv.ptr, inv.ld)
        shard.reset()
    shard.sync()
    factor_s = time.perf_counter() - t0
    if world > 1:
        inv.close()          # unmapping and freeing the replicas (0.9 s at 8 GPUs) is not part of the inverse
    shard.save_precision()

    if world > 1 and args.exchange == "peer":
        # the per-selection exchanges are stores into the peers' mailboxes over NVLink (CUDA IPC), signalled with
        # release/acquire flags inside the two step kernels: no collective launch, k selections enqueued at once
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
    # the clock sampler (NVML init + a thread) starts BEFORE the barrier: anything between the barrier and the first
    # launch is start skew between the ranks, which the earliest rank's event pair would count as step time
    with ClockSampler(dev) as clocks:
        barrier()
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
    step_ms = ctypes.c_double()
    call("vgp_greedy_profile_step_ms", shard.handle, ctypes.byref(step_ms))
    call("vgp_greedy_profile", shard.handle, 0)
    per_rank = None
    if dist:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda:%d" % dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # every rank's own kernel times (the step kernel's includes its wait for the slowest rank's record)
        mine = torch.tensor([kms.value / max(kcount.value, 1), step_ms.value / max(kcount.value, 1)],
                            dtype=torch.float64, device="cuda:%d" % dev)
        allr = torch.empty(2 * world, dtype=torch.float64, device="cuda:%d" % dev)
        dist.all_gather_into_tensor(allr, mine)
        v = allr.cpu().tolist()
        per_rank = {"downdate_ms_avg": [round(x, 5) for x in v[0::2]], "step_kernel_ms_avg": [round(x, 5) for x in v[1::2]]}
    sel, scores = shard.results()
    return {"per_rank": per_rank, "shard": shard, "xd": xd, "kernel": (amp, ls, nugget), "c0": c0, "c1": c1, "ms": ms, "kms": kms.value,
            "kcount": kcount.value, "step_kernel_ms": step_ms.value, "launches": launches, "clocks": clocks.summary(), "sel": sel, "scores": scores,
            "build_ms": build_ms, "factor_s": factor_s, "dist_stats": dist_stats}


def roofline_of(r, n, hbm_peak, peak_kind, steps):
    """`roofline` object of a resident run: the downdate kernel's algorithmic bytes (16 n nloc: the panel read and
    written once) over its mean duration from CUDA events around every launch of the timed region."""
    nloc = r["c1"] - r["c0"]
    algo_bytes = 16.0 * n * nloc
    kernel_ms = r["kms"] / max(r["kcount"], 1)
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else None
    traffic, traffic_src = ncu_traffic(n, nloc)
    out = {"bound": "hbm", "kernel": "downdate_kernel", "achieved": achieved, "peak": hbm_peak,
            "unit": "GB/s", "frac": (achieved / hbm_peak) if achieved else None,
            "peak_kind": "%s copy bandwidth (MEASURED_PEAKS.json)" % peak_kind, "traffic": traffic,
            "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms_avg": kernel_ms,
            "kernel_launches_timed": r["kcount"],
            "kernel_share_of_step": r["kms"] / r["ms"] if r["ms"] > 0 else None,
            "peer_step_kernel_ms_avg": (r["step_kernel_ms"] / max(r["kcount"], 1)) if r.get("step_kernel_ms") else None,
            "whole_step_frac": algo_bytes * steps / (r["ms"] * 1e-3) / 1e9 / hbm_peak if r["ms"] > 0 else None,
            "per_rank": r.get("per_rank")}
    if achieved and achieved > hbm_peak:
        # the denominator the contract names is a COPY (two streams, read here / write there); the downdate rewrites
        # each DRAM page right after reading it, in address order, and streams faster than that copy
        out["frac_of_nominal_hbm3e"] = achieved / 7700.0
        out["note"] = ("above the measured copy figure: an in-place read-modify-write walked in address order (one CTA per "
                       "8 rows x 512 columns, no grid-stride loop) reaches 6.97-7.0 TB/s on this box where cudaMemcpy "
                       "device-to-device reaches 6.68 and the same kernel with a grid-stride loop 6.35 "
                       "(profiles/r02_downdate_sweep.log); nominal HBM3e 7.7 TB/s")
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    from vgposp_b200 import _ffi, greedy
    from vgposp_b200._ffi import call

    dev = local_rank
    torch.cuda.set_device(dev)
    for item in args.option:                                  # every rank sets the same table (checked at connect)
        name, _, value = item.partition("=")
        _ffi.set_option(name, int(value))
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    n, k = args.n, args.k
    stream = torch.cuda.current_stream().cuda_stream
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if \
        os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm_peak, peak_kind = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")

    res = resident_run(args, n, rank, world, dev, dist, stream, max(k, args.steps + args.warmup))
    shard, xd, ms, sel, scores = res["shard"], res["xd"], res["ms"], res["sel"], res["scores"]
    amp, ls, nugget = res["kernel"]
    c0, c1 = res["c0"], res["c1"]
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

    # ---- north-star configuration (BASELINE configs[4]): n = 100 000 on 8 GPUs, measured in the same run ------------
    cfg5 = None
    if world == 8 and not args.no_cfg5 and n != 100000:
        # every rank must take the same decision (the run is collective): 80 GB replica + 3 x 10 GB panels + digit planes
        _ffi.workspace_trim(dev)
        free_b, total_b = ctypes.c_size_t(0), ctypes.c_size_t(0)
        call("vgp_device_info", dev, None, 0, None, ctypes.byref(total_b), ctypes.byref(free_b))
        t = torch.tensor([float(free_b.value)], dtype=torch.float64, device="cuda:%d" % dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if float(t.item()) < 125e9:
            cfg5 = {"skipped": "%.0f GB free on the fullest GPU, 125 GB needed" % (float(t.item()) / 1e9)}
        t = torch.tensor([float(res["factor_s"])], dtype=torch.float64, device="cuda:%d" % dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if cfg5 is None and float(t.item()) > 6.0:      # 8 x the work: keep the whole run within minutes
            cfg5 = {"skipped": "the n = 50 000 inverse took %.1f s: the n = 100 000 one would take minutes" % float(t.item())}
    if world == 8 and not args.no_cfg5 and n != 100000 and cfg5 is None:
        try:
            r5 = resident_run(args, 100000, rank, world, dev, dist, stream, args.steps + args.warmup)
            r5["shard"].close()
            r5["xd"].free()
            cfg5 = {"workload": "greedy_mi_placement_n100000_k%d_expquad_cloud (BASELINE configs[4])" % args.steps,
                    "value": args.steps / (r5["ms"] * 1e-3), "unit": "selections/s", "ms_per_step": r5["ms"] / args.steps,
                    "roofline": roofline_of(r5, 100000, hbm_peak, peak_kind, args.steps),
                    "target": "north_star: >= 70 % of the HBM roofline per selection (>= 229 selections/s)",
                    "setup_s": {"inverse_potrf_potri": r5["factor_s"], "inverse_distribution": r5["dist_stats"]},
                    "gpu_launches": int(r5["launches"]), "selection_head": [int(v) for v in r5["sel"][:8]]}
        except Exception as e:      # noqa: BLE001 -- an extra: never lose the headline line over it
            cfg5 = {"error": repr(e)}
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    nloc = c1 - c0
    line = {
        "metric": "greedy_mi_selections_per_s", "value": args.steps / (ms * 1e-3), "unit": "selections/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args),
        "roofline": roofline_of(res, n, hbm_peak, peak_kind, args.steps),
        "clocks": res["clocks"],
        "e2e": e2e,
        "gpu_launches": int(res["launches"]),
        "setup_s": {"expquad_panel_build": res["build_ms"] * 1e-3, "inverse_potrf_potri": res["factor_s"],
                    "expquad_GBps": 8.0 * n * nloc / (res["build_ms"] * 1e-3) / 1e9 if res["build_ms"] > 0 else None,
                    "inverse_tflops": (float(n) ** 3) / res["factor_s"] / 1e12 if res["factor_s"] > 0 else None,
                    "inverse_distribution": res["dist_stats"]},
        "selection_head": [int(s) for s in sel[:8]], "scores_non_increasing": gaps_ok,
    }
    if cfg5 is not None:
        line["cfg5_n100k"] = cfg5
    line["library"] = dict(_ffi.LOADED, options={k: _ffi.get_option(k) for k in _ffi.OPTIONS})
    parity_failed = False
    if world == 1 and not args.no_cpu:
        n_sample = int(os.environ.get("VGP_BENCH_CPU_N", "8192").split(",")[0])
        base = cpu_greedy(n, n_sample, args.steps, args.warmup, keep=True)
        line["parity"] = gpu_vs_cpu_parity(base.pop("_sample"), dev)
        line["cpu_baseline"] = base
        parity_failed = not line["parity"]["ok"]
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
    if parity_failed:
        raise SystemExit("PARITY FAILURE against the CPU oracle on the n = %d sample: %r" % (line["parity"]["n"],
                                                                                             line["parity"]))


def gpu_vs_cpu_parity(sample, dev):
    """The CUDA path on the SAME covariance the CPU arm just ran (first n_sample points of the workload's cloud):
    every formulation through the one-call C-ABI against the oracle's selections and winning scores."""
    from vgposp_b200 import greedy
    cov, want_sel, want_scores = sample["_cov"], sample["_sel"], sample["_scores"]
    k = len(want_sel)
    out = {"n": sample["n"], "k": k, "min_top2_gap": sample["min_rel_top2_gap"], "formulations": {}}
    worst, equal = 0.0, True
    for form in ("dense", "lazy_precision", "lazy_factor"):
        sel, scores, _, _ = greedy.place_single(cov, k, dev, formulation=form)
        same = [int(v) for v in sel] == [int(v) for v in want_sel]
        err = float(np.max(np.abs(scores - want_scores) / np.abs(want_scores)))
        out["formulations"][form] = {"selection_equal": same, "max_rel_score_err": err}
        equal, worst = equal and same, max(worst, err)
    out["selection_equal"], out["max_rel_score_err"] = equal, worst
    out["ok"] = bool(equal and worst <= 1e-9)
    out["bar"] = "selections bit-exact, winning scores within 1e-9 relative (north_star)"
    return out


def dgemm_peak_tflops(dev):
    """FP64 roofline denominator measured in this run: cuBLAS DGEMM 8192^3 through torch.matmul, best of 5."""
    import torch
    m = 8192
    a = torch.randn(m, m, dtype=torch.float64, device="cuda:%d" % dev)
    b = torch.randn(m, m, dtype=torch.float64, device="cuda:%d" % dev)
    c = torch.empty_like(a)
    best = None
    for i in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
    del a, b, c
    torch.cuda.empty_cache()
    return 2.0 * m ** 3 / (best * 1e-3) / 1e12


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
           "loss_first_last": [losses[0], losses[-1]], "roofline_bound": "fp64 tensor pipe (DMMA)"}
    tr.close()
    peak = dgemm_peak_tflops(dev)
    out["roofline"] = {"bound": "tensor", "achieved": flop / dt / 1e12 / world, "peak": peak, "unit": "TFLOP/s per GPU",
                       "frac": flop / dt / 1e12 / world / peak, "traffic": None,
                       "peak_kind": "cuBLAS DGEMM 8192^3 measured in this run (FP64 tensor pipe; MEASURED_PEAKS.json "
                                    "holds no FP64 figure)"}
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
    """k selections through vgp_placement_host_ex (== placement_algorithm_1(cov_vv, k)) with Sigma in pinned host
    memory.  `value` is HOST WALL CLOCK around the call -- allocation, H2D, factorisation, selections, D2H, release --
    of the second call (device workspace cached by the library after the first); the first call is reported as
    `cold_call_value`; CUDA-event and host breakdowns beside them."""
    from vgposp_b200._ffi import call
    from vgposp_b200.greedy import FORMULATIONS
    import psutil
    n, k = args.n, args.k
    nbytes = 8 * n * n
    if psutil.virtual_memory().available < nbytes * 1.3 + (8 << 30):
        return {"value": None, "unit": "selections/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                "note": "host has less than %.0f GB available for the pinned covariance" % (nbytes * 1.3 / 1e9)}
    host = ctypes.c_void_p()
    call("vgp_host_alloc", nbytes, ctypes.byref(host))
    calls = []
    try:
        # fill the host matrix from the device panel (outside the timed region)
        call("vgp_memcpy2d_d2h", dev, host, n * 8, shard.cov_ptr, shard.ld * 8, n * 8, n, shard.stream)
        shard.sync()
        shard.close()                       # the three panels go back (two of them into the workspace cache)
        for _ in range(2):
            sel = np.full(k, -1, dtype=np.int64)
            sc = np.zeros(k)
            secs, hw = np.zeros(4), np.zeros(4)
            t0 = time.perf_counter()
            call("vgp_placement_host_ex", dev, host, n, n, k, 1e-8, 0.0, FORMULATIONS[args.e2e_formulation],
                 sel.ctypes.data, sc.ctypes.data, None, secs.ctypes.data)
            wall = time.perf_counter() - t0
            call("vgp_placement_host_wall", hw.ctypes.data)
            calls.append({"wall": wall, "events": {"h2d": secs[0], "factorisation": secs[1],
                                                   "selections_and_d2h": secs[2], "total": secs[3]},
                          "host": {"alloc_init": hw[0], "enqueue": hw[1], "release": hw[2], "whole_call": hw[3]},
                          "sel": sel.copy()})
        pageable = None
        if args.e2e_pageable and psutil.virtual_memory().available > nbytes * 1.2 + (8 << 30):
            # the reference's call shape: a plain (pageable) NumPy matrix into placement_algorithm_1(cov_vv, k)
            import vgposp_b200.placement_algorithm2 as alg2
            alg2.DEVICE, alg2.PRINTS = dev, False
            cov_np = np.empty((n, n))
            np.copyto(cov_np, np.ctypeslib.as_array(ctypes.cast(host, ctypes.POINTER(ctypes.c_double)), shape=(n, n)))
            t0 = time.perf_counter()
            got = alg2.placement_algorithm_1(cov_np, k)
            wall = time.perf_counter() - t0
            pageable = {"value": k / wall, "wall": wall, "selection_equal": [int(v) for v in got] ==
                        [int(v) for v in calls[-1]["sel"]],
                        "api": "vgposp_b200.placement_algorithm2.placement_algorithm_1(numpy_array, k)"}
            del cov_np
    finally:
        call("vgp_host_free", host)
    warm, cold = calls[1], calls[0]
    same = all(bool(np.array_equal(c["sel"][:len(expect_sel)], expect_sel[:k])) for c in calls)
    for c in calls:
        del c["sel"]
    if args.e2e_formulation == "dense":
        copied = float(nbytes)
    else:       # lazy formulations copy the lower triangle in 2048-row chunks (rows [r0, r1) x columns [0, r1))
        copied = float(sum((min(r0 + 2048, n) - r0) * min(r0 + 2048, n) * 8 for r0 in range(0, n, 2048)))
    return {"value": k / warm["wall"], "unit": "selections/s", "h2d_bytes_per_step": copied / k, "d2h_bytes_per_step": 16,
            "clock": "host wall clock (time.perf_counter) around the C-ABI call",
            "seconds": warm, "cold_call_value": k / cold["wall"], "cold_call_seconds": cold,
            "pageable_numpy_call": pageable,
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
    return {"value": k / wall_b, "unit": "selections/s",
            "h2d_bytes_per_step": (4.0 * n * (n + 1) if lazy else 8.0 * n * n) / k, "d2h_bytes_per_step": 16,
            "clock": "host wall clock (time.perf_counter) around place(), max over ranks",
            "formulation": secs["formulation"], "row_slab_bounds": [int(b) for b in placer_bounds],
            "seconds": dict(secs_b, total_wall_max_over_ranks=wall_b),
            "first_call": {"value": k / wall, "seconds": dict(secs, total_wall_max_over_ranks=wall,
                                                               connect_once_max_over_ranks=connect)},
            "cold_call_value": k / (wall + connect), "k": k, "selection_equals_resident_run": same,
            "api": "vgposp_b200.greedy.ShardedPlacer(n, k, rank, world, torch.distributed, device).place(row_slab): one "
                   "process per GPU; H2D of the row slabs, NVLink push, distributed factorisation (lazy: potrf + trtri, "
                   "the triangular matrix-vector product of every selection split over the ranks; dense: full inverse + "
                   "precision downdate on column panels), selections, D2H inside the timed region (host wall clock, max "
                   "over ranks).  The constructor (allocation, CUDA IPC mapping of the peers' replicas, peer-access "
                   "enable) is the once-per-process connection, reported as connect_once; value = the steady-state "
                   "call (second call on the connected placer, like N = 1); first_call = the call right after "
                   "connecting; cold_call_value = k / (connect + first call)"}


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
    ap.add_argument("--option", action="append", default=[], metavar="NAME=INT",
                    help="library option for this run (vgp_set_option; e.g. gemm_emulate_slices=0), recorded in config")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-elbo", action="store_true")
    ap.add_argument("--no-lazy", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true", help="8 GPUs: skip the extra n = 100 000 measurement")
    ap.add_argument("--e2e-formulation", default="auto", choices=["auto", "dense", "lazy_precision", "lazy_factor"])
    ap.add_argument("--no-e2e-pageable", dest="e2e_pageable", action="store_false",
                    help="skip the extra e2e call on a pageable NumPy copy of the covariance (20 GB at n = 50k)")
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
