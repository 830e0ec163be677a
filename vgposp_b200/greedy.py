"""Host side of the greedy mutual-information placement: handle wrapper and the sharded step protocol.

`GreedyShard` wraps one `vgp_greedy` handle (one device, columns [c0, c1) of the candidate set).
`ShardedGreedy` runs the per-selection protocol of include/vgposp.h over any number of ranks; the only
data-path collectives are two small all-gathers per selection:

    exchange 1   32 bytes per rank   (score, index, numerator, P_yy) of each rank's local winner
    exchange 2   16 n/G bytes/rank   [w_J | p_J] row segments of the winner

The local arithmetic sits behind a small engine interface (`local_best / select / segments / apply`) so
that the protocol can be exercised on CPU with world_size 2 over gloo (tests inject an oracle-backed
engine); the product engine is `GreedyShard`, i.e. the CUDA library.

Reference semantics: placement_algorithm2.py:128-145 (alg. 1), :151-219 (alg. 2).
"""
import ctypes

import numpy as np

from . import _ffi
from ._ffi import call, c_i64, c_int, c_vp
from .dist_inverse import UPLOAD_LOWER

GUARD_NUMPY = 1e-8        # placement_algorithm2.py:116,198
GUARD_TF_GRAPH = 1e-7     # snippets_a2.py:480
JITTER_TF_GRAPH = 1e-6    # snippets_a2.py:161-163


def _ptr(x):
    """Device address of an int, a torch tensor (`data_ptr()`) or a DeviceArray (`ptr`)."""
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    return x.ptr


def triangle_bounds(n, world):
    """Row slabs with equal shares of the LOWER TRIANGLE (row i has i + 1 entries): r_g = n sqrt(g / G), multiples of
    128 -- the slabs the sharded lazy path uploads, where only the lower triangle of Sigma travels."""
    b = [min(n, int(round(n * (g / world) ** 0.5 / 128.0)) * 128) for g in range(world)] + [n]
    for g in range(1, world + 1):
        b[g] = max(b[g], b[g - 1])
    return b


def shard_bounds(n, world):
    """Contiguous column ranges of the ranks: bounds[g] .. bounds[g + 1]."""
    return [(n * g) // world for g in range(world + 1)]


class GreedyShard:
    """One device's shard: owns the Sigma[:, J] and P[:, J] panels and the incremental state."""

    def __init__(self, n, c0, c1, kmax, device=0, small=GUARD_NUMPY, jitter=0.0, stream=None):
        _ffi.require_device(device)
        self.n, self.c0, self.c1, self.kmax = int(n), int(c0), int(c1), int(kmax)
        self.nloc = self.c1 - self.c0
        self.device, self.stream = device, stream
        h = c_vp()
        call("vgp_greedy_create", ctypes.byref(h), device, self.n, self.c0, self.nloc, self.kmax,
             float(small), float(jitter))
        self.handle = h.value
        cov, prec, ld, n_pad = c_vp(), c_vp(), c_i64(), c_i64()
        call("vgp_greedy_panels", self.handle, ctypes.byref(cov), ctypes.byref(prec), ctypes.byref(ld),
             ctypes.byref(n_pad))
        self.cov_ptr, self.prec_ptr, self.ld, self.n_pad = cov.value, prec.value, ld.value, n_pad.value
        self.single = self.c0 == 0 and self.nloc == self.n

    # ---- loading -----------------------------------------------------------------------------------
    def load_cov_host(self, cov_vv):
        """H2D of this shard's column panel of a host matrix [n, n] (any row stride)."""
        a = np.asarray(cov_vv)
        if a.dtype != np.float64 or a.strides[1] != 8:
            a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == (self.n, self.n), "cov_vv must be [n, n]"
        call("vgp_memcpy2d_h2d", self.device, self.cov_ptr, self.ld * 8, a.ctypes.data + self.c0 * 8,
             a.strides[0], self.nloc * 8, self.n, self.stream)
        call("vgp_stream_sync", self.device, self.stream)

    def load_prec_host(self, prec):
        a = np.ascontiguousarray(prec, dtype=np.float64)
        assert a.shape == (self.n, self.n)
        call("vgp_memcpy2d_h2d", self.device, self.prec_ptr, self.ld * 8, a.ctypes.data + self.c0 * 8,
             a.strides[0], self.nloc * 8, self.n, self.stream)
        call("vgp_stream_sync", self.device, self.stream)

    def load_prec_device(self, full_ptr, full_ld):
        """Copy columns [c0, c1) of a device-resident full precision matrix [n, full_ld] into the panel."""
        call("vgp_memcpy2d_d2d", self.device, self.prec_ptr, self.ld * 8, full_ptr + self.c0 * 8, full_ld * 8,
             self.nloc * 8, self.n, self.stream)

    def build_cov_expquad(self, x_dev_ptr, d, amplitude, length_scale, nugget):
        """Sigma[:, J] = K_expquad(X, X[J]) + nugget on the global diagonal, written into the panel."""
        call("vgp_expquad_matrix", self.device, x_dev_ptr, self.n, x_dev_ptr + self.c0 * d * 8, self.nloc, d,
             float(amplitude), float(length_scale), float(nugget), self.c0, self.cov_ptr, self.ld, self.stream)

    def factor(self):
        """P = Sigma^-1 on device (single shard only)."""
        info = c_int(0)
        call("vgp_greedy_factor", self.handle, ctypes.byref(info), self.stream)

    def reset(self):
        call("vgp_greedy_reset", self.handle, self.stream)

    def save_precision(self):
        call("vgp_greedy_save_precision", self.handle, self.stream)

    def restore_precision(self):
        call("vgp_greedy_restore_precision", self.handle, self.stream)

    def record_scores(self, enable=True):
        call("vgp_greedy_record_scores", self.handle, 1 if enable else 0)

    # ---- engine interface (device pointers in, device pointers out) ---------------------------------
    def local_best(self, rec):
        call("vgp_greedy_local_best", self.handle, _ptr(rec), self.stream)

    def select(self, recs, nrecords):
        call("vgp_greedy_select", self.handle, _ptr(recs), nrecords, self.stream)

    def segments(self, seg, seg_stride):
        call("vgp_greedy_segments", self.handle, _ptr(seg), seg_stride, self.stream)

    def apply(self, gathered, seg_stride, bounds):
        arr = (c_i64 * len(bounds))(*bounds)
        call("vgp_greedy_apply", self.handle, _ptr(gathered), seg_stride, len(bounds) - 1, arr, self.stream)

    def run(self, k):
        call("vgp_greedy_run", self.handle, int(k), self.stream)

    # ---- peer-memory exchange (one box, NVLink; replaces the two all-gathers per selection) ---------
    def comm_create(self, rank, nranks, bounds):
        """Allocate this shard's mailbox.  Returns (ipc_handle bytes [64], mailbox device address)."""
        arr = (c_i64 * len(bounds))(*bounds)
        ipc = (ctypes.c_ubyte * 64)()
        mb = c_vp()
        call("vgp_greedy_comm_create", self.handle, int(rank), int(nranks), arr, ctypes.cast(ipc, c_vp),
             ctypes.byref(mb))
        return bytes(ipc), mb.value

    def comm_connect_pointers(self, mailboxes):
        """Peers are handles of this process: their mailbox device addresses in rank order."""
        arr = (c_vp * len(mailboxes))(*mailboxes)
        call("vgp_greedy_comm_connect", self.handle, ctypes.cast(arr, c_vp), 0)

    def comm_connect_ipc(self, handles):
        """Peers are other processes: their 64-byte IPC handles in rank order."""
        blob = b"".join(handles)
        buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
        call("vgp_greedy_comm_connect", self.handle, ctypes.cast(buf, c_vp), 1)

    def run_peer(self, k):
        call("vgp_greedy_run_peer", self.handle, int(k), self.stream)

    def comm_status(self):
        err = c_int(0)
        call("vgp_greedy_comm_status", self.handle, ctypes.byref(err), self.stream)

    # ---- results -----------------------------------------------------------------------------------
    def results(self):
        count = c_i64(0)
        sel = np.full(self.kmax, -1, dtype=np.int64)
        sc = np.zeros(self.kmax, dtype=np.float64)
        call("vgp_greedy_results", self.handle, ctypes.byref(count), sel.ctypes.data, sc.ctypes.data, self.kmax,
             self.stream)
        return sel[:count.value], sc[:count.value]

    def step_scores(self):
        sel, _ = self.results()
        out = np.empty((len(sel), self.nloc), dtype=np.float64)
        if len(sel):
            call("vgp_greedy_step_scores", self.handle, out.ctypes.data, len(sel), self.stream)
        return out

    def launch_count(self):
        c = c_i64(0)
        call("vgp_greedy_launch_count", self.handle, ctypes.byref(c))
        return c.value

    def sync(self):
        call("vgp_stream_sync", self.device, self.stream)

    def close(self):
        if getattr(self, "handle", None):
            call("vgp_greedy_destroy", self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LazyGreedy:
    """Lazy-column greedy on one device (csrc/lazy.cu): mode 0 keeps P_0 = Sigma^-1 resident, mode 1 only the
    inverse Cholesky factor.  Same results as GreedyShard.run; the precision is never rewritten."""

    def __init__(self, n, kmax, device=0, small=GUARD_NUMPY, jitter=0.0, mode=0, stream=None):
        _ffi.require_device(device)
        self.n, self.kmax, self.device, self.stream, self.mode = int(n), int(kmax), device, stream, int(mode)
        h = c_vp()
        call("vgp_lazy_create", ctypes.byref(h), device, self.n, self.kmax, float(small), float(jitter), self.mode)
        self.handle = h.value
        cov, fac, ld = c_vp(), c_vp(), c_i64()
        call("vgp_lazy_matrices", self.handle, ctypes.byref(cov), ctypes.byref(fac), ctypes.byref(ld))
        self.cov_ptr, self.fac_ptr, self.ld = cov.value, fac.value, ld.value
        self.n_pad = self.ld

    @classmethod
    def from_dist(cls, dist_inverse, kmax, small=GUARD_NUMPY, jitter=0.0):
        """The sharded form (csrc/lazy.cu, vgp_lazy_create_dist) on the ranks of a connected DistInverse: the inverse
        factor is that handle's replica, the triangular matrix-vector product of every selection is split over the
        ranks.  Same calls on every rank; every rank returns the same selection."""
        self = cls.__new__(cls)
        self.n, self.kmax, self.device, self.stream, self.mode = dist_inverse.n, int(kmax), dist_inverse.device, \
            dist_inverse.stream, 1
        h = c_vp()
        call("vgp_lazy_create_dist", ctypes.byref(h), dist_inverse.handle, self.n, self.kmax, float(small), float(jitter))
        self.handle = h.value
        cov, fac, ld = c_vp(), c_vp(), c_i64()
        call("vgp_lazy_matrices", self.handle, ctypes.byref(cov), ctypes.byref(fac), ctypes.byref(ld))
        self.cov_ptr, self.fac_ptr, self.ld = cov.value, fac.value, ld.value
        self.n_pad = self.ld
        return self

    def load_cov_host(self, cov_vv):
        a = np.asarray(cov_vv)
        if a.dtype != np.float64 or a.strides[1] != 8:
            a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == (self.n, self.n), "cov_vv must be [n, n]"
        call("vgp_memcpy2d_h2d", self.device, self.cov_ptr, self.ld * 8, a.ctypes.data, a.strides[0], self.n * 8,
             self.n, self.stream)
        call("vgp_stream_sync", self.device, self.stream)

    def load_cov_device(self, src_ptr, src_ld):
        """Sigma from a device matrix [n, src_ld] (e.g. the replica of the distributed inverse before it is factorised)."""
        if src_ld == self.ld:           # one contiguous copy (3x the speed of the pitched form)
            call("vgp_memcpy_d2d", self.device, self.cov_ptr, src_ptr, self.n * self.ld * 8, self.stream)
        else:
            call("vgp_memcpy2d_d2d", self.device, self.cov_ptr, self.ld * 8, src_ptr, src_ld * 8, self.n * 8, self.n,
                 self.stream)

    def build_cov_expquad(self, x_dev_ptr, d, amplitude, length_scale, nugget):
        call("vgp_expquad_matrix", self.device, x_dev_ptr, self.n, x_dev_ptr, self.n, d, float(amplitude),
             float(length_scale), float(nugget), 0, self.cov_ptr, self.ld, self.stream)

    def factor(self):
        info = c_int(0)
        call("vgp_lazy_factor", self.handle, ctypes.byref(info), self.stream)

    def adopt_factor(self):
        call("vgp_lazy_adopt_factor", self.handle, self.stream)

    def reset(self):
        call("vgp_lazy_reset", self.handle, self.stream)

    def record_scores(self, enable=True):
        call("vgp_lazy_record_scores", self.handle, 1 if enable else 0)

    def set_local(self, cover_spatial, cutoff):
        """Algorithm 3 (snippets_a3.py:43-364): re-evaluate only the index box of half-width `cutoff` around each
        winner on the I0 x I1 x I2 grid; cutoff 0 = exact greedy."""
        i0, i1, i2 = (int(v) for v in cover_spatial)
        call("vgp_lazy_set_local", self.handle, i0, i1, i2, int(cutoff))

    def run(self, k):
        call("vgp_lazy_run", self.handle, int(k), self.stream)

    def results(self):
        count = c_i64(0)
        sel = np.full(self.kmax, -1, dtype=np.int64)
        sc = np.zeros(self.kmax, dtype=np.float64)
        call("vgp_lazy_results", self.handle, ctypes.byref(count), sel.ctypes.data, sc.ctypes.data, self.kmax,
             self.stream)
        return sel[:count.value], sc[:count.value]

    def step_scores(self):
        sel, _ = self.results()
        out = np.empty((len(sel), self.n), dtype=np.float64)
        if len(sel):
            call("vgp_lazy_step_scores", self.handle, out.ctypes.data, len(sel), self.stream)
        return out

    def launch_count(self):
        c = c_i64(0)
        call("vgp_lazy_launch_count", self.handle, ctypes.byref(c))
        return c.value

    def profile(self, enable):
        """Returns (summed trigemv ms, launches) since profiling was last enabled, then switches it."""
        ms, cnt = ctypes.c_double(0), c_i64(0)
        call("vgp_lazy_profile", self.handle, 1 if enable else 0, ctypes.byref(ms), ctypes.byref(cnt))
        return ms.value, cnt.value

    def sync(self):
        call("vgp_stream_sync", self.device, self.stream)

    def close(self):
        if getattr(self, "handle", None):
            call("vgp_lazy_destroy", self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def check_selection(sel):
    """The reference crashes with `ValueError: list.remove(x): x not in list` when no candidate scores above
    -1 (placement_algorithm2.py:144 after argmax_ returned -1); keep that failure mode."""
    if len(sel) and (np.asarray(sel) < 0).any():
        raise ValueError("list.remove(x): x not in list")


FORMULATIONS = {"auto": -1, "dense": 0, "lazy_precision": 1, "lazy_factor": 2}     # VGP_FORMULATION_*


def place_single(cov_vv, k, device=0, small=GUARD_NUMPY, jitter=0.0, want_step_scores=False, formulation="auto",
                 pinv_fallback=True, algorithm=1):
    """placement_algorithm_1/2(cov_vv, k) for a host matrix on one device: one C-ABI call
    (H2D, factorisation, k selections, D2H).  Returns (selection, scores, step_scores or None, seconds).
    formulation: "dense" (precision downdate, north-star formulation), "lazy_precision", "lazy_factor"
    (csrc/lazy.cu) or "auto" (lazy_factor when 35 k < n, else lazy_precision)."""
    _ffi.require_device(device)
    a = np.asarray(cov_vv)
    if a.ndim != 2 or a.shape[0] != a.shape[1]:
        raise ValueError("cov_vv must be a square matrix, got shape %r" % (a.shape,))
    if a.dtype != np.float64 or a.strides[1] != 8 or a.strides[0] % 8:
        a = np.ascontiguousarray(a, dtype=np.float64)
    n = a.shape[0]
    k = int(k)
    if k <= 0:
        return np.zeros(0, np.int64), np.zeros(0), (np.zeros((0, n)) if want_step_scores else None), np.zeros(4)
    if k > n:
        raise ValueError("list.remove(x): x not in list")     # the reference runs out of candidates (:144)
    sel = np.full(k, -1, dtype=np.int64)
    sc = np.zeros(k)
    steps = np.empty((k, n)) if want_step_scores else None
    secs = np.zeros(4)
    try:
        call("vgp_placement_host_ex", device, a.ctypes.data, n, a.strides[0] // 8, k, float(small), float(jitter),
             FORMULATIONS[formulation], sel.ctypes.data, sc.ctypes.data,
             steps.ctypes.data if want_step_scores else None, secs.ctypes.data)
    except _ffi.NotPositiveDefiniteError:
        if jitter != 0.0 or not pinv_fallback:
            raise
        # numerically rank-deficient covariance: the reference's pinv semantics (csrc/pinv.cu).  Raises
        # NotPositiveDefiniteError itself when the matrix is not positive semi-definite either.
        return place_single_pinv(a, k, device, small, want_step_scores, algorithm=algorithm)
    check_selection(sel)
    return sel, sc, steps, secs


def place_single_pinv(cov_vv, k, device=0, small=GUARD_NUMPY, want_step_scores=False, max_rank=0, algorithm=1):
    """placement_algorithm_<algorithm>(cov_vv, k) for a positive semi-definite matrix of numerical rank r < n through
    vgp_placement_host_pinv.  Returns (selection, scores, step_scores or None, seconds) like place_single;
    `place_single_pinv.last_rank` holds r of the last call."""
    _ffi.require_device(device)
    a = np.ascontiguousarray(np.asarray(cov_vv), dtype=np.float64)
    n = a.shape[0]
    if not np.allclose(a, a.T, rtol=0, atol=1e-10 * max(float(np.abs(np.diag(a)).max()), 1e-300)):
        raise ValueError("cov_vv must be symmetric for the pseudo-inverse path")
    k = int(k)
    sel = np.full(k, -1, dtype=np.int64)
    sc = np.zeros(k)
    steps = np.empty((k, n)) if want_step_scores else None
    secs = np.zeros(4)
    rank = _ffi.c_i64(0)
    try:
        call("vgp_placement_host_pinv", device, a.ctypes.data, n, n, k, float(small), int(algorithm), int(max_rank),
             sel.ctypes.data,
             sc.ctypes.data, steps.ctypes.data if want_step_scores else None, ctypes.byref(rank), secs.ctypes.data)
    except _ffi.VgpError as e:
        if "list.remove" in str(e):
            raise ValueError("list.remove(x): x not in list") from None
        raise
    place_single_pinv.last_rank = int(rank.value)
    return sel, sc, steps, secs


def connect_peers_torch(shard, rank, world, dist, device):
    """Give `shard` a mailbox and map every other rank's (one process per GPU, same box): the 64-byte IPC
    handles travel through one all-gather of `dist` (torch.distributed, any backend) -- plumbing, not data path."""
    import torch
    bounds = shard_bounds(shard.n, world)
    ipc, _ = shard.comm_create(rank, world, bounds)
    mine = torch.tensor(list(ipc), dtype=torch.uint8, device=device)
    everyone = torch.empty(64 * world, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(everyone, mine)
    blob = bytes(everyone.cpu().tolist())
    shard.comm_connect_ipc([blob[64 * q:64 * (q + 1)] for q in range(world)])
    dist.barrier()                       # nobody stores into a mailbox that is not mapped and zeroed yet


class ShardedPlacer:
    """placement_algorithm_1/2(cov_vv, k) over the GPUs of one box, one process per GPU: the persistent part.

    Construction is the one-time connection of the ranks (what ncclCommInitRank is to a collective): the replica of
    the distributed inverse and the greedy panels are allocated, every peer's replica and mailbox are mapped through
    CUDA IPC (which also enables peer access between the device pairs) and the padding is initialised.  `place` can
    then be called any number of times; each call is H2D of this rank's row slab, NVLink push into every replica,
    distributed inverse (csrc/dist.cu), the column panels cut from the replica, k selections with the peer-memory
    exchange and D2H of the result.  `dist` is torch.distributed (used only for the IPC handle exchange)."""

    def __init__(self, n, kmax, rank, world, dist, device, stream=None, small=GUARD_NUMPY, jitter=0.0,
                 formulation="auto"):
        """formulation: "lazy" = sharded lazy-column greedy on the inverse Cholesky factor (potrf + trtri, 2/3 of the
        inverse's flops; per selection a sharded triangular matrix-vector product, csrc/lazy.cu) -- needs the replica
        plus a copy of Sigma on every GPU (16 n^2 bytes); "dense" = full inverse + precision downdate on column panels
        (the north-star formulation; 8 n^2 + 24 n^2 / G bytes); "auto" = lazy when it fits in 150 GB."""
        import time
        from .dist_inverse import DistInverse
        t0 = time.perf_counter()
        self.n, self.kmax, self.rank, self.world = int(n), int(kmax), int(rank), int(world)
        self.device, self.stream = device, stream
        self.small, self.jitter = small, jitter
        if formulation == "auto":
            formulation = "lazy" if 16.0 * self.n * self.n < 150e9 else "dense"
        assert formulation in ("lazy", "dense")
        self.formulation = formulation
        # the row slab this rank uploads: `bounds` (equal column panels <-> equal row slabs for the dense path, equal
        # lower-triangle shares for the lazy path)
        self.bounds = shard_bounds(self.n, self.world) if formulation == "dense" else \
            triangle_bounds(self.n, self.world)
        self.r0, self.r1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        dev_name = "cuda:%d" % device
        self.inv = DistInverse(self.n, self.rank, self.world, device, stream=stream)
        self.inv.connect_torch(dist, dev_name)
        if formulation == "dense":
            self.shard = GreedyShard(self.n, self.r0, self.r1, self.kmax, device, small=small, jitter=jitter,
                                     stream=stream)
            connect_peers_torch(self.shard, self.rank, self.world, dist, dev_name)
            self.lazy = None
        else:
            self.shard = None
            self.lazy = LazyGreedy.from_dist(self.inv, self.kmax, small=small, jitter=jitter)
        self.inv.fill_padding()
        call("vgp_stream_sync", device, stream)
        dist.barrier()
        self.connect_seconds = time.perf_counter() - t0

    def place(self, cov_rows, k=None):
        """`cov_rows`: this rank's HOST row slab cov_vv[self.r0:self.r1, :] (self.bounds; a full [n, n] matrix is
        accepted too; the lazy formulation reads columns [0, r1) of it only).  Returns (selection, scores, seconds
        dict)."""
        import time
        k = self.kmax if k is None else int(k)
        assert 0 < k <= self.kmax
        n, r0, r1, inv, shard = self.n, self.r0, self.r1, self.inv, self.shard
        a = np.asarray(cov_rows)
        if a.shape[0] == n and self.world > 1:
            a = a[r0:r1]
        assert a.shape == (r1 - r0, n) and a.dtype == np.float64 and a.strides[1] == 8, \
            "cov_rows must be this rank's row slab"
        secs = {"formulation": self.formulation}
        t1 = time.perf_counter()
        # dense: whole rows (the column panels of Sigma and P are cut from them); lazy: Sigma is symmetric and both the
        # factorisation and the selection kernels read its lower triangle only
        inv.upload_rows(a, r0, r1, n if self.lazy is None else UPLOAD_LOWER)
        inv.barrier()                                     # every slab has landed in every replica
        if self.lazy is None:
            call("vgp_memcpy2d_d2d", self.device, shard.cov_ptr, shard.ld * 8, inv.ptr + r0 * 8, inv.ld * 8,
                 (r1 - r0) * 8, n, self.stream)
        else:
            self.lazy.load_cov_device(inv.ptr, inv.ld)     # Sigma stays read-only beside the factor
            if self.jitter != 0.0:
                inv.add_diag(self.jitter)
        call("vgp_stream_sync", self.device, self.stream)
        secs["h2d_push"] = time.perf_counter() - t1
        t2 = time.perf_counter()
        if self.lazy is None:
            inv.invert()
        else:
            inv.factor_inverse()
        secs["inverse"] = time.perf_counter() - t2
        t3 = time.perf_counter()
        if self.lazy is None:
            shard.load_prec_device(inv.ptr, inv.ld)
            shard.reset()
            shard.run_peer(k)
            shard.comm_status()
            sel, scores = shard.results()
        else:
            self.lazy.adopt_factor()
            self.lazy.run(k)
            sel, scores = self.lazy.results()
            sel, scores = sel[:k], scores[:k]
        secs["selections_and_d2h"] = time.perf_counter() - t3
        secs["total"] = time.perf_counter() - t1
        check_selection(sel)
        return sel, scores, secs

    def close(self):
        if self.shard is not None:
            self.shard.close()
        if self.lazy is not None:
            self.lazy.close()
        self.inv.close()


def place_sharded(cov_rows, n, k, rank, world, dist, device, stream=None, small=GUARD_NUMPY, jitter=0.0):
    """One-shot form: connect, place, disconnect.  Returns (selection, scores, seconds dict); `alloc_connect` is
    the one-time part a long-lived process pays once (ShardedPlacer)."""
    import time
    t0 = time.perf_counter()
    placer = ShardedPlacer(n, k, rank, world, dist, device, stream=stream, small=small, jitter=jitter)
    try:
        sel, scores, secs = placer.place(cov_rows, k)
    finally:
        placer.close()
    secs = dict(secs, alloc_connect=placer.connect_seconds, total=time.perf_counter() - t0)
    return sel, scores, secs


class ShardedGreedy:
    """The per-selection protocol over `world` ranks.  `engine` provides the local arithmetic, `comm`
    the two all-gathers; buffers are torch tensors on the engine's device (cuda with NCCL, cpu with gloo)."""

    def __init__(self, engine, n, rank, world, make_buffer, all_gather):
        self.engine, self.n, self.rank, self.world = engine, int(n), int(rank), int(world)
        self.bounds = shard_bounds(self.n, self.world)
        self.stride = max(b - a for a, b in zip(self.bounds[:-1], self.bounds[1:]))
        self.stride += self.stride % 2                      # keep segments 16-byte aligned
        self.all_gather = all_gather
        self.rec = make_buffer(4)                           # vgp_candidate viewed as 4 x 8 bytes
        self.recs = make_buffer(4 * self.world)
        self.seg = make_buffer(2 * self.stride)
        self.segs = make_buffer(2 * self.stride * self.world)

    def step(self):
        e = self.engine
        e.local_best(self.rec)
        if self.world > 1:
            self.all_gather(self.recs, self.rec)
            recs = self.recs
        else:
            recs = self.rec
        e.select(recs, self.world)
        e.segments(self.seg, self.stride)
        if self.world > 1:
            self.all_gather(self.segs, self.seg)
            segs = self.segs
        else:
            segs = self.seg
        e.apply(segs, self.stride, self.bounds)

    def run(self, k):
        for _ in range(int(k)):
            self.step()
