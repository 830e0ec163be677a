"""Host side of the distributed SPD inverse (csrc/dist.cu): P = Sigma^-1 spread over the GPUs of one box.

Stands for the pseudo-inverses of placement_algorithm2.py:399-413 when the candidate set is sharded
("Setup (Sigma -> P)").  One `DistInverse` per rank (process or thread); every rank owns
a full replica of the matrix, large GEMMs are split by output tile over the ranks and their epilogues store each
tile into every replica over NVLink, so no collective library is on the data path.  torch.distributed is used
only to pass the 128-byte CUDA IPC handles around (`connect_torch`).
"""
import ctypes

import numpy as np

from . import _ffi
from ._ffi import call, c_i64, c_int, c_vp

IPC_BYTES = 128
UPLOAD_LOWER = -1          # VGP_UPLOAD_LOWER


class DistInverse:
    def __init__(self, n, rank, nranks, device=0, stream=None):
        _ffi.require_device(device)
        self.n, self.rank, self.nranks, self.device, self.stream = int(n), int(rank), int(nranks), device, stream
        h = c_vp()
        ipc = (ctypes.c_ubyte * IPC_BYTES)()
        ptrs = (c_vp * 2)()
        call("vgp_dist_create", ctypes.byref(h), device, self.rank, self.nranks, self.n, ctypes.cast(ipc, c_vp), ptrs)
        self.handle = h.value
        self.ipc = bytes(ipc)
        self.pointers = (ptrs[0], ptrs[1])
        m, ld = c_vp(), c_i64()
        call("vgp_dist_matrix", self.handle, ctypes.byref(m), ctypes.byref(ld))
        self.ptr, self.ld = m.value, ld.value
        self.n_pad = self.ld

    # ---- connecting the ranks ----------------------------------------------------------------------
    def connect_pointers(self, pointers):
        """Peers are handles of this process (threads): [(matrix, flags)] device addresses in rank order."""
        flat = (c_vp * (2 * len(pointers)))(*[p for pair in pointers for p in pair])
        call("vgp_dist_connect", self.handle, ctypes.cast(flat, c_vp), 0)

    def connect_ipc(self, handles):
        """Peers are other processes: their 128-byte IPC blobs in rank order."""
        blob = b"".join(handles)
        buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
        call("vgp_dist_connect", self.handle, ctypes.cast(buf, c_vp), 1)

    def connect_torch(self, dist, device):
        import torch
        if self.nranks == 1:
            self.connect_pointers([self.pointers])
            return
        mine = torch.tensor(list(self.ipc), dtype=torch.uint8, device=device)
        everyone = torch.empty(IPC_BYTES * self.nranks, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(everyone, mine)
        blob = bytes(everyone.cpu().tolist())
        self.connect_ipc([blob[IPC_BYTES * q:IPC_BYTES * (q + 1)] for q in range(self.nranks)])
        dist.barrier()

    # ---- filling the replica -----------------------------------------------------------------------
    def fill_padding(self):
        """Zero the replica's padding and put ones on the padding diagonal."""
        n, n_pad = self.n, self.n_pad
        if n_pad == n:
            return
        call("vgp_memset", self.device, self.ptr + n * n_pad * 8, 0, (n_pad - n) * n_pad * 8, self.stream)
        zeros = np.zeros((n, n_pad - n))
        call("vgp_memcpy2d_h2d", self.device, self.ptr + n * 8, n_pad * 8, zeros.ctypes.data, (n_pad - n) * 8,
             (n_pad - n) * 8, n, self.stream)
        ones = np.ones(n_pad - n)
        call("vgp_memcpy2d_h2d", self.device, self.ptr + (n * n_pad + n) * 8, (n_pad + 1) * 8, ones.ctypes.data, 8, 8,
             n_pad - n, self.stream)
        call("vgp_stream_sync", self.device, self.stream)

    def load_host(self, cov, row0=0, row1=None):
        """H2D of rows [row0, row1) of a host matrix [n, n] into this replica."""
        a = np.asarray(cov)
        row1 = self.n if row1 is None else row1
        if row1 > row0:
            call("vgp_memcpy2d_h2d", self.device, self.ptr + row0 * self.ld * 8, self.ld * 8,
                 a.ctypes.data + row0 * a.strides[0], a.strides[0], self.n * 8, row1 - row0, self.stream)
        call("vgp_stream_sync", self.device, self.stream)

    def build_expquad(self, x_dev_ptr, d, amplitude, length_scale, nugget):
        call("vgp_expquad_matrix", self.device, x_dev_ptr, self.n, x_dev_ptr, self.n, d, float(amplitude),
             float(length_scale), float(nugget), 0, self.ptr, self.ld, self.stream)

    def push_rows(self, row0, row1):
        call("vgp_dist_push_rows", self.handle, int(row0), int(row1), self.stream)

    def upload_rows(self, rows_host, row0, row1, ncols=None):
        """Host rows [row0, row1) (array [row1 - row0, >= ncols]) into every replica; ncols = UPLOAD_LOWER for the
        lower triangle only."""
        a = np.asarray(rows_host)
        ncols = self.n if ncols is None else int(ncols)
        assert a.dtype == np.float64 and a.strides[1] == 8 and a.shape[0] == row1 - row0
        assert a.shape[1] >= (row1 if ncols == UPLOAD_LOWER else ncols)
        call("vgp_dist_upload_rows", self.handle, a.ctypes.data, a.strides[0] // 8, int(row0), int(row1), ncols,
             self.stream)

    def barrier(self):
        call("vgp_dist_barrier", self.handle, self.stream)

    # ---- the collective ----------------------------------------------------------------------------
    def invert(self):
        info = c_int(0)
        call("vgp_dist_spd_inverse", self.handle, ctypes.byref(info), self.stream)

    def factor_inverse(self):
        """potrf + trtri only: the replicas hold M = L^-1 afterwards (sharded lazy-column greedy)."""
        info = c_int(0)
        call("vgp_dist_factor_inverse", self.handle, ctypes.byref(info), self.stream)

    def add_diag(self, value):
        call("vgp_dist_add_diag", self.handle, self.n, float(value), self.stream)

    def stats(self):
        g, b = c_i64(0), c_i64(0)
        call("vgp_dist_stats", self.handle, ctypes.byref(g), ctypes.byref(b))
        return {"distributed_gemms": g.value, "barriers": b.value}

    def to_host(self):
        out = np.empty((self.n, self.n))
        call("vgp_memcpy2d_d2h", self.device, out.ctypes.data, self.n * 8, self.ptr, self.ld * 8, self.n * 8, self.n,
             self.stream)
        call("vgp_stream_sync", self.device, self.stream)
        return out

    def close(self):
        if getattr(self, "handle", None):
            call("vgp_dist_destroy", self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
