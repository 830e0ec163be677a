"""ctypes binding of libvgposp.so (include/vgposp.h) plus the array plumbing around it.

Nothing here computes: every numeric result comes from the CUDA library.  If the library is missing, or no
CUDA device is visible, calls raise -- there is no CPU fallback.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvgposp.so")

VGP_OK, VGP_ERR_INVALID, VGP_ERR_CUDA, VGP_ERR_NOT_PD, VGP_ERR_NOMEM, VGP_ERR_STATE = range(6)

c_int, c_i64, c_dbl, c_vp, c_sz = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t
P = ctypes.POINTER


class VgpError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("libvgposp status %d: %s" % (status, message))
        self.status = status


class NotPositiveDefiniteError(VgpError, np.linalg.LinAlgError):
    """The covariance handed to the CUDA path is not SPD (the reference would have taken a pinv)."""


class Candidate(ctypes.Structure):          # vgp_candidate
    _fields_ = [("score", c_dbl), ("index", c_i64), ("num", c_dbl), ("pdiag", c_dbl)]


class VgpTerms(ctypes.Structure):           # vgp_vgp_terms
    _fields_ = [("loss", c_dbl), ("ll", c_dbl), ("tr1", c_dbl), ("tr2", c_dbl), ("kl", c_dbl)]


class TensorView(ctypes.Structure):         # vgp_tensor_view
    _fields_ = [("data", c_vp), ("device_type", ctypes.c_int32), ("device_id", ctypes.c_int32),
                ("ndim", ctypes.c_int32), ("dtype_code", ctypes.c_int32), ("dtype_bits", ctypes.c_int32),
                ("contiguous", ctypes.c_int32), ("shape", c_i64 * 8), ("strides", c_i64 * 8)]


# name -> argtypes, in the order of include/vgposp.h (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "vgp_abi_version": [],
    "vgp_last_error": [],
    "vgp_device_count": [P(c_int)],
    "vgp_device_info": [c_int, ctypes.c_char_p, c_int, P(c_int), P(c_sz), P(c_sz)],
    "vgp_set_option": [c_int, c_i64],
    "vgp_get_option": [c_int, P(c_i64)],
    "vgp_workspace_trim": [c_int, P(c_sz)],
    "vgp_malloc": [c_int, c_sz, P(c_vp)],
    "vgp_free": [c_int, c_vp],
    "vgp_host_alloc": [c_sz, P(c_vp)],
    "vgp_host_free": [c_vp],
    "vgp_memcpy_h2d": [c_int, c_vp, c_vp, c_sz, c_vp],
    "vgp_memcpy_d2h": [c_int, c_vp, c_vp, c_sz, c_vp],
    "vgp_memcpy_d2d": [c_int, c_vp, c_vp, c_sz, c_vp],
    "vgp_memcpy2d_h2d": [c_int, c_vp, c_sz, c_vp, c_sz, c_sz, c_sz, c_vp],
    "vgp_memcpy2d_d2h": [c_int, c_vp, c_sz, c_vp, c_sz, c_sz, c_sz, c_vp],
    "vgp_memcpy2d_d2d": [c_int, c_vp, c_sz, c_vp, c_sz, c_sz, c_sz, c_vp],
    "vgp_memset": [c_int, c_vp, c_int, c_sz, c_vp],
    "vgp_stream_create": [c_int, P(c_vp)],
    "vgp_stream_destroy": [c_int, c_vp],
    "vgp_stream_sync": [c_int, c_vp],
    "vgp_event_record": [c_int, c_vp, P(c_vp)],
    "vgp_event_elapsed_ms": [c_int, c_vp, c_vp, P(ctypes.c_float)],
    "vgp_dlpack_view": [c_vp, P(TensorView)],
    "vgp_expquad_matrix": [c_int, c_vp, c_i64, c_vp, c_i64, c_int, c_dbl, c_dbl, c_dbl, c_i64, c_vp, c_i64, c_vp],
    "vgp_kernel_matrix": [c_int, c_int, c_vp, c_i64, c_vp, c_i64, c_int, c_dbl, c_dbl, c_dbl, c_i64, c_vp, c_i64, c_vp],
    "vgp_dgemm": [c_int, c_int, c_int, c_i64, c_i64, c_i64, c_dbl, c_vp, c_i64, c_vp, c_i64, c_dbl, c_vp, c_i64, c_vp],
    "vgp_potrf": [c_int, c_vp, c_i64, c_i64, P(c_int), c_vp],
    "vgp_spd_inverse": [c_int, c_vp, c_i64, c_i64, P(c_int), c_vp],
    "vgp_trsm": [c_int, c_int, c_int, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp],
    "vgp_gp_logprob": [c_int, c_vp, c_i64, c_int, c_vp, c_dbl, c_dbl, c_dbl, c_dbl, P(c_dbl), c_vp],
    "vgp_gp_regression": [c_int, c_vp, c_i64, c_int, c_vp, c_vp, c_i64, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_vp,
                          c_vp, c_vp],
    "vgp_vgp_optimal_posterior": [c_int, c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_dbl, c_dbl, c_dbl, c_dbl, c_int,
                                  c_vp, c_vp, c_vp],
    "vgp_vgp_loss": [c_int, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_i64, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl,
                     P(VgpTerms), c_vp],
    "vgp_vgp_predict": [c_int, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_i64, c_dbl, c_dbl, c_dbl, c_dbl, c_vp, c_vp,
                        c_vp],
    "vgp_gp_logprob_k": [c_int, c_int, c_vp, c_i64, c_int, c_vp, c_dbl, c_dbl, c_dbl, c_dbl, P(c_dbl), c_vp],
    "vgp_empirical_cov": [c_int, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp],
    "vgp_cov_taper": [c_int, c_vp, c_i64, c_i64, c_vp, c_dbl, c_dbl, c_vp],
    "vgp_gp_logprob_batch_k": [c_int, c_int, c_vp, c_i64, c_int, c_vp, c_vp, c_i64, c_dbl, c_vp, c_vp],
    "vgp_gp_logprob_grad_k": [c_int, c_int, c_vp, c_i64, c_int, c_vp, c_dbl, c_dbl, c_dbl, c_dbl, P(c_dbl), c_vp, c_vp],
    "vgp_gp_regression_k": [c_int, c_int, c_vp, c_i64, c_int, c_vp, c_vp, c_i64, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl,
                            c_vp, c_vp, c_vp],
    "vgp_vgp_optimal_posterior_k": [c_int, c_int, c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_dbl, c_dbl, c_dbl, c_dbl,
                                    c_int, c_vp, c_vp, c_vp],
    "vgp_vgp_loss_k": [c_int, c_int, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_i64, c_dbl, c_dbl, c_dbl, c_dbl,
                       c_dbl, P(VgpTerms), c_vp],
    "vgp_vgp_predict_k": [c_int, c_int, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_i64, c_dbl, c_dbl, c_dbl, c_dbl, c_vp,
                          c_vp, c_vp],
    "vgp_elbo_create": [P(c_vp), c_int, c_vp, c_vp, c_i64, c_int, c_vp, c_i64, c_i64, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl,
                        c_dbl],
    "vgp_elbo_set_kernel": [c_vp, c_int],
    "vgp_elbo_destroy": [c_vp],
    "vgp_elbo_loss_grad": [c_vp, c_vp, c_vp, P(c_dbl), c_vp, c_vp, P(VgpTerms), c_vp],
    "vgp_elbo_step": [c_vp, c_vp, c_vp, P(c_dbl), c_vp],
    "vgp_elbo_set_exchange": [c_vp, c_i64, c_vp, c_vp],
    "vgp_elbo_get_params": [c_vp, c_vp, c_vp, c_vp],
    "vgp_elbo_set_params": [c_vp, c_vp, c_vp, c_vp],
    "vgp_elbo_launch_count": [c_vp, P(c_i64)],
    "vgp_greedy_create": [P(c_vp), c_int, c_i64, c_i64, c_i64, c_i64, c_dbl, c_dbl],
    "vgp_greedy_destroy": [c_vp],
    "vgp_greedy_panels": [c_vp, P(c_vp), P(c_vp), P(c_i64), P(c_i64)],
    "vgp_greedy_factor": [c_vp, P(c_int), c_vp],
    "vgp_greedy_reset": [c_vp, c_vp],
    "vgp_greedy_save_precision": [c_vp, c_vp],
    "vgp_greedy_restore_precision": [c_vp, c_vp],
    "vgp_greedy_local_best": [c_vp, c_vp, c_vp],
    "vgp_greedy_select": [c_vp, c_vp, c_int, c_vp],
    "vgp_greedy_segments": [c_vp, c_vp, c_i64, c_vp],
    "vgp_greedy_apply": [c_vp, c_vp, c_i64, c_int, P(c_i64), c_vp],
    "vgp_greedy_run": [c_vp, c_i64, c_vp],
    "vgp_dist_create": [P(c_vp), c_int, c_int, c_int, c_i64, c_vp, P(c_vp)],
    "vgp_dist_destroy": [c_vp],
    "vgp_dist_matrix": [c_vp, P(c_vp), P(c_i64)],
    "vgp_dist_connect": [c_vp, c_vp, c_int],
    "vgp_dist_push_rows": [c_vp, c_i64, c_i64, c_vp],
    "vgp_gemm_emulated": [c_int, c_int, c_int, c_i64, c_i64, c_i64, c_dbl, c_vp, c_i64, c_vp, c_i64, c_dbl, c_vp,
                          c_i64, c_int, c_int, c_vp],
    "vgp_dist_upload_rows": [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp],
    "vgp_dist_barrier": [c_vp, c_vp],
    "vgp_dist_spd_inverse": [c_vp, P(c_int), c_vp],
    "vgp_dist_factor_inverse": [c_vp, P(c_int), c_vp],
    "vgp_dist_add_diag": [c_vp, c_i64, c_dbl, c_vp],
    "vgp_lazy_create_dist": [P(c_vp), c_vp, c_i64, c_i64, c_dbl, c_dbl],
    "vgp_dist_stats": [c_vp, P(c_i64), P(c_i64)],
    "vgp_greedy_comm_create": [c_vp, c_int, c_int, P(c_i64), c_vp, P(c_vp)],
    "vgp_greedy_comm_connect": [c_vp, c_vp, c_int],
    "vgp_greedy_run_peer": [c_vp, c_i64, c_vp],
    "vgp_greedy_comm_status": [c_vp, P(c_int), c_vp],
    "vgp_greedy_results": [c_vp, P(c_i64), c_vp, c_vp, c_i64, c_vp],
    "vgp_greedy_record_scores": [c_vp, c_int],
    "vgp_greedy_step_scores": [c_vp, c_vp, c_i64, c_vp],
    "vgp_greedy_launch_count": [c_vp, P(c_i64)],
    "vgp_greedy_profile": [c_vp, c_int],
    "vgp_greedy_profile_read": [c_vp, P(c_dbl), P(c_i64)],
    "vgp_greedy_profile_step_ms": [c_vp, P(c_dbl)],
    "vgp_placement_host": [c_int, c_vp, c_i64, c_i64, c_i64, c_dbl, c_dbl, c_vp, c_vp, c_vp, c_vp],
    "vgp_placement_host_ex": [c_int, c_vp, c_i64, c_i64, c_i64, c_dbl, c_dbl, c_int, c_vp, c_vp, c_vp, c_vp],
    "vgp_placement_host_wall": [c_vp],
    "vgp_placement_host_pinv": [c_int, c_vp, c_i64, c_i64, c_i64, c_dbl, c_int, c_i64, c_vp, c_vp, c_vp, P(c_i64), c_vp],
    "vgp_lazy_create": [P(c_vp), c_int, c_i64, c_i64, c_dbl, c_dbl, c_int],
    "vgp_lazy_destroy": [c_vp],
    "vgp_lazy_matrices": [c_vp, P(c_vp), P(c_vp), P(c_i64)],
    "vgp_lazy_factor": [c_vp, P(c_int), c_vp],
    "vgp_lazy_adopt_factor": [c_vp, c_vp],
    "vgp_lazy_reset": [c_vp, c_vp],
    "vgp_lazy_run": [c_vp, c_i64, c_vp],
    "vgp_lazy_results": [c_vp, P(c_i64), c_vp, c_vp, c_i64, c_vp],
    "vgp_lazy_record_scores": [c_vp, c_int],
    "vgp_lazy_set_local": [c_vp, c_i64, c_i64, c_i64, c_i64],
    "vgp_lazy_step_scores": [c_vp, c_vp, c_i64, c_vp],
    "vgp_lazy_launch_count": [c_vp, P(c_i64)],
    "vgp_lazy_profile": [c_vp, c_int, P(c_dbl), P(c_i64)],
}
_RESTYPES = {"vgp_last_error": ctypes.c_char_p}
_NO_STATUS = {"vgp_abi_version", "vgp_last_error"}

# vgp_set_option identifiers (include/vgposp.h)
OPTIONS = {"gemm_emulate_slices": 0, "gemm_emulate_min": 1, "h2d_overlap": 2, "gemm_tile_config": 3,
           "gemm_small_below": 4, "dist_min_tiles": 5, "dist_min_k": 6, "elbo_overlap": 7, "workspace_cache_bytes": 8,
           "dist_emulate_min": 9}

_lib = None
LOADED = {"path": None, "stamp": None, "stamp_matches_sources": None}     # what load() bound; bench/tests print it


def load(build_if_missing=True):
    """Load (once) and type libvgposp.so.  Builds it in-tree when absent or stale and nvcc is available; a build
    that fails while the sources no longer match the existing binary is an error, never a silent stale load."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    if build_if_missing:
        # no-op when the in-tree library matches the sources (stamp); rebuilds after an edit
        try:
            _build.build()
        except Exception:
            if not os.path.exists(LIB_PATH) or _build.nvcc_available():
                raise                    # nvcc is here and the build failed: do not run against the old binary
    if not os.path.exists(LIB_PATH):
        raise VgpError(VGP_ERR_STATE, "libvgposp.so not built (run python -m vgposp_b200.build)")
    path = os.environ.get("VGP_LIB", LIB_PATH)                     # VGP_LIB: A/B comparisons of two builds
    built = _build.built_stamp()
    LOADED.update(path=path, stamp=built, stamp_matches_sources=(built == _build.source_stamp()) if path == LIB_PATH
                  else None)
    if path == LIB_PATH and built != _build.source_stamp():
        import warnings
        warnings.warn("libvgposp.so was built from different sources (stamp %s...): rebuild with python -m "
                      "vgposp_b200.build" % (built or "none")[:12])
    lib = ctypes.CDLL(path)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    if lib.vgp_abi_version() != 1:
        raise VgpError(VGP_ERR_STATE, "ABI version mismatch: library %d, binding 1" % lib.vgp_abi_version())
    _lib = lib
    return lib


def call(name, *args):
    """Call a status-returning entry point; raise on failure."""
    lib = load()
    status = getattr(lib, name)(*args)
    if name in _NO_STATUS:
        return status
    if status != VGP_OK:
        msg = (lib.vgp_last_error() or b"").decode("utf-8", "replace")
        if status == VGP_ERR_NOT_PD:
            raise NotPositiveDefiniteError(status, msg)
        raise VgpError(status, msg)
    return status


def set_option(name, value):
    """vgp_set_option by name (OPTIONS); returns the previous value."""
    old = c_i64(0)
    call("vgp_get_option", OPTIONS[name], ctypes.byref(old))
    call("vgp_set_option", OPTIONS[name], int(value))
    return old.value


def get_option(name):
    v = c_i64(0)
    call("vgp_get_option", OPTIONS[name], ctypes.byref(v))
    return v.value


def apply_env_options():
    """For the measurement tools only (the library itself never reads the environment): VGP_OPT_<NAME>=<int> for
    every name in OPTIONS, e.g. VGP_OPT_GEMM_EMULATE_SLICES=0.  Returns what was applied."""
    applied = {}
    for name in OPTIONS:
        v = os.environ.get("VGP_OPT_" + name.upper())
        if v not in (None, ""):
            set_option(name, int(v))
            applied[name] = int(v)
    return applied


def workspace_trim(device=0):
    """Hand the cached device workspace (matrices of the one-call placement, digit planes) back to the driver."""
    out = c_sz(0)
    call("vgp_workspace_trim", device, ctypes.byref(out))
    return out.value


def device_count():
    n = c_int(0)
    try:
        call("vgp_device_count", ctypes.byref(n))
    except VgpError:
        return 0
    return n.value


def require_device(device=0):
    if device_count() <= device:
        raise VgpError(VGP_ERR_CUDA, "no CUDA device %d visible: vgposp_b200 has no CPU fallback" % device)


# ------------------------------------------------------------------------------------------------------
# device memory
# ------------------------------------------------------------------------------------------------------
class DeviceBuffer:
    """Owning handle on a cudaMalloc allocation made through the C-ABI."""

    def __init__(self, nbytes, device=0):
        self.device = device
        self.nbytes = int(nbytes)
        ptr = c_vp()
        call("vgp_malloc", device, max(self.nbytes, 1), ctypes.byref(ptr))
        self.ptr = ptr.value

    def free(self):
        if getattr(self, "ptr", None):
            try:
                call("vgp_free", self.device, self.ptr)
            finally:
                self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p)


class _CudaArrayView:
    """Minimal __cuda_array_interface__ holder: lets torch alias a device buffer this library owns."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def torch_view_f64(ptr, count, device=0):
    """A torch float64 tensor aliasing `count` doubles at device address `ptr` (zero-copy; plumbing for
    torch.distributed collectives on the library's buffers)."""
    import torch
    return torch.as_tensor(_CudaArrayView(ptr, count), device="cuda:%d" % device)


class DeviceArray:
    """float64 / int64 row-major array in HBM: (buffer or borrowed pointer, shape).  Exposes just enough
    (`ptr`, `shape`, `ld`, copy in/out) for the host layer; it is not an array library."""

    def __init__(self, shape, dtype=np.float64, device=0, ptr=None, owner=None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.device = device
        self.size = int(np.prod(self.shape)) if self.shape else 1
        self.nbytes = self.size * self.dtype.itemsize
        if ptr is None:
            self._buf = DeviceBuffer(self.nbytes, device)
            self.ptr = self._buf.ptr
        else:
            self._buf = owner
            self.ptr = int(ptr)

    @property
    def ld(self):
        return self.shape[-1] if self.shape else 1

    @classmethod
    def from_host(cls, array, device=0, stream=None):
        a = np.ascontiguousarray(array)
        out = cls(a.shape, a.dtype, device)
        if a.nbytes:
            call("vgp_memcpy_h2d", device, out.ptr, a.ctypes.data, a.nbytes, stream)
            call("vgp_stream_sync", device, stream)
        return out

    def to_host(self, stream=None):
        out = np.empty(self.shape, self.dtype)
        if self.nbytes:
            call("vgp_memcpy_d2h", self.device, out.ctypes.data, self.ptr, self.nbytes, stream)
            call("vgp_stream_sync", self.device, stream)
        return out

    def zero_(self, stream=None):
        call("vgp_memset", self.device, self.ptr, 0, self.nbytes, stream)
        return self

    def free(self):
        if self._buf is not None and isinstance(self._buf, DeviceBuffer):
            self._buf.free()
        self.ptr = None


# ------------------------------------------------------------------------------------------------------
# DLPack: zero-copy views of foreign tensors (TensorFlow eager tensors, torch tensors, cupy arrays ...)
# ------------------------------------------------------------------------------------------------------
_DL_CPU, _DL_CUDA, _DL_CUDA_HOST = 1, 2, 3


def _capsule_pointer(capsule):
    get = ctypes.pythonapi.PyCapsule_GetPointer
    get.restype = c_vp
    get.argtypes = [ctypes.py_object, ctypes.c_char_p]
    name = ctypes.pythonapi.PyCapsule_GetName
    name.restype = ctypes.c_char_p
    name.argtypes = [ctypes.py_object]
    nm = name(capsule)
    if nm not in (b"dltensor", b"dltensor_versioned"):
        raise ValueError("not an unconsumed DLPack capsule: %r" % nm)
    if nm == b"dltensor_versioned":
        raise ValueError("versioned DLPack capsules are not supported; export with the legacy protocol")
    return get(capsule, nm)


class ForeignTensor:
    """Zero-copy view of a tensor exported through DLPack.  Keeps the capsule (and so the producer's
    memory) alive for as long as the view lives; the capsule is never consumed."""

    def __init__(self, obj):
        if hasattr(obj, "__dlpack__"):
            capsule = obj.__dlpack__()
        else:
            capsule = obj               # already a capsule (e.g. tf.experimental.dlpack.to_dlpack(t))
        self._capsule = capsule
        self._producer = obj
        view = TensorView()
        call("vgp_dlpack_view", _capsule_pointer(capsule), ctypes.byref(view))
        self.view = view
        self.shape = tuple(view.shape[i] for i in range(view.ndim))
        self.ptr = view.data
        self.on_device = view.device_type == _DL_CUDA
        self.device = view.device_id if self.on_device else None
        self.is_f64 = view.dtype_code == 2 and view.dtype_bits == 64
        self.contiguous = bool(view.contiguous)

    def __del__(self):
        # an unconsumed capsule's destructor calls the DLManagedTensor deleter
        self._capsule = None


def as_device_f64(x, device=0, stream=None):
    """Return (DeviceArray-like with .ptr/.shape, keepalive).  Device-resident float64 contiguous tensors
    (anything with __dlpack__, or a DLPack capsule) are used in place; host data is copied H2D."""
    if isinstance(x, DeviceArray):
        return x
    if hasattr(x, "__dlpack__") and not isinstance(x, np.ndarray) or type(x).__name__ == "PyCapsule":
        ft = ForeignTensor(x)
        if ft.on_device:
            if not (ft.is_f64 and ft.contiguous):
                raise TypeError("device tensors must be contiguous float64 to be used zero-copy")
            if ft.device != device:
                raise ValueError("tensor lives on device %s, call targets device %d" % (ft.device, device))
            return DeviceArray(ft.shape, np.float64, device, ptr=ft.ptr, owner=ft)
        if ft.is_f64 and ft.contiguous and ft.view.device_type in (_DL_CPU, _DL_CUDA_HOST):
            out = DeviceArray(ft.shape, np.float64, device)
            if out.nbytes:
                call("vgp_memcpy_h2d", device, out.ptr, ft.ptr, out.nbytes, stream)
                call("vgp_stream_sync", device, stream)
            return out
    return DeviceArray.from_host(np.asarray(x, dtype=np.float64), device, stream)
