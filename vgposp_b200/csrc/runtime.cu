// Runtime part of the C-ABI: errors, memory, streams, events, DLPack views.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace vgp {

static thread_local std::string g_error;
thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    return e == cudaErrorMemoryAllocation ? VGP_ERR_NOMEM : VGP_ERR_CUDA;
}

// Scratch buffers come from the device's stream-ordered pool (cudaMallocAsync).  By default the pool hands its
// memory back to the driver at every synchronisation, so each blocking call would re-map its scratch; keep up to
// 2 GB cached (large one-off scratch, e.g. a padded copy of a covariance matrix, still goes back).
static void keep_pool_memory(int device) {
    static bool done[64] = {};
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long threshold = 2ull << 30;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    cudaGetLastError();
    done[device] = true;
}

// ---- options ------------------------------------------------------------------------------------------------------
static std::atomic<int64_t> g_options[VGP_OPT_COUNT] = {
    {8},       // VGP_OPT_GEMM_EMULATE_SLICES
    {1024},    // VGP_OPT_GEMM_EMULATE_MIN
    {1},       // VGP_OPT_H2D_OVERLAP
    {-1},      // VGP_OPT_GEMM_TILE_CONFIG
    {74},      // VGP_OPT_GEMM_SMALL_BELOW
    {96},      // VGP_OPT_DIST_MIN_TILES
    {256},     // VGP_OPT_DIST_MIN_K
    {3},       // VGP_OPT_ELBO_OVERLAP
    {-1},      // VGP_OPT_WORKSPACE_CACHE_BYTES (-1: half of the device's memory)
    {-1},      // VGP_OPT_DIST_EMULATE_MIN (-1: by rank count, see dense_gemm)
};
int64_t option(int which) { return which >= 0 && which < VGP_OPT_COUNT ? g_options[which].load() : 0; }

// ---- workspace cache ------------------------------------------------------------------------------------------------
namespace {
constexpr int MAX_DEV = 64;
constexpr size_t GRANULE = 2u << 20;
struct Block {
    void *ptr;
    size_t bytes;
};
std::mutex g_cache_mu;
std::vector<Block> g_cache_free[MAX_DEV];
std::unordered_map<void *, size_t> g_cache_live[MAX_DEV];
size_t g_cache_bytes[MAX_DEV];

int current_device() {
    int d = -1;
    return cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < MAX_DEV ? d : -1;
}
size_t cache_limit(int device) {
    const int64_t v = option(VGP_OPT_WORKSPACE_CACHE_BYTES);
    if (v >= 0) return (size_t)v;
    static size_t half_of_device[MAX_DEV] = {};      // cudaMemGetInfo costs milliseconds with tens of GB mapped: once
    if (half_of_device[device] == 0) {
        size_t f = 0, t = 0;
        if (cudaMemGetInfo(&f, &t) != cudaSuccess) return 0;
        half_of_device[device] = t / 2;
    }
    return half_of_device[device];
}
size_t trim_locked(int device) {
    size_t freed = 0;
    for (Block &b : g_cache_free[device]) {
        cudaFree(b.ptr);
        freed += b.bytes;
    }
    g_cache_free[device].clear();
    g_cache_bytes[device] = 0;
    return freed;
}
}  // namespace

size_t cache_trim_current(size_t *cached_before) {
    const int d = current_device();
    if (d < 0) return 0;
    std::lock_guard<std::mutex> lock(g_cache_mu);
    if (cached_before) *cached_before = g_cache_bytes[d];
    return trim_locked(d);
}

// An address range cudaMalloc has just handed out must not overlap anything the cache believes it owns: that would mean
// a cached block was released behind the cache's back (two owners of one range: silent corruption).  Fatal by design.
static bool overlaps_cache_locked(int d, const void *ptr, size_t bytes) {
    const char *a0 = (const char *)ptr, *a1 = a0 + bytes;
    for (const Block &b : g_cache_free[d])
        if (a0 < (const char *)b.ptr + b.bytes && (const char *)b.ptr < a1) return true;
    for (const auto &kv : g_cache_live[d])
        if (kv.first != ptr && a0 < (const char *)kv.first + kv.second && (const char *)kv.first < a1) return true;
    return false;
}

static cudaError_t checked_new_range(void *ptr, size_t bytes) {
    const int d = current_device();
    if (d < 0 || !ptr) return cudaSuccess;
    std::lock_guard<std::mutex> lock(g_cache_mu);
    g_cache_live[d].erase(ptr);           // an address cudaMalloc hands out is not (or no longer) a cache block
    if (overlaps_cache_locked(d, ptr, bytes)) {
        set_error("internal error: a new device allocation overlaps a block of the workspace cache");
        return cudaErrorUnknown;
    }
    return cudaSuccess;
}

cudaError_t device_malloc(void **ptr, size_t bytes) {
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e == cudaSuccess) return checked_new_range(*ptr, bytes);
    if (e != cudaErrorMemoryAllocation) return e;
    cudaGetLastError();
    const int d = current_device();
    if (d < 0) return e;
    {
        std::lock_guard<std::mutex> lock(g_cache_mu);
        if (trim_locked(d) == 0) return e;
    }
    e = cudaMalloc(ptr, bytes);
    if (e == cudaSuccess) return checked_new_range(*ptr, bytes);
    return e;
}

cudaError_t cache_alloc(void **ptr, size_t bytes) {
    *ptr = nullptr;
    const int d = current_device();
    if (d < 0) return cudaErrorInvalidDevice;
    bytes = (bytes + GRANULE - 1) / GRANULE * GRANULE;
    std::lock_guard<std::mutex> lock(g_cache_mu);
    std::vector<Block> &fr = g_cache_free[d];
    int best = -1;
    for (int i = 0; i < (int)fr.size(); ++i)
        if (fr[i].bytes >= bytes && fr[i].bytes <= bytes + bytes / 8 && (best < 0 || fr[i].bytes < fr[best].bytes)) best = i;
    if (best >= 0) {
        *ptr = fr[best].ptr;
        g_cache_live[d][*ptr] = fr[best].bytes;
        g_cache_bytes[d] -= fr[best].bytes;
        fr.erase(fr.begin() + best);
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        if (trim_locked(d) > 0) e = cudaMalloc(ptr, bytes);
    }
    if (e == cudaSuccess) {
        g_cache_live[d].erase(*ptr);
        if (overlaps_cache_locked(d, *ptr, bytes)) {
            set_error("internal error: a new device allocation overlaps a block of the workspace cache");
            return cudaErrorUnknown;
        }
        g_cache_live[d][*ptr] = bytes;
    }
    return e;
}

void cache_free(void *ptr) {
    if (!ptr) return;
    const int d = current_device();
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lock(g_cache_mu);
        if (d >= 0) {
            auto it = g_cache_live[d].find(ptr);
            if (it != g_cache_live[d].end()) {
                bytes = it->second;
                g_cache_live[d].erase(it);
            }
        }
    }
    if (bytes == 0 || d < 0) {           // not one of ours (or allocated on another device): plain release
        cudaFree(ptr);
        return;
    }
    cudaDeviceSynchronize();             // what cudaFree would have done: nothing in flight still uses the block
    std::lock_guard<std::mutex> lock(g_cache_mu);
    if (g_cache_bytes[d] + bytes <= cache_limit(d)) {
        g_cache_free[d].push_back({ptr, bytes});
        g_cache_bytes[d] += bytes;
    } else {
        cudaFree(ptr);
    }
}

DeviceGuard::DeviceGuard(int device) {
    cudaError_t e = cudaGetDevice(&prev);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaGetDevice (no CUDA device? this library has no CPU fallback)", __FILE__, __LINE__);
        prev = -1;
        return;
    }
    if (prev != device) {
        e = cudaSetDevice(device);
        if (e != cudaSuccess) {
            cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__);
            return;
        }
    }
    keep_pool_memory(device);
    ok = true;
}

DeviceGuard::~DeviceGuard() {
    if (ok && prev >= 0) {
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
}

}  // namespace vgp

using namespace vgp;

extern "C" {

int vgp_abi_version(void) { return VGP_ABI_VERSION; }

const char *vgp_last_error(void) { return g_error.c_str(); }

int vgp_device_count(int *count) {
    VGP_REQUIRE(count, "count is NULL");
    *count = 0;
    VGP_CUDA(cudaGetDeviceCount(count));
    return VGP_OK;
}

int vgp_device_info(int device, char *name, int len, int *sm_count, size_t *total_bytes, size_t *free_bytes) {
    VGP_ENTER(device);
    cudaDeviceProp prop;
    VGP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (name && len > 0) {
        strncpy(name, prop.name, len - 1);
        name[len - 1] = 0;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    size_t f = 0, t = 0;
    VGP_CUDA(cudaMemGetInfo(&f, &t));
    if (total_bytes) *total_bytes = t;
    if (free_bytes) *free_bytes = f;
    return VGP_OK;
}

int vgp_set_option(int option_id, int64_t value) {
    VGP_REQUIRE(option_id >= 0 && option_id < VGP_OPT_COUNT, "unknown option %d", option_id);
    switch (option_id) {
        case VGP_OPT_GEMM_EMULATE_SLICES:
            VGP_REQUIRE(value == 0 || (value >= 2 && value <= 8), "digit planes must be 0 (off) or 2..8");
            break;
        case VGP_OPT_GEMM_EMULATE_MIN:
            VGP_REQUIRE(value >= 128, "smallest emulated product must be >= 128");
            break;
        case VGP_OPT_DIST_EMULATE_MIN:
            VGP_REQUIRE(value == -1 || value >= 128, "smallest distributed emulated product: -1 (auto) or >= 128");
            break;
        case VGP_OPT_GEMM_TILE_CONFIG:
            VGP_REQUIRE(value >= -1 && value <= 2, "tile configuration: -1 auto, 0 base, 1 pair, 2 tma");
            break;
        case VGP_OPT_DIST_MIN_TILES:
        case VGP_OPT_DIST_MIN_K:
        case VGP_OPT_GEMM_SMALL_BELOW:
            VGP_REQUIRE(value >= 0, "negative threshold");
            break;
        default:
            break;
    }
    g_options[option_id].store(value);
    return VGP_OK;
}

int vgp_get_option(int option_id, int64_t *value) {
    VGP_REQUIRE(option_id >= 0 && option_id < VGP_OPT_COUNT && value, "unknown option %d", option_id);
    *value = g_options[option_id].load();
    return VGP_OK;
}

int vgp_workspace_trim(int device, size_t *released_bytes) {
    VGP_ENTER(device);
    VGP_CUDA(cudaDeviceSynchronize());
    const size_t freed = cache_trim_current(nullptr) + emulated_release();
    if (released_bytes) *released_bytes = freed;
    return VGP_OK;
}

int vgp_malloc(int device, size_t bytes, void **ptr_dev) {
    VGP_REQUIRE(ptr_dev, "ptr is NULL");
    VGP_ENTER(device);
    *ptr_dev = nullptr;
    VGP_CUDA(device_malloc(ptr_dev, bytes ? bytes : 1));
    return VGP_OK;
}

int vgp_free(int device, void *ptr_dev) {
    VGP_ENTER(device);
    VGP_CUDA(cudaFree(ptr_dev));
    return VGP_OK;
}

int vgp_host_alloc(size_t bytes, void **ptr_host) {
    VGP_REQUIRE(ptr_host, "ptr is NULL");
    *ptr_host = nullptr;
    VGP_CUDA(cudaHostAlloc(ptr_host, bytes ? bytes : 1, cudaHostAllocDefault));
    return VGP_OK;
}

int vgp_host_free(void *ptr_host) {
    VGP_CUDA(cudaFreeHost(ptr_host));
    return VGP_OK;
}

int vgp_memcpy_h2d(int device, void *dst_dev, const void *src_host, size_t bytes, void *stream) {
    VGP_ENTER(device);
    VGP_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return VGP_OK;
}

int vgp_memcpy_d2h(int device, void *dst_host, const void *src_dev, size_t bytes, void *stream) {
    VGP_ENTER(device);
    VGP_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return VGP_OK;
}

int vgp_memcpy_d2d(int device, void *dst_dev, const void *src_dev, size_t bytes, void *stream) {
    VGP_ENTER(device);
    VGP_CUDA(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return VGP_OK;
}

int vgp_memcpy2d_h2d(int device, void *dst_dev, size_t dpitch, const void *src_host, size_t spitch,
                     size_t width_bytes, size_t rows, void *stream) {
    VGP_ENTER(device);
    VGP_CUDA(cudaMemcpy2DAsync(dst_dev, dpitch, src_host, spitch, width_bytes, rows, cudaMemcpyHostToDevice,
                               (cudaStream_t)stream));
    return VGP_OK;
}

int vgp_memcpy2d_d2h(int device, void *dst_host, size_t dpitch, const void *src_dev, size_t spitch,
                     size_t width_bytes, size_t rows, void *stream) {
    VGP_ENTER(device);
    VGP_CUDA(cudaMemcpy2DAsync(dst_host, dpitch, src_dev, spitch, width_bytes, rows, cudaMemcpyDeviceToHost,
                               (cudaStream_t)stream));
    return VGP_OK;
}

int vgp_memcpy2d_d2d(int device, void *dst_dev, size_t dpitch, const void *src_dev, size_t spitch,
                     size_t width_bytes, size_t rows, void *stream) {
    VGP_ENTER(device);
    VGP_CUDA(cudaMemcpy2DAsync(dst_dev, dpitch, src_dev, spitch, width_bytes, rows, cudaMemcpyDeviceToDevice,
                               (cudaStream_t)stream));
    return VGP_OK;
}

int vgp_memset(int device, void *dst_dev, int value, size_t bytes, void *stream) {
    VGP_ENTER(device);
    VGP_CUDA(cudaMemsetAsync(dst_dev, value, bytes, (cudaStream_t)stream));
    return VGP_OK;
}

int vgp_stream_create(int device, void **stream) {
    VGP_REQUIRE(stream, "stream is NULL");
    VGP_ENTER(device);
    cudaStream_t s;
    VGP_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void *)s;
    return VGP_OK;
}

int vgp_stream_destroy(int device, void *stream) {
    VGP_ENTER(device);
    VGP_CUDA(cudaStreamDestroy((cudaStream_t)stream));
    return VGP_OK;
}

int vgp_stream_sync(int device, void *stream) {
    VGP_ENTER(device);
    VGP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return VGP_OK;
}

int vgp_event_record(int device, void *stream, void **event) {
    VGP_REQUIRE(event, "event is NULL");
    VGP_ENTER(device);
    cudaEvent_t ev;
    VGP_CUDA(cudaEventCreate(&ev));
    VGP_CUDA(cudaEventRecord(ev, (cudaStream_t)stream));
    *event = (void *)ev;
    return VGP_OK;
}

int vgp_event_elapsed_ms(int device, void *start_event, void *stop_event, float *ms) {
    VGP_REQUIRE(start_event && stop_event && ms, "NULL argument");
    VGP_ENTER(device);
    VGP_CUDA(cudaEventSynchronize((cudaEvent_t)stop_event));
    VGP_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start_event, (cudaEvent_t)stop_event));
    cudaEventDestroy((cudaEvent_t)start_event);
    cudaEventDestroy((cudaEvent_t)stop_event);
    return VGP_OK;
}

// ---- DLPack (dlpack.h v0.x layout, restated here so that no external header is needed) -------------
struct DLDevice_ { int32_t device_type; int32_t device_id; };
struct DLDataType_ { uint8_t code; uint8_t bits; uint16_t lanes; };
struct DLTensor_ {
    void *data;
    DLDevice_ device;
    int32_t ndim;
    DLDataType_ dtype;
    int64_t *shape;
    int64_t *strides;
    uint64_t byte_offset;
};
struct DLManagedTensor_ {
    DLTensor_ dl_tensor;
    void *manager_ctx;
    void (*deleter)(DLManagedTensor_ *);
};

int vgp_dlpack_view(const void *dl_managed_tensor, vgp_tensor_view *out) {
    VGP_REQUIRE(dl_managed_tensor && out, "NULL argument");
    const DLTensor_ &t = ((const DLManagedTensor_ *)dl_managed_tensor)->dl_tensor;
    VGP_REQUIRE(t.ndim >= 0 && t.ndim <= 8, "DLPack tensor rank %d not supported", t.ndim);
    VGP_REQUIRE(t.dtype.lanes == 1, "vector dtypes not supported");
    memset(out, 0, sizeof *out);
    out->data = (char *)t.data + t.byte_offset;
    out->device_type = t.device.device_type;
    out->device_id = t.device.device_id;
    out->ndim = t.ndim;
    out->dtype_code = t.dtype.code;
    out->dtype_bits = t.dtype.bits;
    int64_t expect = 1;
    int contiguous = 1;
    for (int i = t.ndim - 1; i >= 0; --i) {
        out->shape[i] = t.shape[i];
        out->strides[i] = t.strides ? t.strides[i] : expect;
        if (t.shape[i] > 1 && out->strides[i] != expect) contiguous = 0;
        expect *= t.shape[i];
    }
    out->contiguous = contiguous;
    return VGP_OK;
}

}  // extern "C"
