// Shared helpers of libvgposp.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/vgposp.h"

namespace vgp {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

// Launch counter: every kernel launch of the library goes through VGP_LAUNCH_CHECK.
extern thread_local int64_t g_launches;

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int device);
    ~DeviceGuard();
};

// Process-wide options (vgp_set_option): the library's only global state; it never reads the environment.
int64_t option(int which);

// cudaMalloc that, when the device is out of memory, first hands the cached workspace blocks back and retries.
cudaError_t device_malloc(void **ptr, size_t bytes);
// Workspace cache of the CURRENT device: the multi-gigabyte matrices of the one-call placement paths are kept across
// calls (allocating and releasing 2 x 20 GB per call cost 0.03 - 3.5 s of host time at n = 50 000, erratically).
// cache_alloc hands out a cached block of at least `bytes` (at most 1/8 larger) or allocates; cache_free waits for
// the device and keeps the block unless the cache limit (VGP_OPT_WORKSPACE_CACHE_BYTES) would be exceeded.
cudaError_t cache_alloc(void **ptr, size_t bytes);
void cache_free(void *ptr);
size_t cache_trim_current(size_t *cached_before);
size_t emulated_release();      // emulated.cu: digit-plane workspace of the current device

constexpr int64_t TILE = 128;   // every dense matrix the library owns is padded to a multiple of this
inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

}  // namespace vgp

#define VGP_CUDA(call)                                                            \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess) return vgp::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define VGP_LAUNCH_CHECK()                                                        \
    do {                                                                          \
        ++vgp::g_launches;                                                        \
        cudaError_t e__ = cudaGetLastError();                                     \
        if (e__ != cudaSuccess) return vgp::cuda_fail(e__, "kernel launch", __FILE__, __LINE__); \
    } while (0)

#define VGP_REQUIRE(cond, ...)                                                    \
    do {                                                                          \
        if (!(cond)) {                                                            \
            vgp::set_error(__VA_ARGS__);                                          \
            return VGP_ERR_INVALID;                                               \
        }                                                                         \
    } while (0)

#define VGP_ENTER(device)                                                         \
    vgp::DeviceGuard guard__(device);                                             \
    if (!guard__.ok) return VGP_ERR_CUDA

#define VGP_TRY(call)                                                             \
    do {                                                                          \
        int s__ = (call);                                                         \
        if (s__ != VGP_OK) return s__;                                            \
    } while (0)
