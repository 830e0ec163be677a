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
