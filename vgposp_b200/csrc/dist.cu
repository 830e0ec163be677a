// Distributed SPD inverse over the GPUs of one box ("Setup (Sigma -> P)").
//
// The reference inverts on one CPU (np.linalg.pinv inside every denominator, placement_algorithm2.py:399-413);
// the single-GPU path here is dense_spd_inverse (potrf + trtri + lauum, dense.cu).  This file spreads that same
// recursion over G ranks, one process (or thread) per GPU:
//
//   * every rank owns a full replica [n_pad][n_pad] of the matrix (20 GB at n = 50k, 80 GB at n = 100k -- sized
//     for 180 GB of HBM3e next to the 10 + 10 GB greedy panels);
//   * every large GEMM of the recursion is split by output tiles over the ranks and its epilogue stores each
//     tile into all replicas through peer-mapped pointers (NVLink/NVSwitch) -- compute and all-gather in one
//     kernel (dense.cu, DistContext); a flag barrier in stream order separates dependent operations;
//   * diagonal-block kernels and small GEMMs run redundantly, so the replicas never diverge, and because every
//     element is produced by the same kernel with the same summation order as on one GPU, the result is
//     bitwise identical to dense_spd_inverse.
//
// Replicas and flag words are plain cudaMalloc allocations, exported with CUDA IPC for other processes.
#include <string.h>

#include <new>

#include "dense.cuh"

using namespace vgp;

#include "dist.cuh"

namespace {
constexpr size_t FLAG_BYTES = 8 * (DIST_MAX + 2);

// The options every rank must agree on: they decide which kernel produces a tile (replicas must stay bitwise equal)
// and how large the buffers behind the replica are.  Never 0 (0 = "peer has not created its handle").
unsigned long long options_fingerprint(int64_t n_pad, int64_t tail_rows) {
    unsigned long long h = 1469598103934665603ull;
    const int64_t words[] = {VGP_ABI_VERSION, n_pad, tail_rows, option(VGP_OPT_GEMM_EMULATE_SLICES),
                             option(VGP_OPT_GEMM_EMULATE_MIN), option(VGP_OPT_DIST_MIN_TILES), option(VGP_OPT_DIST_MIN_K),
                             option(VGP_OPT_GEMM_TILE_CONFIG), option(VGP_OPT_GEMM_SMALL_BELOW),
                             option(VGP_OPT_DIST_EMULATE_MIN)};
    for (int64_t w : words) h = (h ^ (unsigned long long)w) * 1099511628211ull;
    return h | 1ull;
}

int check(vgp_dist *h) {
    if (!h) {
        set_error("dist handle is NULL");
        return VGP_ERR_INVALID;
    }
    return VGP_OK;
}
}  // namespace

extern "C" {

int vgp_dist_create(vgp_dist **handle, int device, int rank, int nranks, int64_t n, void *ipc_out,
                    void **ptrs_out) {
    VGP_REQUIRE(handle, "handle is NULL");
    *handle = nullptr;
    VGP_REQUIRE(nranks >= 1 && nranks <= DIST_MAX && rank >= 0 && rank < nranks, "bad rank %d of %d (max %d ranks)",
                rank, nranks, DIST_MAX);
    VGP_REQUIRE(n > 0, "bad size");
    VGP_ENTER(device);
    vgp_dist *h = new (std::nothrow) vgp_dist();
    VGP_REQUIRE(h, "out of host memory");
    h->device = device;
    h->n_pad = round_up(n, TILE);
    h->ctx.rank = rank;
    h->ctx.nranks = nranks;
    const size_t bytes = (size_t)h->n_pad * h->n_pad * 8;
    // the tail behind the replica (peer-mapped with it): two parity buffers of [row blocks][n_pad] partial sums for the
    // sharded lazy-column greedy (lazy.cu)
    h->tail_rows = 2 * ((h->n_pad + DIST_TAIL_RB - 1) / DIST_TAIL_RB);
    cudaError_t e = device_malloc((void **)&h->matrix, bytes + (size_t)h->tail_rows * h->n_pad * 8);
    if (e == cudaSuccess) e = cudaMemset(h->matrix + (size_t)h->n_pad * h->n_pad, 0, (size_t)h->tail_rows * h->n_pad * 8);
    if (e == cudaSuccess) e = device_malloc((void **)&h->flags, FLAG_BYTES);
    if (e == cudaSuccess) e = cudaMemset(h->flags, 0, FLAG_BYTES);
    h->fingerprint = options_fingerprint(h->n_pad, h->tail_rows);
    if (e == cudaSuccess) e = cudaMemcpy(h->flags + DIST_MAX + 1, &h->fingerprint, 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        int rc = cuda_fail(e, "replica allocation", __FILE__, __LINE__);
        vgp_dist_destroy(h);
        return rc;
    }
    h->ctx.base = h->matrix;
    h->ctx.bytes = bytes;
    // thresholds: DistContext defaults (dense.cuh) -- measured on 8 GPUs; the NVLink egress of the tile stores
    // (64 KB per 128x64 tile and peer) stays below a rank's 900 GB/s share down to k = 256
    // everything that could allocate, free or load a module later happens now, before any rank can be spinning
    int rc0 = dense_preload();
    if (rc0 == VGP_OK) rc0 = h->ws.ensure(h->n_pad / TILE);
    // digit-plane workspace of the int8 products -- not with more than 4 ranks on the default rule (dense_gemm keeps
    // distributed products on the FP64 pipe there, and the local ones below the distribution thresholds are small)
    const bool int8_never = option(VGP_OPT_DIST_EMULATE_MIN) < 0 && nranks > 4;
    if (rc0 == VGP_OK && !int8_never && option(VGP_OPT_GEMM_EMULATE_SLICES) >= 2 &&
        h->n_pad >= option(VGP_OPT_GEMM_EMULATE_MIN))
        rc0 = h->ws.emu.reserve(largest_half(h->n_pad), (int)option(VGP_OPT_GEMM_EMULATE_SLICES), nullptr);
    if (rc0 != VGP_OK) {
        vgp_dist_destroy(h);
        return rc0;
    }
    if (ipc_out) {
        cudaIpcMemHandle_t a, b;
        e = cudaIpcGetMemHandle(&a, h->matrix);
        if (e == cudaSuccess) e = cudaIpcGetMemHandle(&b, h->flags);
        if (e != cudaSuccess) {
            int rc = cuda_fail(e, "cudaIpcGetMemHandle", __FILE__, __LINE__);
            vgp_dist_destroy(h);
            return rc;
        }
        memcpy(ipc_out, &a, sizeof a);
        memcpy((char *)ipc_out + sizeof a, &b, sizeof b);
    }
    if (ptrs_out) {
        ptrs_out[0] = h->matrix;
        ptrs_out[1] = h->flags;
    }
    *handle = h;
    return VGP_OK;
}

int vgp_dist_matrix(vgp_dist *h, double **matrix_dev, int64_t *ld) {
    VGP_TRY(check(h));
    if (matrix_dev) *matrix_dev = h->matrix;
    if (ld) *ld = h->n_pad;
    return VGP_OK;
}

int vgp_dist_connect(vgp_dist *h, const void *peers, int kind) {
    VGP_TRY(check(h));
    VGP_REQUIRE(peers && (kind == 0 || kind == 1), "bad arguments");
    VGP_ENTER(h->device);
    const int G = h->ctx.nranks, me = h->ctx.rank;
    for (int q = 0; q < G; ++q) {
        double *pm;
        unsigned long long *pf;
        if (q == me) {
            pm = h->matrix;
            pf = h->flags;
        } else if (kind == 0) {
            void *const *pp = (void *const *)peers;
            pm = (double *)pp[2 * q];
            pf = (unsigned long long *)pp[2 * q + 1];
            VGP_REQUIRE(pm && pf, "peer %d pointers are NULL", q);
            int peer_dev = -1;
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, pm) == cudaSuccess) peer_dev = at.device;
            if (peer_dev >= 0 && peer_dev != h->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(peer_dev, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
                cudaGetLastError();
            }
        } else {
            cudaIpcMemHandle_t a, b;
            const char *src = (const char *)peers + (size_t)q * 2 * sizeof a;
            memcpy(&a, src, sizeof a);
            memcpy(&b, src + sizeof a, sizeof b);
            void *va = nullptr, *vb = nullptr;
            VGP_CUDA(cudaIpcOpenMemHandle(&va, a, cudaIpcMemLazyEnablePeerAccess));
            VGP_CUDA(cudaIpcOpenMemHandle(&vb, b, cudaIpcMemLazyEnablePeerAccess));
            pm = (double *)va;
            pf = (unsigned long long *)vb;
        }
        if (q != me) {
            // a rank created with other options (or another size) would run other kernels on its tiles, or store into
            // buffers of another size: refuse to connect
            unsigned long long theirs = 0;
            VGP_CUDA(cudaMemcpy(&theirs, pf + DIST_MAX + 1, 8, cudaMemcpyDefault));
            if (theirs != h->fingerprint) {
                set_error("rank %d was created with different options or sizes than rank %d (vgp_set_option must be "
                          "identical on all ranks before vgp_dist_create)", q, me);
                return VGP_ERR_STATE;
            }
        }
        h->peer_matrix[q] = pm;
        h->ctx.delta[q] = pm - h->matrix;
        h->ctx.flags[q] = pf;
        VGP_REQUIRE(h->ctx.delta[q] % 2 == 0, "peer replica is not 16-byte congruent");
    }
    h->ipc = kind;
    h->connected = 1;
    return VGP_OK;
}

// Push rows [r0, r1) of this rank's replica into every other replica (peer copies), e.g. after a host upload
// that only this rank performed.  Follow with vgp_dist_barrier on all ranks.
int vgp_dist_push_rows(vgp_dist *h, int64_t r0, int64_t r1, void *stream) {
    VGP_TRY(check(h));
    VGP_REQUIRE(h->connected, "vgp_dist_connect first");
    VGP_REQUIRE(r0 >= 0 && r1 >= r0 && r1 <= h->n_pad, "bad row range");
    if (r1 == r0) return VGP_OK;
    VGP_ENTER(h->device);
    const size_t off = (size_t)r0 * h->n_pad, count = (size_t)(r1 - r0) * h->n_pad * 8;
    for (int q = 0; q < h->ctx.nranks; ++q) {
        if (q == h->ctx.rank) continue;
        VGP_CUDA(cudaMemcpyAsync(h->peer_matrix[q] + off, h->matrix + off, count, cudaMemcpyDefault,
                                 (cudaStream_t)stream));
    }
    return VGP_OK;
}

/* Rows [r0, r1), columns [0, ncols) of a HOST matrix (pinned for full speed) into this replica and on into every
 * other replica: the upload runs in row chunks on `stream`, the peer copies of a chunk follow it on a side stream while
 * the next chunk uploads, peers visited in a rank-staggered order (no two ranks start on the same destination).
 * ncols = VGP_UPLOAD_LOWER when only the lower triangle of a symmetric matrix is needed: each chunk of rows [c0, c1)
 * travels with columns [0, c1).  Returns with the side stream joined back into `stream`; follow with vgp_dist_barrier
 * on all ranks. */
int vgp_dist_upload_rows(vgp_dist *h, const double *host, int64_t host_ld, int64_t r0, int64_t r1, int64_t ncols,
                         void *stream) {
    VGP_TRY(check(h));
    VGP_REQUIRE(h->connected || h->ctx.nranks == 1, "vgp_dist_connect first");
    const bool lower = ncols == VGP_UPLOAD_LOWER;        // every row chunk [c0, c1) travels with columns [0, c1)
    if (lower) ncols = r1;
    VGP_REQUIRE(r0 >= 0 && r1 >= r0 && r1 <= h->n_pad && ncols >= 0 && ncols <= h->n_pad && host_ld >= ncols,
                "bad row / column range");
    if (r1 == r0 || ncols == 0) return VGP_OK;
    VGP_REQUIRE(host, "host matrix is NULL");
    VGP_ENTER(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (!h->push_stream) {
        VGP_CUDA(cudaStreamCreateWithFlags(&h->push_stream, cudaStreamNonBlocking));
        VGP_CUDA(cudaEventCreateWithFlags(&h->push_event, cudaEventDisableTiming));
    }
    const int64_t chunk = 1024;
    const int G = h->ctx.nranks, me = h->ctx.rank;
    const size_t pitch = (size_t)h->n_pad * 8;
    for (int64_t c0 = r0; c0 < r1; c0 += chunk) {
        const int64_t rows = (r1 - c0 < chunk) ? r1 - c0 : chunk;
        const size_t width = (size_t)(lower ? c0 + rows : ncols) * 8;
        double *mine = h->matrix + (size_t)c0 * h->n_pad;
        VGP_CUDA(cudaMemcpy2DAsync(mine, pitch, host + (size_t)(c0 - r0) * host_ld, (size_t)host_ld * 8, width,
                                   (size_t)rows, cudaMemcpyHostToDevice, s));
        if (G == 1) continue;
        VGP_CUDA(cudaEventRecord(h->push_event, s));
        VGP_CUDA(cudaStreamWaitEvent(h->push_stream, h->push_event, 0));
        for (int i = 1; i < G; ++i) {
            const int q = (me + i) % G;
            VGP_CUDA(cudaMemcpy2DAsync(h->peer_matrix[q] + (size_t)c0 * h->n_pad, pitch, mine, pitch, width,
                                       (size_t)rows, cudaMemcpyDefault, h->push_stream));
        }
    }
    if (G > 1) {
        VGP_CUDA(cudaEventRecord(h->push_event, h->push_stream));
        VGP_CUDA(cudaStreamWaitEvent(s, h->push_event, 0));
    }
    return VGP_OK;
}

// replica[i][i] += value for i < n (every rank on its own replica: e.g. the TF-graph variant's jitter before factorising)
int vgp_dist_add_diag(vgp_dist *h, int64_t n, double value, void *stream) {
    VGP_TRY(check(h));
    VGP_REQUIRE(n >= 0 && n <= h->n_pad, "bad size");
    VGP_ENTER(h->device);
    return dense_add_diag(h->matrix, h->n_pad, n, value, (cudaStream_t)stream);
}

int vgp_dist_barrier(vgp_dist *h, void *stream) {
    VGP_TRY(check(h));
    VGP_REQUIRE(h->connected, "vgp_dist_connect first");
    VGP_ENTER(h->device);
    if (h->ctx.nranks == 1) return VGP_OK;
    return dense_dist_barrier(h->ctx, (cudaStream_t)stream);
}

static int dist_run(vgp_dist *h, int *info_host, void *stream, bool factor_only);

int vgp_dist_spd_inverse(vgp_dist *h, int *info_host, void *stream) { return dist_run(h, info_host, stream, false); }

/* potrf + trtri only: the replicas end up holding M = L^-1 (lower triangle), what the lazy-column greedy with the
 * inverse factor resident needs -- 2/3 of the flops of the inverse. */
int vgp_dist_factor_inverse(vgp_dist *h, int *info_host, void *stream) { return dist_run(h, info_host, stream, true); }

static int dist_run(vgp_dist *h, int *info_host, void *stream, bool factor_only) {
    VGP_TRY(check(h));
    VGP_REQUIRE(h->connected || h->ctx.nranks == 1, "vgp_dist_connect first");
    if (info_host) *info_host = 0;
    VGP_ENTER(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (options_fingerprint(h->n_pad, h->tail_rows) != h->fingerprint) {
        set_error("options changed since vgp_dist_create: set them before creating the handle, identically on all ranks");
        return VGP_ERR_STATE;
    }
    h->ctx.min_tiles = option(VGP_OPT_DIST_MIN_TILES);
    h->ctx.min_k = option(VGP_OPT_DIST_MIN_K);
    if (h->ctx.min_k < 2 * TILE) h->ctx.min_k = 2 * TILE;      // the k = 128 leaf GEMMs update their operand in place
    dense_set_dist(h->ctx.nranks > 1 ? &h->ctx : nullptr);
    int rc = VGP_OK;
    if (h->ctx.nranks > 1) rc = dense_dist_barrier(h->ctx, s);      // every replica is filled
    if (rc == VGP_OK && !factor_only) rc = dense_spd_inverse(h->matrix, h->n_pad, h->n_pad, h->ws, info_host, s);
    if (rc == VGP_OK && factor_only) {
        rc = dense_potrf(h->matrix, h->n_pad, h->n_pad, h->ws, s);
        if (rc == VGP_OK) rc = dense_read_info(h->ws, info_host, s);
        if (rc == VGP_OK) rc = dense_trtri(h->matrix, h->n_pad, h->n_pad, h->ws, s);
    }
    if (rc == VGP_OK && h->ctx.nranks > 1) rc = dense_dist_barrier(h->ctx, s);
    dense_set_dist(nullptr);
    cudaError_t e = cudaStreamSynchronize(s);
    if (rc == VGP_OK && e != cudaSuccess) rc = cuda_fail(e, "distributed inverse", __FILE__, __LINE__);
    if (rc == VGP_OK && h->ctx.nranks > 1) {
        int err = 0;
        VGP_CUDA(cudaMemcpy(&err, h->flags + DIST_MAX, sizeof(int), cudaMemcpyDeviceToHost));
        if (err != 0) {
            set_error("distributed inverse: barrier timed out waiting for rank %d", err - 1);
            rc = VGP_ERR_STATE;
        }
    }
    return rc;
}

int vgp_dist_stats(vgp_dist *h, int64_t *dist_gemms, int64_t *barriers) {
    VGP_TRY(check(h));
    if (dist_gemms) *dist_gemms = h->ctx.dist_gemms;
    if (barriers) *barriers = h->ctx.barriers;
    return VGP_OK;
}

int vgp_dist_destroy(vgp_dist *h) {
    if (!h) return VGP_OK;
    VGP_ENTER(h->device);
    if (h->push_stream) {
        cudaStreamSynchronize(h->push_stream);
        cudaStreamDestroy(h->push_stream);
        cudaEventDestroy(h->push_event);
    }
    if (h->ipc && h->connected) {
        for (int q = 0; q < h->ctx.nranks; ++q) {
            if (q == h->ctx.rank) continue;
            if (h->peer_matrix[q]) cudaIpcCloseMemHandle(h->peer_matrix[q]);
            if (h->ctx.flags[q]) cudaIpcCloseMemHandle(h->ctx.flags[q]);
        }
    }
    if (h->matrix) cudaFree(h->matrix);
    if (h->flags) cudaFree(h->flags);
    h->ws.release();
    delete h;
    return VGP_OK;
}

}  // extern "C"
