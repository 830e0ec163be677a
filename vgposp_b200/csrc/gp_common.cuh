// Shared helpers of the GP / VGP translation units (gp.cu, elbo.cu): padded scratch buffers, deterministic
// scalar reductions, small matrix-vector kernels.  Everything sits in an anonymous namespace: each TU gets
// its own copy.
#pragma once
#include <math.h>

#include "dense.cuh"

namespace vgp {

int expquad_dispatch_public(const double *x1, int64_t n1, const double *x2, int64_t n2, int d, double amplitude,
                            double length_scale, double diag_add, int64_t diag_col0, double *out, int64_t ld,
                            cudaStream_t s);
int kernel_matrix_dispatch(int kind, const double *x1, int64_t n1, const double *x2, int64_t n2, int d,
                           double amplitude, double length_scale, double diag_add, int64_t diag_col0, double *out,
                           int64_t ld, cudaStream_t s);

namespace {

struct Buf {
    double *p = nullptr;
    int64_t rows = 0, cols = 0;
    cudaStream_t s = nullptr;
    int alloc(int64_t r, int64_t c, cudaStream_t stream, bool pad = true) {
        rows = pad ? round_up(r > 0 ? r : 1, TILE) : r;
        cols = pad ? round_up(c > 0 ? c : 1, TILE) : c;
        s = stream;
        VGP_CUDA(cudaMallocAsync((void **)&p, (size_t)rows * cols * 8, s));
        VGP_CUDA(cudaMemsetAsync(p, 0, (size_t)rows * cols * 8, s));
        return VGP_OK;
    }
    ~Buf() {
        if (p) cudaFreeAsync(p, s);
    }
};

// ---- deterministic scalar reductions ---------------------------------------------------------------
constexpr int RED_BLOCKS = 128;

struct SumSqRegion {      // sum over i < rows, j < cols of a[i][j]^2
    const double *a;
    int64_t ld, cols;
    __device__ double operator()(int64_t e) const {
        const double v = a[(e / cols) * ld + e % cols];
        return v * v;
    }
};
struct SumLogDiag {
    const double *a;
    int64_t ld;
    __device__ double operator()(int64_t e) const { return log(a[e * ld + e]); }
};
struct SumSqDiff {        // (y - mu)^2
    const double *y, *mu;
    __device__ double operator()(int64_t e) const {
        const double r = y[e] - mu[e];
        return r * r;
    }
};
struct SumSqStrided {
    const double *a;
    int64_t stride;
    __device__ double operator()(int64_t e) const {
        const double v = a[e * stride];
        return v * v;
    }
};

template <class F>
__global__ void __launch_bounds__(256) reduce_kernel(F f, int64_t count, double *partials, unsigned *counter,
                                                     double *out) {
    __shared__ double sh[256];
    double acc = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < count; e += (int64_t)gridDim.x * 256) acc += f(e);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = sh[0];
    __threadfence();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double t = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) t += __ldcg(partials + b);
        *out = t;
    }
}

struct Reducer {
    double *partials = nullptr;
    unsigned *counter = nullptr;
    double *scal = nullptr;      // device results
    cudaStream_t s = nullptr;
    int init(cudaStream_t stream) {
        s = stream;
        VGP_CUDA(cudaMallocAsync((void **)&partials, RED_BLOCKS * 8, s));
        VGP_CUDA(cudaMallocAsync((void **)&counter, 4, s));
        VGP_CUDA(cudaMallocAsync((void **)&scal, 16 * 8, s));
        VGP_CUDA(cudaMemsetAsync(counter, 0, 4, s));
        VGP_CUDA(cudaMemsetAsync(scal, 0, 16 * 8, s));
        return VGP_OK;
    }
    template <class F>
    int run(F f, int64_t count, int slot) {
        if (count <= 0) return VGP_OK;
        int64_t blocks = (count + 255) / 256;
        if (blocks > RED_BLOCKS) blocks = RED_BLOCKS;
        reduce_kernel<F><<<(unsigned)blocks, 256, 0, s>>>(f, count, partials, counter, scal + slot);
        VGP_LAUNCH_CHECK();
        return VGP_OK;
    }
    int fetch(double *host, int count) {
        VGP_CUDA(cudaMemcpyAsync(host, scal, (size_t)count * 8, cudaMemcpyDeviceToHost, s));
        VGP_CUDA(cudaStreamSynchronize(s));
        return VGP_OK;
    }
    ~Reducer() {
        if (partials) cudaFreeAsync(partials, s);
        if (counter) cudaFreeAsync(counter, s);
        if (scal) cudaFreeAsync(scal, s);
    }
};

// ---- small dense helpers ---------------------------------------------------------------------------
// out[j] = sum_k a[k][j] v[k * vstride]   (thread per column, coalesced over j)
__global__ void __launch_bounds__(256) gemv_t_kernel(const double *a, int64_t ld, int64_t rows, int64_t cols,
                                                     const double *v, int64_t vstride, double *out) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= cols) return;
    double acc = 0.0;
    for (int64_t k = 0; k < rows; ++k) acc = fma(a[k * ld + j], v[k * vstride], acc);
    out[j] = acc;
}

// out[k * ostride] (+)= scale * sum_j a[k][j] v[j]   (block per row, fixed reduction tree)
__global__ void __launch_bounds__(256) gemv_n_kernel(const double *a, int64_t ld, int64_t cols, const double *v,
                                                     double scale, int accumulate, double *out, int64_t ostride) {
    __shared__ double sh[256];
    const int64_t k = blockIdx.x;
    double acc = 0.0;
    for (int64_t j = threadIdx.x; j < cols; j += 256) acc = fma(a[k * ld + j], v[j], acc);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[k * ostride] = accumulate ? out[k * ostride] + scale * sh[0] : scale * sh[0];
}

// var[j] = base - sum_k c[k][j]^2 + sum_k e[k][j]^2   (e may be NULL)
__global__ void __launch_bounds__(256) colvar_kernel(const double *c, const double *e, int64_t ld, int64_t rows,
                                                     int64_t cols, double base, double *var) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= cols) return;
    double sc = 0.0, se = 0.0;
    for (int64_t k = 0; k < rows; ++k) {
        const double v = c[k * ld + j];
        sc = fma(v, v, sc);
        if (e) {
            const double w = e[k * ld + j];
            se = fma(w, w, se);
        }
    }
    var[j] = base - sc + se;
}

// dst[i][j] = a[i][j] * sa + b[i][j] * sb (+ diag on i == j < n)
__global__ void __launch_bounds__(256) axpby_kernel(const double *a, double sa, const double *b, double sb,
                                                    double diag, int64_t n, double *dst, int64_t ld, int64_t count) {
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= count) return;
    const int64_t i = e / ld, j = e % ld;
    double v = a[e] * sa + (b ? b[e] * sb : 0.0);
    if (i == j && i < n) v += diag;
    dst[e] = v;
}

// dst[j][i] = src[i][j] for a square [n][ld] matrix (out of place)
__global__ void __launch_bounds__(256) transpose_kernel(const double *src, double *dst, int64_t n, int64_t ld) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t i0 = (int64_t)blockIdx.y * 32, j0 = (int64_t)blockIdx.x * 32;
    for (int r = ty; r < 32; r += 8) tile[r][tx] = (i0 + r < n && j0 + tx < n) ? src[(i0 + r) * ld + j0 + tx] : 0.0;
    __syncthreads();
    for (int r = ty; r < 32; r += 8)
        if (j0 + r < n && i0 + tx < n) dst[(j0 + r) * ld + i0 + tx] = tile[tx][r];
}

int copy_vec_to_col0(const double *v, int64_t n, Buf &dst) {
    VGP_CUDA(cudaMemcpy2DAsync(dst.p, (size_t)dst.cols * 8, v, 8, 8, (size_t)n, cudaMemcpyDeviceToDevice, dst.s));
    return VGP_OK;
}
int copy_col0_to_vec(const Buf &src, int64_t n, double *v) {
    VGP_CUDA(cudaMemcpy2DAsync(v, 8, src.p, (size_t)src.cols * 8, 8, (size_t)n, cudaMemcpyDeviceToDevice, src.s));
    return VGP_OK;
}
int copy_matrix(const double *src, int64_t lds, int64_t r, int64_t c, Buf &dst) {
    VGP_CUDA(cudaMemcpy2DAsync(dst.p, (size_t)dst.cols * 8, src, (size_t)lds * 8, (size_t)c * 8, (size_t)r,
                               cudaMemcpyDeviceToDevice, dst.s));
    return VGP_OK;
}

// K(x, x) + shift I, padded with identity, factorised in place.
int kernel_cholesky(int kind, const double *x, int64_t n, int d, double amplitude, double length_scale, double shift,
                    Buf &k, DenseWorkspace &ws, cudaStream_t s) {
    VGP_TRY(k.alloc(n, n, s));
    VGP_TRY(kernel_matrix_dispatch(kind, x, n, x, n, d, amplitude, length_scale, shift, 0, k.p, k.cols, s));
    VGP_TRY(pad_identity(k.p, k.cols, n, k.rows, s));
    VGP_TRY(dense_potrf(k.p, k.rows, k.cols, ws, s));
    return dense_read_info(ws, nullptr, s);
}

}  // namespace
}  // namespace vgp
