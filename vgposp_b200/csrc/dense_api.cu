// C-ABI wrappers of the dense layer for caller-owned matrices of any size / leading dimension: operands
// are copied into 128-padded scratch (identity on the padding diagonal where a factorisation needs it),
// the padded kernels run, and the result is copied back.  The greedy and GP paths call dense.cuh directly
// on padded storage they own and never pay for these copies.
#include <math.h>

#include "dense.cuh"

using namespace vgp;

namespace vgp {

struct Padded {
    double *p = nullptr;
    int64_t rows = 0, cols = 0;   // padded
    cudaStream_t s = nullptr;
    int alloc(int64_t r, int64_t c, cudaStream_t stream) {
        rows = round_up(r > 0 ? r : 1, TILE);
        cols = round_up(c > 0 ? c : 1, TILE);
        s = stream;
        VGP_CUDA(cudaMallocAsync((void **)&p, (size_t)rows * cols * 8, s));
        VGP_CUDA(cudaMemsetAsync(p, 0, (size_t)rows * cols * 8, s));
        return VGP_OK;
    }
    int load(const double *src, int64_t ld, int64_t r, int64_t c) {
        if (r == 0 || c == 0) return VGP_OK;
        VGP_CUDA(cudaMemcpy2DAsync(p, (size_t)cols * 8, src, (size_t)ld * 8, (size_t)c * 8, (size_t)r,
                                   cudaMemcpyDeviceToDevice, s));
        return VGP_OK;
    }
    int store(double *dst, int64_t ld, int64_t r, int64_t c) const {
        if (r == 0 || c == 0) return VGP_OK;
        VGP_CUDA(cudaMemcpy2DAsync(dst, (size_t)ld * 8, p, (size_t)cols * 8, (size_t)c * 8, (size_t)r,
                                   cudaMemcpyDeviceToDevice, s));
        return VGP_OK;
    }
    ~Padded() {
        if (p) cudaFreeAsync(p, s);
    }
};

__global__ void pad_diag_kernel(double *a, int64_t ld, int64_t n, int64_t n_pad) {
    const int64_t i = n + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) a[i * ld + i] = 1.0;
}

int pad_identity(double *a, int64_t ld, int64_t n, int64_t n_pad, cudaStream_t s) {
    if (n_pad <= n) return VGP_OK;
    pad_diag_kernel<<<(unsigned)((n_pad - n + 127) / 128), 128, 0, s>>>(a, ld, n, n_pad);
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

// copy the lower triangle (incl. diagonal) of src [n][lds] into dst [n][ldd]
__global__ void __launch_bounds__(256) copy_lower_kernel(const double *src, int64_t lds, double *dst, int64_t ldd,
                                                         int64_t n) {
    for (int64_t i = blockIdx.y; i < n; i += gridDim.y)
        for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j <= i; j += (int64_t)gridDim.x * 256)
            dst[i * ldd + j] = src[i * lds + j];
}

int copy_lower(const double *src, int64_t lds, double *dst, int64_t ldd, int64_t n, cudaStream_t s) {
    const int64_t gx = (n + 255) / 256 < 64 ? (n + 255) / 256 : 64;
    const int64_t gy = n < 4096 ? n : 4096;
    copy_lower_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, s>>>(src, lds, dst, ldd, n);
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

// row i of m [n][ld] (first s columns) minus its mean, in place; one CTA per row, fixed reduction tree
__global__ void __launch_bounds__(256) centre_rows_kernel(double *m, int64_t ld, int64_t s) {
    __shared__ double sh[256];
    double *row = m + (int64_t)blockIdx.x * ld;
    double acc = 0.0;
    for (int64_t j = threadIdx.x; j < s; j += 256) acc += row[j];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    const double mean = sh[0] / (double)s;
    for (int64_t j = threadIdx.x; j < s; j += 256) row[j] -= mean;
}

// c[i][j] *= exp(-(beta delta_ij)^2 / (2 pi)) with delta the Euclidean distance of the integer grid indices of
// locations i and j; factors below `cutoff` become exact zeros (main_architecture_2_sampledistribution.py:376-394)
__global__ void __launch_bounds__(256) taper_kernel(double *c, int64_t ld, int64_t n, const int *__restrict__ idx,
                                                    double beta, double cutoff) {
    const int64_t i = blockIdx.y;
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    const double d0 = (double)(idx[i * 3] - idx[j * 3]), d1 = (double)(idx[i * 3 + 1] - idx[j * 3 + 1]),
                 d2 = (double)(idx[i * 3 + 2] - idx[j * 3 + 2]);
    const double delta = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    const double bd = beta * delta;
    double decay = exp(-(bd * bd) / (2.0 * M_PI));
    if (decay < cutoff) decay = 0.0;
    c[i * ld + j] *= decay;
}

}  // namespace vgp

extern "C" {

/* cov[i][j] = mean_s((m[i][s] - mean_i)(m[j][s] - mean_j)): the biased sample covariance of every pair of rows --
 * np.cov(tracers_loc_i, tracers_loc_j, bias=True)[0, 1] for all (i, j) at once (gp_functions.py:1019-1057;
 * tfp.stats.covariance in main_architecture_2.py:431).  One centring pass, one lower-tile split-K SYRK on the DMMA
 * GEMM, mirrored: exactly symmetric. */
int vgp_empirical_cov(int device, const double *m_dev, int64_t n, int64_t s_samples, int64_t ldm, double *cov_dev,
                      int64_t ldc, void *stream) {
    VGP_REQUIRE(n >= 0 && s_samples > 0 && ldm >= s_samples && ldc >= n, "bad sizes");
    if (n == 0) return VGP_OK;
    VGP_REQUIRE(m_dev && cov_dev, "NULL pointer");
    VGP_ENTER(device);
    cudaStream_t s = (cudaStream_t)stream;
    Padded pm, pc, part;
    VGP_TRY(pm.alloc(n, s_samples, s));
    VGP_TRY(pc.alloc(n, n, s));
    VGP_TRY(pm.load(m_dev, ldm, n, s_samples));
    centre_rows_kernel<<<(unsigned)n, 256, 0, s>>>(pm.p, pm.cols, s_samples);
    VGP_LAUNCH_CHECK();
    const int64_t tiles = (pc.rows / TILE) * (pc.rows / TILE + 1) / 2;
    int splits = (int)(592 / (2 * tiles));                      // about two waves of 128x64 CTAs
    const int64_t max_splits = pm.cols / 256 > 0 ? pm.cols / 256 : 1;
    if (splits > max_splits) splits = (int)max_splits;
    if (splits < 1) splits = 1;
    VGP_TRY(part.alloc(pc.rows * splits, pc.cols, s));
    VGP_TRY(dense_gemm_splitk(0, 1, pc.rows, pc.rows, pm.cols, 1.0 / (double)s_samples, pm.p, pm.cols, pm.p, pm.cols,
                              0.0, pc.p, pc.cols, splits, part.p, s, GEMM_LOWER));
    return pc.store(cov_dev, ldc, n, n);
}

/* In-place taper of a covariance by the reference's decay filter over integer grid indices idx_dev [n, 3] (int32):
 * factor exp(-(beta delta)^2 / (2 pi)), zero below `cutoff` (0.01 in the reference). */
int vgp_cov_taper(int device, double *cov_dev, int64_t n, int64_t ldc, const int *idx_dev, double beta,
                  double cutoff, void *stream) {
    VGP_REQUIRE(n >= 0 && ldc >= n, "bad sizes");
    if (n == 0) return VGP_OK;
    VGP_REQUIRE(cov_dev && idx_dev, "NULL pointer");
    VGP_REQUIRE(n <= 65535, "vgp_cov_taper: n too large for one launch");
    VGP_ENTER(device);
    cudaStream_t s = (cudaStream_t)stream;
    taper_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)n), 256, 0, s>>>(cov_dev, ldc, n, idx_dev, beta, cutoff);
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

int vgp_dgemm(int device, int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, double alpha,
              const double *a_dev, int64_t lda, const double *b_dev, int64_t ldb, double beta, double *c_dev,
              int64_t ldc, void *stream) {
    VGP_REQUIRE(m >= 0 && n >= 0 && k >= 0, "negative size");
    if (m == 0 || n == 0) return VGP_OK;
    VGP_REQUIRE(a_dev && b_dev && c_dev, "NULL pointer");
    const int64_t ar = trans_a ? k : m, ac = trans_a ? m : k, br = trans_b ? n : k, bc = trans_b ? k : n;
    VGP_REQUIRE(lda >= ac && ldb >= bc && ldc >= n, "leading dimension too small");
    {
        // C may not share memory with an operand: tiles of C are written while other CTAs still read A and B
        auto overlaps = [](const double *x, int64_t rows, int64_t cols, int64_t ld, const double *y, int64_t yrows,
                           int64_t ycols, int64_t yld) {
            if (rows == 0 || cols == 0 || yrows == 0 || ycols == 0) return false;
            const double *x1 = x + (rows - 1) * ld + cols, *y1 = y + (yrows - 1) * yld + ycols;
            return x < y1 && y < x1;
        };
        VGP_REQUIRE(!overlaps(c_dev, m, n, ldc, a_dev, ar, ac, lda) && !overlaps(c_dev, m, n, ldc, b_dev, br, bc, ldb),
                    "vgp_dgemm: C overlaps an operand");
    }
    VGP_ENTER(device);
    cudaStream_t s = (cudaStream_t)stream;
    if (m % TILE == 0 && n % TILE == 0 && k % 16 == 0 && k > 0 && lda % 2 == 0 && ldb % 2 == 0 && ldc % 2 == 0 &&
        ((uintptr_t)a_dev % 16) == 0 && ((uintptr_t)b_dev % 16) == 0 && ((uintptr_t)c_dev % 16) == 0)
        return dense_gemm(trans_a, trans_b, m, n, k, alpha, a_dev, lda, b_dev, ldb, beta, c_dev, ldc, GEMM_FULL, s);
    Padded pa, pb, pc;
    VGP_TRY(pa.alloc(ar, ac, s));
    VGP_TRY(pb.alloc(br, bc, s));
    VGP_TRY(pc.alloc(m, n, s));
    VGP_TRY(pa.load(a_dev, lda, ar, ac));
    VGP_TRY(pb.load(b_dev, ldb, br, bc));
    if (beta != 0.0) VGP_TRY(pc.load(c_dev, ldc, m, n));
    const int64_t kp = round_up(k > 0 ? k : 1, TILE);
    VGP_TRY(dense_gemm(trans_a, trans_b, pc.rows, pc.cols, kp, alpha, pa.p, pa.cols, pb.p, pb.cols, beta, pc.p,
                       pc.cols, GEMM_FULL, s));
    return pc.store(c_dev, ldc, m, n);
}

int vgp_potrf(int device, double *a_dev, int64_t n, int64_t lda, int *info_host, void *stream) {
    VGP_REQUIRE(n >= 0 && lda >= n, "bad size");
    if (info_host) *info_host = 0;
    if (n == 0) return VGP_OK;
    VGP_REQUIRE(a_dev, "NULL pointer");
    VGP_ENTER(device);
    cudaStream_t s = (cudaStream_t)stream;
    Padded pa;
    VGP_TRY(pa.alloc(n, n, s));
    VGP_TRY(pa.load(a_dev, lda, n, n));
    VGP_TRY(pad_identity(pa.p, pa.cols, n, pa.rows, s));
    DenseWorkspace ws;
    int rc = dense_potrf(pa.p, pa.rows, pa.cols, ws, s);
    if (rc == VGP_OK) rc = dense_read_info(ws, info_host, s);
    // only the lower triangle is written back: the strict upper triangle of the caller's matrix stays
    if (rc == VGP_OK) rc = copy_lower(pa.p, pa.cols, a_dev, lda, n, s);
    cudaStreamSynchronize(s);
    ws.release();
    return rc;
}

int vgp_spd_inverse(int device, double *a_dev, int64_t n, int64_t lda, int *info_host, void *stream) {
    VGP_REQUIRE(n >= 0 && lda >= n, "bad size");
    if (info_host) *info_host = 0;
    if (n == 0) return VGP_OK;
    VGP_REQUIRE(a_dev, "NULL pointer");
    VGP_ENTER(device);
    cudaStream_t s = (cudaStream_t)stream;
    if (n % TILE == 0 && lda % 2 == 0 && ((uintptr_t)a_dev % 16) == 0) {      // in place, no scratch copy
        DenseWorkspace ws;
        int rc = dense_spd_inverse(a_dev, n, lda, ws, info_host, s);
        cudaStreamSynchronize(s);
        ws.release();
        return rc;
    }
    Padded pa;
    VGP_TRY(pa.alloc(n, n, s));
    VGP_TRY(pa.load(a_dev, lda, n, n));
    VGP_TRY(pad_identity(pa.p, pa.cols, n, pa.rows, s));
    DenseWorkspace ws;
    int rc = dense_spd_inverse(pa.p, pa.rows, pa.cols, ws, info_host, s);
    if (rc == VGP_OK) rc = pa.store(a_dev, lda, n, n);
    cudaStreamSynchronize(s);
    ws.release();
    return rc;
}

int vgp_trsm(int device, int side, int trans, int64_t n, int64_t nrhs, const double *l_dev, int64_t ldl,
             double *b_dev, int64_t ldb, void *stream) {
    VGP_REQUIRE(n >= 0 && nrhs >= 0 && ldl >= n, "bad size");
    if (n == 0 || nrhs == 0) return VGP_OK;
    VGP_REQUIRE(l_dev && b_dev, "NULL pointer");
    const int64_t br = side == 0 ? n : nrhs, bc = side == 0 ? nrhs : n;
    VGP_REQUIRE(ldb >= bc, "ldb too small");
    VGP_ENTER(device);
    cudaStream_t s = (cudaStream_t)stream;
    Padded pl, pb;
    VGP_TRY(pl.alloc(n, n, s));
    VGP_TRY(pb.alloc(br, bc, s));
    // lower triangle of L only: row by row would be n copies; copy the full square then rely on the
    // block kernels reading j <= i only inside diagonal blocks, and zero the strict upper blocks.
    VGP_TRY(pl.load(l_dev, ldl, n, n));
    VGP_TRY(dense_zero_strict_upper(pl.p, pl.rows, pl.cols, s));
    VGP_TRY(pad_identity(pl.p, pl.cols, n, pl.rows, s));
    VGP_TRY(pb.load(b_dev, ldb, br, bc));
    DenseWorkspace ws;
    int rc = dense_trsm(side, trans, pl.rows, side == 0 ? pb.cols : pb.rows, 1.0, pl.p, pl.cols, pb.p, pb.cols, ws,
                        false, s);
    if (rc == VGP_OK) rc = pb.store(b_dev, ldb, br, bc);
    cudaStreamSynchronize(s);
    ws.release();
    return rc;
}

}  // extern "C"
