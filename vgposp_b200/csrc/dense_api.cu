// C-ABI wrappers of the dense layer for caller-owned matrices of any size / leading dimension: operands
// are copied into 128-padded scratch (identity on the padding diagonal where a factorisation needs it),
// the padded kernels run, and the result is copied back.  The greedy and GP paths call dense.cuh directly
// on padded storage they own and never pay for these copies.
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

}  // namespace vgp

extern "C" {

int vgp_dgemm(int device, int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, double alpha,
              const double *a_dev, int64_t lda, const double *b_dev, int64_t ldb, double beta, double *c_dev,
              int64_t ldc, void *stream) {
    VGP_REQUIRE(m >= 0 && n >= 0 && k >= 0, "negative size");
    if (m == 0 || n == 0) return VGP_OK;
    VGP_REQUIRE(a_dev && b_dev && c_dev, "NULL pointer");
    const int64_t ar = trans_a ? k : m, ac = trans_a ? m : k, br = trans_b ? n : k, bc = trans_b ? k : n;
    VGP_REQUIRE(lda >= ac && ldb >= bc && ldc >= n, "leading dimension too small");
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
