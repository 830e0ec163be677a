// Exact-GP and variational-GP arithmetic on top of expquad + dense.
//
// Reference call sites (all delegate to TensorFlow-Probability):
//   exact GP log_prob        gp_functions.py:166-172, 3D_sin_wave.py:161-172
//   GP regression model      gp_functions.py:283-297, 3D_sin_wave.py:262-268
//   optimal q(u)             variational_Gaussian_process_example.py:68-74, main_architecture_2.py:200-206
//   variational_loss / mean  variational_Gaussian_process_example.py:83-99,141-142
// Formulas: SURVEY.md Appendix A.2-A.4 (restated on the CPU in oracle/gp_oracle.py).
//
// Everything is float64.  Matrices live in 128-padded scratch (identity on the padding diagonal of every
// factorised matrix, zeros elsewhere), so padded rows/columns contribute exact zeros to every norm.
// Scalar reductions are two-stage with a fixed tree: results are run-to-run deterministic.
#include <vector>

#include "block128.cuh"
#include "gp_common.cuh"

namespace vgp {
namespace gp_detail {

// Row i of the fused gradient reduction of the exact-GP log marginal likelihood:
//   W_ij = (alpha_i alpha_j - Cinv_ij) / 2,  rowacc[i] = ( sum_j W_ij K_ij,  sum_j W_ij dK_ij/dl,  W_ii )
// K and dK/dl are re-evaluated from the coordinates (nothing but Cinv is read from HBM).  One CTA per row, fixed
// reduction tree: deterministic.
template <int KIND, int D>
__global__ void __launch_bounds__(256) gp_grad_kernel(const double *__restrict__ x, int64_t n,
                                                      const double *__restrict__ cinv, int64_t ld,
                                                      const double *__restrict__ alpha, int64_t astride, double amp2,
                                                      double ls, double *rowacc) {
    __shared__ double sh[3][256];
    const int64_t i = blockIdx.x;
    double xi[D];
#pragma unroll
    for (int k = 0; k < D; ++k) xi[k] = x[i * D + k];
    const double ai = alpha[i * astride];
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += 256) {
        double s2 = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const double t = xi[k] - x[j * D + k];
            s2 = fma(t, t, s2);
        }
        double kv, dk;
        if (KIND == VGP_KERNEL_EXPQUAD) {
            kv = amp2 * exp(-0.5 * s2 / (ls * ls));
            dk = kv * s2 / (ls * ls * ls);
        } else {
            const double r = sqrt(s2);
            if (KIND == VGP_KERNEL_MATERN12) {
                kv = amp2 * exp(-r / ls);
                dk = kv * r / (ls * ls);
            } else if (KIND == VGP_KERNEL_MATERN32) {
                const double z = sqrt(3.0) * r / ls, e = amp2 * exp(-z);
                kv = (1.0 + z) * e;
                dk = z * z * e / ls;
            } else {
                const double z = sqrt(5.0) * r / ls, e = amp2 * exp(-z);
                kv = (1.0 + z + z * z / 3.0) * e;
                dk = (z * z / 3.0) * (1.0 + z) * e / ls;
            }
        }
        const double w = 0.5 * (ai * alpha[j * astride] - cinv[i * ld + j]);
        acc0 = fma(w, kv, acc0);
        acc1 = fma(w, dk, acc1);
        if (j == i) acc2 = w;
    }
    sh[0][threadIdx.x] = acc0;
    sh[1][threadIdx.x] = acc1;
    sh[2][threadIdx.x] = acc2;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) {
#pragma unroll
            for (int q = 0; q < 3; ++q) sh[q][threadIdx.x] += sh[q][threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x < 3) rowacc[i * 3 + threadIdx.x] = sh[threadIdx.x][0];
}

// ---- batched exact-GP likelihood (calc_H, gp_functions.py:864-876): one CTA per hyper-parameter triple ---------
// The n <= 127 observations of the reference's likelihood-surface sweeps (25 points at main.py:418-419) fit one
// 128-block: the CTA builds C = K + (noise + jitter) I straight into the register tile of the elimination kernel, with
// y as row n of the block (the Cholesky of [[C, y], [y^T, big]] has L^-1 y as its row n), factorises the panels that
// hold live rows, and reduces log-determinant and quadratic form -- no global-memory traffic but x, y and 8 bytes out.
template <int KIND, int D>
__global__ void __launch_bounds__(256, 1) gp_logprob_batch_kernel(const double *__restrict__ x, int n,
                                                                  const double *__restrict__ y,
                                                                  const double *__restrict__ params, double jitter,
                                                                  double *out, int *info) {
    __shared__ double colbuf[2][NB];
    __shared__ double xs[NB][D];
    __shared__ double ys[NB];
    __shared__ double red[2][256];
    const int b = blockIdx.x;
    const double amp = params[3 * b], ls = params[3 * b + 1], shift = params[3 * b + 2] + jitter;
    const double amp2 = amp * amp;
    for (int t = threadIdx.x; t < NB * D; t += 256) xs[t / D][t % D] = t / D < n ? x[t] : 0.0;
    for (int t = threadIdx.x; t < NB; t += 256) ys[t] = t < n ? y[t] : 0.0;
    __syncthreads();
    const int ti = threadIdx.x >> 4, tc = threadIdx.x & 15;
    double v[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const int i = ti + 16 * r, c = tc + 16 * s;
            double val = 0.0;
            if (c <= i) {
                if (i < n) {
                    double s2 = 0.0;
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        const double t = xs[i][k] - xs[c][k];
                        s2 = fma(t, t, s2);
                    }
                    if (KIND == VGP_KERNEL_EXPQUAD) {
                        val = amp2 * exp(-0.5 * s2 / (ls * ls));
                    } else {
                        const double rr = sqrt(s2);
                        if (KIND == VGP_KERNEL_MATERN12) {
                            val = amp2 * exp(-rr / ls);
                        } else if (KIND == VGP_KERNEL_MATERN32) {
                            const double z = sqrt(3.0) * rr / ls;
                            val = amp2 * (1.0 + z) * exp(-z);
                        } else {
                            const double z = sqrt(5.0) * rr / ls;
                            val = amp2 * (1.0 + z + z * z / 3.0) * exp(-z);
                        }
                    }
                    if (i == c) val += shift;
                } else if (i == n) {
                    val = c < n ? ys[c] : 1e300;           // row n = y^T, its pivot kept huge and harmless
                } else {
                    val = i == c ? 1.0 : 0.0;              // identity padding
                }
            }
            v[r][s] = val;
        }
    // only the panels that contain rows 0 .. n need eliminating
    int *flag = info + b;
    const int panels = n / 16 + 1;
    potf2_panel<0>(v, colbuf, ti, tc, flag, 0);
    if (panels > 1) potf2_panel<1>(v, colbuf, ti, tc, flag, 0);
    if (panels > 2) potf2_panel<2>(v, colbuf, ti, tc, flag, 0);
    if (panels > 3) potf2_panel<3>(v, colbuf, ti, tc, flag, 0);
    if (panels > 4) potf2_panel<4>(v, colbuf, ti, tc, flag, 0);
    if (panels > 5) potf2_panel<5>(v, colbuf, ti, tc, flag, 0);
    if (panels > 6) potf2_panel<6>(v, colbuf, ti, tc, flag, 0);
    if (panels > 7) potf2_panel<7>(v, colbuf, ti, tc, flag, 0);
    double logdet = 0.0, quad = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int s = 0; s <= r; ++s) {
            const int i = ti + 16 * r, c = tc + 16 * s;
            if (i == c && i < n) logdet += log(v[r][s]);
            if (i == n && c < n) quad = fma(v[r][s], v[r][s], quad);
        }
    red[0][threadIdx.x] = logdet;
    red[1][threadIdx.x] = quad;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) {
            red[0][threadIdx.x] += red[0][threadIdx.x + off];
            red[1][threadIdx.x] += red[1][threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[b] = -0.5 * red[1][0] - red[0][0] - 0.5 * (double)n * log(2.0 * M_PI);
}

template <int KIND>
int launch_gp_batch(int d, const double *x, int n, const double *y, const double *params, int64_t batch,
                    double jitter, double *out, int *info, cudaStream_t s) {
    switch (d) {
#define VGP_CASE(DD)                                                                                              \
    case DD:                                                                                                      \
        gp_logprob_batch_kernel<KIND, DD><<<(unsigned)batch, 256, 0, s>>>(x, n, y, params, jitter, out, info);    \
        break;
        VGP_CASE(1) VGP_CASE(2) VGP_CASE(3) VGP_CASE(4) VGP_CASE(5) VGP_CASE(6) VGP_CASE(7) VGP_CASE(8)
#undef VGP_CASE
        default:
            set_error("feature dimension %d outside [1, 8]", d);
            return VGP_ERR_INVALID;
    }
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

struct ColSum3 {
    const double *a;
    int c;
    __device__ double operator()(int64_t e) const { return a[e * 3 + c]; }
};

template <int KIND>
int launch_gp_grad(int d, const double *x, int64_t n, const double *cinv, int64_t ld, const double *alpha,
                   int64_t astride, double amp2, double ls, double *rowacc, cudaStream_t s) {
    const unsigned g = (unsigned)n;
    switch (d) {
#define VGP_CASE(DD)                                                                                      \
    case DD:                                                                                              \
        gp_grad_kernel<KIND, DD><<<g, 256, 0, s>>>(x, n, cinv, ld, alpha, astride, amp2, ls, rowacc);     \
        break;
        VGP_CASE(1) VGP_CASE(2) VGP_CASE(3) VGP_CASE(4) VGP_CASE(5) VGP_CASE(6) VGP_CASE(7) VGP_CASE(8)
#undef VGP_CASE
        default:
            set_error("feature dimension %d outside [1, 8]", d);
            return VGP_ERR_INVALID;
    }
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

}  // namespace gp_detail
}  // namespace vgp

using namespace vgp;
using namespace vgp::gp_detail;

extern "C" {

/* log N(y | 0, C), C = K + (noise + jitter) I, and its gradient with respect to (amplitude, length_scale,
 * noise_variance): dL/dtheta = tr(W dC/dtheta), W = (alpha alpha^T - C^-1) / 2, alpha = C^-1 y.  What TF's autodiff
 * hands to tf.train.AdamOptimizer(...).minimize(-log_likelihood) in gpf.tf_train_gp_adam (gp_functions.py:179-182,
 * main.py:110). */
int vgp_gp_logprob_grad_k(int device, int kind, const double *x_dev, int64_t n, int d, const double *y_dev,
                          double amplitude, double length_scale, double noise_variance, double jitter,
                          double *logprob_host, double *grads_host, void *stream) {
    VGP_REQUIRE(x_dev && y_dev && logprob_host && grads_host && n > 0, "bad argument");
    VGP_REQUIRE(kind >= VGP_KERNEL_EXPQUAD && kind <= VGP_KERNEL_MATERN52, "unknown kernel kind %d", kind);
    VGP_ENTER(device);
    cudaStream_t s = (cudaStream_t)stream;
    DenseWorkspace ws;
    int rc;
    {
        Buf k, yb, acc;
        Reducer red;
        rc = red.init(s);
        if (rc == VGP_OK) rc = kernel_cholesky(kind, x_dev, n, d, amplitude, length_scale, noise_variance + jitter, k, ws, s);
        if (rc == VGP_OK) rc = yb.alloc(n, 1, s);
        if (rc == VGP_OK) rc = acc.alloc(n, 3, s, false);
        if (rc == VGP_OK) rc = copy_vec_to_col0(y_dev, n, yb);
        if (rc == VGP_OK) rc = dense_trsm(0, 0, k.rows, yb.cols, 1.0, k.p, k.cols, yb.p, yb.cols, ws, true, s);
        if (rc == VGP_OK) rc = red.run(SumSqStrided{yb.p, yb.cols}, n, 0);
        if (rc == VGP_OK) rc = red.run(SumLogDiag{k.p, k.cols}, n, 1);
        if (rc == VGP_OK) rc = dense_trsm(0, 1, k.rows, yb.cols, 1.0, k.p, k.cols, yb.p, yb.cols, ws, true, s);   // alpha
        if (rc == VGP_OK) rc = dense_trtri(k.p, k.rows, k.cols, ws, s);
        if (rc == VGP_OK) rc = dense_lauum(k.p, k.rows, k.cols, ws, s);
        if (rc == VGP_OK) rc = dense_mirror_lower(k.p, k.rows, k.cols, s);                                       // C^-1
        if (rc == VGP_OK) {
            const double a2 = amplitude * amplitude;
            switch (kind) {
                case VGP_KERNEL_EXPQUAD:
                    rc = launch_gp_grad<VGP_KERNEL_EXPQUAD>(d, x_dev, n, k.p, k.cols, yb.p, yb.cols, a2, length_scale, acc.p, s);
                    break;
                case VGP_KERNEL_MATERN12:
                    rc = launch_gp_grad<VGP_KERNEL_MATERN12>(d, x_dev, n, k.p, k.cols, yb.p, yb.cols, a2, length_scale, acc.p, s);
                    break;
                case VGP_KERNEL_MATERN32:
                    rc = launch_gp_grad<VGP_KERNEL_MATERN32>(d, x_dev, n, k.p, k.cols, yb.p, yb.cols, a2, length_scale, acc.p, s);
                    break;
                default:
                    rc = launch_gp_grad<VGP_KERNEL_MATERN52>(d, x_dev, n, k.p, k.cols, yb.p, yb.cols, a2, length_scale, acc.p, s);
            }
        }
        for (int c = 0; c < 3 && rc == VGP_OK; ++c) rc = red.run(ColSum3{acc.p, c}, n, 2 + c);
        double h[5] = {0, 0, 0, 0, 0};
        if (rc == VGP_OK) rc = red.fetch(h, 5);
        if (rc == VGP_OK) {
            *logprob_host = -0.5 * h[0] - h[1] - 0.5 * (double)n * log(2.0 * M_PI);
            grads_host[0] = 2.0 * h[2] / amplitude;
            grads_host[1] = h[3];
            grads_host[2] = h[4];
        }
    }
    cudaStreamSynchronize(s);
    ws.release();
    return rc;
}

int vgp_gp_logprob_k(int device, int kind, const double *x_dev, int64_t n, int d, const double *y_dev, double amplitude,
                   double length_scale, double noise_variance, double jitter, double *logprob_host,
                   void *stream) {
    VGP_REQUIRE(x_dev && y_dev && logprob_host && n > 0, "bad argument");
    VGP_ENTER(device);
    cudaStream_t s = (cudaStream_t)stream;
    DenseWorkspace ws;
    int rc;
    {
        Buf k, yb;
        Reducer red;
        rc = red.init(s);
        if (rc == VGP_OK) rc = kernel_cholesky(kind, x_dev, n, d, amplitude, length_scale, noise_variance + jitter, k, ws, s);
        if (rc == VGP_OK) rc = yb.alloc(n, 1, s);
        if (rc == VGP_OK) rc = copy_vec_to_col0(y_dev, n, yb);
        if (rc == VGP_OK) rc = dense_trsm(0, 0, k.rows, yb.cols, 1.0, k.p, k.cols, yb.p, yb.cols, ws, true, s);
        if (rc == VGP_OK) rc = red.run(SumSqStrided{yb.p, yb.cols}, n, 0);
        if (rc == VGP_OK) rc = red.run(SumLogDiag{k.p, k.cols}, n, 1);
        double h[2] = {0, 0};
        if (rc == VGP_OK) rc = red.fetch(h, 2);
        if (rc == VGP_OK) *logprob_host = -0.5 * h[0] - h[1] - 0.5 * (double)n * log(2.0 * M_PI);
    }
    cudaStreamSynchronize(s);
    ws.release();
    return rc;
}

int vgp_gp_regression_k(int device, int kind, const double *x_dev, int64_t n, int d, const double *y_dev,
                      const double *xt_dev, int64_t t, double amplitude, double length_scale,
                      double noise_variance, double predictive_noise_variance, double divisor_jitter,
                      double *mean_dev, double *var_dev, void *stream) {
    VGP_REQUIRE(x_dev && y_dev && xt_dev && n > 0 && t > 0, "bad argument");
    VGP_ENTER(device);
    cudaStream_t s = (cudaStream_t)stream;
    DenseWorkspace ws;
    int rc;
    {
        Buf k, yb, c;
        rc = kernel_cholesky(kind, x_dev, n, d, amplitude, length_scale, noise_variance + divisor_jitter, k, ws, s);
        if (rc == VGP_OK) rc = c.alloc(n, t, s);
        if (rc == VGP_OK)
            rc = kernel_matrix_dispatch(kind, x_dev, n, xt_dev, t, d, amplitude, length_scale, 0.0, 0, c.p, c.cols, s);
        if (rc == VGP_OK) rc = dense_trsm(0, 0, k.rows, c.cols, 1.0, k.p, k.cols, c.p, c.cols, ws, true, s);
        if (rc == VGP_OK) rc = yb.alloc(n, 1, s);
        if (rc == VGP_OK) rc = copy_vec_to_col0(y_dev, n, yb);
        if (rc == VGP_OK) rc = dense_trsm(0, 0, k.rows, yb.cols, 1.0, k.p, k.cols, yb.p, yb.cols, ws, true, s);
        if (rc == VGP_OK && mean_dev) {
            gemv_t_kernel<<<(unsigned)((t + 255) / 256), 256, 0, s>>>(c.p, c.cols, n, t, yb.p, yb.cols, mean_dev);
            ++g_launches;
        }
        if (rc == VGP_OK && var_dev) {
            colvar_kernel<<<(unsigned)((t + 255) / 256), 256, 0, s>>>(c.p, nullptr, c.cols, n, t,
                                                                    amplitude * amplitude + predictive_noise_variance,
                                                                    var_dev);
            ++g_launches;
        }
        if (rc == VGP_OK) {
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) rc = cuda_fail(e, "gp_regression kernels", __FILE__, __LINE__);
        }
    }
    cudaStreamSynchronize(s);
    ws.release();
    return rc;
}

int vgp_vgp_optimal_posterior_k(int device, int kind, const double *z_dev, int64_t m, const double *x_dev, int64_t n_obs,
                              int d, const double *y_dev, double amplitude, double length_scale,
                              double noise_variance, double jitter, int legacy_scale_orientation,
                              double *loc_dev, double *scale_dev, void *stream) {
    VGP_REQUIRE(z_dev && x_dev && y_dev && loc_dev && scale_dev && m > 0 && n_obs > 0, "bad argument");
    VGP_REQUIRE(noise_variance > 0.0, "observation noise variance must be positive");
    VGP_ENTER(device);
    cudaStream_t s = (cudaStream_t)stream;
    DenseWorkspace ws;
    int rc = VGP_OK;
    {
        const int64_t mp = round_up(m, TILE);
        const int64_t chunk = 16384;          // observations per K_zx slab: [mp][chunk] stays L2-resident
        int sm = 148;
        cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
        const int tiles = (int)((mp / TILE) * (mp / TILE));
        int splits = (2 * sm + tiles - 1) / tiles;
        if (splits > (int)(chunk / 128)) splits = (int)(chunk / 128);
        if (splits < 1) splits = 1;
        Buf kzz, kzx, gram, partial, v, sig, kzz2;
        rc = kzz.alloc(m, m, s);
        if (rc == VGP_OK) rc = kzx.alloc(m, chunk, s);
        if (rc == VGP_OK) rc = gram.alloc(m, m, s);
        if (rc == VGP_OK) rc = partial.alloc((int64_t)splits * mp, mp, s, false);
        if (rc == VGP_OK) rc = v.alloc(m, 1, s);
        if (rc == VGP_OK) rc = sig.alloc(m, m, s);
        if (rc == VGP_OK)
            rc = kernel_matrix_dispatch(kind, z_dev, m, z_dev, m, d, amplitude, length_scale, 0.0, 0, kzz.p, kzz.cols, s);
        for (int64_t c0 = 0; c0 < n_obs && rc == VGP_OK; c0 += chunk) {
            const int64_t cn = n_obs - c0 < chunk ? n_obs - c0 : chunk;
            if (cn < chunk) rc = cudaMemsetAsync(kzx.p, 0, (size_t)kzx.rows * kzx.cols * 8, s) == cudaSuccess
                                     ? VGP_OK
                                     : VGP_ERR_CUDA;
            if (rc == VGP_OK)
                rc = kernel_matrix_dispatch(kind, z_dev, m, x_dev + c0 * d, cn, d, amplitude, length_scale, 0.0, -1 - n_obs,
                                             kzx.p, kzx.cols, s);
            // gram += K_zx K_zx^T ;  v += K_zx y
            if (rc == VGP_OK)
                rc = dense_gemm_splitk(0, 1, mp, mp, chunk, 1.0, kzx.p, kzx.cols, kzx.p, kzx.cols, 1.0, gram.p,
                                       gram.cols, splits, partial.p, s);
            if (rc == VGP_OK) {
                gemv_n_kernel<<<(unsigned)m, 256, 0, s>>>(kzx.p, kzx.cols, cn, y_dev + c0, 1.0, 1, v.p, v.cols);
                ++g_launches;
            }
        }
        // Sigma^-1 = K_zz + gram / noise + jitter I   (padding: identity)
        if (rc == VGP_OK) {
            const int64_t count = sig.rows * sig.cols;
            axpby_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(kzz.p, 1.0, gram.p, 1.0 / noise_variance,
                                                                       jitter, m, sig.p, sig.cols, count);
            ++g_launches;
            rc = pad_identity(sig.p, sig.cols, m, sig.rows, s);
        }
        if (rc == VGP_OK) rc = dense_potrf(sig.p, sig.rows, sig.cols, ws, s);
        if (rc == VGP_OK) rc = dense_read_info(ws, nullptr, s);
        // loc = K_zz (L^-T L^-1 v) / noise
        if (rc == VGP_OK) rc = dense_trsm(0, 0, sig.rows, v.cols, 1.0, sig.p, sig.cols, v.p, v.cols, ws, true, s);
        if (rc == VGP_OK) rc = dense_trsm(0, 1, sig.rows, v.cols, 1.0, sig.p, sig.cols, v.p, v.cols, ws, true, s);
        if (rc == VGP_OK) {
            // v is strided ([mp][128], column 0): gather it into a dense vector first
            Buf vd;
            rc = vd.alloc(mp, 1, s, false);
            if (rc == VGP_OK) rc = copy_col0_to_vec(v, m, vd.p);
            if (rc == VGP_OK) {
                gemv_n_kernel<<<(unsigned)m, 256, 0, s>>>(kzz.p, kzz.cols, m, vd.p, 1.0 / noise_variance, 0, loc_dev, 1);
                ++g_launches;
            }
        }
        // scale: L^-1 K_zz (legacy) or its transpose K_zz L^-T (S = scale scale^T = K_zz Sigma K_zz)
        if (rc == VGP_OK) rc = dense_trsm(0, 0, sig.rows, kzz.cols, 1.0, sig.p, sig.cols, kzz.p, kzz.cols, ws, true, s);
        if (rc == VGP_OK) {
            if (legacy_scale_orientation) {
                cudaError_t e = cudaMemcpy2DAsync(scale_dev, (size_t)m * 8, kzz.p, (size_t)kzz.cols * 8, (size_t)m * 8,
                                                  (size_t)m, cudaMemcpyDeviceToDevice, s);
                if (e != cudaSuccess) rc = cuda_fail(e, "scale copy", __FILE__, __LINE__);
            } else {
                rc = kzz2.alloc(m, m, s);
                if (rc == VGP_OK) {
                    dim3 grid((unsigned)((m + 31) / 32), (unsigned)((m + 31) / 32));
                    transpose_kernel<<<grid, 256, 0, s>>>(kzz.p, kzz2.p, m, kzz.cols);
                    ++g_launches;
                    cudaError_t e = cudaMemcpy2DAsync(scale_dev, (size_t)m * 8, kzz2.p, (size_t)kzz2.cols * 8,
                                                      (size_t)m * 8, (size_t)m, cudaMemcpyDeviceToDevice, s);
                    if (e != cudaSuccess) rc = cuda_fail(e, "scale copy", __FILE__, __LINE__);
                }
            }
        }
        if (rc == VGP_OK) {
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) rc = cuda_fail(e, "optimal posterior kernels", __FILE__, __LINE__);
        }
        cudaStreamSynchronize(s);
    }
    ws.release();
    return rc;
}

int vgp_vgp_loss_k(int device, int kind, const double *z_dev, int64_t m, int d, const double *loc_dev,
                 const double *scale_dev, const double *xb_dev, const double *yb_dev, int64_t b,
                 double amplitude, double length_scale, double noise_variance, double kl_weight, double jitter,
                 vgp_vgp_terms *terms_host, void *stream) {
    VGP_REQUIRE(z_dev && loc_dev && scale_dev && xb_dev && yb_dev && terms_host && m > 0 && b > 0, "bad argument");
    VGP_REQUIRE(noise_variance > 0.0, "observation noise variance must be positive");
    VGP_ENTER(device);
    cudaStream_t s = (cudaStream_t)stream;
    DenseWorkspace ws, ws2;
    int rc;
    {
        Buf l, kzb, c, q, a, e, ssum, mu, qa;
        Reducer red;
        rc = red.init(s);
        // L = chol(K_zz + jitter I)
        if (rc == VGP_OK) rc = kernel_cholesky(kind, z_dev, m, d, amplitude, length_scale, jitter, l, ws, s);
        // q = L^-1 q_loc (kept for the KL), then alpha = L^-T q
        if (rc == VGP_OK) rc = q.alloc(m, 1, s);
        if (rc == VGP_OK) rc = copy_vec_to_col0(loc_dev, m, q);
        if (rc == VGP_OK) rc = dense_trsm(0, 0, l.rows, q.cols, 1.0, l.p, l.cols, q.p, q.cols, ws, true, s);
        if (rc == VGP_OK) rc = red.run(SumSqStrided{q.p, q.cols}, m, 4);                  // |L^-1 mu|^2
        if (rc == VGP_OK) rc = dense_trsm(0, 1, l.rows, q.cols, 1.0, l.p, l.cols, q.p, q.cols, ws, true, s);
        // K_zb, mu_b = K_zb^T alpha, ll
        if (rc == VGP_OK) rc = kzb.alloc(m, b, s);
        if (rc == VGP_OK)
            rc = kernel_matrix_dispatch(kind, z_dev, m, xb_dev, b, d, amplitude, length_scale, 0.0, -1 - b, kzb.p, kzb.cols, s);
        if (rc == VGP_OK) rc = mu.alloc(b, 1, s, false);
        if (rc == VGP_OK) {
            gemv_t_kernel<<<(unsigned)((b + 255) / 256), 256, 0, s>>>(kzb.p, kzb.cols, m, b, q.p, q.cols, mu.p);
            ++g_launches;
            rc = red.run(SumSqDiff{yb_dev, mu.p}, b, 0);                                  // |y - mu|^2
        }
        // C = L^-1 K_zb ; D = L^-T C ; E = q_scale^T D
        if (rc == VGP_OK) rc = dense_trsm(0, 0, l.rows, kzb.cols, 1.0, l.p, l.cols, kzb.p, kzb.cols, ws, true, s);
        if (rc == VGP_OK) rc = red.run(SumSqRegion{kzb.p, kzb.cols, b}, m * b, 1);        // |C|_F^2
        if (rc == VGP_OK) rc = dense_trsm(0, 1, l.rows, kzb.cols, 1.0, l.p, l.cols, kzb.p, kzb.cols, ws, true, s);
        if (rc == VGP_OK) rc = a.alloc(m, m, s);
        if (rc == VGP_OK) rc = copy_matrix(scale_dev, m, m, m, a);
        if (rc == VGP_OK) rc = e.alloc(m, b, s);
        if (rc == VGP_OK)
            rc = dense_gemm(1, 0, a.cols, kzb.cols, a.rows, 1.0, a.p, a.cols, kzb.p, kzb.cols, 0.0, e.p, e.cols,
                            GEMM_FULL, s);
        if (rc == VGP_OK) rc = red.run(SumSqRegion{e.p, e.cols, b}, m * b, 2);            // |A^T D|_F^2
        // KL pieces: |L^-1 A|_F^2, logdet K, logdet S (S = A A^T factorised on its own workspace)
        if (rc == VGP_OK) rc = qa.alloc(m, m, s);
        if (rc == VGP_OK) rc = copy_matrix(scale_dev, m, m, m, qa);
        if (rc == VGP_OK) rc = dense_trsm(0, 0, l.rows, qa.cols, 1.0, l.p, l.cols, qa.p, qa.cols, ws, true, s);
        if (rc == VGP_OK) rc = red.run(SumSqRegion{qa.p, qa.cols, m}, m * m, 3);          // tr(K^-1 S)
        if (rc == VGP_OK) rc = red.run(SumLogDiag{l.p, l.cols}, m, 5);
        if (rc == VGP_OK) rc = ssum.alloc(m, m, s);
        if (rc == VGP_OK)
            rc = dense_gemm(0, 1, a.rows, a.rows, a.cols, 1.0, a.p, a.cols, a.p, a.cols, 0.0, ssum.p, ssum.cols,
                            GEMM_FULL, s);
        if (rc == VGP_OK) rc = pad_identity(ssum.p, ssum.cols, m, ssum.rows, s);
        if (rc == VGP_OK) rc = dense_potrf(ssum.p, ssum.rows, ssum.cols, ws2, s);
        if (rc == VGP_OK) rc = dense_read_info(ws2, nullptr, s);
        if (rc == VGP_OK) rc = red.run(SumLogDiag{ssum.p, ssum.cols}, m, 6);
        double h[8] = {0};
        if (rc == VGP_OK) rc = red.fetch(h, 8);
        if (rc == VGP_OK) {
            const double log2pi = log(2.0 * M_PI);
            vgp_vgp_terms t;
            t.ll = -0.5 * h[0] / noise_variance - 0.5 * (double)b * (log2pi + log(noise_variance));
            t.tr1 = (double)b * amplitude * amplitude - h[1];
            t.tr2 = h[2];
            t.kl = 0.5 * (h[3] + h[4] - (double)m + 2.0 * h[5] - 2.0 * h[6]);
            t.loss = -(t.ll - 0.5 * (t.tr1 + t.tr2) / noise_variance - kl_weight * t.kl);
            *terms_host = t;
        }
        cudaStreamSynchronize(s);
    }
    ws.release();
    ws2.release();
    return rc;
}

int vgp_vgp_predict_k(int device, int kind, const double *z_dev, int64_t m, int d, const double *loc_dev,
                    const double *scale_dev, const double *xt_dev, int64_t t, double amplitude,
                    double length_scale, double predictive_noise_variance, double jitter, double *mean_dev,
                    double *var_dev, void *stream) {
    VGP_REQUIRE(z_dev && loc_dev && xt_dev && m > 0 && t > 0, "bad argument");
    VGP_REQUIRE(!var_dev || scale_dev, "variance needs the variational scale");
    VGP_ENTER(device);
    cudaStream_t s = (cudaStream_t)stream;
    DenseWorkspace ws;
    int rc;
    {
        Buf l, kzt, c, q, a, e;
        rc = kernel_cholesky(kind, z_dev, m, d, amplitude, length_scale, jitter, l, ws, s);
        if (rc == VGP_OK) rc = q.alloc(m, 1, s);
        if (rc == VGP_OK) rc = copy_vec_to_col0(loc_dev, m, q);
        if (rc == VGP_OK) rc = dense_trsm(0, 0, l.rows, q.cols, 1.0, l.p, l.cols, q.p, q.cols, ws, true, s);
        if (rc == VGP_OK) rc = dense_trsm(0, 1, l.rows, q.cols, 1.0, l.p, l.cols, q.p, q.cols, ws, true, s);
        if (rc == VGP_OK) rc = kzt.alloc(m, t, s);
        if (rc == VGP_OK)
            rc = kernel_matrix_dispatch(kind, z_dev, m, xt_dev, t, d, amplitude, length_scale, 0.0, -1 - t, kzt.p, kzt.cols, s);
        if (rc == VGP_OK && mean_dev) {
            gemv_t_kernel<<<(unsigned)((t + 255) / 256), 256, 0, s>>>(kzt.p, kzt.cols, m, t, q.p, q.cols, mean_dev);
            ++g_launches;
        }
        if (rc == VGP_OK && var_dev) {
            rc = dense_trsm(0, 0, l.rows, kzt.cols, 1.0, l.p, l.cols, kzt.p, kzt.cols, ws, true, s);      // C
            if (rc == VGP_OK) rc = c.alloc(m, t, s);
            if (rc == VGP_OK) {
                cudaError_t ce = cudaMemcpyAsync(c.p, kzt.p, (size_t)c.rows * c.cols * 8, cudaMemcpyDeviceToDevice, s);
                if (ce != cudaSuccess) rc = cuda_fail(ce, "copy C", __FILE__, __LINE__);
            }
            if (rc == VGP_OK) rc = dense_trsm(0, 1, l.rows, kzt.cols, 1.0, l.p, l.cols, kzt.p, kzt.cols, ws, true, s);   // D
            if (rc == VGP_OK) rc = a.alloc(m, m, s);
            if (rc == VGP_OK) rc = copy_matrix(scale_dev, m, m, m, a);
            if (rc == VGP_OK) rc = e.alloc(m, t, s);
            if (rc == VGP_OK)
                rc = dense_gemm(1, 0, a.cols, kzt.cols, a.rows, 1.0, a.p, a.cols, kzt.p, kzt.cols, 0.0, e.p, e.cols,
                                GEMM_FULL, s);
            if (rc == VGP_OK) {
                colvar_kernel<<<(unsigned)((t + 255) / 256), 256, 0, s>>>(
                    c.p, e.p, c.cols, m, t, amplitude * amplitude + predictive_noise_variance, var_dev);
                ++g_launches;
            }
        }
        if (rc == VGP_OK) {
            cudaError_t ce = cudaGetLastError();
            if (ce != cudaSuccess) rc = cuda_fail(ce, "vgp_predict kernels", __FILE__, __LINE__);
        }
        cudaStreamSynchronize(s);
    }
    ws.release();
    return rc;
}

/* log_prob of the same observations under `batch` hyper-parameter triples params_host [batch][3] = (amplitude,
 * length_scale, noise_variance): the likelihood-surface sweep of gpf.calc_H (gp_functions.py:864-876; 160 x 160
 * evaluations at main.py:400-401) as ONE launch when n <= 127 (one CTA per triple, the whole factorisation in
 * registers), otherwise one vgp_gp_logprob_k per triple.  Non-positive-definite triples return NaN. */
int vgp_gp_logprob_batch_k(int device, int kind, const double *x_dev, int64_t n, int d, const double *y_dev,
                           const double *params_host, int64_t batch, double jitter, double *logprob_host,
                           void *stream) {
    VGP_REQUIRE(x_dev && y_dev && params_host && logprob_host && n > 0 && batch >= 0, "bad argument");
    VGP_REQUIRE(kind >= VGP_KERNEL_EXPQUAD && kind <= VGP_KERNEL_MATERN52, "unknown kernel kind %d", kind);
    if (batch == 0) return VGP_OK;
    if (n > NB - 1) {
        for (int64_t b = 0; b < batch; ++b) {
            int rc = vgp_gp_logprob_k(device, kind, x_dev, n, d, y_dev, params_host[3 * b], params_host[3 * b + 1],
                                      params_host[3 * b + 2], jitter, logprob_host + b, stream);
            if (rc == VGP_ERR_NOT_PD) logprob_host[b] = nan("");
            else if (rc != VGP_OK) return rc;
        }
        return VGP_OK;
    }
    VGP_ENTER(device);
    cudaStream_t s = (cudaStream_t)stream;
    double *params = nullptr, *out = nullptr;
    int *info = nullptr;
    VGP_CUDA(cudaMallocAsync((void **)&params, (size_t)batch * 24, s));
    VGP_CUDA(cudaMallocAsync((void **)&out, (size_t)batch * 8, s));
    VGP_CUDA(cudaMallocAsync((void **)&info, (size_t)batch * 4, s));
    VGP_CUDA(cudaMemcpyAsync(params, params_host, (size_t)batch * 24, cudaMemcpyHostToDevice, s));
    VGP_CUDA(cudaMemsetAsync(info, 0, (size_t)batch * 4, s));
    int rc;
    switch (kind) {
        case VGP_KERNEL_EXPQUAD: rc = launch_gp_batch<VGP_KERNEL_EXPQUAD>(d, x_dev, (int)n, y_dev, params, batch, jitter, out, info, s); break;
        case VGP_KERNEL_MATERN12: rc = launch_gp_batch<VGP_KERNEL_MATERN12>(d, x_dev, (int)n, y_dev, params, batch, jitter, out, info, s); break;
        case VGP_KERNEL_MATERN32: rc = launch_gp_batch<VGP_KERNEL_MATERN32>(d, x_dev, (int)n, y_dev, params, batch, jitter, out, info, s); break;
        default: rc = launch_gp_batch<VGP_KERNEL_MATERN52>(d, x_dev, (int)n, y_dev, params, batch, jitter, out, info, s);
    }
    if (rc == VGP_OK) {
        std::vector<int> flags((size_t)batch);
        cudaError_t e = cudaMemcpyAsync(logprob_host, out, (size_t)batch * 8, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(flags.data(), info, (size_t)batch * 4, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = cuda_fail(e, "batched log_prob", __FILE__, __LINE__);
        for (int64_t b = 0; rc == VGP_OK && b < batch; ++b)
            if (flags[(size_t)b] != 0 && flags[(size_t)b] <= n) logprob_host[b] = nan("");
    }
    cudaFreeAsync(params, s);
    cudaFreeAsync(out, s);
    cudaFreeAsync(info, s);
    return rc;
}

/* ExponentiatedQuadratic forms (the kernel of variational_Gaussian_process_example.py:55-57). */
int vgp_gp_logprob(int device, const double *x_dev, int64_t n, int d, const double *y_dev, double amplitude,
                   double length_scale, double noise_variance, double jitter, double *logprob_host,
                   void *stream) {
    return vgp_gp_logprob_k(device, VGP_KERNEL_EXPQUAD, x_dev, n, d, y_dev, amplitude, length_scale, noise_variance,
                            jitter, logprob_host, stream);
}
int vgp_gp_regression(int device, const double *x_dev, int64_t n, int d, const double *y_dev,
                      const double *xt_dev, int64_t t, double amplitude, double length_scale,
                      double noise_variance, double predictive_noise_variance, double divisor_jitter,
                      double *mean_dev, double *var_dev, void *stream) {
    return vgp_gp_regression_k(device, VGP_KERNEL_EXPQUAD, x_dev, n, d, y_dev, xt_dev, t, amplitude, length_scale,
                               noise_variance, predictive_noise_variance, divisor_jitter, mean_dev, var_dev, stream);
}
int vgp_vgp_optimal_posterior(int device, const double *z_dev, int64_t m, const double *x_dev, int64_t n_obs,
                              int d, const double *y_dev, double amplitude, double length_scale,
                              double noise_variance, double jitter, int legacy_scale_orientation,
                              double *loc_dev, double *scale_dev, void *stream) {
    return vgp_vgp_optimal_posterior_k(device, VGP_KERNEL_EXPQUAD, z_dev, m, x_dev, n_obs, d, y_dev, amplitude,
                                       length_scale, noise_variance, jitter, legacy_scale_orientation, loc_dev,
                                       scale_dev, stream);
}
int vgp_vgp_loss(int device, const double *z_dev, int64_t m, int d, const double *loc_dev,
                 const double *scale_dev, const double *xb_dev, const double *yb_dev, int64_t b,
                 double amplitude, double length_scale, double noise_variance, double kl_weight, double jitter,
                 vgp_vgp_terms *terms_host, void *stream) {
    return vgp_vgp_loss_k(device, VGP_KERNEL_EXPQUAD, z_dev, m, d, loc_dev, scale_dev, xb_dev, yb_dev, b, amplitude,
                          length_scale, noise_variance, kl_weight, jitter, terms_host, stream);
}
int vgp_vgp_predict(int device, const double *z_dev, int64_t m, int d, const double *loc_dev,
                    const double *scale_dev, const double *xt_dev, int64_t t, double amplitude,
                    double length_scale, double predictive_noise_variance, double jitter, double *mean_dev,
                    double *var_dev, void *stream) {
    return vgp_vgp_predict_k(device, VGP_KERNEL_EXPQUAD, z_dev, m, d, loc_dev, scale_dev, xt_dev, t, amplitude,
                             length_scale, predictive_noise_variance, jitter, mean_dev, var_dev, stream);
}

}  // extern "C"
