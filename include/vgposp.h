/*
 * vgposp.h -- C-ABI of libvgposp.so, the B200 (sm_100a) implementation of VGPosp's GP placement hot path.
 *
 * The reference (DL-WG/VGPosp) has no FFI layer: its boundary for this path is a set of Python module
 * functions.  This header is the boundary a maintainer binds instead (ctypes
 * stub: vgposp_b200/_ffi.py; reference-side patch: INTEGRATION.md).  Each entry point cites the
 * reference interface it stands in for, as file:line under the reference tree.
 *
 * Conventions
 *   - every function returns an int status (VGP_OK == 0); vgp_last_error() gives the thread-local text;
 *   - all matrices are float64, row-major, with an explicit leading dimension in ELEMENTS;
 *   - pointers named *_dev are device pointers on `device`; pointers named *_host are host pointers;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Calls are
 *     asynchronous on that stream unless stated otherwise;
 *   - handles are per device and not thread-safe; the only process-wide state is the option table below
 *     (vgp_set_option) and the per-device workspace cache (vgp_workspace_trim).  The library never reads
 *     environment variables;
 *   - there is no CPU fallback anywhere: without a CUDA device every compute call fails with
 *     VGP_ERR_CUDA.
 */
#ifndef VGPOSP_H
#define VGPOSP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

#define VGP_OK 0
#define VGP_ERR_INVALID 1      /* bad argument (shape, alignment, NULL)                         */
#define VGP_ERR_CUDA 2         /* CUDA runtime/driver error, including "no device"              */
#define VGP_ERR_NOT_PD 3       /* matrix not positive definite (the reference would pinv it)    */
#define VGP_ERR_NOMEM 4
#define VGP_ERR_STATE 5        /* call order / handle state                                     */

#define VGP_ABI_VERSION 1

/* ---------------------------------------------------------------- runtime ---------------------------------- */
int vgp_abi_version(void);
const char *vgp_last_error(void);
int vgp_device_count(int *count);
/* name[<=len], SM count, total/free bytes of HBM */
int vgp_device_info(int device, char *name, int len, int *sm_count, size_t *total_bytes, size_t *free_bytes);

/* Process-wide options.  Every rank of a multi-GPU run must hold the same values (vgp_dist_connect verifies the ones
 * that change numerics or buffer sizes).  Defaults in brackets.
 *   GEMM_EMULATE_SLICES  [8]  products of the O(n^3) factorisations (potrf / trtri / lauum) with m, n >= EMULATE_MIN and
 *                             2 k >= EMULATE_MIN run on the int8 tensor cores (tcgen05.mma kind::i8) after an error-free
 *                             split of the FP64 operands into this many 7-bit digit planes (csrc/emulated.cu): 8 planes
 *                             = 36 exact integer products, result within 1e-14 of the FP64 DMMA product.  0 = every
 *                             product on the FP64 tensor pipe (mma.sync DMMA).
 *   GEMM_EMULATE_MIN     [1024] (one-call placement at n = 50 000: factorisation 1.46 s against 1.51 s at 2048, 1.67 s at 4096)
 *   H2D_OVERLAP          [1]  one-call placement: factorise behind the arriving host matrix (0: copy first)
 *   GEMM_TILE_CONFIG     [-1] measurement knob: force the DMMA tile configuration (0 base, 1 pair, 2 tma)
 *   GEMM_SMALL_BELOW     [74] products with fewer 128 x 64 tiles use the 64 x 64 tile shape
 *   DIST_MIN_TILES / DIST_MIN_K [96 / 256] smallest product the distributed factorisation shares out over the ranks
 *   DIST_EMULATE_MIN     [-1] smallest DISTRIBUTED product on the int8 tensor cores.  Every rank cuts the digit planes
 *                             of the whole operands but multiplies only its share of the tiles, so the break-even
 *                             margin shrinks with the rank count.  -1 = GEMM_EMULATE_MIN with 2 ranks, 4096 with 3-4,
 *                             FP64 pipe only with more (n = 50 000 inverse: 2 ranks 1.54 s int8 vs 2.38 s; 4 ranks
 *                             1.28 s at 4096 vs 1.44 s FP64 pipe; 8 ranks 1.24 s at 4096, 1.13 s at 8192, 1.07 s FP64
 *                             pipe alone -- profiles/r02_dist_inverse_bench_g4.jsonl, _g8.jsonl)
 *   ELBO_OVERLAP         [3]  side streams of the ELBO step (0: single stream)
 *   WORKSPACE_CACHE_BYTES [-1] cap of the per-device workspace cache; -1 = half of the device memory, 0 = no caching */
enum {
    VGP_OPT_GEMM_EMULATE_SLICES = 0,
    VGP_OPT_GEMM_EMULATE_MIN = 1,
    VGP_OPT_H2D_OVERLAP = 2,
    VGP_OPT_GEMM_TILE_CONFIG = 3,
    VGP_OPT_GEMM_SMALL_BELOW = 4,
    VGP_OPT_DIST_MIN_TILES = 5,
    VGP_OPT_DIST_MIN_K = 6,
    VGP_OPT_ELBO_OVERLAP = 7,
    VGP_OPT_WORKSPACE_CACHE_BYTES = 8,
    VGP_OPT_DIST_EMULATE_MIN = 9,
    VGP_OPT_COUNT = 10
};
int vgp_set_option(int option, int64_t value);
int vgp_get_option(int option, int64_t *value);
/* The one-call placement entries (vgp_placement_host*) keep their device matrices (2 x 8 n^2 bytes) and the digit-plane
 * workspace of the emulated products in a per-device cache between calls; this hands all of it back to the driver.
 * Any allocation of the library that runs out of memory trims the cache by itself and retries. */
int vgp_workspace_trim(int device, size_t *released_bytes);

int vgp_malloc(int device, size_t bytes, void **ptr_dev);
int vgp_free(int device, void *ptr_dev);
int vgp_host_alloc(size_t bytes, void **ptr_host);          /* pinned */
int vgp_host_free(void *ptr_host);
int vgp_memcpy_h2d(int device, void *dst_dev, const void *src_host, size_t bytes, void *stream);
int vgp_memcpy_d2h(int device, void *dst_host, const void *src_dev, size_t bytes, void *stream);
int vgp_memcpy_d2d(int device, void *dst_dev, const void *src_dev, size_t bytes, void *stream);
/* strided 2-D copies, widths in bytes (cudaMemcpy2DAsync) -- panels of a host matrix to the device */
int vgp_memcpy2d_h2d(int device, void *dst_dev, size_t dpitch, const void *src_host, size_t spitch,
                     size_t width_bytes, size_t rows, void *stream);
int vgp_memcpy2d_d2h(int device, void *dst_host, size_t dpitch, const void *src_dev, size_t spitch,
                     size_t width_bytes, size_t rows, void *stream);
int vgp_memcpy2d_d2d(int device, void *dst_dev, size_t dpitch, const void *src_dev, size_t spitch,
                     size_t width_bytes, size_t rows, void *stream);
int vgp_memset(int device, void *dst_dev, int value, size_t bytes, void *stream);
int vgp_stream_create(int device, void **stream);
int vgp_stream_destroy(int device, void *stream);
int vgp_stream_sync(int device, void *stream);
/* CUDA-event timing on `stream`: start/stop return opaque events, elapsed gives milliseconds */
int vgp_event_record(int device, void *stream, void **event);
int vgp_event_elapsed_ms(int device, void *start_event, void *stop_event, float *ms); /* syncs on stop; frees both */

/* ---------------------------------------------------------------- DLPack ----------------------------------- */
/* Zero-copy view of a DLPack tensor (DLManagedTensor*, the payload of a "dltensor" capsule such as
 * tf.experimental.dlpack.to_dlpack(t) or torch.utils.dlpack.to_dlpack(t)).  The library never takes
 * ownership and never calls the deleter. */
typedef struct vgp_tensor_view {
    void *data;            /* data pointer + byte_offset already applied */
    int32_t device_type;   /* DLDeviceType: 1 CPU, 2 CUDA, 3 CUDA-pinned host */
    int32_t device_id;
    int32_t ndim;
    int32_t dtype_code;    /* DLDataTypeCode: 0 int, 1 uint, 2 float */
    int32_t dtype_bits;
    int32_t contiguous;    /* 1 if row-major contiguous (or strides == NULL) */
    int64_t shape[8];
    int64_t strides[8];    /* in elements */
} vgp_tensor_view;
int vgp_dlpack_view(const void *dl_managed_tensor, vgp_tensor_view *out);

/* ---------------------------------------------------------------- (1) ExpQuad kernel matrix ---------------- */
/* out[i][j] = amplitude^2 * exp(-|x1_i - x2_j|^2 / (2 length_scale^2)) + (i == diag_col0 + j ? diag_add : 0)
 * Replaces tfkern.ExponentiatedQuadratic(amplitude, length_scale).matrix(x1, x2)
 *   (call sites variational_Gaussian_process_example.py:55-57, 3D_sin_wave.py:158-159, main_tests.py:617-619;
 *    factory hook gp_functions.py:160-163).
 * x1_dev [n1, d], x2_dev [n2, d] row-major, d in [1, 8]; out_dev [n1, ld_out].  When x1_dev == x2_dev,
 * n1 == n2 and diag_col0 == 0 the symmetric path computes each off-diagonal tile once. */
int vgp_expquad_matrix(int device, const double *x1_dev, int64_t n1, const double *x2_dev, int64_t n2, int d,
                       double amplitude, double length_scale, double diag_add, int64_t diag_col0,
                       double *out_dev, int64_t ld_out, void *stream);
/* The same builder for the other stationary kernels the reference instantiates (r = |x1_i - x2_j|):
 *   MATERN12  a^2 exp(-r / l)                                   tfkern.MaternOneHalf, gp_functions.py:160-163
 *   MATERN32  a^2 (1 + z) exp(-z),           z = sqrt(3) r / l  tfkern.MaternThreeHalves
 *   MATERN52  a^2 (1 + z + z^2/3) exp(-z),   z = sqrt(5) r / l  tfkern.MaternFiveHalves, main_architecture_2.py:184
 * EXPQUAD is vgp_expquad_matrix.  Values agree with libm-based float64 evaluation to 2 ulp. */
enum { VGP_KERNEL_EXPQUAD = 0, VGP_KERNEL_MATERN12 = 1, VGP_KERNEL_MATERN32 = 2, VGP_KERNEL_MATERN52 = 3 };
int vgp_kernel_matrix(int device, int kind, const double *x1_dev, int64_t n1, const double *x2_dev, int64_t n2,
                      int d, double amplitude, double length_scale, double diag_add, int64_t diag_col0,
                      double *out_dev, int64_t ld_out, void *stream);

/* ---------------------------------------------------------------- (2) dense float64 factorisations ---------- */
/* C[m,n] = alpha * op(A) * op(B) + beta * C.  trans_a/trans_b: 0 = as stored, 1 = transposed.  Hand-written
 * DMMA (mma.sync f64) kernel; stands for the tf.matmul / LinearOperator matmuls TFP issues inside the VGP loss
 * (variational_Gaussian_process_example.py:68-74,96-99). */
int vgp_dgemm(int device, int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, double alpha,
              const double *a_dev, int64_t lda, const double *b_dev, int64_t ldb, double beta, double *c_dev,
              int64_t ldc, void *stream);
/* In-place lower Cholesky A = L L^T (strict upper triangle is left untouched).  Replaces tf.linalg.cholesky
 * inside tfd.GaussianProcess.log_prob / VariationalGaussianProcess (gp_functions.py:166-172,
 * 3D_sin_wave.py:172).  Synchronous w.r.t. the host only for the final status read:
 * *info_host = 0, or 1 + index of the first non-positive pivot (then returns VGP_ERR_NOT_PD). */
int vgp_potrf(int device, double *a_dev, int64_t n, int64_t lda, int *info_host, void *stream);
/* Full symmetric inverse of an SPD matrix, in place (potrf + trtri + lauum + mirror).  Replaces the
 * pseudo-inverses of placement_algorithm2.py:399-405 on the well-conditioned inputs where pinv == inv.
 * When n is a multiple of 128, lda is even and a_dev is 16-byte aligned the matrix is processed where it
 * lies; otherwise through a padded copy (8 n^2 bytes of scratch). */
int vgp_spd_inverse(int device, double *a_dev, int64_t n, int64_t lda, int *info_host, void *stream);
/* Triangular solves with the lower factor: side 0: op(L) X = B (B is [n, nrhs]), side 1: X op(L) = B
 * (B is [nrhs, n]); trans 0: op(L) = L, 1: op(L) = L^T.  In place on B.  Replaces
 * LinearOperatorLowerTriangular.solve in the VGP loss / GPRM (gp_functions.py:290-296). */
int vgp_trsm(int device, int side, int trans, int64_t n, int64_t nrhs, const double *l_dev, int64_t ldl,
             double *b_dev, int64_t ldb, void *stream);

/* Empirical covariance of n locations from S samples each: cov[i][j] = biased sample
 * covariance of rows i and j of m_dev [n, ldm] -- np.cov(tracers_i, tracers_j, bias=True)[0, 1] for every pair
 * (gp_functions.py:1019-1057 create_cov_matrix; tfp.stats.covariance at main_architecture_2.py:431), as one centring
 * pass plus one lower-tile SYRK on the DMMA GEMM; the result is exactly symmetric. */
int vgp_empirical_cov(int device, const double *m_dev, int64_t n, int64_t s_samples, int64_t ldm, double *cov_dev,
                      int64_t ldc, void *stream);
/* The reference's decay filter on such a covariance, in place: cov[i][j] *= exp(-(beta delta_ij)^2 / (2 pi)), delta
 * the Euclidean distance between the integer grid indices idx_dev [n, 3] (int32) of the two locations; factors below
 * `cutoff` (0.01 there) become exact zeros (main_architecture_2_sampledistribution.py:376-394). */
int vgp_cov_taper(int device, double *cov_dev, int64_t n, int64_t ldc, const int *idx_dev, double beta,
                  double cutoff, void *stream);

/* ---------------------------------------------------------------- exact GP (a2, a3) ------------------------- */
/* log N(y | 0, K + (noise + jitter) I), K ExpQuad on x_dev [n, d].  Replaces
 * gpf.fit_gp(kernel, x, noise).log_prob(y) (gp_functions.py:166-172; 3D_sin_wave.py:161-172).
 * Blocking; result in *logprob_host. */
int vgp_gp_logprob(int device, const double *x_dev, int64_t n, int d, const double *y_dev, double amplitude,
                   double length_scale, double noise_variance, double jitter, double *logprob_host,
                   void *stream);
/* Posterior mean [t] and marginal variance [t] at xt_dev [t, d] given observations.  Replaces
 * gpf.tf_gp_regression_model(...).mean() / .variance() (gp_functions.py:283-297). */
int vgp_gp_regression(int device, const double *x_dev, int64_t n, int d, const double *y_dev,
                      const double *xt_dev, int64_t t, double amplitude, double length_scale,
                      double noise_variance, double predictive_noise_variance, double divisor_jitter,
                      double *mean_dev, double *var_dev, void *stream);

/* ---------------------------------------------------------------- variational GP (a4, a5) ------------------- */
typedef struct vgp_vgp_terms {
    double loss, ll, tr1, tr2, kl;
} vgp_vgp_terms;
/* Titsias-optimal q(u): loc [m], scale [m, m] (S = scale scale^T = Kzz Sigma Kzz).  Replaces
 * tfd.VariationalGaussianProcess.optimal_variational_posterior
 * (variational_Gaussian_process_example.py:68-74; main_architecture_2.py:200-206).  K_zx is generated tile by
 * tile and consumed by the m x m SYRK; the m x N matrix is never stored. */
int vgp_vgp_optimal_posterior(int device, const double *z_dev, int64_t m, const double *x_dev, int64_t n_obs,
                              int d, const double *y_dev, double amplitude, double length_scale,
                              double noise_variance, double jitter, int legacy_scale_orientation,
                              double *loc_dev, double *scale_dev, void *stream);
/* Negative ELBO on a minibatch.  Replaces vgp.variational_loss(observations, observation_index_points,
 * kl_weight) (variational_Gaussian_process_example.py:96-99).  Blocking; pieces in *terms_host. */
int vgp_vgp_loss(int device, const double *z_dev, int64_t m, int d, const double *loc_dev,
                 const double *scale_dev, const double *xb_dev, const double *yb_dev, int64_t b,
                 double amplitude, double length_scale, double noise_variance, double kl_weight, double jitter,
                 vgp_vgp_terms *terms_host, void *stream);
/* Predictive mean [t] / marginal variance [t].  Replaces vgp.mean() / vgp.variance()
 * (variational_Gaussian_process_example.py:141-142; main_architecture_2.py:347-360). */
int vgp_vgp_predict(int device, const double *z_dev, int64_t m, int d, const double *loc_dev,
                    const double *scale_dev, const double *xt_dev, int64_t t, double amplitude,
                    double length_scale, double predictive_noise_variance, double jitter, double *mean_dev,
                    double *var_dev, void *stream);

/* The same five calls for any kernel of vgp_kernel_matrix (`kind` = VGP_KERNEL_*): the reference's exact GP is
 * built on tfkern.MaternOneHalf (main.py:94, gp_functions.py:160-163), its 5-D VGP on tfkern.MaternFiveHalves
 * (main_architecture_2.py:184).  The un-suffixed names above are the ExponentiatedQuadratic forms. */
int vgp_gp_logprob_k(int device, int kind, const double *x_dev, int64_t n, int d, const double *y_dev,
                     double amplitude, double length_scale, double noise_variance, double jitter,
                     double *logprob_host, void *stream);
int vgp_gp_regression_k(int device, int kind, const double *x_dev, int64_t n, int d, const double *y_dev,
                        const double *xt_dev, int64_t t, double amplitude, double length_scale,
                        double noise_variance, double predictive_noise_variance, double divisor_jitter,
                        double *mean_dev, double *var_dev, void *stream);
int vgp_vgp_optimal_posterior_k(int device, int kind, const double *z_dev, int64_t m, const double *x_dev,
                                int64_t n_obs, int d, const double *y_dev, double amplitude, double length_scale,
                                double noise_variance, double jitter, int legacy_scale_orientation,
                                double *loc_dev, double *scale_dev, void *stream);
int vgp_vgp_loss_k(int device, int kind, const double *z_dev, int64_t m, int d, const double *loc_dev,
                   const double *scale_dev, const double *xb_dev, const double *yb_dev, int64_t b,
                   double amplitude, double length_scale, double noise_variance, double kl_weight, double jitter,
                   vgp_vgp_terms *terms_host, void *stream);
int vgp_vgp_predict_k(int device, int kind, const double *z_dev, int64_t m, int d, const double *loc_dev,
                      const double *scale_dev, const double *xt_dev, int64_t t, double amplitude,
                      double length_scale, double predictive_noise_variance, double jitter, double *mean_dev,
                      double *var_dev, void *stream);

/* Exact-GP log marginal likelihood and its gradient with respect to (amplitude, length_scale, noise_variance)
 * [grads_host, 3 doubles]: what TF's autodiff hands to tf.train.AdamOptimizer(lr).minimize(-log_likelihood) in
 * gpf.tf_train_gp_adam (gp_functions.py:179-182; main.py:110, lr 0.1).  Blocking. */
int vgp_gp_logprob_grad_k(int device, int kind, const double *x_dev, int64_t n, int d, const double *y_dev,
                          double amplitude, double length_scale, double noise_variance, double jitter,
                          double *logprob_host, double *grads_host, void *stream);

/* log_prob of the same observations under `batch` hyper-parameter triples params_host [batch][3] = (amplitude,
 * length_scale, noise_variance) -> logprob_host [batch]: the likelihood-surface sweep of gpf.calc_H
 * (gp_functions.py:864-876; 160 x 160 evaluations at main.py:400-401, 60 x 60 at main_GP_fit.py:122-123) as one launch
 * when n <= 127 (one CTA per triple, kernel build + Cholesky + solve in registers), otherwise one vgp_gp_logprob_k
 * per triple.  Triples whose matrix is not positive definite give NaN.  Blocking. */
int vgp_gp_logprob_batch_k(int device, int kind, const double *x_dev, int64_t n, int d, const double *y_dev,
                           const double *params_host, int64_t batch, double jitter, double *logprob_host,
                           void *stream);

/* ---------------------------------------------------------------- VGP training step (a4-a7) ---------------- */
/* Reference-faithful ELBO training (variational_Gaussian_process_example.py:47-102): amplitude = softplus(v[0]),
 * length_scale = offset + softplus(v[1]), noise = softplus(v[2]); (loc, scale) are the Titsias optimum over ALL
 * n_obs observations and are re-derived from the trainable parameters every step; the loss is the minibatch
 * variational_loss with kl_weight = batch / n_obs; tf.train.AdamOptimizer(learning_rate) (beta1 .9, beta2 .999,
 * eps 1e-8) updates (v, Z).  The gradient is hand-derived (csrc/elbo.cu).  x_dev [n_obs, d] and y_dev [n_obs] are
 * borrowed for the lifetime of the handle; z_init_host [m, d]. */
typedef struct vgp_elbo vgp_elbo;
int vgp_elbo_create(vgp_elbo **handle, int device, const double *x_dev, const double *y_dev, int64_t n_obs, int d,
                    const double *z_init_host, int64_t m, int64_t batch, double v_amplitude, double v_length_scale,
                    double v_noise, double length_scale_offset, double jitter, double learning_rate);
/* Kernel family of the training step: VGP_KERNEL_EXPQUAD (default), MATERN32 or MATERN52 -- the reference's VGP trains
 * tfkern.MaternFiveHalves over 5-D (x, y, z, t, p) inputs (main_architecture_2.py:184-249).  Call before the first step. */
int vgp_elbo_set_kernel(vgp_elbo *handle, int kind);
int vgp_elbo_destroy(vgp_elbo *handle);
/* Loss and gradient at the current parameters (blocking).  grads_host[3] = d loss / d v; gradz_host [m, d] and
 * terms_host are optional. */
int vgp_elbo_loss_grad(vgp_elbo *handle, const double *xb_dev, const double *yb_dev, double *loss_host,
                       double *grads_host, double *gradz_host, vgp_vgp_terms *terms_host, void *stream);
/* One training step: loss + gradient + Adam update; *loss_host is the loss before the update (what
 * sess.run([train_op, loss]) returns, :123-125). */
int vgp_elbo_step(vgp_elbo *handle, const double *xb_dev, const double *yb_dev, double *loss_host, void *stream);
/* N-axis sharding of the training step over the GPUs of a box: every rank creates its handle
 * on its slice of the observations and registers a sum-all-reduce; the step then calls it at the two places where a
 * sum over ALL observations is needed -- G = K_zx K_zx^T [m_pad^2] with v = K_zx y [m_pad], and the kernel push-through
 * sums [m (2 + d)] -- on device buffers, in stream order (3 calls, ~2 MB per step at m = 512).  Everything of size m
 * runs replicated, so all ranks hold bitwise identical parameters after every step.  n_total = observations over all
 * ranks (sets kl_weight = B / n_total).  The callback returns 0 on success; every rank must pass the same minibatch. */
typedef int (*vgp_allreduce_fn)(void *ctx, double *buf_dev, int64_t count, void *stream);
int vgp_elbo_set_exchange(vgp_elbo *handle, int64_t n_total, vgp_allreduce_fn fn, void *ctx);
int vgp_elbo_get_params(vgp_elbo *handle, double *v3_host, double *z_host, void *stream);
int vgp_elbo_set_params(vgp_elbo *handle, const double *v3_host, const double *z_host, void *stream);
int vgp_elbo_launch_count(vgp_elbo *handle, int64_t *launches);

/* ---------------------------------------------------------------- (3) greedy MI placement ------------------- */
/* One handle holds one device's shard of the candidate set: columns [c0, c0 + nloc) of Sigma and of the
 * precision P of the unselected set, all n rows.  Replaces the state that
 * placement_algorithm2.py:128-145 (alg. 1) / :151-219 (alg. 2) keep in Python lists.
 *
 * Per selection (all kernels on `stream`, no host synchronisation):
 *   vgp_greedy_local_best   delta_j = guard(num_j, 1/P_jj) over the shard, first strict maximum
 *   (exchange 1: every rank's 32-byte candidate record)
 *   vgp_greedy_select       winner over all records: max score, lowest index on exact ties
 *   vgp_greedy_segments     w_J = (Sigma[y,J] - W^T W[:,y]) / sqrt(num_y),  p_J = P[y,J]
 *   (exchange 2: every rank's [w_J | p_J])
 *   vgp_greedy_apply        num_J -= w_J^2;  P[:,J] -= p p_J^T / p_y;  row/column y zeroed
 * With one shard (c0 == 0, nloc == n) vgp_greedy_run does the whole loop inside the library. */
typedef struct vgp_greedy vgp_greedy;

typedef struct vgp_candidate {      /* exchange-1 record, 32 bytes */
    double score;                   /* -inf when the shard has no candidate */
    int64_t index;                  /* global candidate index, -1 when none */
    double num;                     /* numerator sigma^2(y|A) of that candidate (incl. jitter) */
    double pdiag;                   /* P_yy */
} vgp_candidate;

/* small: guard threshold (1e-8, placement_algorithm2.py:116,198; 1e-7 for the TF-graph variant
 * snippets_a2.py:480); jitter: diagonal shift of Sigma_AA (0; 1e-6 for snippets_a2.py:161-163). */
int vgp_greedy_create(vgp_greedy **handle, int device, int64_t n, int64_t c0, int64_t nloc, int64_t kmax,
                      double small, double jitter);
int vgp_greedy_destroy(vgp_greedy *handle);
/* Leading dimension (elements) and device pointers of the handle-owned panels: Sigma[:, J] and P[:, J],
 * both [n_pad, ld] with rows >= n and columns >= nloc zero-padded by the library. */
int vgp_greedy_panels(vgp_greedy *handle, double **cov_dev, double **prec_dev, int64_t *ld, int64_t *n_pad);
/* Single shard only: build P = Sigma^-1 from the Sigma panel already stored in the handle (potrf + potri
 * on device) and reset the selection state.  Blocking for the status read. */
int vgp_greedy_factor(vgp_greedy *handle, int *info_host, void *stream);
/* Reset selection state (num = diag Sigma (+jitter), W empty, nothing taken), P taken as stored. */
int vgp_greedy_reset(vgp_greedy *handle, void *stream);
/* Snapshot / restore of the precision panel (bench warm-up; 8 n ld bytes of extra HBM). */
int vgp_greedy_save_precision(vgp_greedy *handle, void *stream);
int vgp_greedy_restore_precision(vgp_greedy *handle, void *stream);

int vgp_greedy_local_best(vgp_greedy *handle, vgp_candidate *best_dev, void *stream);
int vgp_greedy_select(vgp_greedy *handle, const vgp_candidate *records_dev, int nrecords, void *stream);
/* seg_dev: [2, seg_stride] doubles: row 0 = w_J, row 1 = p_J (entries >= nloc are zero) */
int vgp_greedy_segments(vgp_greedy *handle, double *seg_dev, int64_t seg_stride, void *stream);
/* gathered_dev: [nranks, 2, seg_stride]; bounds_host[nranks + 1]: column ranges of the ranks */
int vgp_greedy_apply(vgp_greedy *handle, const double *gathered_dev, int64_t seg_stride, int nranks,
                     const int64_t *bounds_host, void *stream);
/* Single shard: k further selections enqueued back to back. */
int vgp_greedy_run(vgp_greedy *handle, int64_t k, void *stream);

/* Peer-memory exchange for the shards of ONE box (the reference has no counterpart -- its
 * loop is single-process Python, placement_algorithm2.py:128-145).  Instead of two collective launches per
 * selection, each rank's kernels store their 32-byte candidate record and their [w_J | p_J] segment straight into
 * every peer's mailbox over NVLink/NVSwitch and signal with system-scope release/acquire flags; the consuming
 * kernels wait on those flags.  No host synchronisation: vgp_greedy_run_peer enqueues k selections at once.
 *   comm_create   allocate this rank's mailbox (bounds_host[nranks + 1] = column ranges of all ranks);
 *                 ipc_handle_out (64 bytes, may be NULL) receives its cudaIpcMemHandle_t for other processes,
 *                 mailbox_dev_out (may be NULL) its device address for handles of the same process;
 *   comm_connect  peers: kind 0 = void *[nranks] device pointers (same process), kind 1 = nranks x 64 bytes of
 *                 IPC handles in rank order (cudaIpcOpenMemHandle, peer access enabled lazily).  All ranks must
 *                 have connected (a host barrier) before any of them runs;
 *   run_peer      k further selections; every rank must request the same k;
 *   comm_status   blocking; VGP_ERR_STATE if a wait timed out (~10 s without the peer's flag). */
int vgp_greedy_comm_create(vgp_greedy *handle, int rank, int nranks, const int64_t *bounds_host,
                           void *ipc_handle_out, void **mailbox_dev_out);
int vgp_greedy_comm_connect(vgp_greedy *handle, const void *peers, int kind);
int vgp_greedy_run_peer(vgp_greedy *handle, int64_t k, void *stream);
int vgp_greedy_comm_status(vgp_greedy *handle, int *error_host, void *stream);
/* Results so far (blocking): indices [count], winning scores [count], relative top-2 gap is not tracked. */
int vgp_greedy_results(vgp_greedy *handle, int64_t *count, int64_t *selection_host, double *scores_host,
                       int64_t capacity, void *stream);
/* Dense per-step score vectors for the steps run so far, [count, nloc] (NaN where already selected);
 * only recorded if enabled before the run.  Lets the host replay alg. 2's lazy cache and its prints
 * (placement_algorithm2.py:183-208). */
int vgp_greedy_record_scores(vgp_greedy *handle, int enable);
int vgp_greedy_step_scores(vgp_greedy *handle, double *scores_host, int64_t capacity_rows, void *stream);
/* Number of kernel launches issued by this handle since creation (bench.py's gpu_launches). */
int vgp_greedy_launch_count(vgp_greedy *handle, int64_t *launches);

/* Timing of the dominant kernel (the precision downdate): when enabled, every downdate launch is bracketed by
 * CUDA events on its stream; profile_read returns their summed duration and count (blocking). */
int vgp_greedy_profile(vgp_greedy *handle, int enable);
int vgp_greedy_profile_step_ms(vgp_greedy *handle, double *total_ms);   /* peer path: the step kernel's share */
int vgp_greedy_profile_read(vgp_greedy *handle, double *total_ms, int64_t *launches);

/* ---------------------------------------------------------------- greedy placement, lazy-column formulation -- */
/* Same selections and scores as the vgp_greedy_* path (placement_algorithm2.py:105-145, :151-219, :371-413), but
 * the precision of the unselected set is never rewritten: each step replays the rank-1 history on the winner's
 * column only (SURVEY.md 8d "lazy-column formulation").  Single device, all candidates.
 *   mode 0  P_0 = Sigma^-1 resident (potrf + trtri + lauum); scores bitwise equal to the dense downdate;
 *   mode 1  only M = L^-1 resident (potrf + trtri = 2/3 of the flops); P_0[:, y] = M^T M e_y per selection by a
 *           triangular matrix-vector kernel that streams rows >= y of M once (HBM-bound).
 * matrices: cov_dev [n_pad][ld] is filled by the caller (rows/columns >= n zero) before vgp_lazy_factor;
 * adopt_factor: the caller filled factor_dev itself (P_0 or M for the handle's mode). */
typedef struct vgp_lazy vgp_lazy;
int vgp_lazy_create(vgp_lazy **handle, int device, int64_t n, int64_t kmax, double small, double jitter, int mode);
/* Sharded lazy-column greedy over the ranks of a connected vgp_dist (one box, one process or thread per GPU): mode 1
 * with M = L^-1 in the replicas; per selection the triangular matrix-vector product is split by 512-row blocks over
 * the ranks, every rank stores its blocks' partial sums into all ranks' buffers over NVLink (the tail behind the
 * replicas) and one flag barrier follows; the O(n t) step kernel runs replicated, summing the blocks in the
 * single-device order -- every rank selects the same winners, scores bitwise those of one device.  Every rank: fill
 * cov_dev (vgp_lazy_matrices) and the replica with Sigma, vgp_dist_factor_inverse, vgp_lazy_adopt_factor, vgp_lazy_run. */
struct vgp_dist;
int vgp_lazy_create_dist(vgp_lazy **handle, struct vgp_dist *dist, int64_t n, int64_t kmax, double small,
                         double jitter);
int vgp_lazy_destroy(vgp_lazy *handle);
int vgp_lazy_matrices(vgp_lazy *handle, double **cov_dev, double **factor_dev, int64_t *ld);
int vgp_lazy_factor(vgp_lazy *handle, int *info_host, void *stream);
int vgp_lazy_adopt_factor(vgp_lazy *handle, void *stream);
int vgp_lazy_reset(vgp_lazy *handle, void *stream);
int vgp_lazy_run(vgp_lazy *handle, int64_t k, void *stream);
int vgp_lazy_results(vgp_lazy *handle, int64_t *count, int64_t *selection_host, double *scores_host,
                     int64_t capacity, void *stream);
/* Algorithm 3, the local-kernel greedy of snippets_a3.sparse_placement_algorithm_3 (snippets_a3.py:43-364): the
 * candidates are the points of an i0 x i1 x i2 grid (index = i2 i1 a + i2 b + c); after each selection only the
 * deltas inside the index box [i - cutoff, i + cutoff) per axis around the winner are re-evaluated, the rest of the
 * cache stays stale; recorded step scores are then the columns of delta_cached_iters.  cutoff = 0: exact greedy. */
int vgp_lazy_set_local(vgp_lazy *handle, int64_t i0, int64_t i1, int64_t i2, int64_t cutoff);
int vgp_lazy_record_scores(vgp_lazy *handle, int enable);
int vgp_lazy_step_scores(vgp_lazy *handle, double *scores_host, int64_t capacity_rows, void *stream);
int vgp_lazy_launch_count(vgp_lazy *handle, int64_t *launches);
/* Returns the summed duration / count of the trigemv launches since profiling was last enabled, then sets it. */
int vgp_lazy_profile(vgp_lazy *handle, int enable, double *total_ms, int64_t *launches);

/* One-call form of placement_algorithm_1/2(cov_vv, k) (placement_algorithm2.py:128,151) for a HOST matrix:
 * H2D, factor, k selections, D2H.  cov_host [n, ld_host] row-major float64 (pageable or pinned).
 * seconds_host (optional, may be NULL): [h2d, factor, select, total] seconds measured with CUDA events.  In the lazy
 * formulations cov_vv must be symmetric and only its lower triangle is read: it is copied in row chunks on a copy
 * stream while the Cholesky of the leading blocks already runs (h2d and factor then both start at the first byte and
 * overlap; total is not their sum).
 * formulation: DENSE = precision downdate (vgp_greedy_*), LAZY_PRECISION / LAZY_FACTOR = vgp_lazy_* mode 0 / 1,
 * AUTO = LAZY_FACTOR when 35 k < n (its cheaper setup wins), else LAZY_PRECISION.  vgp_placement_host == AUTO. */
enum { VGP_FORMULATION_AUTO = -1, VGP_FORMULATION_DENSE = 0, VGP_FORMULATION_LAZY_PRECISION = 1,
       VGP_FORMULATION_LAZY_FACTOR = 2 };
int vgp_placement_host_ex(int device, const double *cov_host, int64_t n, int64_t ld_host, int64_t k, double small,
                          double jitter, int formulation, int64_t *selection_host, double *scores_host,
                          double *step_scores_host, double *seconds_host);
int vgp_placement_host(int device, const double *cov_host, int64_t n, int64_t ld_host, int64_t k, double small,
                       double jitter, int64_t *selection_host, double *scores_host, double *step_scores_host,
                       double *seconds_host);
/* The same call for a positive SEMI-definite cov_vv of numerical rank r < n (e.g. an empirical covariance M M^T / S
 * over n > S locations: main_architecture_2.py:391-444, gp_functions.py:1019-1057), where the reference's
 * np.linalg.pinv (placement_algorithm2.py:399-405) is a true pseudo-inverse and vgp_placement_host* return
 * VGP_ERR_NOT_PD (their Cholesky meets a pivot below 1e-12 of the largest variance).  Conditional variances are taken
 * in the factor space of a pivoted Cholesky (csrc/pinv.cu): sigma^2(y | A) as the distance of f_y from span F_A,
 * sigma^2(y | Abar \ y) = 0 unless y is essential for the span of Abar (then 1 / (Sigma_AbarAbar^+)_yy), guard and
 * first-strict-maximum rule as in the reference (:116-123).  algorithm = 1: placement_algorithm_1 (fresh deltas every
 * selection); 2: placement_algorithm_2 (its lazy cache, :151-219 -- on these inputs the two differ, because a delta can
 * rise from 0 to a positive value between selections).  The whole symmetric matrix is read.  max_rank <= 0: n.
 * rank_host (optional) receives r; seconds_host [4] = [0, H2D + factor, selections, total] (CUDA events).
 * In the reference's own regime (r << n) every delta is 0 and the selection is [0, 1, ..., k - 1], as the reference's. */
int vgp_placement_host_pinv(int device, const double *cov_host, int64_t n, int64_t ld_host, int64_t k, double small,
                            int algorithm, int64_t max_rank, int64_t *selection_host, double *scores_host, double *step_scores_host,
                            int64_t *rank_host, double *seconds_host);
/* Host wall-clock breakdown of the calling thread's last vgp_placement_host_ex (lazy formulations), in seconds:
 * [0] state allocation + initialisation (a hit in the per-device workspace cache after the first call), [1] enqueue of
 * copies, factorisation and selections, [2] release of the state (back into the cache), [3] the whole call. */
int vgp_placement_host_wall(double *seconds4);

/* ---------------------------------------------------------------- distributed SPD inverse (setup, 8e) ------- */
/* P = Sigma^-1 over the GPUs of one box, one rank per GPU (process or thread).  Stands for the same pseudo-inverses
 * as vgp_spd_inverse (placement_algorithm2.py:399-413) when the candidate set is sharded.  Every rank holds a
 * full replica [n_pad][n_pad] (n_pad = n rounded up to 128, ld = n_pad) that the caller fills identically on all
 * ranks (or on one rank followed by vgp_dist_push_rows + vgp_dist_barrier); the large GEMMs of potrf / trtri /
 * lauum are split by output tile over the ranks and each finished tile is stored into every replica by the GEMM
 * epilogue itself (peer stores over NVLink) -- the result in every replica is bitwise the single-GPU one.
 *   create    allocates replica + barrier words; ipc_out (128 bytes, may be NULL) = their two cudaIpcMemHandle_t,
 *             ptrs_out (void *[2], may be NULL) = their device addresses for ranks of the same process;
 *   connect   peers: kind 0 = void *[nranks][2] device pointers, kind 1 = nranks x 128 bytes of IPC handles;
 *   spd_inverse  collective (every rank calls it); blocking; the padding diagonal must hold ones. */
typedef struct vgp_dist vgp_dist;
int vgp_dist_create(vgp_dist **handle, int device, int rank, int nranks, int64_t n, void *ipc_out, void **ptrs_out);
int vgp_dist_destroy(vgp_dist *handle);
int vgp_dist_matrix(vgp_dist *handle, double **matrix_dev, int64_t *ld);
int vgp_dist_connect(vgp_dist *handle, const void *peers, int kind);
int vgp_dist_push_rows(vgp_dist *handle, int64_t row0, int64_t row1, void *stream);
/* Rows [row0, row1), columns [0, ncols) of a host matrix (host_ld doubles per row; pinned memory for full speed) into
 * this replica and every other one: chunked upload with the peer copies of each chunk under the next upload.  ncols =
 * VGP_UPLOAD_LOWER uploads just the lower triangle's share of the rows, chunk by chunk (all that the factorisation and
 * the lazy-column greedy read of a symmetric Sigma). */
#define VGP_UPLOAD_LOWER (-1)
int vgp_dist_upload_rows(vgp_dist *handle, const double *host, int64_t host_ld, int64_t row0, int64_t row1,
                         int64_t ncols, void *stream);
int vgp_dist_barrier(vgp_dist *handle, void *stream);
int vgp_dist_spd_inverse(vgp_dist *handle, int *info_host, void *stream);
/* potrf + trtri only: the replicas end up holding M = L^-1 (lower triangle; 2/3 of the flops of the inverse) -- the
 * state of the sharded lazy-column greedy (vgp_lazy_create_dist).  vgp_dist_add_diag: replica[i][i] += value, i < n. */
int vgp_dist_factor_inverse(vgp_dist *handle, int *info_host, void *stream);
int vgp_dist_add_diag(vgp_dist *handle, int64_t n, double value, void *stream);
int vgp_dist_stats(vgp_dist *handle, int64_t *distributed_gemms, int64_t *barriers);

/* ---------------------------------------------------------------- FP64-class GEMM on the int8 tensor cores ---- */
/* C[m][n] = alpha op(A) op(B) + beta C with FP64-class accuracy, the products taken exactly on tcgen05.mma kind::i8
 * (int32 accumulators in tensor memory) after an error-free split of the operands into `slices` (2..8) 7-bit digit
 * planes (Ozaki scheme; csrc/emulated.cu).  Same operand convention as vgp_dgemm's internals: trans_a == 0: A stored
 * [m][k], 1: [k][m]; trans_b == 0: B stored [k][n], 1: [n][k].  lower != 0: only the 128 x 128 tiles on or below the
 * diagonal.  C must not overlap A or B.  This is what the large products of potrf / trtri / lauum run on by default
 * (VGP_OPT_GEMM_EMULATE_SLICES; the seed inverse behind placement_algorithm2.py:399-413): 59 TFLOP/s FP64-equivalent
 * at 8192^3 against 36 for DGEMM on the FP64 pipe, max relative difference 1e-14 (profiles/r02_*). */
int vgp_gemm_emulated(int device, int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, double alpha,
                      const double *a_dev, int64_t lda, const double *b_dev, int64_t ldb, double beta, double *c_dev,
                      int64_t ldc, int slices, int lower, void *stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* VGPOSP_H */
