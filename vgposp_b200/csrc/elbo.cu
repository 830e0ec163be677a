// VGP ELBO training step, "reference-faithful" mode:
//
//   variational_Gaussian_process_example.py:47-61   amplitude = softplus(v_a), length_scale = 1e-5 + softplus(v_l),
//                                                   noise = softplus(v_s)
//   :68-74   (loc, scale) = optimal_variational_posterior over the FULL training set -- functions of the
//            trainable kernel parameters and inducing points, re-evaluated every step
//   :96-99   loss = variational_loss(minibatch, kl_weight = B / N)
//   :101-102 tf.train.AdamOptimizer(0.01).minimize(loss)  on (v_a, v_l, v_s, Z)
//
// TF obtains the gradient by autodiff through TFP; here it is hand-derived.  With K = k(Z,Z), Q = (K + eI)^-1,
// G = K_zx K_zx^T, v = K_zx y (all N observations), Gb = K_zb K_zb^T, vb = K_zb y_b (minibatch), b = 1/noise,
// M = K + b G + eI, Sigma = M^-1, u = Sigma v, mu = b K u (= loc), S = K Sigma K (= scale scale^T), alpha = Q mu,
// R = Q Gb Q, the loss of SURVEY Appendix A.2 is
//
//   loss = b/2 (y_b.y_b - 2 alpha.vb + alpha^T Gb alpha) + B/2 (log 2pi + log noise)         (-ll)
//        + b/2 (B a^2 - tr(Q Gb)) + b/2 tr(S R)                                             (trace terms)
//        + w/2 (tr(Q S) + mu^T Q mu - m + logdet(K + eI) - 2 logdet K + logdet M)           (w KL)
//
// i.e. the data enter only through m x m sufficient statistics.  The reverse sweep runs in m-space (a dozen
// m^3 GEMMs), yields the adjoints of K, G, v, Gb, vb, and is pushed through the ExpQuad kernel to
// (a, l, Z) by one more GEMM per data block (W = 2 Gbar K_zx) and a fused reduction over K_zx
// (`kernback_kernel`).  The formulas were checked against torch autograd of the forward
// (oracle/gp_oracle_torch.py) to 1e-13 before being written here; tests/test_gpu_elbo.py repeats that check
// against this implementation.
//
// FLOP per step at m = 512, N = 200k: 2 m^2 N (SYRK as GEMM) + 2 m^2 N (W) = 2.1e11 -> FP64 tensor pipe bound.
#include <stdlib.h>

#include <new>

#include "gp_common.cuh"

namespace vgp {
namespace elbo_detail {

constexpr int MAXT = 12;
struct LinComb {                       // out = sum_t c_t * (T_t ? A_t^T : A_t) + sum_r c_r * a_r b_r^T
    const double *mat[MAXT];
    double coef[MAXT];
    int trans[MAXT];
    int nmat;
    const double *ra[4];
    const double *rb[4];
    double rcoef[4];
    int nrank;
};

__global__ void __launch_bounds__(256) lincomb_kernel(LinComb lc, double *out, int64_t mp) {
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= mp * mp) return;
    const int64_t i = e / mp, j = e % mp;
    double acc = 0.0;
    for (int t = 0; t < lc.nmat; ++t) acc += lc.coef[t] * (lc.trans[t] ? lc.mat[t][j * mp + i] : lc.mat[t][e]);
    for (int r = 0; r < lc.nrank; ++r) acc += lc.rcoef[r] * lc.ra[r][i] * lc.rb[r][j];
    out[e] = acc;
}

struct DotRegion {          // sum_{i<m, j<m} a[i][j] b[i][j]
    const double *a, *b;
    int64_t ld, m;
    __device__ double operator()(int64_t e) const {
        const int64_t o = (e / m) * ld + e % m;
        return a[o] * b[o];
    }
};
struct DotVec {
    const double *a, *b;
    __device__ double operator()(int64_t e) const { return a[e] * b[e]; }
};
struct ColSum {             // sum_i acc[i][c] of a [m][stride] table
    const double *a;
    int64_t stride;
    int c;
    __device__ double operator()(int64_t e) const { return a[e * stride + c]; }
};

// y = scale * A x (A [m][ld], dense vectors), block per row
__global__ void __launch_bounds__(256) matvec_kernel(const double *a, int64_t ld, int64_t cols, const double *x,
                                                     double scale, double *y) {
    __shared__ double sh[256];
    const int64_t k = blockIdx.x;
    double acc = 0.0;
    for (int64_t j = threadIdx.x; j < cols; j += 256) acc = fma(a[k * ld + j], x[j], acc);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) y[k] = scale * sh[0];
}

// z = ca * a + cb * b (vectors; b may be NULL)
__global__ void axpby_vec_kernel(const double *a, double ca, const double *b, double cb, double *z, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] = ca * a[i] + (b ? cb * b[i] : 0.0);
}

// Push an adjoint W of a kernel block K_ij = k(z_i, x_j) through the kernel.  t_ij = (W_ij + vbar_i y_j) K_ij;
// rowacc[i] += ( sum_j t_ij,  sum_j q_ij t_ij r_ij^2,  zscale * sum_j q_ij t_ij (x_j - z_i) ).  One CTA per inducing
// point, fixed reduction tree, launches accumulate in stream order: deterministic.
// q carries the kernel family: every stationary kernel here has  dK/dl = q K r^2 / l^3  and  dK/dz_i = q K (x_j - z_i)
// / l^2  with   ExpQuad q = 1;   Matern-3/2 (u = sqrt(3) r / l) q = 3 / (1 + u);   Matern-5/2 (u = sqrt(5) r / l)
// q = 5 (1 + u) / (3 + 3 u + u^2)   -- ratios of the kernel's own polynomial factors, so the stored K is reused and no
// second exponential is needed (main_architecture_2.py:184 trains Matern-5/2 over 5-D inputs).
template <int KIND>
__device__ __forceinline__ double kernel_q(double r2, double ucoef) {
    if (KIND == VGP_KERNEL_EXPQUAD) return 1.0;
    const double u = ucoef * sqrt(r2);
    if (KIND == VGP_KERNEL_MATERN32) return 3.0 / (1.0 + u);
    return 5.0 * (1.0 + u) / (3.0 + u * (3.0 + u));
}

template <int D, int KIND>
__global__ void __launch_bounds__(256) kernback_kernel(const double *__restrict__ w, const double *__restrict__ kmat,
                                                       int64_t ld, const double *__restrict__ z,
                                                       const double *__restrict__ x, int64_t n2,
                                                       const double *__restrict__ vbar,
                                                       const double *__restrict__ yvec, double zscale, double ucoef,
                                                       double *rowacc) {
    __shared__ double sh[2 + D][256];
    const int64_t i = blockIdx.x;
    double zi[D];
#pragma unroll
    for (int k = 0; k < D; ++k) zi[k] = z[i * D + k];
    const double vb = vbar ? vbar[i] : 0.0;
    double acc[2 + D];
#pragma unroll
    for (int k = 0; k < 2 + D; ++k) acc[k] = 0.0;
    for (int64_t j = threadIdx.x; j < n2; j += 256) {
        double wij = w[i * ld + j];
        if (vbar) wij = fma(vb, yvec[j], wij);
        const double t = wij * kmat[i * ld + j];
        double r2 = 0.0, dk[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            dk[k] = x[j * D + k] - zi[k];
            r2 = fma(dk[k], dk[k], r2);
        }
        const double tq = t * kernel_q<KIND>(r2, ucoef);
#pragma unroll
        for (int k = 0; k < D; ++k) acc[2 + k] = fma(tq, dk[k], acc[2 + k]);
        acc[0] += t;
        acc[1] = fma(tq, r2, acc[1]);
    }
#pragma unroll
    for (int k = 0; k < 2 + D; ++k) sh[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) {
#pragma unroll
            for (int k = 0; k < 2 + D; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x < 2 + D) {
        const double s = threadIdx.x >= 2 ? zscale : 1.0;
        rowacc[i * (2 + D) + threadIdx.x] += s * sh[threadIdx.x][0];
    }
}

// tf.train.AdamOptimizer update of Z: grad = rowacc[i][2 + k] * gscale
__global__ void adam_z_kernel(double *z, double *mz, double *vz, const double *rowacc, int64_t m, int d, double gscale,
                              double lr_t, double b1, double b2, double eps) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m * d) return;
    const double g = rowacc[(e / d) * (2 + d) + 2 + e % d] * gscale;
    const double mm = b1 * mz[e] + (1.0 - b1) * g;
    const double vv = b2 * vz[e] + (1.0 - b2) * g * g;
    mz[e] = mm;
    vz[e] = vv;
    z[e] -= lr_t * mm / (sqrt(vv) + eps);
}

__global__ void gradz_kernel(const double *rowacc, int64_t m, int d, double gscale, double *out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < m * d) out[e] = rowacc[(e / d) * (2 + d) + 2 + e % d] * gscale;
}

double softplus_h(double v) { return v > 30.0 ? v : log1p(exp(v)); }
double sigmoid_h(double v) { return 1.0 / (1.0 + exp(-v)); }

enum MatId { K_, Q_, SG_, KINV_, G_, GB_, T1_, R_, T2_, S_, T2R_, RS_, RSQ_, KR_, KRK_, QS_, QSQ_, T2Q_, KQ_, KQK_,
             KBAR_, SGBAR_, TMP_, MBAR_, GBAR2_, GBBAR2_, NMAT_ };
enum VecId { V_, VB_, U_, KU_, MU_, AL_, GBAL_, ALBAR_, QA_, MUBAR_, UBAR_, VBAR_, VBBAR_, NVEC_ };
enum Scal { S_YY, S_ALVB, S_ALGBAL, S_TRQGB, S_TRSR, S_TRQS, S_MUAL, S_LDK, S_LDKT, S_LDM, S_MUBKU, S_TRMBG, S_ACC0,
            S_ACC1, S_N };

}  // namespace elbo_detail
}  // namespace vgp

using namespace vgp;
using namespace vgp::elbo_detail;

struct vgp_elbo {
    int device = 0;
    int64_t n = 0, m = 0, b = 0, mp = 0, np_ = 0, bp = 0;
    int d = 0;
    const double *x = nullptr, *y = nullptr;
    double v[3] = {0, 0, 0};            // unconstrained amplitude / length_scale / noise
    double am[3] = {0, 0, 0}, av[3] = {0, 0, 0};
    double ls_offset = 1e-5, jitter = 1e-6, lr = 0.01, b1 = 0.9, b2 = 0.999, eps = 1e-8;
    int64_t t = 0;
    double *z = nullptr, *mz = nullptr, *vz = nullptr;
    double *mats = nullptr, *vecs = nullptr, *kzx = nullptr, *wzx = nullptr, *kzb = nullptr, *wzb = nullptr;
    double *partial = nullptr, *rowacc = nullptr, *gradz = nullptr;
    int splits_n = 1, splits_b = 1;
    DenseWorkspace ws[3];
    // N-axis sharding: this handle holds n of n_total observations; the two sums over the
    // observations (G = K_zx K_zx^T with v = K_zx y, and the kernel push-through sums) are all-reduced by the caller
    int64_t n_total = 0;
    vgp_allreduce_fn allreduce = nullptr;
    void *allreduce_ctx = nullptr;
    cudaStream_t side[2] = {nullptr, nullptr};      // the two inverses that only need K_zz run beside the N-sized work
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    cudaEvent_t ev_gb = nullptr, ev_bar = nullptr, ev_mini = nullptr;
    double *rowacc2 = nullptr;                      // push-through sums of the minibatch / K_zz parts (side stream)
    int64_t launches = 0;
    int kind = VGP_KERNEL_EXPQUAD;                  // vgp_elbo_set_kernel
    double cur_length_scale = 1.0;                  // of the step in flight (kernback's u = c r / l)
    double last_terms[5] = {0, 0, 0, 0, 0};
    double *mat(int id) const { return mats + (size_t)id * mp * mp; }
    double *vec(int id) const { return vecs + (size_t)id * mp; }
};

namespace vgp {
namespace elbo_detail {

template <int D>
int launch_kernback(vgp_elbo *h, const double *w, const double *kmat, int64_t ld, const double *x2, int64_t n2,
                    const double *vbar, const double *yvec, double zscale, double *acc, cudaStream_t s) {
    const unsigned grid = (unsigned)h->m;
    const double l = h->cur_length_scale;
    if (h->kind == VGP_KERNEL_EXPQUAD)
        kernback_kernel<D, VGP_KERNEL_EXPQUAD><<<grid, 256, 0, s>>>(w, kmat, ld, h->z, x2, n2, vbar, yvec, zscale, 0.0, acc);
    else if (h->kind == VGP_KERNEL_MATERN32)
        kernback_kernel<D, VGP_KERNEL_MATERN32><<<grid, 256, 0, s>>>(w, kmat, ld, h->z, x2, n2, vbar, yvec, zscale,
                                                                      sqrt(3.0) / l, acc);
    else
        kernback_kernel<D, VGP_KERNEL_MATERN52><<<grid, 256, 0, s>>>(w, kmat, ld, h->z, x2, n2, vbar, yvec, zscale,
                                                                      sqrt(5.0) / l, acc);
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

int kernback(vgp_elbo *h, const double *w, const double *kmat, int64_t ld, const double *x2, int64_t n2,
             const double *vbar, const double *yvec, double zscale, double *acc, cudaStream_t s) {
    switch (h->d) {
        case 1: return launch_kernback<1>(h, w, kmat, ld, x2, n2, vbar, yvec, zscale, acc, s);
        case 2: return launch_kernback<2>(h, w, kmat, ld, x2, n2, vbar, yvec, zscale, acc, s);
        case 3: return launch_kernback<3>(h, w, kmat, ld, x2, n2, vbar, yvec, zscale, acc, s);
        case 4: return launch_kernback<4>(h, w, kmat, ld, x2, n2, vbar, yvec, zscale, acc, s);
        case 5: return launch_kernback<5>(h, w, kmat, ld, x2, n2, vbar, yvec, zscale, acc, s);
        case 6: return launch_kernback<6>(h, w, kmat, ld, x2, n2, vbar, yvec, zscale, acc, s);
        case 7: return launch_kernback<7>(h, w, kmat, ld, x2, n2, vbar, yvec, zscale, acc, s);
        case 8: return launch_kernback<8>(h, w, kmat, ld, x2, n2, vbar, yvec, zscale, acc, s);
    }
    return VGP_ERR_INVALID;
}

struct Terms {
    LinComb lc;
    Terms() {
        lc.nmat = 0;
        lc.nrank = 0;
    }
    Terms &add(const double *a, double c, int trans = 0) {
        lc.mat[lc.nmat] = a;
        lc.coef[lc.nmat] = c;
        lc.trans[lc.nmat] = trans;
        ++lc.nmat;
        return *this;
    }
    Terms &sym(const double *a, double c) { return add(a, c, 0).add(a, c, 1); }     // c (A + A^T)
    Terms &rank1(const double *a, const double *b, double c) {
        lc.ra[lc.nrank] = a;
        lc.rb[lc.nrank] = b;
        lc.rcoef[lc.nrank] = c;
        ++lc.nrank;
        return *this;
    }
};

int lincomb(const Terms &t, double *out, int64_t mp, cudaStream_t s) {
    lincomb_kernel<<<(unsigned)((mp * mp + 255) / 256), 256, 0, s>>>(t.lc, out, mp);
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

// *out = sum_{i < m} log L[i][i]: one block, fixed tree (runs on whichever stream factorised L)
__global__ void __launch_bounds__(256) logdiag_kernel(const double *l, int64_t ld, int64_t m, double *out) {
    __shared__ double sh[256];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < m; i += 256) acc += log(l[i * ld + i]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}

// dst = (src + shift I on the valid block, identity on the padding)^-1, log-determinant of the valid block -> slot
int invert(vgp_elbo *h, const double *src, double shift, double *dst, DenseWorkspace &ws, Reducer &red, int slot,
           cudaStream_t s) {
    const int64_t mp = h->mp, count = mp * mp;
    axpby_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(src, 1.0, nullptr, 0.0, shift, h->m, dst, mp, count);
    VGP_LAUNCH_CHECK();
    VGP_TRY(pad_identity(dst, mp, h->m, mp, s));
    VGP_TRY(dense_potrf(dst, mp, mp, ws, s));
    logdiag_kernel<<<1, 256, 0, s>>>(dst, mp, h->m, red.scal + slot);
    VGP_LAUNCH_CHECK();
    VGP_TRY(dense_trtri(dst, mp, mp, ws, s));
    VGP_TRY(dense_lauum(dst, mp, mp, ws, s));
    return dense_mirror_lower(dst, mp, mp, s);
}

int matvec(const double *a, int64_t ld, int64_t rows, int64_t cols, const double *x, double scale, double *y,
           cudaStream_t s) {
    matvec_kernel<<<(unsigned)rows, 256, 0, s>>>(a, ld, cols, x, scale, y);
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

// The whole forward + reverse sweep.  On return (after the final synchronisation) h->rowacc holds the kernel
// push-through sums, `sc` the scalars; loss and the three scalar gradients are assembled on the host.
int loss_and_grad(vgp_elbo *h, const double *xb, const double *yb, double *loss_out, double grads[3],
                  double *z_gscale, cudaStream_t s) {
    const int64_t m = h->m, mp = h->mp, n = h->n, b = h->b;
    const double a = softplus_h(h->v[0]), l = h->ls_offset + softplus_h(h->v[1]), noise = softplus_h(h->v[2]);
    const double beta = 1.0 / noise, w = (double)b / (double)(h->n_total > 0 ? h->n_total : n), eps = h->jitter;
    auto exchange = [&](double *buf, int64_t count) -> int {
        if (!h->allreduce) return VGP_OK;
        if (h->allreduce(h->allreduce_ctx, buf, count, (void *)s) != 0) {
            set_error("ELBO step: the all-reduce callback failed");
            return VGP_ERR_STATE;
        }
        return VGP_OK;
    };
    const int64_t before = g_launches;
    Reducer red;
    VGP_TRY(red.init(s));
    auto M = [&](int id) { return h->mat(id); };
    auto V = [&](int id) { return h->vec(id); };
    auto mm_on = [&](cudaStream_t st, const double *A, const double *B, double *C) {
        return dense_gemm(0, 0, mp, mp, mp, 1.0, A, mp, B, mp, 0.0, C, mp, GEMM_FULL, st);
    };
    auto mm = [&](const double *A, const double *B, double *C) { return mm_on(s, A, B, C); };
    // measurement knob: bit 0 = the K/Q/Gb products on a side stream, bit 1 = the minibatch push-through on a side stream
    const int overlap = (int)option(VGP_OPT_ELBO_OVERLAP);
    cudaStream_t sb = (overlap & 1) ? h->side[0] : s;       // products that need only K, Q, Gb
    cudaStream_t sd = (overlap & 2) ? h->side[1] : s;       // minibatch / K_zz push-through
    VGP_CUDA(cudaMemsetAsync(h->rowacc, 0, (size_t)m * (2 + h->d) * 8, s));
    VGP_CUDA(cudaMemsetAsync(h->rowacc2, 0, (size_t)m * (2 + h->d) * 8, s));
    VGP_CUDA(cudaMemsetAsync(h->vecs, 0, (size_t)NVEC_ * mp * 8, s));

    // ---- kernel blocks and sufficient statistics ---------------------------------------------------
    h->cur_length_scale = l;
    VGP_TRY(kernel_matrix_dispatch(h->kind, h->z, m, h->z, m, h->d, a, l, 0.0, 0, M(K_), mp, s));
    // fork: (K + eps I)^-1 and K^-1 depend on K_zz only; their latency-bound chains of small kernels run on two side
    // streams underneath the K_zx build and the m x m x N SYRK
    VGP_CUDA(cudaEventRecord(h->ev_fork, s));
    VGP_CUDA(cudaStreamWaitEvent(h->side[0], h->ev_fork, 0));
    VGP_CUDA(cudaStreamWaitEvent(h->side[1], h->ev_fork, 0));
    VGP_TRY(invert(h, M(K_), eps, M(Q_), h->ws[0], red, S_LDKT, h->side[0]));
    VGP_TRY(invert(h, M(K_), 0.0, M(KINV_), h->ws[2], red, S_LDK, h->side[1]));
    VGP_CUDA(cudaEventRecord(h->ev_join[1], h->side[1]));
    VGP_TRY(kernel_matrix_dispatch(h->kind, h->z, m, h->x, n, h->d, a, l, 0.0, -1 - n, h->kzx, h->np_, s));
    VGP_TRY(kernel_matrix_dispatch(h->kind, h->z, m, xb, b, h->d, a, l, 0.0, -1 - b, h->kzb, h->bp, s));
    // The two observation-sized products (G = K_zx K_zx^T here, W = 2 G_bar K_zx below: 2.1e11 flop of the step's 2.2e11)
    // run on the int8 tensor cores like the factorisations' large products (emulated.cu; k(x, y) <= a^2 bounds every
    // entry of K_zx, so its row exponents need no pass over the data); VGP_OPT_GEMM_EMULATE_SLICES = 0 or a small
    // problem keeps them on the FP64 pipe.
    const int emu = (int)option(VGP_OPT_GEMM_EMULATE_SLICES);
    const bool emulate = emu >= 2 && (double)mp * (double)mp * (double)h->np_ >= 4e9;
    if (emulate) {
        VGP_TRY(emulated_gemm_splitk(h->ws[1].emu, 0, 1, mp, mp, h->np_, 1.0, h->kzx, h->np_, h->kzx, h->np_, 0.0, M(G_), mp,
                                     emu, 1, a * a, a * a, s));
        VGP_TRY(dense_mirror_lower(M(G_), mp, mp, s));
    } else {
        VGP_TRY(dense_gemm_splitk(0, 1, mp, mp, h->np_, 1.0, h->kzx, h->np_, h->kzx, h->np_, 0.0, M(G_), mp, h->splits_n,
                                  h->partial, s, GEMM_LOWER));
    }
    VGP_TRY(dense_gemm_splitk(0, 1, mp, mp, h->bp, 1.0, h->kzb, h->bp, h->kzb, h->bp, 0.0, M(GB_), mp, h->splits_b,
                              h->partial, s, GEMM_LOWER));
    // the six m^3 products that need only K, Q and Gb follow (K + eps I)^-1 on its side stream, underneath the
    // inversion of M on the main stream
    VGP_CUDA(cudaEventRecord(h->ev_gb, s));
    if (sb != s) {
        VGP_CUDA(cudaStreamWaitEvent(sb, h->ev_gb, 0));
    } else {
        VGP_CUDA(cudaEventRecord(h->ev_join[0], h->side[0]));      // Q alone; the products then follow on `s`
        VGP_CUDA(cudaStreamWaitEvent(s, h->ev_join[0], 0));
    }
    auto six = [&](cudaStream_t st) -> int {
        VGP_TRY(mm_on(st, M(Q_), M(GB_), M(T1_)));
        VGP_TRY(mm_on(st, M(T1_), M(Q_), M(R_)));            // R = Q Gb Q
        VGP_TRY(mm_on(st, M(K_), M(R_), M(KR_)));
        VGP_TRY(mm_on(st, M(KR_), M(K_), M(KRK_)));
        VGP_TRY(mm_on(st, M(K_), M(Q_), M(KQ_)));
        return mm_on(st, M(KQ_), M(K_), M(KQK_));
    };
    VGP_TRY(six(sb));
    if (sb != s) VGP_CUDA(cudaEventRecord(h->ev_join[0], h->side[0]));
    VGP_TRY(matvec(h->kzx, h->np_, m, n, h->y, 1.0, V(V_), s));
    VGP_TRY(exchange(M(G_), mp * mp));                  // sums over all observations, one 8 m^2 byte all-reduce
    VGP_TRY(exchange(V(V_), mp));
    VGP_TRY(matvec(h->kzb, h->bp, m, b, yb, 1.0, V(VB_), s));
    VGP_TRY(red.run(DotVec{yb, yb}, b, S_YY));

    // ---- inverses -----------------------------------------------------------------------------------
    {   // M = K + beta G + eps I
        const int64_t count = mp * mp;
        axpby_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(M(K_), 1.0, M(G_), beta, 0.0, m, M(TMP_), mp, count);
        VGP_LAUNCH_CHECK();
    }
    VGP_TRY(invert(h, M(TMP_), eps, M(SG_), h->ws[1], red, S_LDM, s));
    VGP_CUDA(cudaStreamWaitEvent(s, h->ev_join[0], 0));        // join
    VGP_CUDA(cudaStreamWaitEvent(s, h->ev_join[1], 0));

    // ---- forward vectors ----------------------------------------------------------------------------
    VGP_TRY(matvec(M(SG_), mp, m, m, V(V_), 1.0, V(U_), s));           // u = Sigma v
    VGP_TRY(matvec(M(K_), mp, m, m, V(U_), 1.0, V(KU_), s));           // K u
    VGP_TRY(matvec(M(K_), mp, m, m, V(U_), beta, V(MU_), s));          // mu = beta K u
    VGP_TRY(matvec(M(Q_), mp, m, m, V(MU_), 1.0, V(AL_), s));          // alpha = Q mu
    VGP_TRY(matvec(M(GB_), mp, m, m, V(AL_), 1.0, V(GBAL_), s));       // Gb alpha
    VGP_TRY(red.run(DotVec{V(AL_), V(VB_)}, m, S_ALVB));
    VGP_TRY(red.run(DotVec{V(AL_), V(GBAL_)}, m, S_ALGBAL));
    VGP_TRY(red.run(DotVec{V(MU_), V(AL_)}, m, S_MUAL));

    // ---- forward matrices ---------------------------------------------------------------------------
    VGP_TRY(mm(M(SG_), M(K_), M(T2_)));           // T2 = Sigma K
    VGP_TRY(mm(M(K_), M(T2_), M(S_)));            // S = K Sigma K
    VGP_TRY(red.run(DotRegion{M(Q_), M(GB_), mp, m}, m * m, S_TRQGB));
    VGP_TRY(red.run(DotRegion{M(S_), M(R_), mp, m}, m * m, S_TRSR));
    VGP_TRY(red.run(DotRegion{M(Q_), M(S_), mp, m}, m * m, S_TRQS));

    // ---- reverse sweep in m-space -------------------------------------------------------------------
    VGP_TRY(mm(M(T2_), M(R_), M(T2R_)));
    VGP_TRY(mm(M(R_), M(S_), M(RS_)));
    VGP_TRY(mm(M(RS_), M(Q_), M(RSQ_)));
    VGP_TRY(mm(M(Q_), M(S_), M(QS_)));
    VGP_TRY(mm(M(QS_), M(Q_), M(QSQ_)));
    VGP_TRY(mm(M(T2_), M(Q_), M(T2Q_)));
    {   // alpha_bar = beta (Gb alpha - vb);  qa = Q alpha_bar;  mu_bar = w alpha + qa;  u_bar = beta K mu_bar
        const unsigned gb = (unsigned)((mp + 255) / 256);
        axpby_vec_kernel<<<gb, 256, 0, s>>>(V(GBAL_), beta, V(VB_), -beta, V(ALBAR_), m);
        VGP_LAUNCH_CHECK();
        VGP_TRY(matvec(M(Q_), mp, m, m, V(ALBAR_), 1.0, V(QA_), s));
        axpby_vec_kernel<<<gb, 256, 0, s>>>(V(AL_), w, V(QA_), 1.0, V(MUBAR_), m);
        VGP_LAUNCH_CHECK();
        VGP_TRY(matvec(M(K_), mp, m, m, V(MUBAR_), beta, V(UBAR_), s));
        VGP_TRY(red.run(DotVec{V(MUBAR_), V(KU_)}, m, S_MUBKU));
        axpby_vec_kernel<<<gb, 256, 0, s>>>(V(AL_), -beta, nullptr, 0.0, V(VBBAR_), m);     // vb_bar = -beta alpha
        VGP_LAUNCH_CHECK();
    }
    // Sigma_bar = beta/2 K R K + w/2 K Q K + 1/2 (u_bar v^T + v u_bar^T)
    VGP_TRY(lincomb(Terms().add(M(KRK_), 0.5 * beta).add(M(KQK_), 0.5 * w).rank1(V(UBAR_), V(V_), 0.5)
                        .rank1(V(V_), V(UBAR_), 0.5), M(SGBAR_), mp, s));
    // M_bar = w/2 Sigma - Sigma Sigma_bar Sigma
    VGP_TRY(mm(M(SG_), M(SGBAR_), M(TMP_)));
    VGP_TRY(mm(M(TMP_), M(SG_), M(MBAR_)));
    VGP_TRY(lincomb(Terms().add(M(SG_), 0.5 * w).add(M(MBAR_), -1.0), M(TMP_), mp, s));      // TMP = M_bar
    VGP_TRY(red.run(DotRegion{M(TMP_), M(G_), mp, m}, m * m, S_TRMBG));
    VGP_TRY(matvec(M(SG_), mp, m, m, V(UBAR_), 1.0, V(VBAR_), s));                             // v_bar = Sigma u_bar
    // K_bar
    VGP_TRY(lincomb(Terms().add(M(R_), 0.5 * beta).sym(M(T2R_), 0.5 * beta).sym(M(RSQ_), -0.5 * beta)
                        .add(M(QSQ_), -0.5 * w).sym(M(T2Q_), 0.5 * w).add(M(Q_), 0.5 * w).add(M(KINV_), -w)
                        .add(M(TMP_), 1.0)
                        .rank1(V(AL_), V(AL_), -0.5 * w).rank1(V(QA_), V(AL_), -0.5).rank1(V(AL_), V(QA_), -0.5)
                        .rank1(V(MUBAR_), V(U_), 0.5 * beta),
                    M(KBAR_), mp, s));
    // second symmetric half of mu_bar u^T needs a 5th rank-1 term: fold it in with a follow-up pass
    VGP_TRY(lincomb(Terms().add(M(KBAR_), 1.0).rank1(V(U_), V(MUBAR_), 0.5 * beta), M(KBAR_), mp, s));
    // 2 G_bar = 2 beta M_bar ;  2 Gb_bar = beta alpha alpha^T - beta Q + beta Q S Q
    VGP_TRY(lincomb(Terms().add(M(TMP_), 2.0 * beta), M(GBAR2_), mp, s));
    VGP_TRY(lincomb(Terms().add(M(Q_), -beta).add(M(QSQ_), beta).rank1(V(AL_), V(AL_), beta), M(GBBAR2_), mp, s));

    // ---- push through the kernel ---------------------------------------------------------------------
    // the minibatch and K_zz parts run on a side stream (own accumulator) underneath the N-sized product
    VGP_CUDA(cudaEventRecord(h->ev_bar, s));
    VGP_CUDA(cudaStreamWaitEvent(sd, h->ev_bar, 0));
    VGP_TRY(dense_gemm(0, 0, mp, h->bp, mp, 1.0, M(GBBAR2_), mp, h->kzb, h->bp, 0.0, h->wzb, h->bp, GEMM_FULL, sd));
    VGP_TRY(kernback(h, h->wzb, h->kzb, h->bp, xb, b, V(VBBAR_), yb, 1.0, h->rowacc2, sd));
    VGP_TRY(kernback(h, M(KBAR_), M(K_), mp, h->z, m, nullptr, nullptr, 2.0, h->rowacc2, sd));
    VGP_CUDA(cudaEventRecord(h->ev_mini, sd));
    if (emulate) {
        VGP_TRY(emulated_gemm(h->ws[1].emu, 0, 0, mp, h->np_, mp, 1.0, M(GBAR2_), mp, h->kzx, h->np_, 0.0, h->wzx, h->np_, emu, 0,
                              s, nullptr, 0.0, a * a));
    } else {
        VGP_TRY(dense_gemm(0, 0, mp, h->np_, mp, 1.0, M(GBAR2_), mp, h->kzx, h->np_, 0.0, h->wzx, h->np_, GEMM_FULL, s));
    }
    VGP_TRY(kernback(h, h->wzx, h->kzx, h->np_, h->x, n, V(VBAR_), h->y, 1.0, h->rowacc, s));
    VGP_TRY(exchange(h->rowacc, m * (2 + h->d)));       // the observation-sized part of the push-through sums
    VGP_CUDA(cudaStreamWaitEvent(s, h->ev_mini, 0));
    {
        const int64_t cnt = m * (2 + h->d);
        axpby_vec_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(h->rowacc, 1.0, h->rowacc2, 1.0, h->rowacc, cnt);
        VGP_LAUNCH_CHECK();
    }
    VGP_TRY(red.run(ColSum{h->rowacc, 2 + h->d, 0}, m, S_ACC0));
    VGP_TRY(red.run(ColSum{h->rowacc, 2 + h->d, 1}, m, S_ACC1));

    double sc[S_N];
    VGP_TRY(red.fetch(sc, S_N));
    for (int i = 0; i < 3; ++i) {
        int info = 0;
        VGP_CUDA(cudaMemcpy(&info, h->ws[i].info, sizeof(int), cudaMemcpyDeviceToHost));
        if (info != 0) {
            set_error("ELBO step: %s is not positive definite (pivot %d)",
                      i == 0 ? "K_zz + jitter" : (i == 1 ? "K_zz + K_zx K_xz / noise" : "K_zz"), info - 1);
            return VGP_ERR_NOT_PD;
        }
    }
    const double log2pi = log(2.0 * M_PI);
    const double quad = sc[S_YY] - 2.0 * sc[S_ALVB] + sc[S_ALGBAL];
    const double ll = -0.5 * beta * quad - 0.5 * (double)b * (log2pi + log(noise));
    const double tr1 = (double)b * a * a - sc[S_TRQGB];
    const double tr2 = sc[S_TRSR];
    const double kl = 0.5 * (sc[S_TRQS] + sc[S_MUAL] - (double)m + 2.0 * sc[S_LDKT] - 4.0 * sc[S_LDK] + 2.0 * sc[S_LDM]);
    const double loss = -ll + 0.5 * beta * (tr1 + tr2) + w * kl;
    h->last_terms[0] = loss;
    h->last_terms[1] = ll;
    h->last_terms[2] = tr1;
    h->last_terms[3] = tr2;
    h->last_terms[4] = kl;
    const double beta_bar = 0.5 * quad - 0.5 * (double)b / beta + 0.5 * tr1 + 0.5 * tr2 + sc[S_MUBKU] + sc[S_TRMBG];
    const double s_bar = -beta_bar / (noise * noise);
    const double ga = 2.0 * sc[S_ACC0] / a + 2.0 * a * (0.5 * beta * (double)b);
    const double gl = sc[S_ACC1] / (l * l * l);
    grads[0] = ga * sigmoid_h(h->v[0]);
    grads[1] = gl * sigmoid_h(h->v[1]);
    grads[2] = s_bar * sigmoid_h(h->v[2]);
    *z_gscale = 1.0 / (l * l);
    *loss_out = loss;
    h->launches += g_launches - before;
    return VGP_OK;
}

}  // namespace elbo_detail
}  // namespace vgp

extern "C" {

int vgp_elbo_create(vgp_elbo **handle, int device, const double *x_dev, const double *y_dev, int64_t n_obs, int d,
                    const double *z_init_host, int64_t m, int64_t batch, double v_amplitude, double v_length_scale,
                    double v_noise, double length_scale_offset, double jitter, double learning_rate) {
    VGP_REQUIRE(handle && x_dev && y_dev && z_init_host, "NULL argument");
    *handle = nullptr;
    VGP_REQUIRE(n_obs > 0 && m > 0 && batch > 0 && d >= 1 && d <= 8, "bad sizes");
    VGP_ENTER(device);
    vgp_elbo *h = new (std::nothrow) vgp_elbo();
    VGP_REQUIRE(h, "out of host memory");
    h->device = device;
    h->n = n_obs;
    h->m = m;
    h->b = batch;
    h->d = d;
    h->mp = round_up(m, TILE);
    h->np_ = round_up(n_obs, TILE);
    h->bp = round_up(batch, TILE);
    h->x = x_dev;
    h->y = y_dev;
    h->v[0] = v_amplitude;
    h->v[1] = v_length_scale;
    h->v[2] = v_noise;
    h->ls_offset = length_scale_offset;
    h->jitter = jitter;
    h->lr = learning_rate;
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
    // the two Gram matrices are symmetric: only the lower tiles are computed; split K so that the CTAs fill
    // (just under) two waves of the SMs
    const int tiles = (int)((h->mp / TILE) * (h->mp / TILE + 1) / 2);
    auto pick = [&](int64_t k) {
        int s = (2 * sm) / tiles;
        const int64_t maxs = k / 256 > 0 ? k / 256 : 1;
        if (s > maxs) s = (int)maxs;
        return s < 1 ? 1 : s;
    };
    h->splits_n = pick(h->np_);
    h->splits_b = pick(h->bp);
    const int smax = h->splits_n > h->splits_b ? h->splits_n : h->splits_b;
    const size_t mm = (size_t)h->mp * h->mp * 8;
    struct {
        void **p;
        size_t bytes;
    } allocs[] = {
        {(void **)&h->z, (size_t)m * d * 8},       {(void **)&h->mz, (size_t)m * d * 8},
        {(void **)&h->vz, (size_t)m * d * 8},      {(void **)&h->mats, (size_t)NMAT_ * mm},
        {(void **)&h->vecs, (size_t)NVEC_ * h->mp * 8},
        {(void **)&h->kzx, (size_t)h->mp * h->np_ * 8}, {(void **)&h->wzx, (size_t)h->mp * h->np_ * 8},
        {(void **)&h->kzb, (size_t)h->mp * h->bp * 8},  {(void **)&h->wzb, (size_t)h->mp * h->bp * 8},
        {(void **)&h->partial, (size_t)smax * mm}, {(void **)&h->rowacc, (size_t)m * (2 + d) * 8},
        {(void **)&h->gradz, (size_t)m * d * 8},         {(void **)&h->rowacc2, (size_t)m * (2 + d) * 8},
    };
    for (auto &al : allocs) {
        cudaError_t e = device_malloc(al.p, al.bytes);
        if (e == cudaSuccess) e = cudaMemset(*al.p, 0, al.bytes);
        if (e != cudaSuccess) {
            int rc = cuda_fail(e, "cudaMalloc (ELBO state)", __FILE__, __LINE__);
            vgp_elbo_destroy(h);
            return rc;
        }
    }
    cudaError_t e = cudaMemcpy(h->z, z_init_host, (size_t)m * d * 8, cudaMemcpyHostToDevice);
    for (auto &st : h->side)
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    for (auto &ev : h->ev_join)
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (cudaEvent_t *ev : {&h->ev_gb, &h->ev_bar, &h->ev_mini})
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        int rc = cuda_fail(e, "Z upload / side streams", __FILE__, __LINE__);
        vgp_elbo_destroy(h);
        return rc;
    }
    *handle = h;
    return VGP_OK;
}

/* Kernel family of the training step (default EXPQUAD): MATERN32 and MATERN52 as well -- the reference's VGP trains
 * tfkern.MaternFiveHalves over 5-D inputs (main_architecture_2.py:184-249).  MATERN12 is refused: its gradient with
 * respect to the inducing points is singular on the diagonal of K_zz (r = 0). */
int vgp_elbo_set_kernel(vgp_elbo *h, int kind) {
    VGP_REQUIRE(h, "handle is NULL");
    VGP_REQUIRE(kind == VGP_KERNEL_EXPQUAD || kind == VGP_KERNEL_MATERN32 || kind == VGP_KERNEL_MATERN52,
                "ELBO training supports EXPQUAD, MATERN32 and MATERN52 (kind %d: the gradient of Matern-1/2 with respect "
                "to the inducing points is singular at r = 0)", kind);
    h->kind = kind;
    return VGP_OK;
}

int vgp_elbo_destroy(vgp_elbo *h) {
    if (!h) return VGP_OK;
    VGP_ENTER(h->device);
    void *ptrs[] = {h->z, h->mz, h->vz, h->mats, h->vecs, h->kzx, h->wzx, h->kzb, h->wzb, h->partial, h->rowacc, h->gradz,
                    h->rowacc2};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (auto &w : h->ws) w.release();
    for (auto &st : h->side)
        if (st) cudaStreamDestroy(st);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    for (auto &e : h->ev_join)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : {h->ev_gb, h->ev_bar, h->ev_mini})
        if (e) cudaEventDestroy(e);
    delete h;
    return VGP_OK;
}

int vgp_elbo_loss_grad(vgp_elbo *h, const double *xb_dev, const double *yb_dev, double *loss_host,
                       double *grads_host, double *gradz_host, vgp_vgp_terms *terms_host, void *stream) {
    VGP_REQUIRE(h && xb_dev && yb_dev && loss_host, "NULL argument");
    VGP_ENTER(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    double g[3], zs = 0.0;
    VGP_TRY(loss_and_grad(h, xb_dev, yb_dev, loss_host, g, &zs, s));
    if (grads_host)
        for (int i = 0; i < 3; ++i) grads_host[i] = g[i];
    if (gradz_host) {
        const int64_t cnt = h->m * h->d;
        gradz_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(h->rowacc, h->m, h->d, zs, h->gradz);
        VGP_LAUNCH_CHECK();
        VGP_CUDA(cudaMemcpyAsync(gradz_host, h->gradz, (size_t)cnt * 8, cudaMemcpyDeviceToHost, s));
        VGP_CUDA(cudaStreamSynchronize(s));
    }
    if (terms_host) {
        terms_host->loss = h->last_terms[0];
        terms_host->ll = h->last_terms[1];
        terms_host->tr1 = h->last_terms[2];
        terms_host->tr2 = h->last_terms[3];
        terms_host->kl = h->last_terms[4];
    }
    return VGP_OK;
}

int vgp_elbo_step(vgp_elbo *h, const double *xb_dev, const double *yb_dev, double *loss_host, void *stream) {
    VGP_REQUIRE(h && xb_dev && yb_dev && loss_host, "NULL argument");
    VGP_ENTER(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    double g[3], zs = 0.0;
    VGP_TRY(loss_and_grad(h, xb_dev, yb_dev, loss_host, g, &zs, s));
    ++h->t;
    const double lr_t = h->lr * sqrt(1.0 - pow(h->b2, (double)h->t)) / (1.0 - pow(h->b1, (double)h->t));
    for (int i = 0; i < 3; ++i) {
        h->am[i] = h->b1 * h->am[i] + (1.0 - h->b1) * g[i];
        h->av[i] = h->b2 * h->av[i] + (1.0 - h->b2) * g[i] * g[i];
        h->v[i] -= lr_t * h->am[i] / (sqrt(h->av[i]) + h->eps);
    }
    const int64_t cnt = h->m * h->d;
    adam_z_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(h->z, h->mz, h->vz, h->rowacc, h->m, h->d, zs, lr_t,
                                                               h->b1, h->b2, h->eps);
    VGP_LAUNCH_CHECK();
    ++h->launches;
    return VGP_OK;
}

int vgp_elbo_set_exchange(vgp_elbo *h, int64_t n_total, vgp_allreduce_fn fn, void *ctx) {
    VGP_REQUIRE(h, "NULL handle");
    VGP_REQUIRE(n_total >= h->n, "n_total %lld is smaller than this handle's %lld observations", (long long)n_total,
                (long long)h->n);
    h->n_total = n_total;
    h->allreduce = fn;
    h->allreduce_ctx = ctx;
    return VGP_OK;
}

int vgp_elbo_get_params(vgp_elbo *h, double *v3_host, double *z_host, void *stream) {
    VGP_REQUIRE(h, "NULL handle");
    VGP_ENTER(h->device);
    if (v3_host)
        for (int i = 0; i < 3; ++i) v3_host[i] = h->v[i];
    if (z_host) {
        VGP_CUDA(cudaMemcpyAsync(z_host, h->z, (size_t)h->m * h->d * 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        VGP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    }
    return VGP_OK;
}

int vgp_elbo_set_params(vgp_elbo *h, const double *v3_host, const double *z_host, void *stream) {
    VGP_REQUIRE(h, "NULL handle");
    VGP_ENTER(h->device);
    if (v3_host)
        for (int i = 0; i < 3; ++i) h->v[i] = v3_host[i];
    if (z_host) {
        VGP_CUDA(cudaMemcpyAsync(h->z, z_host, (size_t)h->m * h->d * 8, cudaMemcpyHostToDevice, (cudaStream_t)stream));
        VGP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    }
    return VGP_OK;
}

int vgp_elbo_launch_count(vgp_elbo *h, int64_t *launches) {
    VGP_REQUIRE(h && launches, "NULL argument");
    *launches = h->launches;
    return VGP_OK;
}

}  // extern "C"
