// (1) ExponentiatedQuadratic kernel-matrix builder.
//
// out[i][j] = a^2 exp(-|x1_i - x2_j|^2 / (2 l^2)) (+ diag_add on the shifted diagonal).
// Stands in for tfkern.ExponentiatedQuadratic(...).matrix(x1, x2) (reference call sites:
// variational_Gaussian_process_example.py:55-57, 3D_sin_wave.py:158-159, main_tests.py:617-619).
//
// The op is HBM-write-bound (8 n1 n2 bytes out, 8 d (n1 + n2) bytes in): no tensor cores, each thread
// produces two adjacent float64 and issues one 128-bit store; a warp writes 512 contiguous bytes per row.
// The squared distance is the direct sum of squared coordinate differences (not |x|^2+|y|^2-2xy), which
// keeps full float64 accuracy for near-by points.  The float64 exp() costs ~40 issue slots per element,
// enough to make the full build FP64-issue-bound rather than write-bound; the symmetric path therefore
// evaluates each off-diagonal 64x64 tile once and writes it twice (once transposed through shared memory).
#include <math.h>

#include "common.cuh"

namespace vgp {

struct ExpQuadArgs {
    const double *x1;
    const double *x2;
    int64_t n1, n2;
    double amp2;
    double neg_half_inv_l2;
    double diag_add;
    int64_t diag_col0;
    double *out;
    int64_t ld;
};

template <int D>
__device__ __forceinline__ double eq_value(const double (&a)[D], const double (&b)[D], double amp2, double c) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const double t = a[k] - b[k];
        s = fma(t, t, s);
    }
    return amp2 * exp(s * c);
}

// ---- general rectangular path: CTA tile 32 rows x 512 columns, 256 threads, 2 columns per thread ------
constexpr int EQ_ROWS = 32;
constexpr int EQ_COLS = 512;

template <int D, bool VEC>
__global__ void __launch_bounds__(256) expquad_rect_kernel(ExpQuadArgs p) {
    __shared__ double xs[EQ_ROWS][D];
    const int64_t row0 = (int64_t)blockIdx.y * EQ_ROWS;
    const int64_t col = (int64_t)blockIdx.x * EQ_COLS + 2 * threadIdx.x;
    for (int t = threadIdx.x; t < EQ_ROWS * D; t += 256) {
        const int64_t r = row0 + t / D;
        xs[t / D][t % D] = r < p.n1 ? p.x1[r * D + t % D] : 0.0;
    }
    double b0[D], b1[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        b0[k] = col < p.n2 ? p.x2[col * D + k] : 0.0;
        b1[k] = col + 1 < p.n2 ? p.x2[(col + 1) * D + k] : 0.0;
    }
    __syncthreads();
    if (col >= p.n2) return;
    const int rows = (int)min((int64_t)EQ_ROWS, p.n1 - row0);
#pragma unroll 4
    for (int r = 0; r < rows; ++r) {
        double a[D];
#pragma unroll
        for (int k = 0; k < D; ++k) a[k] = xs[r][k];
        double v0 = eq_value<D>(a, b0, p.amp2, p.neg_half_inv_l2);
        double v1 = eq_value<D>(a, b1, p.amp2, p.neg_half_inv_l2);
        const int64_t i = row0 + r;
        if (i == p.diag_col0 + col) v0 += p.diag_add;
        if (i == p.diag_col0 + col + 1) v1 += p.diag_add;
        double *dst = p.out + i * p.ld + col;
        if (VEC && col + 1 < p.n2) {
            *reinterpret_cast<double2 *>(dst) = make_double2(v0, v1);
        } else {
            dst[0] = v0;
            if (col + 1 < p.n2) dst[1] = v1;
        }
    }
}

// ---- symmetric path: one CTA per 64x64 tile pair (I <= J) -------------------------------------------
constexpr int SQ = 64;

template <int D, bool VEC>
__global__ void __launch_bounds__(256) expquad_sym_kernel(ExpQuadArgs p, int ntiles) {
    __shared__ double xi[SQ][D];
    __shared__ double xj[SQ][D];
    __shared__ double tile[SQ][SQ + 1];
    // linear block id -> (I, J) with I <= J, row-major over the upper triangle
    int64_t b = blockIdx.x;
    int I = (int)floor((2.0 * ntiles + 1.0 - sqrt((2.0 * ntiles + 1.0) * (2.0 * ntiles + 1.0) - 8.0 * (double)b)) * 0.5);
    // guard against floating-point rounding of the closed form
    while ((int64_t)I * ntiles - (int64_t)I * (I - 1) / 2 > b) --I;
    while ((int64_t)(I + 1) * ntiles - (int64_t)(I + 1) * I / 2 <= b) ++I;
    const int J = I + (int)(b - ((int64_t)I * ntiles - (int64_t)I * (I - 1) / 2));
    const int64_t r0 = (int64_t)I * SQ, c0 = (int64_t)J * SQ;
    const int64_t n = p.n1;
    for (int t = threadIdx.x; t < SQ * D; t += 256) {
        const int64_t r = r0 + t / D, c = c0 + t / D;
        xi[t / D][t % D] = r < n ? p.x1[r * D + t % D] : 0.0;
        xj[t / D][t % D] = c < n ? p.x1[c * D + t % D] : 0.0;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int jc = 2 * tx;
    double b0[D], b1[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        b0[k] = xj[jc][k];
        b1[k] = xj[jc + 1][k];
    }
#pragma unroll
    for (int rr = 0; rr < SQ / 8; ++rr) {
        const int r = ty + 8 * rr;
        double a[D];
#pragma unroll
        for (int k = 0; k < D; ++k) a[k] = xi[r][k];
        double v0 = eq_value<D>(a, b0, p.amp2, p.neg_half_inv_l2);
        double v1 = eq_value<D>(a, b1, p.amp2, p.neg_half_inv_l2);
        const int64_t i = r0 + r, j = c0 + jc;
        if (i == j) v0 += p.diag_add;
        if (i == j + 1) v1 += p.diag_add;
        tile[r][jc] = v0;
        tile[r][jc + 1] = v1;
        if (i < n && j < n) {
            double *dst = p.out + i * p.ld + j;
            if (VEC && j + 1 < n) {
                *reinterpret_cast<double2 *>(dst) = make_double2(v0, v1);
            } else {
                dst[0] = v0;
                if (j + 1 < n) dst[1] = v1;
            }
        }
    }
    if (I == J) return;
    __syncthreads();
    // mirrored tile: out[c0 + j][r0 + i] = tile[i][j].  A warp writes 2 x 256 contiguous bytes of one
    // output row with 64-bit stores; lane <-> i keeps the column reads of `tile` bank-conflict free
    // (row stride 65 doubles), which a 128-bit (i, i+1) pairing would not.
#pragma unroll
    for (int rr = 0; rr < SQ / 8; ++rr) {
        const int j = ty + 8 * rr;
        const int64_t row = c0 + j;
        if (row >= n) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = tx + 32 * h;
            if (r0 + i < n) p.out[row * p.ld + r0 + i] = tile[i][j];
        }
    }
}

template <int D>
static int launch_expquad(const ExpQuadArgs &p, bool symmetric, cudaStream_t s) {
    const bool vec = (p.ld % 2 == 0) && (((uintptr_t)p.out) % 16 == 0);
    if (symmetric) {
        const int nt = (int)((p.n1 + SQ - 1) / SQ);
        const int64_t blocks = (int64_t)nt * (nt + 1) / 2;
        if (blocks > 0x7fffffffLL) {
            set_error("expquad: matrix too large for one launch");
            return VGP_ERR_INVALID;
        }
        if (vec)
            expquad_sym_kernel<D, true><<<(unsigned)blocks, 256, 0, s>>>(p, nt);
        else
            expquad_sym_kernel<D, false><<<(unsigned)blocks, 256, 0, s>>>(p, nt);
    } else {
        dim3 grid((unsigned)((p.n2 + EQ_COLS - 1) / EQ_COLS), (unsigned)((p.n1 + EQ_ROWS - 1) / EQ_ROWS));
        if (grid.y > 65535u) {
            // fold very tall matrices into several launches over row slabs
            const int64_t slab = 65535LL * EQ_ROWS;
            for (int64_t r = 0; r < p.n1; r += slab) {
                ExpQuadArgs q = p;
                q.x1 = p.x1 + r * D;
                q.n1 = min(slab, p.n1 - r);
                q.out = p.out + r * p.ld;
                q.diag_col0 = p.diag_col0 - r;
                VGP_TRY(launch_expquad<D>(q, false, s));
            }
            return VGP_OK;
        }
        if (vec)
            expquad_rect_kernel<D, true><<<grid, 256, 0, s>>>(p);
        else
            expquad_rect_kernel<D, false><<<grid, 256, 0, s>>>(p);
    }
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

int expquad_dispatch(const ExpQuadArgs &p, int d, bool symmetric, cudaStream_t s) {
    switch (d) {
        case 1: return launch_expquad<1>(p, symmetric, s);
        case 2: return launch_expquad<2>(p, symmetric, s);
        case 3: return launch_expquad<3>(p, symmetric, s);
        case 4: return launch_expquad<4>(p, symmetric, s);
        case 5: return launch_expquad<5>(p, symmetric, s);
        case 6: return launch_expquad<6>(p, symmetric, s);
        case 7: return launch_expquad<7>(p, symmetric, s);
        case 8: return launch_expquad<8>(p, symmetric, s);
    }
    set_error("expquad: feature dimension %d outside [1, 8]", d);
    return VGP_ERR_INVALID;
}

int expquad_dispatch_public(const double *x1, int64_t n1, const double *x2, int64_t n2, int d, double amplitude,
                            double length_scale, double diag_add, int64_t diag_col0, double *out, int64_t ld,
                            cudaStream_t s) {
    if (n1 == 0 || n2 == 0) return VGP_OK;
    ExpQuadArgs p;
    p.x1 = x1;
    p.x2 = x2;
    p.n1 = n1;
    p.n2 = n2;
    p.amp2 = amplitude * amplitude;
    p.neg_half_inv_l2 = -0.5 / (length_scale * length_scale);
    p.diag_add = diag_add;
    p.diag_col0 = diag_col0;
    p.out = out;
    p.ld = ld;
    const bool symmetric = (x1 == x2) && (n1 == n2) && diag_col0 == 0 && n1 >= 2 * SQ;
    return expquad_dispatch(p, d, symmetric, s);
}

}  // namespace vgp

extern "C" int vgp_expquad_matrix(int device, const double *x1_dev, int64_t n1, const double *x2_dev, int64_t n2,
                                  int d, double amplitude, double length_scale, double diag_add,
                                  int64_t diag_col0, double *out_dev, int64_t ld_out, void *stream) {
    VGP_REQUIRE(n1 >= 0 && n2 >= 0, "negative size");
    VGP_REQUIRE(ld_out >= n2, "ld_out %lld < n2 %lld", (long long)ld_out, (long long)n2);
    VGP_REQUIRE(length_scale > 0.0, "length_scale must be positive");
    VGP_REQUIRE(d >= 1 && d <= 8, "feature dimension %d outside [1, 8]", d);
    if (n1 == 0 || n2 == 0) return VGP_OK;
    VGP_REQUIRE(x1_dev && x2_dev && out_dev, "NULL pointer");
    VGP_ENTER(device);
    return vgp::expquad_dispatch_public(x1_dev, n1, x2_dev, n2, d, amplitude, length_scale, diag_add, diag_col0,
                                        out_dev, ld_out, (cudaStream_t)stream);
}
