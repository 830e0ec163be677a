// (1) Stationary kernel-matrix builder: ExponentiatedQuadratic and the Matern family.
//
//   ExpQuad      out[i][j] = a^2 exp(-|x1_i - x2_j|^2 / (2 l^2))          tfkern.ExponentiatedQuadratic(...).matrix
//   Matern 1/2   a^2 exp(-r / l)                                           gp_functions.py:160-163 (create_cov_kernel)
//   Matern 3/2   a^2 (1 + z) exp(-z),            z = sqrt(3) r / l
//   Matern 5/2   a^2 (1 + z + z^2 / 3) exp(-z),  z = sqrt(5) r / l         main_architecture_2.py:184
// (+ diag_add on the shifted diagonal).  Reference call sites of the ExpQuad form:
// variational_Gaussian_process_example.py:55-57, 3D_sin_wave.py:158-159, main_tests.py:617-619.
//
// The op is HBM-write-bound (8 n1 n2 bytes out, 8 d (n1 + n2) bytes in): no tensor cores, each thread produces two
// adjacent float64 and issues one 128-bit store; a warp writes 512 contiguous bytes per row.  The squared distance
// is the direct sum of squared coordinate differences (not |x|^2 + |y|^2 - 2 x.y), which keeps full float64 accuracy
// for near-by points.
//
// FP64 issue budget.  B200 retires 64 FP64 lanes / clk / SM, i.e. ~22 FP64 instructions per 8-byte element at the
// measured HBM write rate.  libm's exp() plus the distance is 26, which made the first version FP64-issue-bound.
// `exp_nonpos` below is 12: arguments are never positive here, so the reduction is t = k ln2/32 + r with a
// Cody-Waite pair (two FMAs, exact), a degree-6 polynomial on |r| <= ln2/64, one multiply by the table entry
// a^2 2^(j/32) (32 entries in shared memory, amplitude folded in) and an integer add on the exponent field.
// Measured against long-double exp over [-650, 0]: <= 1.94 ulp; below -650 the result is 0.  The symmetric path additionally evaluates each off-diagonal 64x64 tile once and writes it twice
// (once transposed through shared memory); its CTAs walk 16x16 super-tiles so that both the direct and the mirrored
// stores of concurrently running CTAs fall into the same 1024 x 1024 block of the output (8 KB row pieces in L2
// instead of scattered 512-byte pieces).
#include <math.h>

#include "common.cuh"

namespace vgp {

enum { KIND_EXPQUAD = VGP_KERNEL_EXPQUAD, KIND_MATERN12 = VGP_KERNEL_MATERN12, KIND_MATERN32 = VGP_KERNEL_MATERN32,
       KIND_MATERN52 = VGP_KERNEL_MATERN52 };

struct ExpQuadArgs {
    const double *x1;
    const double *x2;
    int64_t n1, n2;
    double amp2;
    double c;                // ExpQuad: -1 / (2 l^2) (multiplies r^2);  Matern nu: -sqrt(2 nu) / l (multiplies r)
    double diag_add;
    int64_t diag_col0;
    double *out;
    int64_t ld;
};

constexpr int EXP_TAB = 32;
constexpr double EXP_INV_L = 46.16624130844683;            // 32 / ln 2
// Constants of exp_nonpos live in constant memory so that every FP64 instruction takes them as a c[bank][offset]
// operand; as literals the compiler rebuilt each 64-bit immediate with two UMOVs per use, one issue slot per DFMA
// in a kernel that ncu showed to be issue-bound (82 % issue-active).
__constant__ double EXP_K[10] = {
    6755399441055744.0,          // [0] MAGIC = 1.5 * 2^52: adding it rounds to the nearest integer
    -0.02166084939249829,        // [1] -ln 2 / 32, float64 head
    -7.247021293269686e-19,      // [2]             and tail
    1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0,      // [3..8] Taylor coefficients, degree 6 .. 1
    1.0 / 3.0};                  // [9] Matern 5/2

// tab[j] = amp2 * 2^(j / 32)
__device__ __forceinline__ void exp_table_fill(double *tab, double amp2) {
    if (threadIdx.x < EXP_TAB) tab[threadIdx.x] = amp2 * exp2((double)threadIdx.x * (1.0 / EXP_TAB));
}

// amp2 * exp(t) for t <= 0 (amp2 within [2^-60, 2^60], checked by the launcher).  `kd` = MAGIC + (an integer within
// 1 of 32 t / ln 2), formed by the caller with one FMA from whatever t was computed from.  Branch-free: arguments
// below -650 (results under 2^-937 a^2, far below anything a covariance carries) and -inf give exactly 0; NaN
// propagates through the polynomial.  <= 1.94 ulp against long-double exp on [-650, 0].
__device__ __forceinline__ double exp_nonpos(double t, double kd, const double *tab) {
    const int k = __double2loint(kd);                       // k <= 0
    kd -= EXP_K[0];
    double r = fma(kd, EXP_K[1], t);
    r = fma(kd, EXP_K[2], r);
    double p = EXP_K[3];
    p = fma(p, r, EXP_K[4]);
    p = fma(p, r, EXP_K[5]);
    p = fma(p, r, EXP_K[6]);
    p = fma(p, r, EXP_K[7]);
    p = fma(p, r, EXP_K[8]);
    p = fma(p, r, EXP_K[8]);
    const double v = tab[k & (EXP_TAB - 1)] * p;
    // v * 2^floor(k / 32) through the exponent field; integer tests on the high word of t (t <= 0: the more
    // negative, the larger the unsigned word) keep the range check off the FP64 pipe
    const unsigned th = (unsigned)__double2hiint(t);
    const bool tiny = th > 0xC0845000u && (th < 0xFFF00000u || (th == 0xFFF00000u && __double2loint(t) == 0));
    const int hi = __double2hiint(v) + ((k & ~(EXP_TAB - 1)) << 15);
    return __hiloint2double(tiny ? 0 : hi, tiny ? 0 : __double2loint(v));
}

template <int KIND, int D, bool FAST>
__device__ __forceinline__ double kernel_value(const double (&a)[D], const double (&b)[D], double amp2, double c,
                                               double cz, const double *tab) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const double t = a[k] - b[k];
        s = fma(t, t, s);
    }
    if (KIND == KIND_EXPQUAD) {
        const double t = s * c;
        return FAST ? exp_nonpos(t, fma(s, cz, EXP_K[0]), tab) : amp2 * exp(t);
    }
    const double t = sqrt(s) * c;                           // -z
    const double e = FAST ? exp_nonpos(t, fma(t, EXP_INV_L, EXP_K[0]), tab) : amp2 * exp(t);
    if (KIND == KIND_MATERN12) return e;
    if (KIND == KIND_MATERN32) return (EXP_K[8] - t) * e;
    return fma(t, fma(t, EXP_K[9], -EXP_K[8]), EXP_K[8]) * e;      // 1 + z + z^2 / 3
}

// ---- general rectangular path: CTA tile 32 rows x 512 columns, 256 threads, 2 columns per thread ------
constexpr int EQ_ROWS = 32;
constexpr int EQ_COLS = 512;

// FULL: the tile lies inside the matrix and 128-bit stores are legal -- no bounds tests in the loop
template <int KIND, int D, bool FAST, bool FULL>
__device__ __forceinline__ void rect_tile(const ExpQuadArgs &p, const double (*xs)[D], const double *tab,
                                          int64_t row0, int64_t col, int rows, bool vec) {
    double b0[D], b1[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        b0[k] = (FULL || col < p.n2) ? p.x2[col * D + k] : 0.0;
        b1[k] = (FULL || col + 1 < p.n2) ? p.x2[(col + 1) * D + k] : 0.0;
    }
    const double cz = p.c * EXP_INV_L;
    double *dst = p.out + row0 * p.ld + col;
#pragma unroll 4
    for (int r = 0; r < (FULL ? EQ_ROWS : rows); ++r, dst += p.ld) {
        double a[D];
#pragma unroll
        for (int k = 0; k < D; ++k) a[k] = xs[r][k];
        const double v0 = kernel_value<KIND, D, FAST>(a, b0, p.amp2, p.c, cz, tab);
        const double v1 = kernel_value<KIND, D, FAST>(a, b1, p.amp2, p.c, cz, tab);
        if (FULL || (vec && col + 1 < p.n2)) {
            *reinterpret_cast<double2 *>(dst) = make_double2(v0, v1);
        } else {
            dst[0] = v0;
            if (col + 1 < p.n2) dst[1] = v1;
        }
    }
}

template <int KIND, int D, bool FAST>
__global__ void __launch_bounds__(256) kernel_rect_kernel(ExpQuadArgs p, int vec) {
    __shared__ double xs[EQ_ROWS][D];
    __shared__ double tab[EXP_TAB];
    const int64_t row0 = (int64_t)blockIdx.y * EQ_ROWS;
    const int64_t col = (int64_t)blockIdx.x * EQ_COLS + 2 * threadIdx.x;
    for (int t = threadIdx.x; t < EQ_ROWS * D; t += 256) {
        const int64_t r = row0 + t / D;
        xs[t / D][t % D] = r < p.n1 ? p.x1[r * D + t % D] : 0.0;
    }
    if (FAST) exp_table_fill(tab, p.amp2);
    __syncthreads();
    if (col >= p.n2) return;
    const int rows = (int)min((int64_t)EQ_ROWS, p.n1 - row0);
    const bool full = vec && rows == EQ_ROWS && ((int64_t)blockIdx.x + 1) * EQ_COLS <= p.n2;      // CTA-uniform
    if (full)
        rect_tile<KIND, D, FAST, true>(p, xs, tab, row0, col, rows, true);
    else
        rect_tile<KIND, D, FAST, false>(p, xs, tab, row0, col, rows, vec != 0);
    // diagonal shift, outside the hot loop: the thread that owns column j also wrote out[diag_col0 + j][j]
    if (p.diag_add != 0.0) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int64_t i = p.diag_col0 + col + e;
            if (col + e < p.n2 && i >= row0 && i < row0 + rows) p.out[i * p.ld + col + e] += p.diag_add;
        }
    }
}

// ---- symmetric path: one CTA per 64x64 tile pair (I <= J), CTAs ordered by 16x16 super-tiles -----------
constexpr int SQ = 64;
constexpr int SQ_LD = SQ + 2;       // 16-byte aligned rows for 128-bit stores; (SQ_LD / 2) odd: a quarter-warp reading one
                                    // column pair of 8 consecutive rows with 128-bit loads covers all 32 banks
constexpr int SUPER = 16;

template <int KIND, int D, bool FAST, bool FULL>
__device__ __forceinline__ void sym_tile(const ExpQuadArgs &p, const double (*xi)[D], const double (*xj)[D],
                                         double (*tile)[SQ_LD], const double *tab, int64_t r0, int64_t c0, bool diag,
                                         bool vec) {
    const int64_t n = p.n1;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int jc = 2 * tx;
    const double cz = p.c * EXP_INV_L;
    double b0[D], b1[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        b0[k] = xj[jc][k];
        b1[k] = xj[jc + 1][k];
    }
    double *dst = p.out + (r0 + ty) * p.ld + c0 + jc;
#pragma unroll
    for (int rr = 0; rr < SQ / 8; ++rr, dst += 8 * p.ld) {
        const int r = ty + 8 * rr;
        double a[D];
#pragma unroll
        for (int k = 0; k < D; ++k) a[k] = xi[r][k];
        double v0 = kernel_value<KIND, D, FAST>(a, b0, p.amp2, p.c, cz, tab);
        double v1 = kernel_value<KIND, D, FAST>(a, b1, p.amp2, p.c, cz, tab);
        if (diag) {                                         // I == J: r0 == c0
            if (r == jc) v0 += p.diag_add;
            if (r == jc + 1) v1 += p.diag_add;
        }
        *reinterpret_cast<double2 *>(&tile[r][jc]) = make_double2(v0, v1);
        if (FULL) {
            *reinterpret_cast<double2 *>(dst) = make_double2(v0, v1);
        } else if (r0 + r < n && c0 + jc < n) {
            if (vec && c0 + jc + 1 < n) {
                *reinterpret_cast<double2 *>(dst) = make_double2(v0, v1);
            } else {
                dst[0] = v0;
                if (c0 + jc + 1 < n) dst[1] = v1;
            }
        }
    }
    if (diag) return;
    __syncthreads();
    // mirrored tile: out[c0 + j][r0 + i] = tile[i][j].  Lane <-> i: one 128-bit shared load brings two output rows
    // (j, j + 1) of column i, each row of the output receives 256 contiguous bytes per warp store.
#pragma unroll
    for (int q = 0; q < SQ / 16; ++q) {
        const int j = 2 * (ty + 8 * q);
        double *m0 = p.out + (c0 + j) * p.ld + r0 + tx;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = tx + 32 * h;
            const double2 v = *reinterpret_cast<const double2 *>(&tile[i][j]);
            if (FULL) {
                m0[32 * h] = v.x;
                m0[p.ld + 32 * h] = v.y;
            } else if (r0 + i < n) {
                if (c0 + j < n) m0[32 * h] = v.x;
                if (c0 + j + 1 < n) m0[p.ld + 32 * h] = v.y;
            }
        }
    }
}

template <int KIND, int D, bool FAST>
__global__ void __launch_bounds__(256) kernel_sym_kernel(ExpQuadArgs p, int ntiles, int nsuper, int vec) {
    __shared__ double xi[SQ][D];
    __shared__ double xj[SQ][D];
    __shared__ __align__(16) double tile[SQ][SQ_LD];
    __shared__ double tab[EXP_TAB];
    // linear super-block id -> (SI, SJ) with SI <= SJ, row-major over the upper triangle of super-tiles
    const int64_t sb = blockIdx.x / (SUPER * SUPER);
    const int local = (int)(blockIdx.x % (SUPER * SUPER));
    int SI = (int)floor((2.0 * nsuper + 1.0 - sqrt((2.0 * nsuper + 1.0) * (2.0 * nsuper + 1.0) - 8.0 * (double)sb)) * 0.5);
    // guard against floating-point rounding of the closed form
    while ((int64_t)SI * nsuper - (int64_t)SI * (SI - 1) / 2 > sb) --SI;
    while ((int64_t)(SI + 1) * nsuper - (int64_t)(SI + 1) * SI / 2 <= sb) ++SI;
    const int SJ = SI + (int)(sb - ((int64_t)SI * nsuper - (int64_t)SI * (SI - 1) / 2));
    const int I = SI * SUPER + local / SUPER, J = SJ * SUPER + local % SUPER;
    if (I >= ntiles || J >= ntiles || I > J) return;
    const int64_t r0 = (int64_t)I * SQ, c0 = (int64_t)J * SQ;
    const int64_t n = p.n1;
    for (int t = threadIdx.x; t < SQ * D; t += 256) {
        const int64_t r = r0 + t / D, c = c0 + t / D;
        xi[t / D][t % D] = r < n ? p.x1[r * D + t % D] : 0.0;
        xj[t / D][t % D] = c < n ? p.x1[c * D + t % D] : 0.0;
    }
    if (FAST) exp_table_fill(tab, p.amp2);
    __syncthreads();
    const bool full = vec && c0 + SQ <= n;                  // (r0 <= c0) CTA-uniform
    if (full)
        sym_tile<KIND, D, FAST, true>(p, xi, xj, tile, tab, r0, c0, I == J, true);
    else
        sym_tile<KIND, D, FAST, false>(p, xi, xj, tile, tab, r0, c0, I == J, vec != 0);
}

template <int KIND, int D, bool FAST>
static int launch_kernel_matrix(const ExpQuadArgs &p, bool symmetric, cudaStream_t s) {
    const int vec = ((p.ld % 2 == 0) && (((uintptr_t)p.out) % 16 == 0)) ? 1 : 0;
    if (symmetric) {
        const int nt = (int)((p.n1 + SQ - 1) / SQ);
        const int ns = (nt + SUPER - 1) / SUPER;
        const int64_t blocks = (int64_t)ns * (ns + 1) / 2 * SUPER * SUPER;
        if (blocks > 0x7fffffffLL) {
            set_error("kernel matrix: too large for one launch");
            return VGP_ERR_INVALID;
        }
        kernel_sym_kernel<KIND, D, FAST><<<(unsigned)blocks, 256, 0, s>>>(p, nt, ns, vec);
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
                VGP_TRY((launch_kernel_matrix<KIND, D, FAST>(q, false, s)));
            }
            return VGP_OK;
        }
        kernel_rect_kernel<KIND, D, FAST><<<grid, 256, 0, s>>>(p, vec);
    }
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

template <int KIND, int D>
static int launch_kind_dim(const ExpQuadArgs &p, bool symmetric, cudaStream_t s) {
    // the exponent-field scaling of exp_nonpos needs a^2 2^(j/32) p(r) to stay a normal number after * 2^-937
    const bool fast = p.amp2 >= 0x1p-60 && p.amp2 <= 0x1p60;
    return fast ? launch_kernel_matrix<KIND, D, true>(p, symmetric, s)
                : launch_kernel_matrix<KIND, D, false>(p, symmetric, s);
}

template <int KIND>
static int launch_kind(const ExpQuadArgs &p, int d, bool symmetric, cudaStream_t s) {
    switch (d) {
        case 1: return launch_kind_dim<KIND, 1>(p, symmetric, s);
        case 2: return launch_kind_dim<KIND, 2>(p, symmetric, s);
        case 3: return launch_kind_dim<KIND, 3>(p, symmetric, s);
        case 4: return launch_kind_dim<KIND, 4>(p, symmetric, s);
        case 5: return launch_kind_dim<KIND, 5>(p, symmetric, s);
        case 6: return launch_kind_dim<KIND, 6>(p, symmetric, s);
        case 7: return launch_kind_dim<KIND, 7>(p, symmetric, s);
        case 8: return launch_kind_dim<KIND, 8>(p, symmetric, s);
    }
    set_error("kernel matrix: feature dimension %d outside [1, 8]", d);
    return VGP_ERR_INVALID;
}

int kernel_matrix_dispatch(int kind, const double *x1, int64_t n1, const double *x2, int64_t n2, int d,
                           double amplitude, double length_scale, double diag_add, int64_t diag_col0, double *out,
                           int64_t ld, cudaStream_t s) {
    if (n1 == 0 || n2 == 0) return VGP_OK;
    ExpQuadArgs p;
    p.x1 = x1;
    p.x2 = x2;
    p.n1 = n1;
    p.n2 = n2;
    p.amp2 = amplitude * amplitude;
    p.diag_add = diag_add;
    p.diag_col0 = diag_col0;
    p.out = out;
    p.ld = ld;
    const bool symmetric = (x1 == x2) && (n1 == n2) && diag_col0 == 0 && n1 >= 2 * SQ;
    switch (kind) {
        case KIND_EXPQUAD:
            p.c = -0.5 / (length_scale * length_scale);
            return launch_kind<KIND_EXPQUAD>(p, d, symmetric, s);
        case KIND_MATERN12:
            p.c = -1.0 / length_scale;
            return launch_kind<KIND_MATERN12>(p, d, symmetric, s);
        case KIND_MATERN32:
            p.c = -sqrt(3.0) / length_scale;
            return launch_kind<KIND_MATERN32>(p, d, symmetric, s);
        case KIND_MATERN52:
            p.c = -sqrt(5.0) / length_scale;
            return launch_kind<KIND_MATERN52>(p, d, symmetric, s);
    }
    set_error("kernel matrix: unknown kernel kind %d", kind);
    return VGP_ERR_INVALID;
}

int expquad_dispatch_public(const double *x1, int64_t n1, const double *x2, int64_t n2, int d, double amplitude,
                            double length_scale, double diag_add, int64_t diag_col0, double *out, int64_t ld,
                            cudaStream_t s) {
    return kernel_matrix_dispatch(KIND_EXPQUAD, x1, n1, x2, n2, d, amplitude, length_scale, diag_add, diag_col0, out,
                                  ld, s);
}

}  // namespace vgp

extern "C" int vgp_kernel_matrix(int device, int kind, const double *x1_dev, int64_t n1, const double *x2_dev,
                                 int64_t n2, int d, double amplitude, double length_scale, double diag_add,
                                 int64_t diag_col0, double *out_dev, int64_t ld_out, void *stream) {
    VGP_REQUIRE(n1 >= 0 && n2 >= 0, "negative size");
    VGP_REQUIRE(ld_out >= n2, "ld_out %lld < n2 %lld", (long long)ld_out, (long long)n2);
    VGP_REQUIRE(length_scale > 0.0, "length_scale must be positive");
    VGP_REQUIRE(d >= 1 && d <= 8, "feature dimension %d outside [1, 8]", d);
    VGP_REQUIRE(kind >= VGP_KERNEL_EXPQUAD && kind <= VGP_KERNEL_MATERN52, "unknown kernel kind %d", kind);
    if (n1 == 0 || n2 == 0) return VGP_OK;
    VGP_REQUIRE(x1_dev && x2_dev && out_dev, "NULL pointer");
    VGP_ENTER(device);
    return vgp::kernel_matrix_dispatch(kind, x1_dev, n1, x2_dev, n2, d, amplitude, length_scale, diag_add, diag_col0,
                                       out_dev, ld_out, (cudaStream_t)stream);
}

extern "C" int vgp_expquad_matrix(int device, const double *x1_dev, int64_t n1, const double *x2_dev, int64_t n2,
                                  int d, double amplitude, double length_scale, double diag_add,
                                  int64_t diag_col0, double *out_dev, int64_t ld_out, void *stream) {
    return vgp_kernel_matrix(device, VGP_KERNEL_EXPQUAD, x1_dev, n1, x2_dev, n2, d, amplitude, length_scale, diag_add,
                             diag_col0, out_dev, ld_out, stream);
}
