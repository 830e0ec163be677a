// EXPERIMENTAL, opt-in (VGP_TRSM_SLAB=<width>), not validated on hardware yet: slab kernels for the triangular solves.
//
// The recursive solves of dense.cu bottom out in products with k <= 128 (m x 128 x 128 leaves and first-level updates):
// 8 107 of the 9 262 launches of potrf + trtri at n = 50 000, 4 % of the flops, 0.26 s of latency-bound waves that every
// rank of the distributed factorisation repeats (tools/factor_schedule_model.py).  The solves are
// independent by rows (X op(L) = alpha B) or by columns (L X = alpha B).  Here one CTA owns a 128-wide slab of the
// right-hand sides and walks the whole triangle of width n <= W in ONE launch:
//     right, transposed   X_j = (alpha B_j - sum_{i<j} X_i L_ji^T) inv(L_jj)^T        j ascending
//     right, plain        X_j = (alpha B_j - sum_{i>j} X_i L_ij)   inv(L_jj)          j descending
//     left,  plain        X_j = inv(L_jj) (alpha B_j - sum_{i<j} L_ji X_i)            j ascending
// Every step is the 128 x 128 tile product of gemm_kernel (cp.async ring, DMMA 8x8x4, same shared-memory layout); the
// partial result T_j goes through global memory (the CTA's own tile, L2-resident) to become the A / B operand of the
// multiplication by the cached inverse of the diagonal block.  Earlier X_i are re-read from L2: 16 flop per byte, the
// ratio the GEMM itself runs at.  In the distributed factorisation the slabs are dealt round-robin to the ranks and every
// finished X_j tile is stored into all replicas, like the GEMM epilogue does.
#include "dense.cuh"

namespace vgp {
namespace slab {

constexpr int TM = 128, TN = 128, TK = 16, STAGES = 3, THREADS = 256, WARPS_N = 4;
constexpr int LDK = TK + 4, LDB = TN + 4;                   // padded rows: conflict-free fragment loads (see dense.cu)
constexpr int A_DOUBLES = TM * LDK;
constexpr int B_DOUBLES = TN * LDK > TK * LDB ? TN * LDK : TK * LDB;
constexpr int STAGE_DOUBLES = A_DOUBLES + B_DOUBLES;
constexpr int SMEM = STAGES * STAGE_DOUBLES * 8;
constexpr int NB = 128;

struct Args {
    double *b;                 // right-hand sides, solved in place
    int64_t ldb;
    const double *l;           // lower-triangular block [n][n]
    int64_t ldl;
    const double *dinv;        // [nb][128][128]: inverses of its diagonal blocks
    int nb;                    // n / 128
    int slabs;                 // 128-row (right forms) or 128-column (left form) slabs of B
    double alpha;
    int dist_n, rank;          // distributed: slab = rank + dist_n * blockIdx.x, X tiles stored into every replica
    int64_t delta[DIST_MAX];
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// acc += A[128][16 KT] * op(B): A k-contiguous at `a` (row stride lda); B k-contiguous ([128 n][k], BKC) or k-strided
// ([k][128 n]).  All threads of the CTA call it; shared memory is free again when it returns.
template <bool BKC>
__device__ __forceinline__ void product(const double *a, int64_t lda, const double *b, int64_t ldb, int KT,
                                        double (&acc)[8][4][2], double *smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp / WARPS_N) * 64, wn0 = (warp % WARPS_N) * 32;
    auto load_stage = [&](int stage, int kt) {
        double *sa = smem + (size_t)stage * STAGE_DOUBLES;
        double *sb = sa + A_DOUBLES;
        const int64_t k0 = (int64_t)kt * TK;
#pragma unroll
        for (int it = 0; it < TM * TK / 2 / THREADS; ++it) {
            const int c = tid + it * THREADS;
            const int row = c / (TK / 2), kc = c % (TK / 2);
            cp_async16(sa + row * LDK + 2 * kc, a + row * lda + k0 + 2 * kc);
        }
#pragma unroll
        for (int it = 0; it < TN * TK / 2 / THREADS; ++it) {
            const int c = tid + it * THREADS;
            if (BKC) {
                const int row = c / (TK / 2), kc = c % (TK / 2);
                cp_async16(sb + row * LDK + 2 * kc, b + row * ldb + k0 + 2 * kc);
            } else {
                const int kr = c / (TN / 2), nc = c % (TN / 2);
                cp_async16(sb + kr * LDB + 2 * nc, b + (k0 + kr) * ldb + 2 * nc);
            }
        }
    };
    __syncthreads();                     // the previous product's fragment loads are done; global writes of the CTA visible
#pragma unroll
    for (int st = 0; st < STAGES - 1; ++st) {
        if (st < KT) load_stage(st, st);
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (kt + STAGES - 1 < KT) load_stage((kt + STAGES - 1) % STAGES, kt + STAGES - 1);
        cp_async_commit();
        const double *sa = smem + (size_t)(kt % STAGES) * STAGE_DOUBLES;
        const double *sb = sa + A_DOUBLES;
#pragma unroll
        for (int kk = 0; kk < TK; kk += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) af[i] = sa[(wm0 + 8 * i + g) * LDK + kk + t];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = BKC ? sb[(wn0 + 8 * j + g) * LDK + kk + t] : sb[(kk + t) * LDB + wn0 + 8 * j + g];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();                     // every thread's copies have landed and been consumed: the sources may be overwritten
}

__device__ __forceinline__ void zero(double (&acc)[8][4][2]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}

// c <- alpha c - acc (this rank's tile only: an intermediate)
__device__ __forceinline__ void store_partial(double *c, int64_t ldc, double alpha, const double (&acc)[8][4][2]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp / WARPS_N) * 64, wn0 = (warp % WARPS_N) * 32;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double2 *dst = reinterpret_cast<double2 *>(c + (int64_t)(wm0 + 8 * i + g) * ldc + wn0 + 8 * j + 2 * t);
            double2 o = *dst;
            o.x = fma(alpha, o.x, -acc[i][j][0]);
            o.y = fma(alpha, o.y, -acc[i][j][1]);
            *dst = o;
        }
}

// c <- acc, into every replica
__device__ __forceinline__ void store_result(const Args &p, double *c, int64_t ldc, const double (&acc)[8][4][2]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp / WARPS_N) * 64, wn0 = (warp % WARPS_N) * 32;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double2 *dst = reinterpret_cast<double2 *>(c + (int64_t)(wm0 + 8 * i + g) * ldc + wn0 + 8 * j + 2 * t);
            const double2 o = make_double2(acc[i][j][0], acc[i][j][1]);
            if (p.dist_n > 0) {
                for (int q = 0; q < p.dist_n; ++q) *(dst + (p.delta[q] >> 1)) = o;
            } else {
                *dst = o;
            }
        }
}

// FORM 0: X L^T = alpha B (rows of B independent); 1: X L = alpha B (rows); 2: L X = alpha B (columns of B independent)
template <int FORM>
__global__ void __launch_bounds__(THREADS, 1) trsm_slab_kernel(Args p) {
    extern __shared__ __align__(16) double smem[];
    const int slab = p.dist_n > 0 ? p.rank + p.dist_n * (int)blockIdx.x : (int)blockIdx.x;
    if (slab >= p.slabs) return;
    double acc[8][4][2];
    for (int step = 0; step < p.nb; ++step) {
        const int j = FORM == 1 ? p.nb - 1 - step : step;
        const double *w = p.dinv + (int64_t)j * NB * NB;
        double *c;                                         // the tile being solved: B_j of this slab
        zero(acc);
        if (FORM == 0) {
            double *bs = p.b + (int64_t)slab * NB * p.ldb;
            c = bs + (int64_t)j * NB;
            if (j > 0) product<true>(bs, p.ldb, p.l + (int64_t)j * NB * p.ldl, p.ldl, j * (NB / TK), acc, smem);
        } else if (FORM == 1) {
            double *bs = p.b + (int64_t)slab * NB * p.ldb;
            c = bs + (int64_t)j * NB;
            if (step > 0)
                product<false>(bs + (int64_t)(j + 1) * NB, p.ldb, p.l + (int64_t)(j + 1) * NB * p.ldl + (int64_t)j * NB, p.ldl,
                               step * (NB / TK), acc, smem);
        } else {
            double *bc = p.b + (int64_t)slab * NB;
            c = bc + (int64_t)j * NB * p.ldb;
            if (j > 0) product<false>(p.l + (int64_t)j * NB * p.ldl, p.ldl, bc, p.ldb, j * (NB / TK), acc, smem);
        }
        store_partial(c, p.ldb, p.alpha, acc);             // T_j = alpha B_j - sum, in place
        zero(acc);
        if (FORM == 0)
            product<true>(c, p.ldb, w, NB, NB / TK, acc, smem);            // T_j inv(L_jj)^T
        else if (FORM == 1)
            product<false>(c, p.ldb, w, NB, NB / TK, acc, smem);           // T_j inv(L_jj)
        else
            product<false>(w, NB, c, p.ldb, NB / TK, acc, smem);           // inv(L_jj) T_j
        store_result(p, c, p.ldb, acc);
    }
    if (p.dist_n > 0) __threadfence_system();
}

template <int FORM>
static int launch(const Args &p, cudaStream_t s) {
    static bool configured[64] = {};
    int dev = 0;
    VGP_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        VGP_CUDA(cudaFuncSetAttribute(trsm_slab_kernel<FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured[dev] = true;
    }
    const int blocks = p.dist_n > 0 ? (p.slabs + p.dist_n - 1) / p.dist_n : p.slabs;
    if (blocks == 0) return VGP_OK;
    trsm_slab_kernel<FORM><<<(unsigned)blocks, THREADS, SMEM, s>>>(p);
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

}  // namespace slab

// form 0 / 1: B is [m][n], X op(L) = alpha B;  form 2: B is [n][m], L X = alpha B.  n, m multiples of 128; dinv holds
// the inverses of L's n / 128 diagonal blocks.  dist != nullptr: slabs shared out over the ranks, results stored into
// every replica (the caller brackets the launch with dense_dist_barrier).
int slab_trsm(int form, int64_t m, int64_t n, double alpha, const double *l, int64_t ldl, const double *dinv, double *b,
              int64_t ldb, const DistContext *dist, cudaStream_t s) {
    VGP_REQUIRE(form >= 0 && form <= 2 && m % slab::NB == 0 && n % slab::NB == 0 && dinv, "slab_trsm: bad arguments");
    VGP_REQUIRE(ldb % 2 == 0 && ldl % 2 == 0 && ((uintptr_t)b & 15) == 0 && ((uintptr_t)l & 15) == 0 &&
                    ((uintptr_t)dinv & 15) == 0,
                "slab_trsm: operands must be 16-byte aligned with even leading dimensions");
    if (m == 0 || n == 0) return VGP_OK;
    slab::Args p;
    p.b = b;
    p.ldb = ldb;
    p.l = l;
    p.ldl = ldl;
    p.dinv = dinv;
    p.nb = (int)(n / slab::NB);
    p.slabs = (int)(m / slab::NB);
    p.alpha = alpha;
    p.dist_n = 0;
    p.rank = 0;
    if (dist && dist->nranks > 1) {
        p.dist_n = dist->nranks;
        p.rank = dist->rank;
        for (int q = 0; q < dist->nranks; ++q) p.delta[q] = dist->delta[q];
    }
    if (form == 0) return slab::launch<0>(p, s);
    if (form == 1) return slab::launch<1>(p, s);
    return slab::launch<2>(p, s);
}

}  // namespace vgp
