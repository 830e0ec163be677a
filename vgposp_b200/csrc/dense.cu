// (2) Dense float64 layer: DMMA GEMM + recursive Cholesky / triangular inverse / X^T X / triangular solves.
//
// Stands in for the LAPACK/Eigen calls the reference reaches through TensorFlow: tf.linalg.cholesky inside
// tfd.GaussianProcess.log_prob (gp_functions.py:166-172, 3D_sin_wave.py:172), the triangular solves and
// matmuls of the VGP loss (variational_Gaussian_process_example.py:68-74,96-99), and the (pseudo-)inverse
// behind every greedy denominator (placement_algorithm2.py:399-413).
//
// Design
//   * Every matrix is row-major with all dimensions multiples of 128 (callers pad with an identity block),
//     so no kernel has edge cases.
//   * All O(n^3) work funnels into one GEMM, float64 tensor-core MMA (mma.sync.m8n8k4.f64 -> DMMA; tcgen05 has no
//     f64 kind, and the larger f64 mma shapes decompose into DMMA.8x8x4 on sm_100a), in two producers:
//       - `gemm_tma_kernel`: operands staged by TMA (cp.async.bulk.tensor.2d + mbarrier, 128-byte swizzle, 4 stages),
//         CTA tile 128x64 (two CTAs per SM) or 64x64 for latency-bound small products -- the default;
//       - `gemm_kernel`: a 3-deep cp.async ring with padded conflict-free layouts, CTA tile 128x64 or 128x128 -- for
//         large products with a k-strided A and for the in-place products of the triangular solves.
//   * Cholesky, L^-1 and L^-T L^-1 are the classic recursive (2x2 block) formulations: each level is
//     one or two large GEMMs on rectangular off-diagonal blocks plus recursion on the diagonal blocks;
//     128x128 diagonal blocks are handled by single-CTA shared-memory kernels.  Diagonal-block solves
//     multiply by an explicitly inverted 128x128 block (one more GEMM call, in place).
#include <cuda.h>          // CUtensorMap types only: the encoder is fetched through cudaGetDriverEntryPoint
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <thread>

#include "dense.cuh"
#include "block128.cuh"

namespace vgp {

// =====================================================================================================
// GEMM
// =====================================================================================================
// Alignment every operand dimension must have (tile shapes below divide it).
constexpr int BM = 128, BN = 128, BK = 16;

// Tile configuration of gemm_kernel.  Warp tile is always 64 x 32 (8 x 4 DMMA 8x8x4 accumulators, 128 registers);
// operands sit in shared memory with padded rows so that the 64-bit fragment loads of a warp (8 rows x 4 k) and the
// 128-bit cp.async stores are bank-conflict free:
//   k-contiguous operand tile  [rows][BK + 4]    row stride = 32 B mod 128 B
//   k-strided   operand tile   [BK][rows + 4]    row stride = 32 B mod 128 B
template <int BM_, int BN_, int BK_, int STAGES_, int MINB_>
struct GemmCfg {
    static constexpr int TM = BM_, TN = BN_, TK = BK_, STAGES = STAGES_, MINB = MINB_;
    static constexpr int WARPS_M = TM / 64, WARPS_N = TN / 32, THREADS = 32 * WARPS_M * WARPS_N;
    static constexpr int LDK = TK + 4, LDA = TM + 4, LDB = TN + 4;
    static constexpr int A_DOUBLES = TM * LDK > TK * LDA ? TM * LDK : TK * LDA;
    static constexpr int B_DOUBLES = TN * LDK > TK * LDB ? TN * LDK : TK * LDB;
    static constexpr int STAGE_DOUBLES = A_DOUBLES + B_DOUBLES;
    static constexpr int SMEM = STAGES * STAGE_DOUBLES * 8;
    static_assert(TM % 64 == 0 && TN % 32 == 0 && TK % 4 == 0 && BM % TM == 0 && BN % TN == 0 && TM % TN == 0, "tile");
    static_assert((TM * TK / 2) % THREADS == 0 && (TN * TK / 2) % THREADS == 0, "load split");
};
// Measured on B200 (profiles/r01_gemm_sweep.json): the single-CTA 128x128 tile leaves the DMMA pipe idle 12.5 % of the
// time (ncu: barrier + scoreboard stalls hit all 8 warps at once), 32.5 TFLOP/s at 8192^3; two 128x64 CTAs per SM
// drift out of phase and cover each other's barrier / fill stalls: 35.0-35.6 TFLOP/s (cuBLAS DGEMM: 35.5-36.2), and
// twice the tiles for the small products of the recursions.
using CfgPair = GemmCfg<128, 64, 16, 3, 2>;     // default: 4 warps, 90 KB, two CTAs per SM
using CfgBase = GemmCfg<128, 128, 16, 3, 1>;    // 8 warps, 120 KB, one CTA per SM: in-place products (see in_place)

struct GemmArgs {
    const double *a;
    const double *b;
    double *c;
    int64_t lda, ldb, ldc;
    int64_t m, n, k;
    double alpha, beta;
    int lower;
    int in_place;            // bit 0: C aliases A (n = 128: needs 128-wide tiles, CfgBase); bit 1: C aliases B (m = 128:
                             // needs 128-row tiles -- every configuration except tma::Small)
    int64_t k_split;         // k range handled by one blockIdx.z slice (== k without split-K)
    int64_t c_split_stride;  // element offset between the partial outputs of consecutive slices
    // distributed mode (dist_n > 0): this launch covers the linear tile range [tile0, tile0 + gridDim.x) of the
    // whole product and stores every finished tile into all dist_n replicas (c + delta[q])
    int dist_n;
    int64_t tile0, tiles_n;
    int64_t delta[DIST_MAX];
};

thread_local DistContext *g_dist = nullptr;
void dense_set_dist(DistContext *ctx) { g_dist = ctx; }

thread_local RowGate *g_gate = nullptr;
void dense_set_gate(RowGate *gate) { g_gate = gate; }

// `rows` storage rows starting at p are about to be touched by a launch on `s`
static int gate_wait(const double *p, int64_t rows, cudaStream_t s) {
    RowGate *g = g_gate;
    if (!g || g->waited >= g->nchunks || p < g->base || p >= g->base + g->rows * g->ld) return VGP_OK;
    const int64_t last = (p - g->base) / g->ld + rows - 1;
    int chunk = (int)(last / g->chunk_rows);
    if (chunk >= g->nchunks) chunk = g->nchunks - 1;
    if (chunk >= g->waited) {
        if (g->recorded) {
            while (g->recorded->load(std::memory_order_acquire) <= chunk) {
                if (g->failed && g->failed->load(std::memory_order_acquire)) {
                    set_error("the host-to-device copy of cov_vv failed");
                    return VGP_ERR_CUDA;
                }
                std::this_thread::yield();
            }
        }
        VGP_CUDA(cudaStreamWaitEvent(s, g->events[chunk], 0));
        g->waited = chunk + 1;
    }
    return VGP_OK;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct DistFlags {
    unsigned long long *f[DIST_MAX];
};
// Cross-rank barrier in stream order: everything enqueued before it on every rank (including peer stores) is
// complete and visible before anything enqueued after it on any rank starts.
__global__ void dist_barrier_kernel(DistFlags fl, int rank, int nranks, unsigned long long seq) {
    const int q = threadIdx.x;
    if (q >= nranks) return;
    __threadfence_system();
    st_release_sys_u64(fl.f[q] + rank, seq);
    const unsigned long long *mine = fl.f[rank] + q;
    const long long t0 = clock64();
    while (ld_acquire_sys_u64(mine) < seq) {
        if (clock64() - t0 > 40000000000LL) {               // ~20 s: a peer died
            atomicExch(reinterpret_cast<int *>(fl.f[rank] + DIST_MAX), 1 + q);
            break;
        }
    }
}

int dense_dist_barrier(DistContext &ctx, cudaStream_t s) {
    DistFlags fl;
    for (int q = 0; q < DIST_MAX; ++q) fl.f[q] = q < ctx.nranks ? ctx.flags[q] : nullptr;
    dist_barrier_kernel<<<1, 32, 0, s>>>(fl, ctx.rank, ctx.nranks, ++ctx.seq);
    VGP_LAUNCH_CHECK();
    ++ctx.barriers;
    return VGP_OK;
}

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

template <class C, bool AKC, bool BKC>
__global__ void __launch_bounds__(C::THREADS, C::MINB) gemm_kernel(GemmArgs p) {
    extern __shared__ __align__(16) double smem[];
    constexpr int TM = C::TM, TN = C::TN, TK = C::TK, STAGES = C::STAGES, THREADS = C::THREADS;
    constexpr int LDK = C::LDK, LDA = C::LDA, LDB = C::LDB;
    constexpr int RATIO = TM / TN;                              // tile columns per tile row on the diagonal
    int tm, tn;
    if (p.lower) {
        // lower mode: row tm holds the RATIO * (tm + 1) tiles that touch the lower triangle (128-block granularity)
        const int64_t b = p.tile0 + blockIdx.x;
        int r = (int)((sqrt(8.0 * (double)b / RATIO + 1.0) - 1.0) * 0.5);
        while ((int64_t)RATIO * r * (r + 1) / 2 > b) --r;
        while ((int64_t)RATIO * (r + 1) * (r + 2) / 2 <= b) ++r;
        tm = r;
        tn = (int)(b - (int64_t)RATIO * r * (r + 1) / 2);
    } else if (p.dist_n > 0) {
        const int64_t b = p.tile0 + blockIdx.x;
        tm = (int)(b / p.tiles_n);
        tn = (int)(b % p.tiles_n);
    } else {
        tn = blockIdx.x;
        tm = blockIdx.y;
    }
    const int64_t m0 = (int64_t)tm * TM, n0 = (int64_t)tn * TN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp / C::WARPS_N) * 64, wn0 = (warp % C::WARPS_N) * 32;
    const int64_t kbase = (int64_t)blockIdx.z * p.k_split;
    const int64_t klen = p.k - kbase < p.k_split ? p.k - kbase : p.k_split;
    const int KT = (int)(klen / TK);
    double *cout = p.c + (int64_t)blockIdx.z * p.c_split_stride;

    auto load_stage = [&](int stage, int kt) {
        double *sa = smem + (size_t)stage * C::STAGE_DOUBLES;
        double *sb = sa + C::A_DOUBLES;
        const int64_t k0 = kbase + (int64_t)kt * TK;
#pragma unroll
        for (int it = 0; it < TM * TK / 2 / THREADS; ++it) {
            const int c = tid + it * THREADS;
            if (AKC) {
                const int row = c / (TK / 2), kc = c % (TK / 2);
                cp_async16(sa + row * LDK + 2 * kc, p.a + (m0 + row) * p.lda + k0 + 2 * kc);
            } else {
                const int kr = c / (TM / 2), mc = c % (TM / 2);
                cp_async16(sa + kr * LDA + 2 * mc, p.a + (k0 + kr) * p.lda + m0 + 2 * mc);
            }
        }
#pragma unroll
        for (int it = 0; it < TN * TK / 2 / THREADS; ++it) {
            const int c = tid + it * THREADS;
            if (BKC) {
                const int row = c / (TK / 2), kc = c % (TK / 2);
                cp_async16(sb + row * LDK + 2 * kc, p.b + (n0 + row) * p.ldb + k0 + 2 * kc);
            } else {
                const int kr = c / (TN / 2), nc = c % (TN / 2);
                cp_async16(sb + kr * LDB + 2 * nc, p.b + (k0 + kr) * p.ldb + n0 + 2 * nc);
            }
        }
    };

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

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
        const double *sa = smem + (size_t)(kt % STAGES) * C::STAGE_DOUBLES;
        const double *sb = sa + C::A_DOUBLES;
#pragma unroll
        for (int kk = 0; kk < TK; kk += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                af[i] = AKC ? sa[(wm0 + 8 * i + g) * LDK + kk + t] : sa[(kk + t) * LDA + wm0 + 8 * i + g];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                bf[j] = BKC ? sb[(wn0 + 8 * j + g) * LDK + kk + t] : sb[(kk + t) * LDB + wn0 + 8 * j + g];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = m0 + wm0 + 8 * i + g;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double2 *dst = reinterpret_cast<double2 *>(cout + r * p.ldc + n0 + wn0 + 8 * j + 2 * t);
            double2 o;
            if (p.beta != 0.0) {
                const double2 old = *dst;
                o.x = fma(p.alpha, acc[i][j][0], p.beta * old.x);
                o.y = fma(p.alpha, acc[i][j][1], p.beta * old.y);
            } else {
                o.x = p.alpha * acc[i][j][0];
                o.y = p.alpha * acc[i][j][1];
            }
            if (p.dist_n > 0) {
                for (int q = 0; q < p.dist_n; ++q) *(dst + (p.delta[q] >> 1)) = o;     // every replica, peers over NVLink
            } else {
                *dst = o;
            }
        }
    }
    if (p.dist_n > 0) {             // one system-scope fence per CTA after its barrier (fences are cumulative)
        __syncthreads();
        if (threadIdx.x == 0) __threadfence_system();
    }
}

// -----------------------------------------------------------------------------------------------------
// TMA-fed variant for two k-contiguous operands (A [m][k], B [n][k]: the Cholesky's TRSM / SYRK updates and the
// ELBO's K_zx K_zx^T).  One elected thread arms an mbarrier with the byte count of a stage and issues two
// cp.async.bulk.tensor.2d loads (A box 16 k x 128 rows, B box 16 k x 64 rows); the tiles land densely in shared memory
// under the 128-byte hardware swizzle (16-byte chunk c of row r sits at chunk c ^ (r & 7)); every thread waits on the
// barrier's phase instead of cp.async.wait_group.  The 8 x 4 float64 DMMA fragment (lane g,t reads row g, k t) is
// conflict-free under that swizzle only if the rows of a half-warp differ in bits 1-2 of (r & 7): fragment i therefore
// takes the rows 16 (i / 2) + 2 g + (i % 2) of the warp tile (even / odd rows of a 16-row group) instead of 8
// consecutive ones -- pure bookkeeping, the accumulators just belong to permuted rows / columns, and the epilogue
// reassembles four consecutive columns per thread from the two fragments of a group (two 128-bit stores).
// Same tile shape, warp tile and two-CTAs-per-SM residency as CfgPair; 4 stages of 24 KB.
// -----------------------------------------------------------------------------------------------------
namespace tma {
constexpr int TK = 16, STAGES = 4, THREADS = 128;
// CTA tile TM x TN, four warps (2 x 2) of WM x WN.  Big: the throughput shape.  Small: four times the CTAs for the
// same product and half the DMMA chain per warp and k-tile -- for products too small to fill the SMs with Big tiles
// (the m^3 products of the ELBO's reverse sweep, the leaves of the recursive factorisations), which are bound by the
// latency of one CTA's k loop, not by the pipe.
template <int TM_, int TN_, int MINB_>
struct Shape {
    static constexpr int TM = TM_, TN = TN_, MINB = MINB_, WM = TM_ / 2, WN = TN_ / 2, FI = WM / 8, FJ = WN / 8;
    static constexpr int A_BYTES = TM * TK * 8, B_BYTES = TN * TK * 8, STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + STAGES * 8;
    static_assert(WM % 16 == 0 && WN % 16 == 0 && STAGE_BYTES % 1024 == 0, "fragment groups are 16 rows / columns");
};
using Big = Shape<128, 64, 2>;
using Small = Shape<64, 64, 3>;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    const long long t0 = clock64();
    for (;;) {
        unsigned done;
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if (clock64() - t0 > 8000000000LL) __trap();      // ~4 s: a lost transaction fails the launch instead of hanging
    }
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *map, int c0, int c1, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
}  // namespace tma

// AKC / BKC: operand stored k-contiguous ([rows][k]: one box of 16 k x TM / TN rows).  Otherwise it is stored [k][rows]
// and arrives as boxes of 16 k x 16 rows (2 KB each, the 128-byte rows now run along m / n); a fragment then takes the
// rows {0,1,8,9,2,3,10,11} (+4 for the odd fragment) of its box, which keeps the 16 lanes of a half-warp on 16 distinct
// 8-byte words -- and for B, pairs of accumulators still belong to adjacent columns.
template <class S, bool AKC, bool BKC>
__global__ void __launch_bounds__(tma::THREADS, S::MINB)
    gemm_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GemmArgs p) {
    using namespace tma;
    constexpr int TM = S::TM, TN = S::TN, FI = S::FI, FJ = S::FJ, A_BYTES = S::A_BYTES, STAGE_BYTES = S::STAGE_BYTES;
    extern __shared__ unsigned char raw[];
    const unsigned raw_addr = smem_u32(raw);
    unsigned char *sm = raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);      // SWIZZLE_128B wants 1024-byte alignment
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + STAGES * STAGE_BYTES);
    constexpr int RATIO = TM / TN;
    int tm, tn;
    if (p.lower) {
        const int64_t b = p.tile0 + blockIdx.x;
        int r = (int)((sqrt(8.0 * (double)b / RATIO + 1.0) - 1.0) * 0.5);
        while ((int64_t)RATIO * r * (r + 1) / 2 > b) --r;
        while ((int64_t)RATIO * (r + 1) * (r + 2) / 2 <= b) ++r;
        tm = r;
        tn = (int)(b - (int64_t)RATIO * r * (r + 1) / 2);
    } else if (p.dist_n > 0) {
        const int64_t b = p.tile0 + blockIdx.x;
        tm = (int)(b / p.tiles_n);
        tn = (int)(b % p.tiles_n);
    } else {
        tn = blockIdx.x;
        tm = blockIdx.y;
    }
    const int64_t m0 = (int64_t)tm * TM, n0 = (int64_t)tn * TN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp >> 1) * S::WM, wn0 = (warp & 1) * S::WN;
    const int64_t kbase = (int64_t)blockIdx.z * p.k_split;
    const int64_t klen = p.k - kbase < p.k_split ? p.k - kbase : p.k_split;
    const int KT = (int)(klen / TK);
    double *cout = p.c + (int64_t)blockIdx.z * p.c_split_stride;

    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < STAGES; ++st) mbar_init(smem_u32(&bars[st]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int kt) {        // thread 0 only
        const int stage = kt % STAGES;
        const unsigned bar = smem_u32(&bars[stage]);
        const unsigned dst = smem_u32(sm + stage * STAGE_BYTES);
        const int k0 = (int)(kbase + (int64_t)kt * TK);
        mbar_expect_tx(bar, STAGE_BYTES);
        if (AKC) {
            tma_load_2d(dst, &map_a, k0, (int)m0, bar);
        } else {
#pragma unroll
            for (int bx = 0; bx < TM / 16; ++bx) tma_load_2d(dst + bx * 2048, &map_a, (int)m0 + 16 * bx, k0, bar);
        }
        if (BKC) {
            tma_load_2d(dst + A_BYTES, &map_b, k0, (int)n0, bar);
        } else {
#pragma unroll
            for (int bx = 0; bx < TN / 16; ++bx) tma_load_2d(dst + A_BYTES + bx * 2048, &map_b, (int)n0 + 16 * bx, k0, bar);
        }
    };
    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < STAGES - 1; ++st)
            if (st < KT) issue(st);
    }

    double acc[FI][FJ][2];
#pragma unroll
    for (int i = 0; i < FI; ++i)
#pragma unroll
        for (int j = 0; j < FJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // byte offsets inside a swizzled tile: row r (128 bytes), 16-byte chunk ((k >> 1) ^ (r & 7)), half (k & 1).
    // For every fragment of this thread (r & 7) = 2 (g & 3) + (i & 1): the chunk is ((kk / 2) ^ gx) | ((t >> 1) ^ e).
    const int gx = 2 * (g & 3);
    const unsigned a_row = AKC ? (unsigned)((wm0 + 2 * g) * 128 + (t & 1) * 8)
                               : (unsigned)((wm0 / 16) * 2048 + t * 128 + (g & 1) * 8);
    // row of the warp tile that fragment i's accumulators belong to (see the two operand layouts above)
    const int perm_g = ((g >> 1) & 1) * 8 + (g >> 2) * 2 + (g & 1);           // {0,1,8,9,2,3,10,11}[g]
    auto frag_row = [&](int i) { return AKC ? 16 * (i >> 1) + 2 * g + (i & 1) : 16 * (i >> 1) + perm_g + 4 * (i & 1); };
    const unsigned b_row = BKC ? (unsigned)(A_BYTES + (wn0 + 2 * g) * 128 + (t & 1) * 8)
                               : (unsigned)(A_BYTES + (wn0 / 16) * 2048 + t * 128 + (g & 1) * 8);
    const int pg = ((g >> 1) & 1) * 4 + (g >> 2);          // (column >> 1) of this lane inside its 16-column box

    for (int kt = 0; kt < KT; ++kt) {
        const int stage = kt % STAGES;
        __syncthreads();                                   // everyone is done with k-tile kt - 1: its stage is free
        if (tid == 0 && kt + STAGES - 1 < KT) issue(kt + STAGES - 1);
        mbar_wait(smem_u32(&bars[stage]), (unsigned)((kt / STAGES) & 1));
        const unsigned char *st = sm + stage * STAGE_BYTES;
#pragma unroll
        for (int kk = 0; kk < TK; kk += 4) {
            const int ck = (kk / 2) ^ gx;
            const unsigned off0 = (unsigned)((ck | (t >> 1)) << 4), off1 = (unsigned)((ck | ((t >> 1) ^ 1)) << 4);
            double af[FI], bf[FJ];
            // k-strided boxes: row kk + t of the box, chunk ((column >> 1) ^ (row & 7)); (kk + t) & 7 = (kk & 4) | t
            const int kt7 = (kk & 4) | t;
            const unsigned o0 = (unsigned)(kk * 128 + ((pg ^ kt7) << 4)), o1 = (unsigned)(kk * 128 + (((pg | 2) ^ kt7) << 4));
            if (AKC) {
#pragma unroll
                for (int i = 0; i < FI; ++i)
                    af[i] = *reinterpret_cast<const double *>(st + a_row + (i >> 1) * 2048 + (i & 1) * 128 +
                                                              ((i & 1) ? off1 : off0));
            } else {
#pragma unroll
                for (int i = 0; i < FI; ++i)
                    af[i] = *reinterpret_cast<const double *>(st + a_row + (i >> 1) * 2048 + ((i & 1) ? o1 : o0));
            }
            if (BKC) {
#pragma unroll
                for (int j = 0; j < FJ; ++j)
                    bf[j] = *reinterpret_cast<const double *>(st + b_row + (j >> 1) * 2048 + (j & 1) * 128 +
                                                              ((j & 1) ? off1 : off0));
            } else {
#pragma unroll
                for (int j = 0; j < FJ; ++j)
                    bf[j] = *reinterpret_cast<const double *>(st + b_row + (j >> 1) * 2048 + ((j & 1) ? o1 : o0));
            }
#pragma unroll
            for (int i = 0; i < FI; ++i)
#pragma unroll
                for (int j = 0; j < FJ; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    if (!BKC) {
        // B fragment j = 2 q + e of box q holds, for this thread, the adjacent columns {0, 8, 2, 10}[t] + 4 e (+ 0, 1)
#pragma unroll
        for (int i = 0; i < FI; ++i) {
            const int64_t r = m0 + wm0 + frag_row(i);
#pragma unroll
            for (int j = 0; j < FJ; ++j) {
                const int col = 16 * (j >> 1) + ((t & 1) * 8 + (t >> 1) * 2) + 4 * (j & 1);
                double2 *dst = reinterpret_cast<double2 *>(cout + r * p.ldc + n0 + wn0 + col);
                double2 o;
                if (p.beta != 0.0) {
                    const double2 old = *dst;
                    o.x = fma(p.alpha, acc[i][j][0], p.beta * old.x);
                    o.y = fma(p.alpha, acc[i][j][1], p.beta * old.y);
                } else {
                    o.x = p.alpha * acc[i][j][0];
                    o.y = p.alpha * acc[i][j][1];
                }
                if (p.dist_n > 0) {
                    for (int d = 0; d < p.dist_n; ++d) *(dst + (p.delta[d] >> 1)) = o;
                } else {
                    *dst = o;
                }
            }
        }
        if (p.dist_n > 0) {
            __syncthreads();
            if (threadIdx.x == 0) __threadfence_system();
        }
        return;
    }

    // epilogue: fragment i holds rows 16 (i / 2) + 2 g + (i % 2); the two column fragments of a 16-column group hold
    // columns 4 t + {0, 2} and 4 t + {1, 3}: four consecutive columns per thread
#pragma unroll
    for (int i = 0; i < FI; ++i) {
        const int64_t r = m0 + wm0 + frag_row(i);
#pragma unroll
        for (int q = 0; q < FJ / 2; ++q) {
            double2 *dst = reinterpret_cast<double2 *>(cout + r * p.ldc + n0 + wn0 + 16 * q + 4 * t);
            double2 o0, o1;
            if (p.beta != 0.0) {
                const double2 old0 = dst[0], old1 = dst[1];
                o0.x = fma(p.alpha, acc[i][2 * q][0], p.beta * old0.x);
                o0.y = fma(p.alpha, acc[i][2 * q + 1][0], p.beta * old0.y);
                o1.x = fma(p.alpha, acc[i][2 * q][1], p.beta * old1.x);
                o1.y = fma(p.alpha, acc[i][2 * q + 1][1], p.beta * old1.y);
            } else {
                o0.x = p.alpha * acc[i][2 * q][0];
                o0.y = p.alpha * acc[i][2 * q + 1][0];
                o1.x = p.alpha * acc[i][2 * q][1];
                o1.y = p.alpha * acc[i][2 * q + 1][1];
            }
            if (p.dist_n > 0) {
                for (int d = 0; d < p.dist_n; ++d) {                 // every replica, peers over NVLink
                    double2 *rep = dst + (p.delta[d] >> 1);
                    rep[0] = o0;
                    rep[1] = o1;
                }
            } else {
                dst[0] = o0;
                dst[1] = o1;
            }
        }
    }
    if (p.dist_n > 0) {             // one system-scope fence per CTA after its barrier (fences are cumulative)
        __syncthreads();
        if (threadIdx.x == 0) __threadfence_system();
    }
}

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
    static TensorMapEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (TensorMapEncodeFn)ptr;
    }
    return fn;
}

// k-contiguous operand [rows][ld] -> tensor map with boxes of 16 k x box_rows rows, 128-byte swizzle
static int make_operand_map(CUtensorMap *map, const double *base, int64_t rows, int64_t k, int64_t ld, int box_rows) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    VGP_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    const cuuint32_t box[2] = {(cuuint32_t)tma::TK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VGP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return VGP_OK;
}

template <class S, bool AKC, bool BKC>
static int gemm_launch_tma(const GemmArgs &p, cudaStream_t s) {
    static bool configured[64] = {};
    int dev = 0;
    VGP_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        VGP_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<S, AKC, BKC>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM));
        configured[dev] = true;
    }
    alignas(64) CUtensorMap ma, mb;
    if (AKC)
        VGP_TRY(make_operand_map(&ma, p.a, p.m, p.k, p.lda, S::TM));
    else
        VGP_TRY(make_operand_map(&ma, p.a, p.k, p.m, p.lda, 16));        // [k][m]: boxes of 16 m x 16 k
    if (BKC)
        VGP_TRY(make_operand_map(&mb, p.b, p.n, p.k, p.ldb, S::TN));
    else
        VGP_TRY(make_operand_map(&mb, p.b, p.k, p.n, p.ldb, 16));        // [k][n]: boxes of 16 n x 16 k
    const int64_t tm = p.m / S::TM, tn = p.n / S::TN;
    const unsigned splits = (unsigned)((p.k + p.k_split - 1) / p.k_split);
    constexpr int RATIO = S::TM / S::TN;
    if (p.dist_n > 0) {
        GemmArgs q = p;
        const int64_t total = p.lower ? (int64_t)RATIO * tm * (tm + 1) / 2 : tm * tn;
        const int64_t rank = p.tile0, nranks = p.tiles_n;
        const int64_t each = total / nranks, rem = total % nranks;
        q.tile0 = rank * each + (rank < rem ? rank : rem);
        const int64_t mine = each + (rank < rem ? 1 : 0);
        q.tiles_n = tn;
        if (mine > 0) {
            gemm_tma_kernel<S, AKC, BKC><<<dim3((unsigned)mine, 1, 1), tma::THREADS, S::SMEM, s>>>(ma, mb, q);
            VGP_LAUNCH_CHECK();
        }
        return VGP_OK;
    }
    if (p.lower) {
        const int64_t blocks = (int64_t)RATIO * tm * (tm + 1) / 2;
        gemm_tma_kernel<S, AKC, BKC><<<dim3((unsigned)blocks, 1, splits), tma::THREADS, S::SMEM, s>>>(ma, mb, p);
    } else {
        gemm_tma_kernel<S, AKC, BKC><<<dim3((unsigned)tn, (unsigned)tm, splits), tma::THREADS, S::SMEM, s>>>(ma, mb, p);
    }
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

// Tile configuration for one product.  The triangular solves multiply in place: C aliases A with n = 128 (then one CTA
// must own all 128 columns of its row block: only the 128x128 tile), or C aliases B with m = 128 (then one CTA must
// own all 128 rows of its column block: every shape with 128-row tiles, i.e. not tma::Small -- a 64-row tile there let
// two CTAs overwrite rows the other was still reading, an intermittent wrong K^-1 in the ELBO step until it was
// excluded).
// VGP_OPT_GEMM_TILE_CONFIG = 0 base | 1 pair | 2 tma forces one for every product that may legally use it
// (measurement knob; -1 = automatic).
enum GemmChoice { CHOICE_BASE = 0, CHOICE_PAIR = 1, CHOICE_TMA = 2 };
static int gemm_forced_choice() { return (int)option(VGP_OPT_GEMM_TILE_CONFIG); }
static GemmChoice gemm_choose(const GemmArgs &p) {
    if (p.in_place & 1) return CHOICE_BASE;
    const int f = gemm_forced_choice();
    if (f >= 0) return (GemmChoice)f;
    return CHOICE_TMA;              // TMA producer where both operands are k-contiguous, CfgPair (cp.async) otherwise
}

template <class C, bool AKC, bool BKC>
static int gemm_launch_cfg(const GemmArgs &p, cudaStream_t s) {
    static bool configured[64] = {};
    int dev = 0;
    VGP_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        VGP_CUDA(cudaFuncSetAttribute(gemm_kernel<C, AKC, BKC>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        configured[dev] = true;
    }
    const int64_t tm = p.m / C::TM, tn = p.n / C::TN;
    const unsigned splits = (unsigned)((p.k + p.k_split - 1) / p.k_split);
    if (p.dist_n > 0) {
        // p.tiles_n carries this rank's share (rank, nranks packed by dense_gemm) on entry, the tile-column count in
        // the kernel: the linear tile range is cut here because the tile count depends on the configuration
        GemmArgs q = p;
        const int64_t total = p.lower ? (int64_t)(C::TM / C::TN) * tm * (tm + 1) / 2 : tm * tn;
        const int64_t rank = p.tile0, nranks = p.tiles_n;
        const int64_t each = total / nranks, rem = total % nranks;
        q.tile0 = rank * each + (rank < rem ? rank : rem);
        const int64_t mine = each + (rank < rem ? 1 : 0);
        q.tiles_n = tn;
        if (mine > 0) {
            gemm_kernel<C, AKC, BKC><<<dim3((unsigned)mine, 1, 1), C::THREADS, C::SMEM, s>>>(q);
            VGP_LAUNCH_CHECK();
        }
        return VGP_OK;
    }
    if (p.lower) {
        const int64_t blocks = (int64_t)(C::TM / C::TN) * tm * (tm + 1) / 2;
        gemm_kernel<C, AKC, BKC><<<dim3((unsigned)blocks, 1, splits), C::THREADS, C::SMEM, s>>>(p);
    } else {
        dim3 grid((unsigned)tn, (unsigned)tm, splits);
        gemm_kernel<C, AKC, BKC><<<grid, C::THREADS, C::SMEM, s>>>(p);
    }
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

template <bool AKC, bool BKC>
static int gemm_launch(const GemmArgs &p, cudaStream_t s) {
    const GemmChoice c = gemm_choose(p);
    if (c == CHOICE_TMA && tensor_map_encoder()) {                         // TMA producer, all four operand layouts
        // (a driver without cuTensorMapEncodeTiled falls through to the cp.async ring)
        // fewer Big tiles than half the SMs: the product is bound by the latency of one CTA's k loop -> Small tiles
        // (the lower-tile mode stays Big: its tile enumeration assumes TM = 128 row blocks)
        int64_t big_tiles = (p.m / tma::Big::TM) * (p.n / tma::Big::TN) * ((p.k + p.k_split - 1) / p.k_split);
        if (p.dist_n > 0) big_tiles /= p.tiles_n;         // distributed product: this rank's share (tiles_n = ranks here)
        const int64_t small_below = option(VGP_OPT_GEMM_SMALL_BELOW);
        if (!p.lower && !p.in_place && big_tiles < small_below) return gemm_launch_tma<tma::Small, AKC, BKC>(p, s);
        // a k-strided A arrives as eight 2 KB boxes per stage: measured equal to (B k-strided) or 1.5 % behind
        // (B k-contiguous) the cp.async ring on large products, so those keep the ring; 34.1-34.6 vs 34.3-35.1 TFLOP/s
        if (AKC) return gemm_launch_tma<tma::Big, AKC, BKC>(p, s);
    }
    if (c != CHOICE_BASE) return gemm_launch_cfg<CfgPair, AKC, BKC>(p, s);
    return gemm_launch_cfg<CfgBase, AKC, BKC>(p, s);
}

static int gemm_dispatch(int trans_a, int trans_b, const GemmArgs &p, cudaStream_t s) {
    // operand is k-contiguous when: A not transposed / B transposed
    if (!trans_a && trans_b) return gemm_launch<true, true>(p, s);
    if (!trans_a && !trans_b) return gemm_launch<true, false>(p, s);
    if (trans_a && !trans_b) return gemm_launch<false, false>(p, s);
    return gemm_launch<false, true>(p, s);
}

// Digit planes of the int8 tensor-core product (emulated.cu) for the large products; 0 = FP64 DMMA only.  Only the
// factorisation entry points (potrf / trtri / lauum and what is built on them) switch it on for their own products:
// thread-local, set by EmulateScope.
static thread_local int g_emulate_slices = 0;
static thread_local EmuWorkspace *g_emu_ws = nullptr;
struct EmulateScope {
    int prev;
    EmuWorkspace *prev_ws;
    explicit EmulateScope(DenseWorkspace &ws) : prev(g_emulate_slices), prev_ws(g_emu_ws) {
        g_emulate_slices = (int)option(VGP_OPT_GEMM_EMULATE_SLICES);
        g_emu_ws = &ws.emu;
    }
    ~EmulateScope() {
        g_emulate_slices = prev;
        g_emu_ws = prev_ws;
    }
};
static int emulate_slices() { return g_emulate_slices; }

int dense_gemm(int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, double alpha, const double *a,
               int64_t lda, const double *b, int64_t ldb, double beta, double *c, int64_t ldc, GemmTiles tiles,
               cudaStream_t s) {
    if (m == 0 || n == 0) return VGP_OK;
    VGP_REQUIRE(m % BM == 0 && n % BN == 0 && k % BK == 0 && k > 0, "dense_gemm: unpadded size %lldx%lldx%lld",
                (long long)m, (long long)n, (long long)k);
    VGP_REQUIRE(lda % 2 == 0 && ldb % 2 == 0 && ldc % 2 == 0, "dense_gemm: odd leading dimension");
    VGP_REQUIRE(((uintptr_t)a % 16) == 0 && ((uintptr_t)b % 16) == 0 && ((uintptr_t)c % 16) == 0,
                "dense_gemm: operands must be 16-byte aligned");
    VGP_REQUIRE(tiles == GEMM_FULL || m == n, "dense_gemm: lower-tile mode needs a square C");
    VGP_REQUIRE(m / BM <= 65535, "dense_gemm: too many row tiles");
    // in-place products of the triangular solves: C == A needs one tile column (n = 128, 128-wide tiles), C == B one
    // tile row (m = 128, true for both configurations) -- then every CTA reads exactly the region it overwrites
    VGP_REQUIRE(c != a || (!trans_a && n == BN), "dense_gemm: C aliases A with n = %lld", (long long)n);
    VGP_REQUIRE(c != b || (!trans_b && m == BM), "dense_gemm: C aliases B with m = %lld", (long long)m);
    GemmArgs p{a, b, c, lda, ldb, ldc, m, n, k, alpha, beta, tiles == GEMM_LOWER ? 1 : 0,
               (c == a ? 1 : 0) | (c == b ? 2 : 0), k, 0, 0, 0, 0, {0}};
    if (g_gate) {
        VGP_TRY(gate_wait(a, trans_a ? k : m, s));
        VGP_TRY(gate_wait(b, trans_b ? n : k, s));
        VGP_TRY(gate_wait(c, m, s));
    }
    DistContext *dc = g_dist;
    // large products of the factorisations: int8 tensor cores (emulated.cu), exact integer partial products
    const int emu_slices = emulate_slices();
    const int64_t emu_min = option(VGP_OPT_GEMM_EMULATE_MIN);
    const bool emulate = emu_slices >= 2 && g_emu_ws && m >= emu_min && n >= emu_min && 2 * k >= emu_min && c != a &&
                         c != b;
    if (emulate && !(dc && dc->nranks > 1))
        return emulated_gemm(*g_emu_ws, trans_a, trans_b, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, emu_slices,
                             tiles == GEMM_LOWER ? 1 : 0, s);
    if (dc && dc->nranks > 1) {
        const int64_t tm = m / BM, tn = n / BN;
        const int64_t total = tiles == GEMM_LOWER ? tm * (tm + 1) / 2 : tm * tn;
        const char *c0 = (const char *)c, *c1 = (const char *)(c + (m - 1) * ldc + n);
        const bool inside = c0 >= (const char *)dc->base && c1 <= (const char *)dc->base + dc->bytes;
        if (inside && total >= dc->min_tiles && k >= dc->min_k) {
            p.dist_n = dc->nranks;
            p.tile0 = dc->rank;                                   // (rank, nranks): gemm_launch_cfg cuts the tile range
            p.tiles_n = dc->nranks;
            for (int q = 0; q < dc->nranks; ++q) p.delta[q] = dc->delta[q];
            ++dc->dist_gemms;
            VGP_TRY(dense_dist_barrier(*dc, s));                  // every rank is done with all earlier work
            // every rank cuts the digit planes of the WHOLE operands and multiplies 1/nranks of the tiles, so the int8
            // path loses its margin as the rank count grows.  Measured at n = 50 000 (profiles/r02_dist_inverse_bench_
            // g4/g8.jsonl): 2 ranks, inverse 1.54 s against 2.38 s on the FP64 pipe; 4 ranks, 1.28 s with products from
            // 4096 up against 1.41 s (from 1024) and 1.44 s (FP64 pipe); 8 ranks, every threshold loses to the FP64
            // pipe alone (1.07 s).  VGP_OPT_DIST_EMULATE_MIN = -1 follows that table.
            int64_t dist_emu_min = option(VGP_OPT_DIST_EMULATE_MIN);
            if (dist_emu_min < 0) dist_emu_min = dc->nranks <= 2 ? emu_min : dc->nranks <= 4 ? 4096 : INT64_MAX;
            if (emulate && m >= dist_emu_min && n >= dist_emu_min && 2 * k >= dist_emu_min) {
                VGP_TRY(emulated_gemm(*g_emu_ws, trans_a, trans_b, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc,
                                      emu_slices, tiles == GEMM_LOWER ? 1 : 0, s, dc));
                return dense_dist_barrier(*dc, s);
            }
            VGP_TRY(gemm_dispatch(trans_a, trans_b, p, s));
            return dense_dist_barrier(*dc, s);                    // every tile has landed in every replica
        }
    }
    return gemm_dispatch(trans_a, trans_b, p, s);
}

// C = beta C + sum_z partial_z, fixed summation order; `lower`: only tiles with tile_row >= tile_col were computed
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const double *partial, int64_t stride, int splits,
                                                            double beta, double *c, int64_t ldc, int64_t m,
                                                            int64_t n, int lower) {
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= m * n) return;
    const int64_t i = e / n, j = e % n;
    if (lower && i / BM < j / BN) return;
    double acc = 0.0;
    for (int z = 0; z < splits; ++z) acc += partial[(int64_t)z * stride + i * n + j];
    c[i * ldc + j] = beta != 0.0 ? fma(beta, c[i * ldc + j], acc) : acc;
}

// Split-K form for short-and-wide products (m, n small, k huge: the m x m SYRK over all N observations).
// `partial` holds splits * m * n doubles.  tiles == GEMM_LOWER (square C, symmetric product A A^T): only the
// lower tiles are computed and the strict upper triangle of C is filled by mirroring (beta must be 0).
int dense_gemm_splitk(int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, double alpha, const double *a,
                      int64_t lda, const double *b, int64_t ldb, double beta, double *c, int64_t ldc, int splits,
                      double *partial, cudaStream_t s, GemmTiles tiles) {
    if (m == 0 || n == 0) return VGP_OK;
    VGP_REQUIRE(m % BM == 0 && n % BN == 0 && k % BK == 0 && k > 0, "dense_gemm_splitk: unpadded size");
    VGP_REQUIRE(splits >= 1 && partial, "dense_gemm_splitk: bad split");
    VGP_REQUIRE(tiles == GEMM_FULL || (m == n && beta == 0.0), "dense_gemm_splitk: lower mode needs square C, beta 0");
    int64_t k_split = round_up((k + splits - 1) / splits, BK);
    const int real_splits = (int)((k + k_split - 1) / k_split);
    GemmArgs p{a, b, partial, lda, ldb, n, m, n, k, alpha, 0.0, tiles == GEMM_LOWER ? 1 : 0, 0, k_split, m * n,
               0, 0, 0, {0}};
    VGP_TRY(gemm_dispatch(trans_a, trans_b, p, s));
    splitk_reduce_kernel<<<(unsigned)((m * n + 255) / 256), 256, 0, s>>>(partial, m * n, real_splits, beta, c, ldc, m, n,
                                                                        tiles == GEMM_LOWER ? 1 : 0);
    VGP_LAUNCH_CHECK();
    if (tiles == GEMM_LOWER) return dense_mirror_lower(c, m, ldc, s);
    return VGP_OK;
}

// =====================================================================================================
// small element-wise helpers
// =====================================================================================================
__global__ void add_diag_kernel(double *a, int64_t ld, int64_t n, double v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i * ld + i] += v;
}

int dense_add_diag(double *a, int64_t ld, int64_t n, double value, cudaStream_t s) {
    if (n == 0) return VGP_OK;
    add_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a, ld, n, value);
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

// a[i][j] = 0 for j > i
__global__ void __launch_bounds__(256) zero_upper_kernel(double *a, int64_t n, int64_t ld) {
    for (int64_t i = blockIdx.y; i < n; i += gridDim.y)
        for (int64_t j = i + 1 + (int64_t)blockIdx.x * 256 + threadIdx.x; j < n; j += (int64_t)gridDim.x * 256)
            a[i * ld + j] = 0.0;
}

int dense_zero_strict_upper(double *a, int64_t n, int64_t ld, cudaStream_t s) {
    const int64_t gx = (n + 255) / 256 < 64 ? (n + 255) / 256 : 64;
    const int64_t gy = n < 4096 ? n : 4096;
    zero_upper_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, s>>>(a, n, ld);
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

// upper = lower^T, 32x32 tiles through shared memory
__global__ void __launch_bounds__(256) mirror_kernel(double *a, int64_t n, int64_t ld) {
    __shared__ double tile[32][33];
    const int64_t b = blockIdx.x;
    int r = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
    while ((int64_t)r * (r + 1) / 2 > b) --r;
    while ((int64_t)(r + 1) * (r + 2) / 2 <= b) ++r;
    const int ti = r, tj = (int)(b - (int64_t)r * (r + 1) / 2);   // ti >= tj
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int rr = ty; rr < 32; rr += 8) {
        const int64_t i = (int64_t)ti * 32 + rr, j = (int64_t)tj * 32 + tx;
        tile[rr][tx] = (i < n && j < n) ? a[i * ld + j] : 0.0;
    }
    __syncthreads();
    for (int rr = ty; rr < 32; rr += 8) {
        const int64_t i = (int64_t)tj * 32 + rr, j = (int64_t)ti * 32 + tx;   // transposed position
        if (i < n && j < n && j > i) a[i * ld + j] = tile[tx][rr];
    }
}

int dense_mirror_lower(double *a, int64_t n, int64_t ld, cudaStream_t s) {
    const int64_t t = (n + 31) / 32;
    const int64_t blocks = t * (t + 1) / 2;
    VGP_REQUIRE(blocks <= 0x7fffffffLL, "mirror: matrix too large");
    mirror_kernel<<<(unsigned)blocks, 256, 0, s>>>(a, n, ld);
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

// =====================================================================================================
// 128 x 128 diagonal-block kernels (one CTA, block resident in shared memory)
// =====================================================================================================
constexpr int SLD = NB + 1;
constexpr int BLOCK_SMEM = NB * SLD * 8;    // 132 096 B

// The three diagonal-block kernels share one register layout: 256 threads as a 16 x 16 grid, thread (ti, tc)
// owns the 8 x 8 cyclic sub-tile {(ti + 16 r, tc + 16 s)} of the 128 x 128 block in registers.  The cyclic
// distribution keeps triangular work balanced; per elimination step a thread reads 16 values from shared
// memory (broadcast / conflict-free) and issues 64 FMAs.

// Lower Cholesky of one diagonal block, in place; strict upper triangle zeroed.  Right-looking, the pivot
// column travels through a double-buffered 128-entry shared array: one barrier per column.  The column loop is
// split into 8 panels of 16 columns with the panel index a template parameter: which register holds column j,
// which rows / columns lie behind the pivot (and need no update) are then compile-time facts, and the trailing
// update shrinks with the panel -- the first version carried them as run-time selects (374 FSEL per column, 200 us
// per block, issue-bound; ncu launch list of the ELBO step).
__global__ void __launch_bounds__(256, 1) potf2_kernel(double *a, int64_t ld, int *info, int row_offset,
                                                       double pivot_floor) {
    __shared__ double colbuf[2][NB];
    const int ti = threadIdx.x >> 4, tc = threadIdx.x & 15;
    double v[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const int i = ti + 16 * r, c = tc + 16 * s;
            v[r][s] = c <= i ? a[(int64_t)i * ld + c] : 0.0;
        }
    potf2_panel<0>(v, colbuf, ti, tc, info, row_offset, pivot_floor);
    potf2_panel<1>(v, colbuf, ti, tc, info, row_offset, pivot_floor);
    potf2_panel<2>(v, colbuf, ti, tc, info, row_offset, pivot_floor);
    potf2_panel<3>(v, colbuf, ti, tc, info, row_offset, pivot_floor);
    potf2_panel<4>(v, colbuf, ti, tc, info, row_offset, pivot_floor);
    potf2_panel<5>(v, colbuf, ti, tc, info, row_offset, pivot_floor);
    potf2_panel<6>(v, colbuf, ti, tc, info, row_offset, pivot_floor);
    potf2_panel<7>(v, colbuf, ti, tc, info, row_offset, pivot_floor);
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const int i = ti + 16 * r, c = tc + 16 * s;
            a[(int64_t)i * ld + c] = c <= i ? v[r][s] : 0.0;
        }
}

// out = inv(L) for one lower 128-block (strict upper of `out` zeroed); `transpose` writes inv(L)^T.
// Forward substitution on the identity, all 128 right-hand sides at once: after row k of X is final, every
// later row gets the rank-1 correction  B[i][:] -= L[i][k] X[k][:].  L sits in shared memory, B/X in registers.
// Same panel split as potf2: X is lower triangular, so row k only has columns <= k, and only rows > k change.
template <int RK>
__device__ __forceinline__ void trtri_panel(double (&x)[8][8], const double *sm, double (*rowbuf)[NB], int ti,
                                            int tc) {
#pragma unroll 1
    for (int kk = 0; kk < 16; ++kk) {
        const int k = 16 * RK + kk, kb = k & 1;
        if (ti == kk) {
#pragma unroll
            for (int s = 0; s <= RK; ++s) rowbuf[kb][tc + 16 * s] = x[RK][s];
        }
        __syncthreads();
        const double dinv = 1.0 / sm[k * SLD + k];
        double xr[8], lk[8];
#pragma unroll
        for (int s = 0; s <= RK; ++s) xr[s] = rowbuf[kb][tc + 16 * s] * dinv;
#pragma unroll
        for (int r = RK; r < 8; ++r) lk[r] = (r > RK || ti > kk) ? sm[(ti + 16 * r) * SLD + k] : 0.0;    // i > k
#pragma unroll
        for (int r = RK; r < 8; ++r)
#pragma unroll
            for (int s = 0; s <= RK; ++s) x[r][s] = fma(-lk[r], xr[s], x[r][s]);
        if (ti == kk) {
#pragma unroll
            for (int s = 0; s <= RK; ++s) x[RK][s] = xr[s];
        }
    }
}

__global__ void __launch_bounds__(256, 1) trtri_kernel(const double *l, int64_t ldl, double *out, int64_t ldo,
                                                       int transpose) {
    extern __shared__ __align__(16) double sm[];          // L, [128][129]
    __shared__ double rowbuf[2][NB];
    const int ti = threadIdx.x >> 4, tc = threadIdx.x & 15;
    for (int e = threadIdx.x; e < NB * NB; e += 256) {
        const int i = e >> 7, j = e & 127;
        sm[i * SLD + j] = j <= i ? l[(int64_t)i * ldl + j] : 0.0;
    }
    double x[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int s = 0; s < 8; ++s) x[r][s] = (ti + 16 * r == tc + 16 * s) ? 1.0 : 0.0;
    __syncthreads();
    trtri_panel<0>(x, sm, rowbuf, ti, tc);
    trtri_panel<1>(x, sm, rowbuf, ti, tc);
    trtri_panel<2>(x, sm, rowbuf, ti, tc);
    trtri_panel<3>(x, sm, rowbuf, ti, tc);
    trtri_panel<4>(x, sm, rowbuf, ti, tc);
    trtri_panel<5>(x, sm, rowbuf, ti, tc);
    trtri_panel<6>(x, sm, rowbuf, ti, tc);
    trtri_panel<7>(x, sm, rowbuf, ti, tc);
    // `out` may alias `l`: every thread finished reading L from global memory before the first barrier
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const int i = ti + 16 * r, c = tc + 16 * s;
            const double val = c <= i ? x[r][s] : 0.0;
            if (!transpose) {
                out[(int64_t)i * ldo + c] = val;
            } else {
                out[(int64_t)c * ldo + i] = val;        // inv(L)^T: upper triangular, zero strict lower
            }
        }
}

// block = X^T X for one lower 128-block X, written as a full symmetric block.
__global__ void __launch_bounds__(256, 1) lauum_kernel(double *x, int64_t ld) {
    extern __shared__ __align__(16) double sm[];
    const int ti = threadIdx.x >> 4, tc = threadIdx.x & 15;
    for (int e = threadIdx.x; e < NB * NB; e += 256) {
        const int i = e >> 7, j = e & 127;
        sm[i * SLD + j] = j <= i ? x[(int64_t)i * ld + j] : 0.0;
    }
    __syncthreads();
    double p[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int s = 0; s < 8; ++s) p[r][s] = 0.0;
#pragma unroll 2
    for (int k = 0; k < NB; ++k) {
        double xi[8], xc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) xi[r] = sm[k * SLD + ti + 16 * r];
#pragma unroll
        for (int s = 0; s < 8; ++s) xc[s] = sm[k * SLD + tc + 16 * s];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int s = 0; s < 8; ++s) p[r][s] = fma(xi[r], xc[s], p[r][s]);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int s = 0; s < 8; ++s) x[(int64_t)(ti + 16 * r) * ld + tc + 16 * s] = p[r][s];
}

static int block_kernels_configure() {
    static bool configured[64] = {};
    int dev = 0;
    VGP_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
                VGP_CUDA(cudaFuncSetAttribute(trtri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BLOCK_SMEM));
        VGP_CUDA(cudaFuncSetAttribute(lauum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BLOCK_SMEM));
        configured[dev] = true;
    }
    return VGP_OK;
}

static int block_trtri(const double *l, int64_t ldl, double *out, int64_t ldo, int transpose, cudaStream_t s) {
    trtri_kernel<<<1, 256, BLOCK_SMEM, s>>>(l, ldl, out, ldo, transpose);
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

// =====================================================================================================
// workspace
// =====================================================================================================
int DenseWorkspace::ensure(int64_t nblocks) {
    int dev = 0;
    VGP_CUDA(cudaGetDevice(&dev));
    if (winv && device != dev) release();
    device = dev;
    if (!winv) {
        VGP_CUDA(cache_alloc((void **)&winv, (size_t)NB * NB * 8));
        VGP_CUDA(cache_alloc((void **)&info, sizeof(int)));
        VGP_CUDA(cudaMemset(info, 0, sizeof(int)));
        VGP_TRY(block_kernels_configure());
    }
    if (nblocks > dinv_blocks) {
        cache_free(dinv);
        dinv = nullptr;
        dinv_blocks = 0;
        VGP_CUDA(cache_alloc((void **)&dinv, (size_t)nblocks * NB * NB * 8));
        dinv_blocks = nblocks;
    }
    return VGP_OK;
}

void DenseWorkspace::release() {
    cache_free(winv);
    cache_free(info);
    cache_free(dinv);
    winv = nullptr;
    info = nullptr;
    dinv = nullptr;
    emu.release();
    dinv_blocks = 0;
    device = -1;
}

// Pivots at or below this count as "not positive definite" in dense_potrf (thread-local; 0 = only non-positive pivots).
// The one-call placement entries set it to 1e-12 x the largest diagonal entry of cov_vv: below that the Cholesky of
// a covariance is numerically rank deficient and the reference's pinv semantics apply (pinv.cu).
static thread_local double g_pivot_floor = 0.0;
void dense_set_pivot_floor(double floor) { g_pivot_floor = floor > 0.0 ? floor : 0.0; }

static inline int64_t split(int64_t n) {
    int64_t n1 = (n / NB / 2) * NB;
    return n1 < NB ? NB : n1;
}

// Explicit inverse of the 128-block at `l`: from the cache filled by potrf (`dinv`, one [128][128] slab per
// diagonal block) when there is one, else computed into ws.winv.
static int block_inverse(const double *l, int64_t ldl, const double *dinv, DenseWorkspace &ws, cudaStream_t s,
                         const double **w) {
    if (dinv) {
        *w = dinv;
        return VGP_OK;
    }
    *w = ws.winv;
    return block_trtri(l, ldl, ws.winv, NB, 0, s);
}
static inline const double *dinv_at(const double *dinv, int64_t rows) {
    return dinv ? dinv + (rows / NB) * NB * NB : nullptr;
}

// =====================================================================================================
// triangular solves (recursive), all in place on B.  `dinv`: cached inverses of L's diagonal blocks or NULL.
// =====================================================================================================
// X L^T = alpha B,  B [m][n]
static int trsm_right_t(int64_t m, int64_t n, double alpha, const double *l, int64_t ldl, const double *dinv,
                        double *b, int64_t ldb, DenseWorkspace &ws, cudaStream_t s) {
    if (n == NB) {
        const double *w;
        VGP_TRY(block_inverse(l, ldl, dinv, ws, s, &w));
        return dense_gemm(0, 1, m, NB, NB, alpha, b, ldb, w, NB, 0.0, b, ldb, GEMM_FULL, s);
    }
    const int64_t n1 = split(n), n2 = n - n1;
    VGP_TRY(trsm_right_t(m, n1, alpha, l, ldl, dinv, b, ldb, ws, s));
    // B2 = alpha B2 - X1 Lb^T
    VGP_TRY(dense_gemm(0, 1, m, n2, n1, -1.0, b, ldb, l + n1 * ldl, ldl, alpha, b + n1, ldb, GEMM_FULL, s));
    return trsm_right_t(m, n2, 1.0, l + n1 * ldl + n1, ldl, dinv_at(dinv, n1), b + n1, ldb, ws, s);
}

// X L = alpha B,  B [m][n]
static int trsm_right_n(int64_t m, int64_t n, double alpha, const double *l, int64_t ldl, const double *dinv,
                        double *b, int64_t ldb, DenseWorkspace &ws, cudaStream_t s) {
    if (n == NB) {
        const double *w;
        VGP_TRY(block_inverse(l, ldl, dinv, ws, s, &w));
        return dense_gemm(0, 0, m, NB, NB, alpha, b, ldb, w, NB, 0.0, b, ldb, GEMM_FULL, s);
    }
    const int64_t n1 = split(n), n2 = n - n1;
    VGP_TRY(trsm_right_n(m, n2, alpha, l + n1 * ldl + n1, ldl, dinv_at(dinv, n1), b + n1, ldb, ws, s));
    // B1 = alpha B1 - X2 Lb
    VGP_TRY(dense_gemm(0, 0, m, n1, n2, -1.0, b + n1, ldb, l + n1 * ldl, ldl, alpha, b, ldb, GEMM_FULL, s));
    return trsm_right_n(m, n1, 1.0, l, ldl, dinv, b, ldb, ws, s);
}

// L X = alpha B,  B [n][nrhs]
static int trsm_left_n(int64_t n, int64_t nrhs, double alpha, const double *l, int64_t ldl, const double *dinv,
                       double *b, int64_t ldb, DenseWorkspace &ws, cudaStream_t s) {
    if (n == NB) {
        const double *w;
        VGP_TRY(block_inverse(l, ldl, dinv, ws, s, &w));
        return dense_gemm(0, 0, NB, nrhs, NB, alpha, w, NB, b, ldb, 0.0, b, ldb, GEMM_FULL, s);
    }
    const int64_t n1 = split(n), n2 = n - n1;
    VGP_TRY(trsm_left_n(n1, nrhs, alpha, l, ldl, dinv, b, ldb, ws, s));
    // B2 = alpha B2 - Lb X1
    VGP_TRY(dense_gemm(0, 0, n2, nrhs, n1, -1.0, l + n1 * ldl, ldl, b, ldb, alpha, b + n1 * ldb, ldb, GEMM_FULL, s));
    return trsm_left_n(n2, nrhs, 1.0, l + n1 * ldl + n1, ldl, dinv_at(dinv, n1), b + n1 * ldb, ldb, ws, s);
}

// L^T X = alpha B,  B [n][nrhs]
static int trsm_left_t(int64_t n, int64_t nrhs, double alpha, const double *l, int64_t ldl, const double *dinv,
                       double *b, int64_t ldb, DenseWorkspace &ws, cudaStream_t s) {
    if (n == NB) {
        const double *w;
        VGP_TRY(block_inverse(l, ldl, dinv, ws, s, &w));
        return dense_gemm(1, 0, NB, nrhs, NB, alpha, w, NB, b, ldb, 0.0, b, ldb, GEMM_FULL, s);
    }
    const int64_t n1 = split(n), n2 = n - n1;
    VGP_TRY(trsm_left_t(n2, nrhs, alpha, l + n1 * ldl + n1, ldl, dinv_at(dinv, n1), b + n1 * ldb, ldb, ws, s));
    // B1 = alpha B1 - Lb^T X2
    VGP_TRY(dense_gemm(1, 0, n1, nrhs, n2, -1.0, l + n1 * ldl, ldl, b + n1 * ldb, ldb, alpha, b, ldb, GEMM_FULL, s));
    return trsm_left_t(n1, nrhs, 1.0, l, ldl, dinv, b, ldb, ws, s);
}

int dense_trsm(int side, int trans, int64_t n, int64_t nrhs, double alpha, const double *l, int64_t ldl,
               double *b, int64_t ldb, DenseWorkspace &ws, bool use_cached_inverses, cudaStream_t s) {
    if (n == 0 || nrhs == 0) return VGP_OK;
    VGP_REQUIRE(n % NB == 0 && nrhs % NB == 0, "dense_trsm: unpadded size n=%lld nrhs=%lld", (long long)n,
                (long long)nrhs);
    VGP_TRY(ws.ensure(0));
    const double *dinv = nullptr;
    if (use_cached_inverses) {
        VGP_REQUIRE(ws.dinv && ws.dinv_blocks >= n / NB, "dense_trsm: no cached diagonal-block inverses");
        dinv = ws.dinv;
    }
    if (side == 0) return trans ? trsm_left_t(n, nrhs, alpha, l, ldl, dinv, b, ldb, ws, s)
                                : trsm_left_n(n, nrhs, alpha, l, ldl, dinv, b, ldb, ws, s);
    return trans ? trsm_right_t(nrhs, n, alpha, l, ldl, dinv, b, ldb, ws, s)
                 : trsm_right_n(nrhs, n, alpha, l, ldl, dinv, b, ldb, ws, s);
}

// =====================================================================================================
// Cholesky, inverse of the factor, X^T X
// =====================================================================================================
static int potrf_rec(double *a, int64_t n, int64_t ld, int64_t row_offset, double *dinv, DenseWorkspace &ws,
                     cudaStream_t s) {
    if (n == NB) {
        VGP_TRY(gate_wait(a, NB, s));
        potf2_kernel<<<1, 256, 0, s>>>(a, ld, ws.info, (int)row_offset, g_pivot_floor);
        VGP_LAUNCH_CHECK();
        return block_trtri(a, ld, dinv, NB, 0, s);          // cache inv(L_ii) for every later solve
    }
    const int64_t n1 = split(n), n2 = n - n1;
    double *a21 = a + n1 * ld, *a22 = a21 + n1;
    double *dinv2 = dinv + (n1 / NB) * NB * NB;
    VGP_TRY(potrf_rec(a, n1, ld, row_offset, dinv, ws, s));
    VGP_TRY(trsm_right_t(n2, n1, 1.0, a, ld, dinv, a21, ld, ws, s));                               // A21 <- A21 L11^-T
    VGP_TRY(dense_gemm(0, 1, n2, n2, n1, -1.0, a21, ld, a21, ld, 1.0, a22, ld, GEMM_LOWER, s));    // A22 -= A21 A21^T
    VGP_TRY(potrf_rec(a22, n2, ld, row_offset + n1, dinv2, ws, s));
    return VGP_OK;
}

int dense_potrf(double *a, int64_t n, int64_t ld, DenseWorkspace &ws, cudaStream_t s) {
    EmulateScope emulate_large_products(ws);
    VGP_REQUIRE(n > 0 && n % NB == 0 && ld >= n && ld % 2 == 0, "dense_potrf: unpadded size %lld (ld %lld)",
                (long long)n, (long long)ld);
    VGP_TRY(ws.ensure(n / NB));
    if (emulate_slices() >= 2 && n >= option(VGP_OPT_GEMM_EMULATE_MIN)) VGP_TRY(ws.emu.reserve(largest_half(n), emulate_slices(), s));
    VGP_CUDA(cudaMemsetAsync(ws.info, 0, sizeof(int), s));
    return potrf_rec(a, n, ld, 0, ws.dinv, ws, s);
}

// copy a cached [128][128] inverse into the matrix (it already has a zero strict upper triangle)
__global__ void __launch_bounds__(256) block_copy_kernel(const double *src, double *dst, int64_t ldd) {
    for (int e = blockIdx.x * 256 + threadIdx.x; e < NB * NB; e += gridDim.x * 256)
        dst[(int64_t)(e >> 7) * ldd + (e & 127)] = src[e];
}

static int trtri_rec(double *l, int64_t n, int64_t ld, const double *dinv, DenseWorkspace &ws, cudaStream_t s) {
    if (n == NB) {
        block_copy_kernel<<<16, 256, 0, s>>>(dinv, l, ld);
        VGP_LAUNCH_CHECK();
        return VGP_OK;
    }
    const int64_t n1 = split(n), n2 = n - n1;
    double *lb = l + n1 * ld, *lc = lb + n1;
    VGP_TRY(trsm_right_n(n2, n1, 1.0, l, ld, dinv, lb, ld, ws, s));                     // Lb <- Lb La^-1
    VGP_TRY(trsm_left_n(n2, n1, -1.0, lc, ld, dinv_at(dinv, n1), lb, ld, ws, s));       // Lb <- -Lc^-1 Lb
    VGP_TRY(trtri_rec(l, n1, ld, dinv, ws, s));
    return trtri_rec(lc, n2, ld, dinv_at(dinv, n1), ws, s);
}

// Needs the diagonal-block inverses cached by dense_potrf on the same workspace.
int dense_trtri(double *l, int64_t n, int64_t ld, DenseWorkspace &ws, cudaStream_t s) {
    EmulateScope emulate_large_products(ws);
    VGP_REQUIRE(n > 0 && n % NB == 0, "dense_trtri: unpadded size");
    VGP_REQUIRE(ws.dinv && ws.dinv_blocks >= n / NB, "dense_trtri: call dense_potrf on this workspace first");
    return trtri_rec(l, n, ld, ws.dinv, ws, s);
}

// B <- X^T B for lower X [n][n] (diagonal 128-blocks have a zero strict upper triangle), B [n][ncols]
static int trmm_left_t(int64_t n, int64_t ncols, const double *x, int64_t ldx, double *b, int64_t ldb,
                       cudaStream_t s) {
    if (n == NB) return dense_gemm(1, 0, NB, ncols, NB, 1.0, x, ldx, b, ldb, 0.0, b, ldb, GEMM_FULL, s);
    const int64_t n1 = split(n), n2 = n - n1;
    VGP_TRY(trmm_left_t(n1, ncols, x, ldx, b, ldb, s));
    // B1 += Xb^T B2
    VGP_TRY(dense_gemm(1, 0, n1, ncols, n2, 1.0, x + n1 * ldx, ldx, b + n1 * ldb, ldb, 1.0, b, ldb, GEMM_FULL, s));
    return trmm_left_t(n2, ncols, x + n1 * ldx + n1, ldx, b + n1 * ldb, ldb, s);
}

static int lauum_rec(double *x, int64_t n, int64_t ld, cudaStream_t s) {
    if (n == NB) {
        lauum_kernel<<<1, 256, BLOCK_SMEM, s>>>(x, ld);
        VGP_LAUNCH_CHECK();
        return VGP_OK;
    }
    const int64_t n1 = split(n), n2 = n - n1;
    double *xb = x + n1 * ld, *xc = xb + n1;
    VGP_TRY(lauum_rec(x, n1, ld, s));                                                           // P11 = Xa^T Xa
    VGP_TRY(dense_gemm(1, 0, n1, n1, n2, 1.0, xb, ld, xb, ld, 1.0, x, ld, GEMM_LOWER, s));      // P11 += Xb^T Xb
    VGP_TRY(trmm_left_t(n2, n1, xc, ld, xb, ld, s));                                            // P21 = Xc^T Xb
    return lauum_rec(xc, n2, ld, s);
}

int dense_lauum(double *x, int64_t n, int64_t ld, DenseWorkspace &ws, cudaStream_t s) {
    EmulateScope emulate_large_products(ws);
    VGP_REQUIRE(n > 0 && n % NB == 0, "dense_lauum: unpadded size");
    VGP_TRY(ws.ensure(0));
    return lauum_rec(x, n, ld, s);
}

// Load every kernel of this file now (CUDA loads modules lazily, and a lazy load can wait on kernels that are
// spinning on a peer's flag) and set the opt-in shared-memory sizes.
template <class C, bool AKC, bool BKC>
static int gemm_preload_cfg() {
    cudaFuncAttributes fa;
    VGP_CUDA(cudaFuncSetAttribute(gemm_kernel<C, AKC, BKC>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    VGP_CUDA(cudaFuncGetAttributes(&fa, gemm_kernel<C, AKC, BKC>));
    return VGP_OK;
}
template <bool AKC, bool BKC>
static int gemm_preload_one() {
    VGP_TRY((gemm_preload_cfg<CfgBase, AKC, BKC>()));
    return gemm_preload_cfg<CfgPair, AKC, BKC>();
}
template <class S, bool AKC, bool BKC>
static int tma_preload() {
    cudaFuncAttributes fa;
    VGP_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<S, AKC, BKC>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM));
    VGP_CUDA(cudaFuncGetAttributes(&fa, gemm_tma_kernel<S, AKC, BKC>));
    return VGP_OK;
}
int dense_preload() {
    cudaFuncAttributes fa;
    VGP_TRY((gemm_preload_one<true, true>()));
    VGP_TRY((gemm_preload_one<true, false>()));
    VGP_TRY((gemm_preload_one<false, false>()));
    VGP_TRY((gemm_preload_one<false, true>()));
    VGP_CUDA(cudaFuncSetAttribute(trtri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BLOCK_SMEM));
    VGP_CUDA(cudaFuncSetAttribute(lauum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BLOCK_SMEM));
    VGP_TRY((tma_preload<tma::Big, true, true>()));
    VGP_TRY((tma_preload<tma::Big, true, false>()));
    VGP_TRY((tma_preload<tma::Big, false, true>()));
    VGP_TRY((tma_preload<tma::Big, false, false>()));
    VGP_TRY((tma_preload<tma::Small, true, true>()));
    VGP_TRY((tma_preload<tma::Small, true, false>()));
    VGP_TRY((tma_preload<tma::Small, false, true>()));
    VGP_TRY((tma_preload<tma::Small, false, false>()));
    VGP_CUDA(cudaFuncGetAttributes(&fa, potf2_kernel));
    VGP_CUDA(cudaFuncGetAttributes(&fa, trtri_kernel));
    VGP_CUDA(cudaFuncGetAttributes(&fa, lauum_kernel));
    VGP_CUDA(cudaFuncGetAttributes(&fa, block_copy_kernel));
    VGP_CUDA(cudaFuncGetAttributes(&fa, mirror_kernel));
    VGP_CUDA(cudaFuncGetAttributes(&fa, zero_upper_kernel));
    VGP_CUDA(cudaFuncGetAttributes(&fa, add_diag_kernel));
    VGP_CUDA(cudaFuncGetAttributes(&fa, splitk_reduce_kernel));
    VGP_CUDA(cudaFuncGetAttributes(&fa, dist_barrier_kernel));
    if (option(VGP_OPT_GEMM_EMULATE_SLICES) >= 2) VGP_TRY(emulated_preload());
    return VGP_OK;
}

int dense_read_info(DenseWorkspace &ws, int *info_host, cudaStream_t s) {
    int info = 0;
    VGP_CUDA(cudaMemcpyAsync(&info, ws.info, sizeof(int), cudaMemcpyDeviceToHost, s));
    VGP_CUDA(cudaStreamSynchronize(s));
    if (info_host) *info_host = info;
    if (info != 0) {
        set_error("matrix is not positive definite: non-positive (or numerically zero) pivot at row %d "
                  "(the reference takes a pseudo-inverse here: vgp_placement_host_pinv; this path needs an SPD input)",
                  info - 1);
        return VGP_ERR_NOT_PD;
    }
    return VGP_OK;
}

int dense_spd_inverse(double *a, int64_t n, int64_t ld, DenseWorkspace &ws, int *info_host, cudaStream_t s) {
    VGP_TRY(dense_potrf(a, n, ld, ws, s));
    VGP_TRY(dense_read_info(ws, info_host, s));
    VGP_TRY(dense_trtri(a, n, ld, ws, s));
    VGP_TRY(dense_lauum(a, n, ld, ws, s));
    return dense_mirror_lower(a, n, ld, s);
}

}  // namespace vgp
