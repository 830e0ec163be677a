// FP64 GEMM on the int8 tensor cores (tcgen05.mma kind::i8, int32 accumulators in TMEM): the large products of the
// factorisations behind the placement path (potrf / trtri / lauum: SURVEY.md section 8d, "Seed inverse").
//
// Those products are bound by the FP64 tensor pipe, which tops out at ~37 TFLOP/s on B200 (DMMA, 97 % pipe-active in
// dense.cu's kernel: profiles/r01_ncu_gemm_summary.txt); tcgen05 has no f64 kind, but its int8 path is two orders of
// magnitude wider.  Ozaki's error-free splitting carries an FP64 product on it: every operand row is scaled by a power
// of two to |x| < 1 and cut into s signed digits of W = 7 bits,
//     A[i][:] = 2^ea[i] * sum_t QA_t[i][:] 2^(-W (t + 1)),      QA_t in [-127, 127],
// (columns of B likewise), the digit planes are multiplied exactly in int32,
//     C = 2^(ea[i] + eb[j]) * sum_{g < s} 2^(-W (g + 2)) * sum_{t <= g} QA_t QB_{g-t}^T,
// and the s group sums are combined in FP64, smallest group first.  s = 8 (36 integer products): within 1e-14 of the
// exact product (native FP64: 1.5e-15); the result does not depend on tile shape, k split or rank count (integer sums).
//
// Layout: digit planes Q[t][rows_pad][k_pad] int8, k-contiguous, fetched by TMA (SWIZZLE_64B boxes of 64 k).  One CTA per
// 128 x 64 tile of C keeps the s group accumulators side by side in tensor memory (s * 64 <= 512 columns) and, per
// 64-byte k block, loads the s plane tiles of A and of B once (16 TMA boxes) to feed all s (s + 1) / 2 plane pairs:
// warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2..9 = epilogue (TMEM -> registers, FP64
// recombination, 2^(ea + eb), alpha / beta).  int32 cannot overflow while s * k * 127^2 < 2^31: the host splits k into
// chunks of 8192.
//
// Measured on B200 (profiles/r02_first_contact_experimental_paths.md): 59 TFLOP/s FP64-equivalent at 8192^3 against 36
// for cuBLAS DGEMM; an earlier group-by-group kernel (128 x 128 tiles, one accumulator) reached 47 and was removed.
#include <cuda.h>

#include "dense.cuh"

namespace vgp {
namespace emu {

constexpr int W = 7;
constexpr int S_MAX = 8;
constexpr int BM = 128, BN = 128, BKB = 128;          // padding granularity of the digit planes (rows, rows, k bytes)
constexpr int UMMA_K = 32;                            // k per tcgen05.mma.kind::i8
constexpr int EPI_WARPS = 8, THREADS = 64 + 32 * EPI_WARPS;
constexpr int64_t K_CHUNK = 8192;

struct GemmArgs {
    int64_t m, n;                 // real extent of C (stores are guarded)
    int64_t m_pad, n_pad;         // rows per digit plane
    int kblocks;                  // k_pad / BKB
    int kb_split;                 // split-K: 64-byte k blocks per blockIdx.z (0: no split); each split writes its own
    int64_t split_stride;         // partial C at c + blockIdx.z * split_stride (beta must be 0 then)
    int s;
    const int *ea, *eb;           // [m_pad], [n_pad]
    double *c;
    int64_t ldc;
    double alpha, beta;
    int lower;                    // skip tiles strictly above the diagonal
    // distributed form (dist.cu): tiles dealt round-robin to the ranks (linear tile = rank + dist_n * blockIdx.x), every
    // finished tile stored into all dist_n replicas (delta[q] = replica q minus the local one, in doubles)
    int dist_n, rank, tiles_n;
    int local_only;               // distributed form: store the tile into this rank's replica only (push_tiles follows)
    int64_t delta[DIST_MAX];
};

// Tile coordinates of this CTA; false when it has no tile.  CTAs are dispatched in linear order, and a wave of 148 of them
// should share operand strips in L2: tiles are walked in super-columns of SUPER tile columns, rows outermost inside one
// (16 x ~9 tiles per wave: ~150 MB of digit planes per wave instead of one strip of A and ALL of B -- the row-major walk
// read 34.9 GB from DRAM for 1.07 GB of planes at 8192^3, L2 hit rate 61 %: profiles/r02_ncu_emulated_gemm.txt).
constexpr int SUPER = 16;
constexpr int v2_BN = 64;                             // tile width of the GEMM kernel (v2::BN2, declared below)
__device__ __forceinline__ bool tile_of_cta(const GemmArgs &p, int tiles_m, int &tm, int &tn) {
    const int64_t lin = p.dist_n > 0 ? (int64_t)p.rank + (int64_t)p.dist_n * blockIdx.x
                                     : (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
    if (lin >= (int64_t)tiles_m * p.tiles_n) return false;
    const int64_t per_super = (int64_t)SUPER * tiles_m;
    const int sc = (int)(lin / per_super);
    const int64_t rem = lin % per_super;
    const int width = p.tiles_n - sc * SUPER < SUPER ? p.tiles_n - sc * SUPER : SUPER;
    tm = (int)(rem / width);
    tn = sc * SUPER + (int)(rem % width);
    return true;
}
__device__ __forceinline__ void store_one(const GemmArgs &p, double *dst, double o) {
    if (p.dist_n > 0 && !p.local_only) {
        for (int q = 0; q < p.dist_n; ++q) dst[p.delta[q]] = o;
    } else {
        *dst = o;
    }
}

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
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
        if (clock64() - t0 > 8000000000LL) __trap();      // ~4 s: fail the launch instead of hanging the device
    }
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *map, int c0, int c1, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile(
        "{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
// ROW_BYTES = 128: SWIZZLE_128B (layout code 2), 64: SWIZZLE_64B (code 4); 8-row groups are 8 * ROW_BYTES apart
template <int ROW_BYTES>
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr) {
    static_assert(ROW_BYTES == 128 || ROW_BYTES == 64, "swizzle span");
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3fff);            // start address, 16-byte units
    d |= (unsigned long long)1 << 16;                                // leading byte offset: unused for swizzled k-major
    d |= (unsigned long long)((8 * ROW_BYTES) >> 4) << 32;           // stride byte offset between 8-row groups
    d |= (unsigned long long)1 << 46;                                // descriptor version (sm_100)
    d |= (unsigned long long)(ROW_BYTES == 128 ? 2 : 4) << 61;
    return d;
}
// instruction descriptor, kind::i8: D = s32, A = B = signed 8 bit, both k-major, N x M
__host__ __device__ constexpr unsigned idesc_i8(int m, int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24);
}

// x * 2^e, exact like scalbn: the power of two is built from its bit pattern (one multiply instead of a libm call per
// output element; the epilogue of a k = 512 tile was longer than its MMAs)
__device__ __forceinline__ double mul_pow2(double x, int e) {
    if (e >= -1022 && e <= 1023) return x * __longlong_as_double((long long)(e + 1023) << 52);
    return scalbn(x, e);
}

__device__ __forceinline__ void umma_i8(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc,
                                        unsigned accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {      // arrives on `bar` once all MMAs issued so far are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned addr, int (&v)[32]) {   // this warp's 32 lanes x 32 columns
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(addr)
        : "memory");
}

// ------------------------------------------------------------------------------------------------ the GEMM
// All digit planes of a k block resident: 96 KB per stage feed 72 MMAs (42 bytes of shared-memory fill per MMA cycle).
namespace v2 {
constexpr int BN2 = 64, BKB2 = 64, STAGES2 = 2, S2_MAX = 8;
static_assert(BN2 == v2_BN, "push_tiles_kernel walks the same tiles");
constexpr int A2 = BM * BKB2, B2 = BN2 * BKB2;                    // one plane tile of A / B
constexpr int STAGE2 = S2_MAX * (A2 + B2);
constexpr int BAR2 = (2 * STAGES2 + 1) * 8 + 8;
constexpr int SMEM2 = STAGES2 * STAGE2 + 1024 + BAR2 + BN2 * 4;

__global__ void __launch_bounds__(THREADS, 1)
    emu_gemm_resident_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GemmArgs p) {
    extern __shared__ unsigned char raw[];
    const unsigned raw_addr = smem_u32(raw);
    unsigned char *sm = raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + STAGES2 * STAGE2);
    unsigned long long *full = bars, *empty = bars + STAGES2, *tfull = bars + 2 * STAGES2;
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(bars + 2 * STAGES2 + 1);
    int *eb_tile = reinterpret_cast<int *>(sm + STAGES2 * STAGE2 + BAR2);

    int tm, tn;
    if (!tile_of_cta(p, (int)(p.m_pad / BM), tm, tn)) return;
    if (p.lower && tn * BN2 > tm * BM + (BM - 1)) return;           // wholly above the diagonal
    const int64_t m0 = (int64_t)tm * BM, n0 = (int64_t)tn * BN2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.s, KB_ALL = p.kblocks * (BKB / BKB2);
    // split-K (blockIdx.z): this CTA's k blocks are [KB0, KB0 + KB)
    const int KB0 = p.kb_split > 0 ? (int)blockIdx.z * p.kb_split : 0;
    const int KB = p.kb_split > 0 ? (KB_ALL - KB0 < p.kb_split ? KB_ALL - KB0 : p.kb_split) : KB_ALL;
    if (KB <= 0) return;

    if (tid == 0) {
        for (int i = 0; i < STAGES2; ++i) {
            mbar_init(smem_u32(&full[i]), 1);
            mbar_init(smem_u32(&empty[i]), 1);
        }
        mbar_init(smem_u32(tfull), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid >= 64 && tid < 64 + BN2) eb_tile[tid - 64] = p.eb[n0 + tid - 64];
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            for (int kb = 0; kb < KB; ++kb) {
                const int stage = kb % STAGES2;
                if (kb >= STAGES2) mbar_wait(smem_u32(&empty[stage]), ((kb / STAGES2) - 1) & 1);
                const unsigned bar = smem_u32(&full[stage]);
                const unsigned dst = smem_u32(sm + stage * STAGE2);
                mbar_expect_tx(bar, (unsigned)(S * (A2 + B2)));
                for (int t = 0; t < S; ++t) {
                    tma_load_2d(dst + t * A2, &map_a, (KB0 + kb) * BKB2, (int)((int64_t)t * p.m_pad + m0), bar);
                    tma_load_2d(dst + S2_MAX * A2 + t * B2, &map_b, (KB0 + kb) * BKB2, (int)((int64_t)t * p.n_pad + n0), bar);
                }
            }
        }
    } else if (warp == 1) {
        for (int kb = 0; kb < KB; ++kb) {
            const int stage = kb % STAGES2;
            mbar_wait(smem_u32(&full[stage]), (kb / STAGES2) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (elect_one()) {
                const unsigned a_addr = smem_u32(sm + stage * STAGE2), b_addr = a_addr + S2_MAX * A2;
                // Plane t of A against planes u = 0 .. S-1-t of B: their accumulators (groups t + u) are NEIGHBOURS in tensor
                // memory and the B planes are neighbours in shared memory (64 rows of 64 bytes each, 8-row groups 512 bytes
                // apart throughout), so up to four of them form ONE operand of N = 256: 12 wide MMAs per k step instead of
                // 36 narrow ones, and A_t is read from shared memory once per wide MMA -- 104 bytes per MMA cycle instead of
                // 192 against a 128-byte port (the N = 64 form kept the tensor pipe 55 % busy: profiles/r02_ncu_emulated_gemm.txt).
                for (int t = 0; t < S; ++t)
                    for (int u0 = 0; t + u0 < S; u0 += 4) {
                        const int cnt = S - t - u0 < 4 ? S - t - u0 : 4;
                        const unsigned idesc = idesc_i8(BM, cnt * BN2);
#pragma unroll
                        for (int k = 0; k < BKB2 / UMMA_K; ++k)
                            umma_i8(tmem_base + (unsigned)((t + u0) * BN2), umma_desc<64>(a_addr + t * A2 + k * UMMA_K),
                                    umma_desc<64>(b_addr + u0 * B2 + k * UMMA_K), idesc,
                                    (kb == 0 && t == 0 && k == 0) ? 0u : 1u);      // t = 0 touches every group first
                    }
                umma_commit(smem_u32(&empty[stage]));
                if (kb == KB - 1) umma_commit(smem_u32(tfull));
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;            // TMEM lane quarter; 32-column half of the tile
        const int row = q * 32 + lane;
        double acc[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = 0.0;
        mbar_wait(smem_u32(tfull), 0);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        for (int g = S - 1; g >= 0; --g) {
            const double scale = scalbn(1.0, -W * (g + 2));
            int v[32];
            tmem_ld32(tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(g * BN2 + half * 32), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] = fma((double)v[c], scale, acc[c]);
        }
        // Every MMA has completed (tfull), so the operand stages are free: each warp turns its 32 x 32 block (one ROW per
        // thread out of tensor memory) through 8.4 KB of them and stores one row segment of 32 consecutive doubles per
        // instruction -- 256 contiguous bytes instead of 32 scattered 16-byte pieces.  That is what the replicas on the
        // other GPUs need: the scattered form turned every piece into its own NVLink packet, and the distributed
        // factorisation took 19 s on 8 GPUs (1.0 s on 2) instead of well under a second.
        const int64_t i = m0 + row;
        const int ea = i < p.m_pad ? p.ea[i] : 0;
        double *buf = reinterpret_cast<double *>(sm) + (warp - 2) * (32 * 33);
#pragma unroll
        for (int c = 0; c < 32; ++c) buf[lane * 33 + c] = p.alpha * mul_pow2(acc[c], ea + eb_tile[half * 32 + c]);
        __syncwarp();
        const int64_t j = n0 + half * 32 + lane;                    // this lane's column
        double *cbase = p.c + (int64_t)blockIdx.z * p.split_stride;
        for (int r = 0; r < 32; ++r) {
            const int64_t ir = m0 + q * 32 + r;
            if (ir >= p.m) break;
            if (j >= p.n || (p.lower && j / BM > ir / BM)) continue;    // keep to the 128 x 128 tiles on or below the diagonal
            double o = buf[r * 33 + lane];
            double *dst = cbase + ir * p.ldc + j;
            if (p.beta != 0.0) o = fma(p.beta, *dst, o);
            store_one(p, dst, o);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (p.dist_n > 0 && tid == 0) __threadfence_system();      // the tile's stores into the peers' replicas (cumulative)
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(512) : "memory");
    }
}
}  // namespace v2

// Distributed form: the tiles this rank has just computed (same tile walk as the GEMM kernel), copied from its replica
// into every other replica.  blockIdx.y = peer.  Done by a kernel of its own -- thousands of small CTAs keep NVLink busy
// -- because the GEMM kernel runs one CTA per SM with a serial epilogue: storing every tile to seven peers from there
// left the tensor pipe waiting on remote stores (distributed potrf + trtri + lauum at n = 50k on 8 GPUs: 3.8 s; with the
// scattered 16-byte stores of the first epilogue: 19 s).
__global__ void __launch_bounds__(256) push_tiles_kernel(GemmArgs p) {
    int tm, tn;
    if (!tile_of_cta(p, (int)(p.m_pad / BM), tm, tn)) return;
    if (p.lower && tn * v2_BN > tm * BM + (BM - 1)) return;
    int q = blockIdx.y;
    if (q >= p.rank) ++q;                                           // skip this rank
    const int64_t m0 = (int64_t)tm * BM, n0 = (int64_t)tn * v2_BN;
    // 128 rows x 64 doubles: a row is 32 double2, 8 rows per pass
    const int cpair = threadIdx.x & 31, rsub = threadIdx.x >> 5;
    const int64_t j = n0 + 2 * cpair;
    for (int r = rsub; r < BM && j < p.n; r += 8) {
        const int64_t i = m0 + r;
        if (i >= p.m) break;
        if (p.lower && j / BM > i / BM) continue;
        const double *src = p.c + i * p.ldc + j;
        double *dst = const_cast<double *>(src) + p.delta[q];
        if (j + 1 < p.n)
            *reinterpret_cast<double2 *>(dst) = *reinterpret_cast<const double2 *>(src);
        else
            *dst = *src;
    }
    __syncthreads();
    if (threadIdx.x == 0) __threadfence_system();
}

// ------------------------------------------------------------------------------------------------ digit planes
// Element (r, k) of the operand lives at x[r * rs + k * ks] (one of rs, ks is 1).
// e[r] = exponent with |x[r][:]| 2^-e < 1 (0 for an all-zero row).  One CTA per 32 rows.
__global__ void __launch_bounds__(256) row_exponent_kernel(const double *__restrict__ x, int64_t rs, int64_t ks, int64_t rows,
                                                           int64_t k, int64_t rows_pad, int *__restrict__ e) {
    __shared__ double part[8][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    if (ks == 1) {                                   // k-contiguous: a warp walks along k, four rows per warp
        for (int j = 0; j < 4; ++j) {
            const int64_t r = r0 + warp * 4 + j;
            double mx = 0.0;
            if (r < rows)
                for (int64_t c = lane; c < k; c += 32) mx = fmax(mx, fabs(x[r * rs + c]));
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) part[0][warp * 4 + j] = mx;
        }
        __syncthreads();
    } else {                                         // row-contiguous: lanes along the rows, warps interleaved along k
        const int64_t r = r0 + lane;
        double mx = 0.0;
        if (r < rows)
            for (int64_t c = warp; c < k; c += 8) mx = fmax(mx, fabs(x[r * rs + c * ks]));
        part[warp][lane] = mx;
        __syncthreads();
        if (warp == 0) {
            for (int w = 1; w < 8; ++w) mx = fmax(mx, part[w][lane]);
            part[0][lane] = mx;
        }
        __syncthreads();
    }
    if (threadIdx.x < 32) {
        const int64_t r = r0 + threadIdx.x;
        if (r < rows_pad) {
            const double mx = part[0][threadIdx.x];
            e[r] = (r < rows && mx > 0.0) ? ilogb(mx) + 1 : 0;
        }
    }
}

// q[t][r][c] (t < s, planes of rows_pad x k_pad bytes) = digit t of x[r][c] 2^-e[r]; zeros in the padding.
// One CTA per 32 rows x 128 k: the tile is read coalesced in either operand layout, every thread cuts its 16 values into
// digits, the digits go through shared memory (the tile's own storage, reused) so that every plane leaves as 16-byte
// stores along k -- the first version stored single bytes (3.0 TB/s of combined traffic on the 0.82 GB K_zx of the ELBO
// step: 0.53 ms per operand).
__global__ void __launch_bounds__(256) digit_planes_kernel(const double *__restrict__ x, int64_t rs, int64_t ks, int64_t rows,
                                                           int64_t k, int64_t rows_pad, int64_t k_pad,
                                                           const int *__restrict__ e, int s, signed char *__restrict__ q) {
    __shared__ __align__(16) double tile[32][129];                 // 33 024 bytes; later [8 planes][32 rows][128] bytes
    const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 128;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (ks == 1) {
        for (int rr = warp; rr < 32; rr += 8)
            for (int cc = lane; cc < 128; cc += 32) {
                const int64_t r = r0 + rr, c = c0 + cc;
                tile[rr][cc] = (r < rows && c < k) ? x[r * rs + c] : 0.0;
            }
    } else {
        for (int cc = warp; cc < 128; cc += 8) {
            const int64_t r = r0 + lane, c = c0 + cc;
            tile[lane][cc] = (r < rows && c < k) ? x[r * rs + c * ks] : 0.0;
        }
    }
    __syncthreads();
    double v[4][4];                                                // rows warp + 8 a, columns lane + 32 b
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t r = r0 + warp + 8 * a;
        const int er = r < rows_pad ? e[r] : 0;
        const bool plain = er > -1000 && er < 1000;
        const double scale = plain ? scalbn(1.0, -er) : 1.0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double t = tile[warp + 8 * a][lane + 32 * b];
            v[a][b] = plain ? t * scale : scalbn(t, -er);          // |v| < 1, exact
        }
    }
    __syncthreads();
    signed char *qt = reinterpret_cast<signed char *>(&tile[0][0]);
    for (int t = 0; t < s; ++t)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                double w = v[a][b] * 128.0;
                const int d = __double2int_rz(w);                  // |d| <= 127
                v[a][b] = w - (double)d;
                qt[(t * 32 + warp + 8 * a) * 128 + lane + 32 * b] = (signed char)d;
            }
    __syncthreads();
    const size_t plane = (size_t)rows_pad * (size_t)k_pad;
    const uint4 *src = reinterpret_cast<const uint4 *>(qt);
    for (int idx = threadIdx.x; idx < s * 256; idx += 256) {       // 8 uint4 per row, 32 rows per plane
        const int t = idx >> 8, rr = (idx >> 3) & 31, seg = idx & 7;
        const int64_t r = r0 + rr;
        if (r < rows_pad)
            *reinterpret_cast<uint4 *>(q + (size_t)t * plane + (size_t)r * (size_t)k_pad + (size_t)(c0 + seg * 16)) = src[idx];
    }
}

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn encoder() {
    static TensorMapEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = (TensorMapEncodeFn)ptr;
    }
    return fn;
}

static int plane_map(CUtensorMap *map, const signed char *base, int64_t rows_total, int64_t k_pad, int box_k, int box_rows) {
    TensorMapEncodeFn enc = encoder();
    VGP_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)k_pad, (cuuint64_t)rows_total};
    const cuuint64_t strides[1] = {(cuuint64_t)k_pad};
    const cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<signed char *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VGP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return VGP_OK;
}

__global__ void constant_exponent_kernel(int *e, int64_t rows, int64_t rows_pad, int value) {
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (r < rows_pad) e[r] = r < rows ? value : 0;
}

// bound > 0: every |x| <= bound is known (a kernel matrix: k(x, y) <= amplitude^2), no pass over the operand for the
// row exponents.
static int slice_operand(const double *x, int64_t rs, int64_t ks, int64_t rows, int64_t k, int64_t rows_pad, int64_t k_pad,
                         int s, int *e, signed char *q, cudaStream_t st, double bound = 0.0) {
    if (bound > 0.0)
        constant_exponent_kernel<<<(unsigned)((rows_pad + 255) / 256), 256, 0, st>>>(e, rows, rows_pad, ilogb(bound) + 1);
    else
        row_exponent_kernel<<<(unsigned)(rows_pad / 32), 256, 0, st>>>(x, rs, ks, rows, k, rows_pad, e);
    VGP_LAUNCH_CHECK();
    digit_planes_kernel<<<dim3((unsigned)(k_pad / 128), (unsigned)(rows_pad / 32)), 256, 0, st>>>(x, rs, ks, rows, k, rows_pad,
                                                                                                   k_pad, e, s, q);
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

// grow-only: the factorisations issue hundreds of products of shrinking size; the first large one sizes the planes
static int grow(void **ptr, size_t *have, size_t want, cudaStream_t st) {
    if (*have >= want) return VGP_OK;
    VGP_CUDA(cudaStreamSynchronize(st));                // earlier products may still read the old planes
    cache_free(*ptr);
    *ptr = nullptr;
    *have = 0;
    VGP_CUDA(cache_alloc(ptr, want));
    *have = want;
    return VGP_OK;
}
static int grow_for(EmuWorkspace &ws, int64_t m_pad, int64_t n_pad, int64_t kc, int slices, cudaStream_t st) {
    VGP_TRY(grow((void **)&ws.qa, &ws.qa_bytes, (size_t)slices * m_pad * kc, st));
    VGP_TRY(grow((void **)&ws.qb, &ws.qb_bytes, (size_t)slices * n_pad * kc, st));
    VGP_TRY(grow((void **)&ws.ea, &ws.ea_bytes, (size_t)m_pad * 4, st));
    return grow((void **)&ws.eb, &ws.eb_bytes, (size_t)n_pad * 4, st);
}

}  // namespace emu

int EmuWorkspace::reserve(int64_t rows, int slices, cudaStream_t st) {
    using namespace emu;
    if (slices < 2 || slices > S_MAX || rows <= 0) return VGP_OK;
    const int64_t r = round_up(rows, BM);
    return grow_for(*this, r, r, r < K_CHUNK ? r : K_CHUNK, slices, st);
}

void EmuWorkspace::release() {
    for (void *p : {(void *)qa, (void *)qb, (void *)ea, (void *)eb, (void *)partial}) cache_free(p);
    *this = EmuWorkspace();
}

int emulated_preload() {
    using namespace emu;
    int device = 0;
    VGP_CUDA(cudaGetDevice(&device));
    VGP_REQUIRE(device >= 0 && device < 64, "device ordinal out of range");
    static bool configured[64] = {};
    if (!configured[device]) {
        VGP_CUDA(cudaFuncSetAttribute(v2::emu_gemm_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::SMEM2));
        cudaFuncAttributes fa;
        VGP_CUDA(cudaFuncGetAttributes(&fa, row_exponent_kernel));
        VGP_CUDA(cudaFuncGetAttributes(&fa, digit_planes_kernel));
        // every kernel a distributed product may launch is loaded NOW: a first launch loads its module lazily, which can
        // wait for the device to drain -- not while another rank's barrier kernel is spinning on it
        VGP_CUDA(cudaFuncGetAttributes(&fa, push_tiles_kernel));
        VGP_CUDA(cudaFuncGetAttributes(&fa, constant_exponent_kernel));
        configured[device] = true;
    }
    return VGP_OK;
}

// Asynchronous on `st`; same operand convention as dense_gemm.  C must not alias A or B (with k > K_CHUNK the second
// chunk's planes would be cut from an already updated operand).
// C = beta C + sum_z partial_z, fixed summation order; lower: only the 128 x 128 tiles on or below the diagonal
__global__ void __launch_bounds__(256) emu_splitk_reduce_kernel(const double *partial, int64_t stride, int splits, double beta,
                                                                double *c, int64_t ldc, int64_t m, int64_t n, int lower) {
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= m * n) return;
    const int64_t i = e / n, j = e % n;
    if (lower && j / emu::BM > i / emu::BM) return;
    double acc = 0.0;
    for (int z = 0; z < splits; ++z) acc += partial[(int64_t)z * stride + i * n + j];
    c[i * ldc + j] = beta != 0.0 ? fma(beta, c[i * ldc + j], acc) : acc;
}

// Short-and-wide products (m, n small, k huge: K_zx K_zx^T over all observations): the k range is cut into pieces of
// <= 8192 (the int32 bound), every piece is a blockIdx.z of ONE launch writing its own partial C, summed afterwards in
// a fixed order.  The operands are cut into digit planes once, over the whole k range.  a == b with (trans_a, trans_b)
// = (0, 1) is a SYRK: one set of planes serves both sides.  bound_a / bound_b > 0: known bounds on |A|, |B| (see
// slice_operand).  lower: only the tiles on or below the diagonal (the caller mirrors).
int emulated_gemm_splitk(EmuWorkspace &ws, int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, double alpha,
                         const double *a, int64_t lda, const double *b, int64_t ldb, double beta, double *c, int64_t ldc,
                         int slices, int lower, double bound_a, double bound_b, cudaStream_t st) {
    using namespace emu;
    VGP_REQUIRE(slices >= 2 && slices <= S_MAX, "slices must be in [2, %d]", S_MAX);
    VGP_REQUIRE((const double *)c != a && (const double *)c != b, "emulated_gemm_splitk: C aliases an operand");
    VGP_REQUIRE(m > 0 && n > 0 && k > 0 && m % 2 == 0 && n % 2 == 0, "emulated_gemm_splitk: bad sizes");
    VGP_TRY(emulated_preload());
    const int64_t m_pad = round_up(m, BM), n_pad = round_up(n, BN), k_pad = round_up(k, BKB);
    const bool syrk = a == b && lda == ldb && m == n && trans_a == 0 && trans_b == 1;
    const int splits = (int)((k_pad + K_CHUNK - 1) / K_CHUNK);
    VGP_TRY(grow((void **)&ws.qa, &ws.qa_bytes, (size_t)slices * m_pad * k_pad, st));
    VGP_TRY(grow((void **)&ws.ea, &ws.ea_bytes, (size_t)m_pad * 4, st));
    if (!syrk) {
        VGP_TRY(grow((void **)&ws.qb, &ws.qb_bytes, (size_t)slices * n_pad * k_pad, st));
        VGP_TRY(grow((void **)&ws.eb, &ws.eb_bytes, (size_t)n_pad * 4, st));
    }
    VGP_TRY(grow((void **)&ws.partial, &ws.partial_bytes, (size_t)splits * m * n * 8, st));
    const int64_t a_rs = trans_a ? 1 : lda, a_ks = trans_a ? lda : 1;
    const int64_t b_rs = trans_b ? ldb : 1, b_ks = trans_b ? 1 : ldb;
    VGP_TRY(slice_operand(a, a_rs, a_ks, m, k, m_pad, k_pad, slices, ws.ea, ws.qa, st, bound_a));
    if (!syrk) VGP_TRY(slice_operand(b, b_rs, b_ks, n, k, n_pad, k_pad, slices, ws.eb, ws.qb, st, bound_b));
    alignas(64) CUtensorMap ma, mb;
    VGP_TRY(plane_map(&ma, ws.qa, (int64_t)slices * m_pad, k_pad, v2::BKB2, BM));
    VGP_TRY(plane_map(&mb, syrk ? ws.qa : ws.qb, (int64_t)slices * n_pad, k_pad, v2::BKB2, v2::BN2));
    GemmArgs p;
    p.m = m;
    p.n = n;
    p.m_pad = m_pad;
    p.n_pad = n_pad;
    p.kblocks = (int)(k_pad / BKB);
    p.s = slices;
    p.ea = ws.ea;
    p.eb = syrk ? ws.ea : ws.eb;
    p.c = reinterpret_cast<double *>(ws.partial);
    p.ldc = n;
    p.alpha = alpha;
    p.beta = 0.0;
    p.lower = lower;
    p.dist_n = 0;
    p.rank = 0;
    p.local_only = 0;
    p.kb_split = (int)(K_CHUNK / v2::BKB2);
    p.split_stride = m * n;
    p.tiles_n = (int)(n_pad / v2::BN2);
    dim3 grid((unsigned)(n_pad / v2::BN2), (unsigned)(m_pad / BM), (unsigned)splits);
    v2::emu_gemm_resident_kernel<<<grid, THREADS, v2::SMEM2, st>>>(ma, mb, p);
    VGP_LAUNCH_CHECK();
    emu_splitk_reduce_kernel<<<(unsigned)((m * n + 255) / 256), 256, 0, st>>>(reinterpret_cast<double *>(ws.partial), m * n,
                                                                             splits, beta, c, ldc, m, n, lower);
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

int emulated_gemm(EmuWorkspace &ws, int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, double alpha,
                  const double *a, int64_t lda, const double *b, int64_t ldb, double beta, double *c, int64_t ldc,
                  int slices, int lower, cudaStream_t st, const DistContext *dc, double bound_a, double bound_b) {
    using namespace emu;
    VGP_REQUIRE(slices >= 2 && slices <= S_MAX, "slices must be in [2, %d]", S_MAX);
    VGP_REQUIRE(ldc % 2 == 0 && ((uintptr_t)c & 15) == 0, "C must be 16-byte aligned with an even leading dimension");
    VGP_REQUIRE((const double *)c != a && (const double *)c != b, "emulated_gemm: C aliases an operand");
    if (m == 0 || n == 0) return VGP_OK;
    VGP_TRY(emulated_preload());
    const int64_t m_pad = round_up(m, BM), n_pad = round_up(n, BN);
    const int64_t kc_max = k < K_CHUNK ? round_up(k > 0 ? k : 1, BKB) : K_CHUNK;
    VGP_TRY(grow_for(ws, m_pad, n_pad, kc_max, slices, st));
    const int64_t a_rs = trans_a ? 1 : lda, a_ks = trans_a ? lda : 1;
    const int64_t b_rs = trans_b ? ldb : 1, b_ks = trans_b ? 1 : ldb;
    for (int64_t k0 = 0; k0 < k || k0 == 0; k0 += K_CHUNK) {
        const int64_t kc = k - k0 < K_CHUNK ? k - k0 : K_CHUNK;
        const int64_t k_pad = round_up(kc > 0 ? kc : 1, BKB);
        VGP_TRY(slice_operand(a + k0 * a_ks, a_rs, a_ks, m, kc, m_pad, k_pad, slices, ws.ea, ws.qa, st, bound_a));
        VGP_TRY(slice_operand(b + k0 * b_ks, b_rs, b_ks, n, kc, n_pad, k_pad, slices, ws.eb, ws.qb, st, bound_b));
        alignas(64) CUtensorMap ma, mb;
        VGP_TRY(plane_map(&ma, ws.qa, (int64_t)slices * m_pad, k_pad, v2::BKB2, BM));
        VGP_TRY(plane_map(&mb, ws.qb, (int64_t)slices * n_pad, k_pad, v2::BKB2, v2::BN2));
        GemmArgs p;
        p.m = m;
        p.n = n;
        p.m_pad = m_pad;
        p.n_pad = n_pad;
        p.kblocks = (int)(k_pad / BKB);
        p.s = slices;
        p.ea = ws.ea;
        p.eb = ws.eb;
        p.c = c;
        p.ldc = ldc;
        p.alpha = alpha;
        p.beta = k0 == 0 ? beta : 1.0;
        p.lower = lower;
        p.dist_n = 0;
        p.rank = 0;
        p.local_only = 0;
        p.kb_split = 0;
        p.split_stride = 0;
        const int64_t tiles_n = n_pad / v2::BN2, tiles_m = m_pad / BM;
        p.tiles_n = (int)tiles_n;
        dim3 grid((unsigned)tiles_n, (unsigned)tiles_m);
        if (dc && dc->nranks > 1) {          // every rank slices the whole operands (O(n^2)); the tiles are shared out
            p.dist_n = dc->nranks;
            p.rank = dc->rank;
            for (int q = 0; q < dc->nranks; ++q) p.delta[q] = dc->delta[q];
            grid = dim3((unsigned)((tiles_n * tiles_m + dc->nranks - 1) / dc->nranks), 1);
            p.local_only = dc->nranks > 2 ? 1 : 0;       // two ranks: the epilogue's peer stores are cheap enough
        }
        v2::emu_gemm_resident_kernel<<<grid, THREADS, v2::SMEM2, st>>>(ma, mb, p);
        VGP_LAUNCH_CHECK();
        const bool last_chunk = k0 + K_CHUNK >= k;
        if (p.dist_n > 0 && p.local_only && last_chunk) {
            push_tiles_kernel<<<dim3(grid.x, (unsigned)(dc->nranks - 1)), 256, 0, st>>>(p);
            VGP_LAUNCH_CHECK();
        }
        if (k == 0) break;
    }
    return VGP_OK;
}

}  // namespace vgp

using namespace vgp;

static thread_local EmuWorkspace g_public_ws[16];
namespace vgp {
size_t emulated_release() {         // vgp_workspace_trim: this thread's planes of the current device go back to the cache
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess || device < 0) return 0;
    EmuWorkspace &ws = g_public_ws[device & 15];
    const size_t bytes = ws.qa_bytes + ws.qb_bytes + ws.ea_bytes + ws.eb_bytes;
    ws.release();
    return bytes;
}
}  // namespace vgp

/* C[m][n] = alpha op(A) op(B) + beta C in FP64 accuracy class, the products on the int8 tensor cores
 * (see the header of this file).  trans_a == 0: A stored [m][k], 1: [k][m]; trans_b == 0: B stored [k][n], 1: [n][k]
 * (dense_gemm's convention).  slices in [2, 8]; lower != 0 computes only the 128 x 128 tiles on or below the diagonal.
 * The digit planes (slices * (m + n) * min(k, 8192) bytes) live in a grow-only workspace per host thread and device.
 * Asynchronous on `stream`. */
int vgp_gemm_emulated(int device, int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, double alpha,
                      const double *a_dev, int64_t lda, const double *b_dev, int64_t ldb, double beta, double *c_dev,
                      int64_t ldc, int slices, int lower, void *stream) {
    VGP_REQUIRE(m >= 0 && n >= 0 && k >= 0, "negative size");
    VGP_REQUIRE(a_dev && b_dev && c_dev, "NULL matrix");
    VGP_ENTER(device);
    // the public entry keeps its digit planes per host thread and device until vgp_workspace_trim
    return emulated_gemm(g_public_ws[device & 15], trans_a, trans_b, m, n, k, alpha, a_dev, lda, b_dev, ldb, beta, c_dev,
                         ldc, slices, lower, (cudaStream_t)stream, nullptr);
}
