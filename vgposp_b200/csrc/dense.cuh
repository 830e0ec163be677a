// Internal interface of the dense float64 layer (dense.cu).  All matrices row-major; every dimension
// handed to these functions is a multiple of vgp::TILE (128) -- callers pad (identity on the padding
// diagonal), which removes every edge case from the kernels.
#pragma once
#include <atomic>

#include "common.cuh"

namespace vgp {

// Digit planes and row exponents of the int8 tensor-core products (emulated.cu); grown on demand, blocks from the
// per-device workspace cache.
struct EmuWorkspace {
    signed char *qa = nullptr, *qb = nullptr;
    int *ea = nullptr, *eb = nullptr;
    size_t qa_bytes = 0, qb_bytes = 0, ea_bytes = 0, eb_bytes = 0;
    void *partial = nullptr;            // split-K partial products (emulated_gemm_splitk)
    size_t partial_bytes = 0;
    // Size for products of up to rows x rows x 8192 once, up front: growing later waits for the device, which inside
    // a distributed factorisation (other ranks' barrier kernels possibly spinning on this device) must not happen.
    int reserve(int64_t rows, int slices, cudaStream_t s);
    void release();
};

// the larger half of the recursive 2 x 2 split (dense.cu `split`): the largest product dimension of a factorisation
inline int64_t largest_half(int64_t n) { return n - (n / TILE / 2) * TILE; }

struct DenseWorkspace {
    int device = -1;
    EmuWorkspace emu;
    double *winv = nullptr;     // TILE x TILE scratch: explicit inverse of a diagonal block
    int *info = nullptr;        // device flag: 0, or 1 + row of the first non-positive pivot
    double *dinv = nullptr;     // [dinv_blocks][TILE][TILE]: inverses of the diagonal blocks of the last potrf
    int64_t dinv_blocks = 0;
    int ensure(int64_t nblocks);
    void release();
};

// Replicated-matrix distribution of the O(n^3) work over the GPUs of one box (setup row).
// Every rank holds a full replica of the matrix being factorised.  A GEMM whose output lies inside the replica
// and that is large enough is split by output tiles over the ranks, and the epilogue of the GEMM kernel stores
// each finished tile into EVERY replica (peer stores over NVLink/NVSwitch): the all-gather is fused into the
// GEMM, tile by tile.  Everything else (diagonal-block kernels, small GEMMs) runs redundantly on every rank.
// Each output element is produced by exactly one rank with the same kernel and summation order as on one GPU,
// so all replicas stay bitwise identical to the single-GPU result.
constexpr int DIST_MAX = 8;
struct DistContext {
    int rank = 0, nranks = 1;
    double *base = nullptr;                 // this rank's replica
    size_t bytes = 0;
    int64_t delta[DIST_MAX] = {0};          // replica base of rank q minus `base`, in doubles
    unsigned long long *flags[DIST_MAX] = {nullptr};   // barrier flags of rank q (peer-mapped): u64[DIST_MAX] + error
    unsigned long long seq = 0;
    // distribute a GEMM only above these.  Measured at n = 50k on 8 GPUs: 296 / 512 (98 distributed products) inverts in
    // 1.26 s, 96 / 256 (1157 products, two flag barriers each) in 1.07 s (profiles/r01_bench_n50k_g8_thresholds.json)
    int64_t min_tiles = 96, min_k = 256;
    int64_t dist_gemms = 0, barriers = 0;
};
void dense_set_dist(DistContext *ctx);      // thread-local; nullptr switches distribution off
int dense_dist_barrier(DistContext &ctx, cudaStream_t s);
int dense_preload();                        // load all dense kernels, set shared-memory opt-ins (current device)

// Row gate: the matrix at `base` ([rows][ld]) is still arriving -- row chunk c (rows [c * chunk_rows, (c + 1) *
// chunk_rows), columns up to the end of the chunk: the lower triangle by chunks) is complete once events[c], recorded
// in chunk order on the stream that copies it, has fired.  While a gate is set, every dense_* launch that touches rows
// of that matrix first makes its stream wait for the last chunk it needs, so a factorisation started right away runs
// behind the copy front instead of behind the whole copy (the recursive Cholesky works top-left to bottom-right).
struct RowGate {
    const double *base = nullptr;
    int64_t ld = 0, rows = 0, chunk_rows = 0;
    cudaEvent_t *events = nullptr;
    int nchunks = 0, waited = 0;            // chunks [0, waited) are already ordered before the compute stream
    // When another host thread issues the copies (pageable source: its copies block the issuing thread), the number
    // of chunk events recorded so far; a wait on chunk c first waits on the host until recorded > c (waiting on an
    // event that has not been recorded yet is a no-op).  nullptr: all events are recorded before the gate is set.
    const std::atomic<int> *recorded = nullptr;
    const std::atomic<int> *failed = nullptr;  // set by the copying thread on an error: waits give up
};
void dense_set_gate(RowGate *gate);         // thread-local; nullptr switches it off
void dense_set_pivot_floor(double floor);   // thread-local; pivots <= floor fail dense_potrf (0: only non-positive ones)

enum GemmTiles { GEMM_FULL = 0, GEMM_LOWER = 1 };   // LOWER: only tiles with tile_row >= tile_col

// C[m,n] = alpha op(A) op(B) + beta C.  trans_a == 0: A stored [m][k]; 1: stored [k][m].
// trans_b == 0: B stored [k][n]; 1: stored [n][k].
int dense_gemm(int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, double alpha, const double *a,
               int64_t lda, const double *b, int64_t ldb, double beta, double *c, int64_t ldc, GemmTiles tiles,
               cudaStream_t s);

// emulated.cu: the same product on the int8 tensor cores (tcgen05.mma kind::i8) after an error-free split into `slices`
// 7-bit digit planes.  dense_gemm routes the large products of potrf / trtri / lauum through it
// (VGP_OPT_GEMM_EMULATE_SLICES, VGP_OPT_GEMM_EMULATE_MIN).
int emulated_gemm(EmuWorkspace &ws, int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, double alpha,
                  const double *a, int64_t lda, const double *b, int64_t ldb, double beta, double *c, int64_t ldc,
                  int slices, int lower, cudaStream_t s, const DistContext *dist = nullptr, double bound_a = 0.0,
                  double bound_b = 0.0);
// split-K form for short-and-wide products (one launch, partial sums per 8192 of k, fixed-order reduction); a == b with
// (trans_a, trans_b) = (0, 1) shares one set of digit planes; bound_* > 0: known bounds on |A|, |B| (kernel matrices)
int emulated_gemm_splitk(EmuWorkspace &ws, int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, double alpha,
                         const double *a, int64_t lda, const double *b, int64_t ldb, double beta, double *c, int64_t ldc,
                         int slices, int lower, double bound_a, double bound_b, cudaStream_t s);
int emulated_preload();                     // load the kernels, set the shared-memory opt-in (current device)

int dense_gemm_splitk(int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, double alpha, const double *a,
                      int64_t lda, const double *b, int64_t ldb, double beta, double *c, int64_t ldc, int splits,
                      double *partial, cudaStream_t s, GemmTiles tiles = GEMM_FULL);

int dense_add_diag(double *a, int64_t ld, int64_t n, double value, cudaStream_t s);
int dense_zero_strict_upper(double *a, int64_t n, int64_t ld, cudaStream_t s);
int dense_mirror_lower(double *a, int64_t n, int64_t ld, cudaStream_t s);

// In-place lower Cholesky (strict upper triangle of off-diagonal blocks untouched, of diagonal
// 128-blocks zeroed).  Asynchronous; failures are recorded in ws.info.
int dense_potrf(double *a, int64_t n, int64_t ld, DenseWorkspace &ws, cudaStream_t s);
// In-place inverse of the lower factor / in-place lower part of X^T X.
int dense_trtri(double *l, int64_t n, int64_t ld, DenseWorkspace &ws, cudaStream_t s);
int dense_lauum(double *x, int64_t n, int64_t ld, DenseWorkspace &ws, cudaStream_t s);
// potrf + trtri + lauum + mirror; synchronises and reports *info_host (VGP_ERR_NOT_PD when non-zero).
int dense_spd_inverse(double *a, int64_t n, int64_t ld, DenseWorkspace &ws, int *info_host, cudaStream_t s);
int dense_read_info(DenseWorkspace &ws, int *info_host, cudaStream_t s);

// Triangular solves with a lower factor L [n][n]; B in place.
//   side 0 (left):  op(L) X = alpha B,  B is [n][nrhs]
//   side 1 (right): X op(L) = alpha B,  B is [nrhs][n]
// use_cached_inverses: L is the factor produced by the last dense_potrf on `ws` (its diagonal-block
// inverses are reused); otherwise each diagonal block is inverted on the fly.
int dense_trsm(int side, int trans, int64_t n, int64_t nrhs, double alpha, const double *l, int64_t ldl,
               double *b, int64_t ldb, DenseWorkspace &ws, bool use_cached_inverses, cudaStream_t s);

int pad_identity(double *a, int64_t ld, int64_t n, int64_t n_pad, cudaStream_t s);

}  // namespace vgp
