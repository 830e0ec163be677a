// (3') Greedy mutual-information placement, lazy-column formulation ("lazy-column").
//
// Same selections and scores as greedy.cu (reference: placement_algorithm2.py:105-145, :151-219, :371-413), but
// the precision of the unselected set is never rewritten.  Step t only needs
//     diag(P_t)      -- the denominators of every candidate,              d_j  <- d_j - p_j^2 / p_y
//     P_t[:, y]      -- the column of the winner,                         P_0[:, y] - sum_s (p_s p_s[y]) / p_s[y_s]
// so P_0 stays read-only and the rank-1 history is replayed on the one column that is needed.  Every fused
// multiply-add is issued in the order the dense downdate (greedy.cu, downdate_kernel) applies it, so with P_0
// resident (mode 0) scores are bitwise those of the dense formulation.
//
// Two sources for P_0[:, y]:
//   mode 0  P_0 = Sigma^-1 resident (potrf + trtri + lauum): one row read, 8 n bytes per selection;
//   mode 1  only M = L^-1 resident (potrf + trtri, 2/3 of the flops of the inverse): P_0[:, y] = M^T (M e_y),
//           a triangular matrix-vector product that streams the rows i >= y of M once -- 4 (n^2 - y^2) bytes,
//           HBM-bound (`trigemv_kernel`), reduced over 512-row blocks in a fixed order (deterministic).
//
// Per selection: [trigemv_kernel (mode 1)] + lazy_step_kernel, which applies the previous winner (new panel rows
// U[t-1], W[t-1], diagonal and numerator updates) and scores every candidate for the next arg-max in one launch.
// Algorithmic bytes of lazy_step_kernel at step t: 8 n (2 (t-1) + 6).
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <new>
#include <atomic>
#include <functional>
#include <thread>
#include <chrono>
#include <vector>

#include "common.cuh"
#include "dense.cuh"
#include "dist.cuh"

using namespace vgp;

struct vgp_lazy {
    int device = 0;
    int64_t n = 0, n_pad = 0, kmax = 0;
    double small_ = 0, jitter = 0;
    int mode = 0;
    double *cov = nullptr, *fac = nullptr;          // [n_pad][n_pad]
    double *d = nullptr, *d0 = nullptr, *num = nullptr;
    int *taken = nullptr;
    double *U = nullptr, *W = nullptr, *inv = nullptr;
    double *partial = nullptr;
    int64_t row_blocks = 0;
    vgp_candidate *partials = nullptr, *cur = nullptr;      // cur[2], double-buffered by step parity
    unsigned *counter = nullptr;
    int64_t *sel = nullptr;
    double *sel_score = nullptr, *step_scores = nullptr;
    int record = 0, factored = 0;
    // sharded form (vgp_lazy_create_dist): `fac` is the replica of a vgp_dist, `partial` its peer-visible tail (two
    // parity buffers); the row blocks of the triangular matrix-vector product are dealt round-robin to the ranks
    vgp_dist *dist = nullptr;
    int own_fac = 1, own_partial = 1;
    int64_t partial_parity_stride = 0;                // doubles between the two parity buffers (0: single buffer)
    int64_t loc_i1 = 0, loc_i2 = 0, loc_cutoff = 0;   // algorithm 3: grid strides I1, I2 and the index-box half-width
    double *cache = nullptr;                          // algorithm 3: the (partly stale) delta cache [n_pad]
    int64_t t = 0;
    std::atomic<int64_t> launches{0};      // the one-call path counts from two host threads (pageable source)
    int blocks = 0;
    DenseWorkspace ws;
    int profile = 0;
    cudaEvent_t pe[2] = {nullptr, nullptr};
    double prof_ms = 0;
    int64_t prof_count = 0;
};

namespace {

constexpr int RB = 512;                 // rows and columns of one trigemv tile
constexpr int CHUNK = 512;              // history entries staged in shared memory at a time
constexpr double NEG_INF = -INFINITY;

__device__ __forceinline__ bool better(double s, int64_t i, double so, int64_t io) {
    return so > s || (so == s && io < i);      // larger score wins, lower index breaks exact ties (:121)
}

__device__ __forceinline__ void warp_argmax(double &s, int64_t &i, int &slot) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double so = __shfl_xor_sync(0xffffffffu, s, off);
        const int64_t io = __shfl_xor_sync(0xffffffffu, i, off);
        const int sl = __shfl_xor_sync(0xffffffffu, slot, off);
        if (better(s, i, so, io)) {
            s = so;
            i = io;
            slot = sl;
        }
    }
}

__device__ __forceinline__ void block_argmax(double &s, int64_t &i, int &slot) {
    __shared__ double ss[8];
    __shared__ int64_t si[8];
    __shared__ int sl[8];
    warp_argmax(s, i, slot);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        ss[w] = s;
        si[w] = i;
        sl[w] = slot;
    }
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        s = l < nw ? ss[l] : NEG_INF;
        i = l < nw ? si[l] : INT64_MAX;
        slot = l < nw ? sl[l] : -1;
        warp_argmax(s, i, slot);
    }
}

// partial[rb][c] = sum over the rows i of row block rb, i >= y, c <= i, of  M[i][c] * (SQUARE ? M[i][c] : M[i][y])
struct TrigemvPeers {          // sharded form: this rank computes the row blocks rb % nranks == rank and stores them
    int rank = 0, nranks = 1;  // into the partial buffer of every rank (delta[q] = that buffer minus the local one)
    int64_t delta[DIST_MAX] = {0};
};

template <bool SQUARE>
__global__ void __launch_bounds__(256) trigemv_kernel(const double *__restrict__ m, int64_t ld, int64_t n_pad,
                                                      const vgp_candidate *cur, double *__restrict__ partial,
                                                      TrigemvPeers peers) {
    const int rb = blockIdx.y, cb = blockIdx.x;
    if (cb > rb) return;                                       // above the diagonal
    if (peers.nranks > 1 && rb % peers.nranks != peers.rank) return;
    int64_t y = 0;
    if (!SQUARE) {
        y = cur->index;
        if (y < 0 || (int64_t)(rb + 1) * RB <= y) return;      // M[i][y] = 0 for i < y
    }
    const int64_t row0 = (int64_t)rb * RB;
    const int rows = (int)min((int64_t)RB, n_pad - row0);
    __shared__ double v[RB];
    if (!SQUARE) {
        for (int r = threadIdx.x; r < RB; r += 256) {
            const int64_t i = row0 + r;
            v[r] = (r < rows && i >= y) ? m[i * ld + y] : 0.0;
        }
        __syncthreads();
    }
    const int64_t c = (int64_t)cb * RB + 2 * threadIdx.x;
    if (c >= n_pad) return;
    const bool diag = cb == rb;
    int r_begin = 0;
    if (!SQUARE && y > row0) r_begin = (int)(y - row0) & ~7;
    double ax = 0.0, ay = 0.0;
    const double *base = m + row0 * ld + c;
    for (int r = r_begin; r < rows; r += 8) {
        double2 e[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            e[u] = (r + u < rows) ? *reinterpret_cast<const double2 *>(base + (int64_t)(r + u) * ld) : make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (diag) {                                        // strict upper triangle holds leftovers of Sigma
                const int64_t i = row0 + r + u;
                if (c > i) e[u].x = 0.0;
                if (c + 1 > i) e[u].y = 0.0;
            }
            if (SQUARE) {
                ax = fma(e[u].x, e[u].x, ax);
                ay = fma(e[u].y, e[u].y, ay);
            } else {
                const double vi = (r + u < rows) ? v[r + u] : 0.0;
                ax = fma(e[u].x, vi, ax);
                ay = fma(e[u].y, vi, ay);
            }
        }
    }
    double2 *dst = reinterpret_cast<double2 *>(partial + (int64_t)rb * n_pad + c);
    if (peers.nranks > 1) {
        for (int q = 0; q < peers.nranks; ++q) *(dst + (peers.delta[q] >> 1)) = make_double2(ax, ay);   // peers: NVLink
    } else {
        *dst = make_double2(ax, ay);
    }
}

// d0[j] = sum over row blocks b >= j / RB of partial[b][j]   (mode 1: column norms of M = diag of M^T M)
__global__ void __launch_bounds__(256) colnorm_reduce_kernel(const double *__restrict__ partial, int64_t n_pad,
                                                             int64_t row_blocks, double *d0) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= n_pad) return;
    double s = 0.0;
    for (int64_t b = j / RB; b < row_blocks; ++b) s += partial[b * n_pad + j];
    d0[j] = s;
}

__global__ void __launch_bounds__(256) diag_gather_kernel(const double *__restrict__ a, int64_t ld, int64_t n_pad,
                                                          double *d0) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j < n_pad) d0[j] = a[j * ld + j];
}

__global__ void __launch_bounds__(256) lazy_reset_kernel(const double *__restrict__ cov, int64_t ld, int64_t n,
                                                         int64_t n_pad, double jitter, const double *__restrict__ d0,
                                                         double *d, double *num, int *taken) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= n_pad) return;
    d[j] = d0[j];
    num[j] = j < n ? cov[j * ld + j] + jitter : 0.0;
    taken[j] = j < n ? 0 : 1;
}

struct StepArgs {
    const double *cov, *fac, *partial;
    int64_t ld, n, n_pad, row_blocks;
    int mode;
    double *d, *num;
    int *taken;
    double *U, *W, *inv;
    int64_t t;                      // index of the selection this launch decides; t > 0: apply selection t - 1 first
    double small_, jitter;
    vgp_candidate *partials, *cur;
    unsigned *counter;
    int64_t *sel;
    double *sel_score, *step_row;
    double *cache;                  // algorithm 3 (local re-evaluation) when non-NULL
    int64_t loc_i1, loc_i2, loc_cutoff;
};

// Apply the winner of step t-1 (if any) to column j, then score j for step t; last block picks the winner.
__global__ void __launch_bounds__(256) lazy_step_kernel(StepArgs a) {
    __shared__ double uy[CHUNK], wy[CHUNK], iv[CHUNK];
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool live = j < a.n;
    double dj = 0.0, nj = 0.0;
    int tk = 1;
    if (live) {
        dj = a.d[j];
        nj = a.num[j];
        tk = a.taken[j];
    }
    if (a.t > 0) {
        const vgp_candidate prev = a.cur[(a.t - 1) & 1];
        const int64_t y = prev.index;
        if (y >= 0) {
            const int64_t hist = a.t - 1;          // rows of U / W already stored
            double p = 0.0, acc = 0.0;
            if (live) {
                if (a.mode == 0) {
                    p = a.fac[y * a.ld + j];       // P_0 is symmetric: row y is column y
                } else {
                    const int64_t b0 = (y > j ? y : j) / RB;
                    for (int64_t b = b0; b < a.row_blocks; ++b) p += a.partial[b * a.n_pad + j];
                }
                acc = j <= y ? a.cov[y * a.ld + j] : a.cov[j * a.ld + y];      // Sigma is symmetric: lower triangle only
                if (j == y) acc += a.jitter;
            }
            for (int64_t s0 = 0; s0 < hist; s0 += CHUNK) {
                const int cnt = (int)min((int64_t)CHUNK, hist - s0);
                __syncthreads();
                for (int s = threadIdx.x; s < cnt; s += 256) {
                    uy[s] = a.U[(s0 + s) * a.n_pad + y];
                    wy[s] = a.W[(s0 + s) * a.n_pad + y];
                    iv[s] = a.inv[s0 + s];
                }
                __syncthreads();
                if (live) {
                    const double *up = a.U + s0 * a.n_pad + j, *wp = a.W + s0 * a.n_pad + j;
#pragma unroll 8
                    for (int s = 0; s < cnt; ++s) {
                        const double us = up[(int64_t)s * a.n_pad], ws = wp[(int64_t)s * a.n_pad];
                        p = fma(-(us * uy[s]), iv[s], p);      // == downdate_kernel applied at step s0 + s
                        acc = fma(-ws, wy[s], acc);            // == segments_kernel
                    }
                }
            }
            const double inv_y = 1.0 / prev.pdiag;
            if (live) {
                const double w = acc / sqrt(prev.num);
                a.U[hist * a.n_pad + j] = p;
                a.W[hist * a.n_pad + j] = w;
                dj = fma(-(p * p), inv_y, dj);
                nj = fma(-w, w, nj);
                if (j == y) {
                    tk = 1;
                    a.taken[j] = 1;
                    a.inv[hist] = inv_y;
                }
                a.d[j] = dj;
                a.num[j] = nj;
            }
        }
    }
    // ---- score for step t (placement_algorithm2.py:105-125) -------------------------------------------------
    double s = NEG_INF;
    int64_t idx = INT64_MAX;
    if (live && !a.cache) {
        if (!tk) {
            const double den = 1.0 / dj - a.jitter;
            const double nom = nj - a.jitter;
            double sc = nom / den;
            if (fabs(den) < a.small_ || fabs(nom) < a.small_) sc = 0.0;     // :116-119
            if (a.step_row) a.step_row[j] = sc;
            if (sc > -1.0) {
                s = sc;
                idx = j;
            }
        } else if (a.step_row) {
            a.step_row[j] = nan("");
        }
    } else if (live) {
        // Algorithm 3 (snippets_a3.py:43-364): the arg-max runs over a cache of deltas of which only the entries
        // inside the index box around the previous winner are refreshed; entries of selected points are 0.
        double c = 0.0;
        if (!tk) {
            bool refresh = a.t == 0;                                        // first pass: every delta (whD, :71-120)
            if (!refresh) {
                const int64_t y = a.cur[(a.t - 1) & 1].index, s0 = a.loc_i1 * a.loc_i2;
                const int64_t y0 = y / s0, y1 = (y - y0 * s0) / a.loc_i2, y2 = y - y0 * s0 - y1 * a.loc_i2;
                const int64_t j0 = j / s0, j1 = (j - j0 * s0) / a.loc_i2, j2 = j - j0 * s0 - j1 * a.loc_i2;
                const int64_t c_ = a.loc_cutoff;                            // [i - cutoff, i + cutoff) per axis, :231-262
                refresh = j0 >= y0 - c_ && j0 < y0 + c_ && j1 >= y1 - c_ && j1 < y1 + c_ && j2 >= y2 - c_ &&
                          j2 < y2 + c_;
            }
            if (refresh) {
                const double den = 1.0 / dj - a.jitter;
                const double nom = nj - a.jitter;
                c = nom / den;
                if (fabs(den) < a.small_ || fabs(nom) < a.small_) c = 0.0;
            } else {
                c = a.cache[j];
            }
            if (c > -1.0) {
                s = c;
                idx = j;
            }
        }
        a.cache[j] = c;
        if (a.step_row) a.step_row[j] = c;                                  // delta_cached_iters[:, t]
    }
    __shared__ int64_t win_idx;
    __shared__ double win_score;
    {
        double rs = s;
        int64_t ri = idx;
        int slot = 0;
        block_argmax(rs, ri, slot);
        if (threadIdx.x == 0) {
            win_idx = ri;
            win_score = rs;
        }
    }
    __syncthreads();
    if (win_idx == INT64_MAX) {
        if (threadIdx.x == 0) a.partials[blockIdx.x] = vgp_candidate{NEG_INF, -1, 0.0, 0.0};
    } else if (idx == win_idx) {
        a.partials[blockIdx.x] = vgp_candidate{win_score, idx, nj, dj};
    }
    __threadfence();
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) last = atomicInc(a.counter, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    double rs = NEG_INF;
    int64_t ri = INT64_MAX;
    int slot = -1;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += 256) {
        const double cs = __ldcg(&a.partials[b].score);
        const int64_t ci = __ldcg((const long long *)&a.partials[b].index);
        if (ci >= 0 && better(rs, ri, cs, ci)) {
            rs = cs;
            ri = ci;
            slot = b;
        }
    }
    __syncthreads();
    block_argmax(rs, ri, slot);
    if (threadIdx.x == 0) {
        vgp_candidate w{NEG_INF, -1, 0.0, 0.0};
        if (slot >= 0) {
            w.score = rs;
            w.index = ri;
            w.num = __ldcg(&a.partials[slot].num);
            w.pdiag = __ldcg(&a.partials[slot].pdiag);
        }
        a.cur[a.t & 1] = w;
        a.sel[a.t] = w.index;
        a.sel_score[a.t] = w.score;
    }
}

__global__ void lazy_pad_identity_kernel(double *a, int64_t ld, int64_t n, int64_t n_pad) {
    const int64_t i = n + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) a[i * ld + i] = 1.0;
}

TrigemvPeers peers_of(const vgp_lazy *h, int parity) {
    TrigemvPeers p;
    if (h->dist && h->dist->ctx.nranks > 1) {
        p.rank = h->dist->ctx.rank;
        p.nranks = h->dist->ctx.nranks;
        (void)parity;                       // the parity offset is the same in every replica: it is part of `partial`
        for (int q = 0; q < p.nranks; ++q) p.delta[q] = h->dist->ctx.delta[q];
    }
    return p;
}

// every rank's blocks have landed in every rank's buffer (stream order, all ranks)
int peers_barrier(vgp_lazy *h, cudaStream_t s) {
    if (!h->dist || h->dist->ctx.nranks == 1) return VGP_OK;
    ++h->launches;
    return dense_dist_barrier(h->dist->ctx, s);
}

int check(vgp_lazy *h) {
    if (!h) {
        set_error("lazy greedy handle is NULL");
        return VGP_ERR_INVALID;
    }
    return VGP_OK;
}

}  // namespace

#define L_LAUNCH_CHECK(h)   \
    do {                    \
        ++(h)->launches;    \
        VGP_LAUNCH_CHECK(); \
    } while (0)

extern "C" {

static int lazy_create(vgp_lazy **handle, int device, int64_t n, int64_t kmax, double small, double jitter, int mode,
                       vgp_dist *dist);

int vgp_lazy_create(vgp_lazy **handle, int device, int64_t n, int64_t kmax, double small, double jitter, int mode) {
    return lazy_create(handle, device, n, kmax, small, jitter, mode, nullptr);
}

/* Sharded form over the ranks of a connected vgp_dist (one box): the inverse factor M = L^-1 is the replica every
 * rank already holds after vgp_dist_factor_inverse; the only O(n^2) work per selection, the triangular matrix-vector
 * product M^T (M e_y), is split by 512-row blocks over the ranks, each rank storing its blocks' partial sums into
 * every rank's buffer over NVLink (the tail behind the replicas), one flag barrier per selection.  The O(n t) step
 * kernel runs replicated and sums the blocks in the single-device order, so every rank selects the same winner and
 * scores are bitwise those of one device.  Call order: fill cov_dev (vgp_lazy_matrices) with Sigma and the replicas
 * with Sigma, vgp_dist_factor_inverse, vgp_lazy_adopt_factor, vgp_lazy_run -- the same calls on every rank. */
int vgp_lazy_create_dist(vgp_lazy **handle, vgp_dist *dist, int64_t n, int64_t kmax, double small, double jitter) {
    VGP_REQUIRE(dist, "dist handle is NULL");
    VGP_REQUIRE(dist->connected || dist->ctx.nranks == 1, "vgp_dist_connect first");
    VGP_REQUIRE(round_up(n, TILE) == dist->n_pad, "size does not match the dist handle");
    return lazy_create(handle, dist->device, n, kmax, small, jitter, 1, dist);
}

static int lazy_create(vgp_lazy **handle, int device, int64_t n, int64_t kmax, double small, double jitter, int mode,
                       vgp_dist *dist) {
    VGP_REQUIRE(handle, "handle is NULL");
    *handle = nullptr;
    VGP_REQUIRE(n > 0 && kmax > 0 && kmax <= n, "bad sizes n=%lld kmax=%lld", (long long)n, (long long)kmax);
    VGP_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (precision resident) or 1 (inverse factor resident)");
    VGP_ENTER(device);
    vgp_lazy *h = new (std::nothrow) vgp_lazy();
    VGP_REQUIRE(h, "out of host memory");
    h->dist = dist;
    h->device = device;
    h->n = n;
    h->kmax = kmax;
    h->small_ = small;
    h->jitter = jitter;
    h->mode = mode;
    h->n_pad = round_up(n, TILE);
    h->row_blocks = (h->n_pad + RB - 1) / RB;
    h->blocks = (int)((n + 255) / 256);
    const size_t mat = (size_t)h->n_pad * h->n_pad * 8, vec = (size_t)h->n_pad * 8;
    struct {
        void **p;
        size_t bytes;
    } allocs[] = {
        {(void **)&h->cov, mat},
        {(void **)&h->fac, dist ? 0 : mat},
        {(void **)&h->d, vec},
        {(void **)&h->d0, vec},
        {(void **)&h->num, vec},
        {(void **)&h->taken, (size_t)h->n_pad * 4},
        {(void **)&h->U, (size_t)kmax * vec},
        {(void **)&h->W, (size_t)kmax * vec},
        {(void **)&h->inv, (size_t)kmax * 8},
        {(void **)&h->partial, dist ? 0 : (size_t)h->row_blocks * vec},
        {(void **)&h->partials, (size_t)h->blocks * sizeof(vgp_candidate)},
        {(void **)&h->cur, 2 * sizeof(vgp_candidate)},
        {(void **)&h->counter, sizeof(unsigned)},
        {(void **)&h->sel, (size_t)kmax * 8},
        {(void **)&h->sel_score, (size_t)kmax * 8},
    };
    for (auto &al : allocs) {
        if (al.bytes == 0) continue;
        // everything comes from (and goes back to) the per-device workspace cache (common.cuh): the one-call path
        // neither allocates nor frees device memory after its first call (cudaFree alone cost 0.2 - 0.9 s there)
        cudaError_t e = cache_alloc(al.p, al.bytes);
        if (e != cudaSuccess) {
            int rc = cuda_fail(e, "cudaMalloc (lazy greedy state)", __FILE__, __LINE__);
            vgp_lazy_destroy(h);
            return rc;
        }
    }
    if (dist) {
        VGP_REQUIRE(RB == DIST_TAIL_RB && dist->tail_rows >= 2 * h->row_blocks, "dist tail too small");
        h->fac = dist->matrix;
        h->partial = dist->matrix + (size_t)h->n_pad * h->n_pad;
        h->partial_parity_stride = h->row_blocks * h->n_pad;
        h->own_fac = h->own_partial = 0;
    }
    cudaMemset(h->cov, 0, mat);
    cudaMemset(h->counter, 0, sizeof(unsigned));
    if (!dist) cudaMemset(h->partial, 0, (size_t)h->row_blocks * vec);
    cudaMemset(h->sel, 0xff, (size_t)kmax * 8);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        int rc = cuda_fail(e, "lazy greedy state init", __FILE__, __LINE__);
        vgp_lazy_destroy(h);
        return rc;
    }
    *handle = h;
    return VGP_OK;
}

int vgp_lazy_destroy(vgp_lazy *h) {
    if (!h) return VGP_OK;
    VGP_ENTER(h->device);
    void *ptrs[] = {h->d, h->d0, h->num, h->taken, h->U, h->W, h->inv,
                    h->own_partial ? h->partial : nullptr, h->partials, h->cur, h->counter, h->sel, h->sel_score,
                    h->step_scores, h->cache};
    for (void *p : ptrs) cache_free(p);
    cache_free(h->cov);
    if (h->own_fac) cache_free(h->fac);
    for (auto &e : h->pe)
        if (e) cudaEventDestroy(e);
    h->ws.release();
    delete h;
    return VGP_OK;
}

int vgp_lazy_matrices(vgp_lazy *h, double **cov_dev, double **factor_dev, int64_t *ld) {
    VGP_TRY(check(h));
    if (cov_dev) *cov_dev = h->cov;
    if (factor_dev) *factor_dev = h->fac;
    if (ld) *ld = h->n_pad;
    return VGP_OK;
}

int vgp_lazy_reset(vgp_lazy *h, void *stream) {
    VGP_TRY(check(h));
    if (!h->factored) {
        set_error("vgp_lazy_factor (or vgp_lazy_adopt_factor) first");
        return VGP_ERR_STATE;
    }
    VGP_ENTER(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    lazy_reset_kernel<<<(unsigned)((h->n_pad + 255) / 256), 256, 0, s>>>(h->cov, h->n_pad, h->n, h->n_pad, h->jitter,
                                                                       h->d0, h->d, h->num, h->taken);
    L_LAUNCH_CHECK(h);
    VGP_CUDA(cudaMemsetAsync(h->sel, 0xff, (size_t)h->kmax * 8, s));
    h->t = 0;
    return VGP_OK;
}

// The factor buffer already holds P_0 (mode 0, full symmetric) or M = L^-1 (mode 1, lower triangle): derive d0.
int vgp_lazy_adopt_factor(vgp_lazy *h, void *stream) {
    VGP_TRY(check(h));
    VGP_ENTER(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned vb = (unsigned)((h->n_pad + 255) / 256);
    if (h->mode == 0) {
        diag_gather_kernel<<<vb, 256, 0, s>>>(h->fac, h->n_pad, h->n_pad, h->d0);
        L_LAUNCH_CHECK(h);
    } else {
        dim3 grid((unsigned)h->row_blocks, (unsigned)h->row_blocks);
        trigemv_kernel<true><<<grid, 256, 0, s>>>(h->fac, h->n_pad, h->n_pad, nullptr, h->partial, peers_of(h, 0));
        L_LAUNCH_CHECK(h);
        VGP_TRY(peers_barrier(h, s));
        colnorm_reduce_kernel<<<vb, 256, 0, s>>>(h->partial, h->n_pad, h->row_blocks, h->d0);
        L_LAUNCH_CHECK(h);
    }
    h->factored = 1;
    return vgp_lazy_reset(h, stream);
}

// fac <- the matrix to factorise, rows [r0, r1) (columns [0, r1): the lower triangle by row chunks):
// Sigma + jitter on the diagonal, identity on the padding diagonal
static int lazy_stage_rows(vgp_lazy *h, int64_t r0, int64_t r1, cudaStream_t s) {
    VGP_CUDA(cudaMemcpy2DAsync(h->fac + r0 * h->n_pad, (size_t)h->n_pad * 8, h->cov + r0 * h->n_pad,
                               (size_t)h->n_pad * 8, (size_t)r1 * 8, (size_t)(r1 - r0), cudaMemcpyDeviceToDevice, s));
    const int64_t live = (r1 < h->n ? r1 : h->n) - r0;
    if (h->jitter != 0.0 && live > 0) {
        VGP_TRY(dense_add_diag(h->fac + r0 * h->n_pad + r0, h->n_pad, live, h->jitter, s));
        ++h->launches;
    }
    if (r1 > h->n) {
        const int64_t extra = h->n_pad - h->n;
        lazy_pad_identity_kernel<<<(unsigned)((extra + 127) / 128), 128, 0, s>>>(h->fac, h->n_pad, h->n, h->n_pad);
        L_LAUNCH_CHECK(h);
    }
    return VGP_OK;
}

static int lazy_factor_staged(vgp_lazy *h, int *info_host, cudaStream_t s);

int vgp_lazy_factor(vgp_lazy *h, int *info_host, void *stream) {
    VGP_TRY(check(h));
    VGP_REQUIRE(!h->dist, "sharded handle: factorise with vgp_dist_factor_inverse, then vgp_lazy_adopt_factor");
    VGP_ENTER(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    VGP_TRY(lazy_stage_rows(h, 0, h->n_pad, s));
    return lazy_factor_staged(h, info_host, s);
}

// potrf + trtri (+ lauum + mirror in mode 0) of the staged matrix, then the initial denominators
static int lazy_factor_staged(vgp_lazy *h, int *info_host, cudaStream_t s) {
    void *stream = (void *)s;
    const int64_t before = g_launches;
    int rc = dense_potrf(h->fac, h->n_pad, h->n_pad, h->ws, s);
    if (rc == VGP_OK) rc = dense_read_info(h->ws, info_host, s);
    if (rc == VGP_OK) rc = dense_trtri(h->fac, h->n_pad, h->n_pad, h->ws, s);
    if (rc == VGP_OK && h->mode == 0) {
        rc = dense_lauum(h->fac, h->n_pad, h->n_pad, h->ws, s);
        if (rc == VGP_OK) rc = dense_mirror_lower(h->fac, h->n_pad, h->n_pad, s);
    }
    h->launches += g_launches - before;
    VGP_TRY(rc);
    return vgp_lazy_adopt_factor(h, stream);
}

int vgp_lazy_run(vgp_lazy *h, int64_t k, void *stream) {
    VGP_TRY(check(h));
    if (!h->factored) {
        set_error("vgp_lazy_factor first");
        return VGP_ERR_STATE;
    }
    VGP_REQUIRE(k >= 0 && h->t + k <= h->kmax, "k = %lld exceeds kmax = %lld (already %lld)", (long long)k,
                (long long)h->kmax, (long long)h->t);
    VGP_ENTER(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (h->profile && !h->pe[0]) {
        VGP_CUDA(cudaEventCreate(&h->pe[0]));
        VGP_CUDA(cudaEventCreate(&h->pe[1]));
    }
    for (int64_t i = 0; i < k; ++i) {
        if (h->t > 0 && h->mode == 1) {
            dim3 grid((unsigned)h->row_blocks, (unsigned)h->row_blocks);
            if (h->profile) VGP_CUDA(cudaEventRecord(h->pe[0], s));
            double *part = h->partial + (h->t & 1) * h->partial_parity_stride;      // peers may still read the other one
            trigemv_kernel<false><<<grid, 256, 0, s>>>(h->fac, h->n_pad, h->n_pad, h->cur + ((h->t - 1) & 1), part,
                                                       peers_of(h, (int)(h->t & 1)));
            L_LAUNCH_CHECK(h);
            VGP_TRY(peers_barrier(h, s));
            if (h->profile) {
                float ms = 0.f;
                VGP_CUDA(cudaEventRecord(h->pe[1], s));
                VGP_CUDA(cudaEventSynchronize(h->pe[1]));
                VGP_CUDA(cudaEventElapsedTime(&ms, h->pe[0], h->pe[1]));
                h->prof_ms += ms;
                ++h->prof_count;
            }
        }
        StepArgs a;
        a.cov = h->cov;
        a.fac = h->fac;
        a.partial = h->partial + (h->t & 1) * h->partial_parity_stride;
        a.ld = h->n_pad;
        a.n = h->n;
        a.n_pad = h->n_pad;
        a.row_blocks = h->row_blocks;
        a.mode = h->mode;
        a.d = h->d;
        a.num = h->num;
        a.taken = h->taken;
        a.U = h->U;
        a.W = h->W;
        a.inv = h->inv;
        a.t = h->t;
        a.small_ = h->small_;
        a.jitter = h->jitter;
        a.partials = h->partials;
        a.cur = h->cur;
        a.counter = h->counter;
        a.sel = h->sel;
        a.sel_score = h->sel_score;
        a.step_row = (h->record && h->step_scores) ? h->step_scores + h->t * h->n : nullptr;
        a.cache = h->loc_cutoff > 0 ? h->cache : nullptr;
        a.loc_i1 = h->loc_i1;
        a.loc_i2 = h->loc_i2;
        a.loc_cutoff = h->loc_cutoff;
        lazy_step_kernel<<<h->blocks, 256, 0, s>>>(a);
        L_LAUNCH_CHECK(h);
        ++h->t;
    }
    return VGP_OK;
}

int vgp_lazy_results(vgp_lazy *h, int64_t *count, int64_t *selection_host, double *scores_host, int64_t capacity,
                     void *stream) {
    VGP_TRY(check(h));
    VGP_ENTER(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t c = h->t < capacity ? h->t : capacity;
    if (count) *count = h->t;
    if (selection_host && c > 0)
        VGP_CUDA(cudaMemcpyAsync(selection_host, h->sel, (size_t)c * 8, cudaMemcpyDeviceToHost, s));
    if (scores_host && c > 0)
        VGP_CUDA(cudaMemcpyAsync(scores_host, h->sel_score, (size_t)c * 8, cudaMemcpyDeviceToHost, s));
    VGP_CUDA(cudaStreamSynchronize(s));
    return VGP_OK;
}

/* Algorithm 3 (snippets_a3.sparse_placement_algorithm_3): candidates are the points of an I0 x I1 x I2 grid
 * (index = I2 I1 i0 + I2 i1 + i2); after each selection only the deltas inside the index box
 * [i - cutoff, i + cutoff) per axis around the winner are re-evaluated, every other entry of the cache keeps its stale
 * value.  cutoff = 0 switches back to the exact greedy.  Call before vgp_lazy_run (and after any reset). */
int vgp_lazy_set_local(vgp_lazy *h, int64_t i0, int64_t i1, int64_t i2, int64_t cutoff) {
    VGP_TRY(check(h));
    VGP_REQUIRE(cutoff >= 0, "negative cutoff");
    if (cutoff == 0) {
        h->loc_cutoff = 0;
        return VGP_OK;
    }
    VGP_REQUIRE(i0 > 0 && i1 > 0 && i2 > 0 && i0 * i1 * i2 == h->n, "grid %lld x %lld x %lld does not have n = %lld points",
                (long long)i0, (long long)i1, (long long)i2, (long long)h->n);
    VGP_REQUIRE(h->t == 0, "set the local mode before the first selection");
    VGP_ENTER(h->device);
    if (!h->cache) VGP_CUDA(cache_alloc((void **)&h->cache, (size_t)h->n_pad * 8));
    h->loc_i1 = i1;
    h->loc_i2 = i2;
    h->loc_cutoff = cutoff;
    return VGP_OK;
}

int vgp_lazy_record_scores(vgp_lazy *h, int enable) {
    VGP_TRY(check(h));
    VGP_ENTER(h->device);
    if (enable && !h->step_scores) VGP_CUDA(cache_alloc((void **)&h->step_scores, (size_t)h->kmax * h->n * 8));
    h->record = enable ? 1 : 0;
    return VGP_OK;
}

int vgp_lazy_step_scores(vgp_lazy *h, double *scores_host, int64_t capacity_rows, void *stream) {
    VGP_TRY(check(h));
    VGP_REQUIRE(scores_host, "scores_host is NULL");
    if (!h->step_scores) {
        set_error("score recording was not enabled");
        return VGP_ERR_STATE;
    }
    VGP_ENTER(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t rows = h->t < capacity_rows ? h->t : capacity_rows;
    if (rows > 0)
        VGP_CUDA(cudaMemcpyAsync(scores_host, h->step_scores, (size_t)rows * h->n * 8, cudaMemcpyDeviceToHost, s));
    VGP_CUDA(cudaStreamSynchronize(s));
    return VGP_OK;
}

int vgp_lazy_launch_count(vgp_lazy *h, int64_t *launches) {
    VGP_TRY(check(h));
    VGP_REQUIRE(launches, "launches is NULL");
    *launches = h->launches;
    return VGP_OK;
}

// Timing of trigemv_kernel (mode 1): events around every launch, read back per launch (serialises the stream --
// for measurement runs only).  total_ms / launches since the last enable.
int vgp_lazy_profile(vgp_lazy *h, int enable, double *total_ms, int64_t *launches) {
    VGP_TRY(check(h));
    if (total_ms) *total_ms = h->prof_ms;
    if (launches) *launches = h->prof_count;
    h->profile = enable ? 1 : 0;
    if (enable) {
        h->prof_ms = 0;
        h->prof_count = 0;
    }
    return VGP_OK;
}

int vgp_placement_host_dense(int device, const double *cov_host, int64_t n, int64_t ld_host, int64_t k, double small,
                             double jitter, int64_t *selection_host, double *scores_host, double *step_scores_host,
                             double *seconds_host);

// host wall-clock breakdown of this thread's last vgp_placement_host_ex call (lazy formulations):
// [0] state allocation + init, [1] enqueue, [2] release, [3] whole call
static thread_local double g_call_stats[4] = {0, 0, 0, 0};
int vgp_placement_host_wall(double *seconds4) {
    VGP_REQUIRE(seconds4, "NULL argument");
    for (int i = 0; i < 4; ++i) seconds4[i] = g_call_stats[i];
    return VGP_OK;
}

int vgp_placement_host_ex(int device, const double *cov_host, int64_t n, int64_t ld_host, int64_t k, double small,
                          double jitter, int formulation, int64_t *selection_host, double *scores_host,
                          double *step_scores_host, double *seconds_host) {
    VGP_REQUIRE(cov_host && selection_host, "NULL argument");
    VGP_REQUIRE(n > 0 && ld_host >= n && k > 0 && k <= n, "bad sizes n=%lld ld=%lld k=%lld", (long long)n,
                (long long)ld_host, (long long)k);
    VGP_REQUIRE(formulation >= VGP_FORMULATION_AUTO && formulation <= VGP_FORMULATION_LAZY_FACTOR,
                "unknown formulation %d", formulation);
    if (formulation == VGP_FORMULATION_AUTO)
        // the inverse factor saves n^3/3 flop of setup and costs ~8 n^2 / 3 bytes of HBM traffic per selection
        formulation = k * 35 < n ? VGP_FORMULATION_LAZY_FACTOR : VGP_FORMULATION_LAZY_PRECISION;
    if (formulation == VGP_FORMULATION_DENSE)
        return vgp_placement_host_dense(device, cov_host, n, ld_host, k, small, jitter, selection_host, scores_host,
                                        step_scores_host, seconds_host);
    VGP_ENTER(device);
    const auto wall0 = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count();
    };
    for (double &v : g_call_stats) v = 0.0;
    vgp_lazy *h = nullptr;
    VGP_TRY(vgp_lazy_create(&h, device, n, k, small, jitter, formulation == VGP_FORMULATION_LAZY_FACTOR ? 1 : 0));
    g_call_stats[0] = since(wall0);                  // state allocation (workspace cache hit or cudaMalloc) + init
    {   // numerical rank deficiency counts as "not positive definite": pivots below 1e-12 of the largest variance
        double scale = 0.0;
        for (int64_t i = 0; i < n; ++i) scale = std::max(scale, fabs(cov_host[i * ld_host + i]));
        dense_set_pivot_floor(1e-12 * scale);
    }
    cudaStream_t s = nullptr;
    cudaEvent_t ev[4];
    for (auto &e : ev) cudaEventCreate(&e);
    int rc = VGP_OK;
    auto fail = [&](int code) {
        dense_set_pivot_floor(0.0);
        for (auto &e : ev) cudaEventDestroy(e);
        const auto t = std::chrono::steady_clock::now();
        vgp_lazy_destroy(h);
        g_call_stats[2] = since(t);                  // release (back into the workspace cache)
        g_call_stats[3] = since(wall0);
        return code;
    };
    // H2D of the lower triangle in row chunks on a copy stream; the factorisation starts at once on `s` and every
    // launch waits only for the last chunk it touches (RowGate, dense.cuh): the copy hides behind the Cholesky of
    // the leading blocks.  Sigma is symmetric -- the strict upper triangle is never read (lazy_step_kernel).
    constexpr int64_t CHUNK_ROWS = 2048;
    const int nchunks = (int)((h->n_pad + CHUNK_ROWS - 1) / CHUNK_ROWS);
    cudaStream_t cs = nullptr;
    std::vector<cudaEvent_t> chunk_ev((size_t)nchunks, nullptr);
    cudaError_t ce = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
    if (ce != cudaSuccess) return fail(cuda_fail(ce, "copy stream", __FILE__, __LINE__));
    std::function<void()> join_before_fail = [] {};
    auto fail2 = [&](int code) {
        dense_set_gate(nullptr);
        join_before_fail();                              // the copy helper (pageable source) must be done with `h`
        cudaStreamSynchronize(cs);
        for (auto &e : chunk_ev)
            if (e) cudaEventDestroy(e);
        cudaStreamDestroy(cs);
        return fail(code);
    };
    for (auto &e : chunk_ev)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess)
            return fail2(cuda_fail(cudaGetLastError(), "chunk events", __FILE__, __LINE__));
    cudaEventRecord(ev[0], s);
    cudaStreamWaitEvent(cs, ev[0], 0);
    // A pageable source (a plain NumPy array: what the reference's callers pass) makes every copy block the issuing
    // thread.  The copies are then issued by a helper thread while this one enqueues the factorisation; the gate waits on
    // the host until the chunk's event has been recorded, then on the device for the event (pinned source: this thread
    // issues everything up front, as before).
    cudaPointerAttributes pattr;
    const bool pageable = cudaPointerGetAttributes(&pattr, cov_host) != cudaSuccess || pattr.type == cudaMemoryTypeUnregistered;
    cudaGetLastError();
    std::atomic<int> recorded{0}, copy_failed{0};
    int copy_rc = VGP_OK;
    auto copy_all = [&]() -> void {
        if (pageable) cudaSetDevice(device);
        for (int c = 0; c < nchunks; ++c) {
            const int64_t r0 = (int64_t)c * CHUNK_ROWS, r1 = std::min(r0 + CHUNK_ROWS, h->n_pad);
            const int64_t rows = std::min(r1, n) - r0, cols = std::min(r1, n);
            if (rows > 0) {
                cudaError_t e = cudaMemcpy2DAsync(h->cov + r0 * h->n_pad, (size_t)h->n_pad * 8, cov_host + r0 * ld_host,
                                                  (size_t)ld_host * 8, (size_t)cols * 8, (size_t)rows, cudaMemcpyHostToDevice,
                                                  cs);
                if (e != cudaSuccess) {
                    copy_rc = cuda_fail(e, "H2D of cov_vv", __FILE__, __LINE__);
                    copy_failed.store(1, std::memory_order_release);
                    return;
                }
            }
            const int rc1 = lazy_stage_rows(h, r0, r1, cs);
            if (rc1 != VGP_OK) {
                copy_rc = rc1;
                copy_failed.store(1, std::memory_order_release);
                return;
            }
            cudaEventRecord(chunk_ev[(size_t)c], cs);
            recorded.store(c + 1, std::memory_order_release);
        }
        cudaEventRecord(ev[1], cs);                 // the last byte has arrived (overlaps the factorisation)
    };
    std::thread copier;
    auto join_copier = [&]() {
        if (copier.joinable()) copier.join();
    };
    join_before_fail = join_copier;
    if (pageable) {
        copier = std::thread(copy_all);             // h->launches is atomic: both threads count launches
    } else {
        copy_all();
        if (copy_rc != VGP_OK) return fail2(copy_rc);
    }
    RowGate gate;
    gate.base = h->fac;
    gate.ld = h->n_pad;
    gate.rows = h->n_pad;
    gate.chunk_rows = CHUNK_ROWS;
    gate.events = chunk_ev.data();
    gate.nchunks = nchunks;
    if (pageable) {
        gate.recorded = &recorded;
        gate.failed = &copy_failed;
    }
    if (option(VGP_OPT_H2D_OVERLAP) == 0) {              // measurement knob: finish the copy before factorising
        join_copier();
        gate.waited = nchunks, cudaStreamWaitEvent(s, chunk_ev[(size_t)nchunks - 1], 0);
    }
    dense_set_gate(&gate);
    int info = 0;
    rc = lazy_factor_staged(h, &info, s);
    dense_set_gate(nullptr);
    join_copier();
    if (rc == VGP_OK && copy_rc != VGP_OK) {
        if (pageable) set_error("the host-to-device copy of cov_vv failed (status %d)", copy_rc);
        rc = copy_rc;
    }
    if (rc != VGP_OK) return fail2(rc);
    cudaStreamWaitEvent(s, chunk_ev[(size_t)nchunks - 1], 0);     // (already implied; keeps `cov` ordered before the steps)
    cudaEventRecord(ev[2], s);
    if (step_scores_host) {
        rc = vgp_lazy_record_scores(h, 1);
        if (rc != VGP_OK) return fail2(rc);
    }
    rc = vgp_lazy_run(h, k, s);
    if (rc != VGP_OK) return fail2(rc);
    int64_t count = 0;
    rc = vgp_lazy_results(h, &count, selection_host, scores_host, k, s);
    if (rc != VGP_OK) return fail2(rc);
    if (step_scores_host) {
        rc = vgp_lazy_step_scores(h, step_scores_host, k, s);
        if (rc != VGP_OK) return fail2(rc);
    }
    cudaEventRecord(ev[3], s);
    g_call_stats[1] = since(wall0) - g_call_stats[0];         // host time to enqueue everything
    ce = cudaEventSynchronize(ev[3]);
    if (ce != cudaSuccess) return fail2(cuda_fail(ce, "placement", __FILE__, __LINE__));
    if (seconds_host) {
        float ms;
        cudaEventElapsedTime(&ms, ev[0], ev[1]);
        seconds_host[0] = ms * 1e-3;
        cudaEventElapsedTime(&ms, ev[0], ev[2]);      // from the first byte: the copy runs underneath
        seconds_host[1] = ms * 1e-3;
        cudaEventElapsedTime(&ms, ev[2], ev[3]);
        seconds_host[2] = ms * 1e-3;
        cudaEventElapsedTime(&ms, ev[0], ev[3]);
        seconds_host[3] = ms * 1e-3;
    }
    return fail2(VGP_OK);
}

int vgp_placement_host(int device, const double *cov_host, int64_t n, int64_t ld_host, int64_t k, double small,
                       double jitter, int64_t *selection_host, double *scores_host, double *step_scores_host,
                       double *seconds_host) {
    return vgp_placement_host_ex(device, cov_host, n, ld_host, k, small, jitter, VGP_FORMULATION_AUTO, selection_host,
                                 scores_host, step_scores_host, seconds_host);
}

}  // extern "C"
