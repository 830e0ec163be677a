// (3) Krause-style greedy mutual-information placement, incremental form.
//
// Reference semantics: placement_algorithm2.py:105-145 (alg. 1), :151-219 (alg. 2), :371-413
// (nominator / denominator through pinv).  Here the per-candidate score
//     delta_y = sigma^2(y | A) / sigma^2(y | Abar \ y)
// is held incrementally: the numerator through a growing conditioning panel W (rows = selections,
// num_j = Sigma_jj - sum_s W[s][j]^2), the denominator as 1 / P_jj where P is the precision of the
// still-unselected set, downdated by the rank-1 Schur step P -= p p^T / p_y when y leaves the set.
//
// Data layout in HBM (one handle = one shard = columns J = [c0, c0 + nloc) of the candidate set):
//     cov   [n_pad][ld]   Sigma[:, J]     row-major, read one row segment per selection
//     prec  [n_pad][ld]   P[:, J]         row-major, read + written once per selection (the hot loop)
//     wfull [kmax][n_pad] conditioning panel rows (full length: W[s][y] is needed for any later y)
//     pfull [n_pad], ploc [ld], num [ld], taken [ld]
// The dominant kernel is `downdate_kernel`: 16 * n_pad * ld algorithmic bytes per launch, pure
// streaming read-modify-write with 128-bit accesses -> HBM roofline.
#include <math.h>
#include <string.h>

#include <new>
#include <utility>
#include <vector>

#include "common.cuh"
#include "dense.cuh"

using namespace vgp;

struct vgp_greedy {
    int device = 0;
    int64_t n = 0, c0 = 0, nloc = 0, kmax = 0;
    int64_t n_pad = 0, ld = 0;
    double small_ = 0, jitter = 0;
    double *cov = nullptr, *prec = nullptr, *prec_saved = nullptr;
    double *num = nullptr, *wfull = nullptr, *pfull = nullptr, *ploc = nullptr, *seg = nullptr;
    int *taken = nullptr;
    vgp_candidate *partials = nullptr, *cur = nullptr, *best = nullptr;
    unsigned *counter = nullptr;
    int64_t *sel = nullptr;
    double *sel_score = nullptr;
    double *step_scores = nullptr;
    int record = 0;
    int64_t t = 0;          // selections enqueued so far
    int64_t launches = 0;
    int score_blocks = 0;
    int sm_count = 148;
    DenseWorkspace ws;
    int profile = 0;                                    // CUDA events around every downdate launch
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_step_events;     // peer path: around peer_step_kernel
    // peer-memory exchange (one box, NVLink): every rank owns a mailbox its peers store into
    char *mailbox = nullptr;                            // this rank's mailbox (plain cudaMalloc: IPC-exportable)
    size_t mailbox_bytes = 0;
    int comm_rank = -1, comm_nranks = 0, comm_ipc = 0;
    int64_t comm_stride = 0;
    size_t comm_seg_off = 0;
    int64_t comm_bounds[65] = {0};
    char *peer_mb[64] = {nullptr};                      // peers' mailboxes mapped into this process
    unsigned *counter2 = nullptr;
    unsigned long long epoch = 0;                       // exchange sequence number, never reset
};

namespace {

constexpr int MAX_RANKS = 64;
struct Bounds {
    int64_t b[MAX_RANKS + 1];
    int nranks;
};

struct Cand {
    double score;
    int64_t idx;
    int slot;
};

__device__ __forceinline__ bool better(double s, int64_t i, double so, int64_t io) {
    // does (so, io) beat (s, i)?  larger score wins, lower index breaks exact ties
    return so > s || (so == s && io < i);
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

// Block-wide arg-max; result valid in thread 0.
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
        s = l < nw ? ss[l] : -INFINITY;
        i = l < nw ? si[l] : INT64_MAX;
        slot = l < nw ? sl[l] : -1;
        warp_argmax(s, i, slot);
    }
}

constexpr double NEG_INF = -INFINITY;

// delta_j for every local candidate + first strict maximum (placement_algorithm2.py:105-125).
__global__ void __launch_bounds__(256) score_kernel(const double *__restrict__ prec, int64_t ld, int64_t c0,
                                                    int64_t nloc, const double *__restrict__ num,
                                                    const int *__restrict__ taken, double small_, double jitter,
                                                    vgp_candidate *partials, unsigned *counter,
                                                    vgp_candidate *best, double *step_row) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    double s = NEG_INF, nm = 0.0, pd = 0.0;
    int64_t idx = INT64_MAX;
    if (j < nloc) {
        pd = prec[(c0 + j) * ld + j];
        nm = num[j];
        if (!taken[j]) {
            const double den = 1.0 / pd - jitter;        // sigma^2(y | Abar \ y)
            const double nom = nm - jitter;              // sigma^2(y | A)
            double d = nom / den;
            if (fabs(den) < small_ || fabs(nom) < small_) d = 0.0;    // :116-119
            if (step_row) step_row[j] = d;
            if (d > -1.0) {                               // running best starts at -1, strict '<' (:106,:121)
                s = d;
                idx = c0 + j;
            }
        } else if (step_row) {
            step_row[j] = nan("");
        }
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
        if (threadIdx.x == 0) partials[blockIdx.x] = vgp_candidate{NEG_INF, -1, 0.0, 0.0};
    } else if (idx == win_idx) {
        partials[blockIdx.x] = vgp_candidate{win_score, idx, nm, pd};
    }
    __threadfence();
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    double rs = NEG_INF;
    int64_t ri = INT64_MAX;
    int slot = -1;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += 256) {
        // other CTAs' partials: read through L2 (written before their fence + counter increment)
        const double cs = __ldcg(&partials[b].score);
        const int64_t ci = __ldcg((const long long *)&partials[b].index);
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
            w.num = __ldcg(&partials[slot].num);
            w.pdiag = __ldcg(&partials[slot].pdiag);
        }
        *best = w;
    }
}

// Winner over all shards' records: max score, lowest index on exact ties (single ascending scan with
// strict '<', placement_algorithm2.py:121).
__global__ void select_kernel(const vgp_candidate *records, int nrecords, vgp_candidate *cur, int64_t *sel,
                              double *sel_score, int64_t t) {
    if (threadIdx.x != 0) return;
    vgp_candidate w{NEG_INF, -1, 0.0, 0.0};
    for (int r = 0; r < nrecords; ++r) {
        const vgp_candidate c = records[r];
        if (c.index < 0) continue;
        if (w.index < 0 || c.score > w.score || (c.score == w.score && c.index < w.index)) w = c;
    }
    *cur = w;
    sel[t] = w.index;
    sel_score[t] = w.score;
}

// w_J = (Sigma'[y, J] - sum_s W[s][J] W[s][y]) / sqrt(num_y),  p_J = P[y, J]
__global__ void __launch_bounds__(256) segments_kernel(const double *__restrict__ cov,
                                                       const double *__restrict__ prec, int64_t ld, int64_t c0,
                                                       int64_t nloc, const double *__restrict__ wfull,
                                                       int64_t n_pad, int64_t t, double jitter,
                                                       const vgp_candidate *cur, double *seg, int64_t stride) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= stride) return;
    const int64_t y = cur->index;
    double w = 0.0, p = 0.0;
    if (y >= 0 && j < nloc) {
        double acc = cov[y * ld + j];
        if (c0 + j == y) acc += jitter;
        for (int64_t s = 0; s < t; ++s) acc = fma(-wfull[s * n_pad + c0 + j], wfull[s * n_pad + y], acc);
        w = acc / sqrt(cur->num);
        p = prec[y * ld + j];
    }
    seg[j] = w;
    seg[stride + j] = p;
}

// Unpack the gathered [rank][w | p] segments into full-length rows and update the local numerators.
__global__ void __launch_bounds__(256) unpack_kernel(const double *__restrict__ gathered, int64_t stride,
                                                     Bounds bounds, int64_t n, int64_t n_pad, int64_t c0,
                                                     int64_t nloc, int64_t ld, double *wrow, double *pfull,
                                                     double *ploc, double *num, int *taken,
                                                     const vgp_candidate *cur) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t y = cur->index;
    if (y < 0) return;
    if (i < n_pad) {
        double w = 0.0, p = 0.0;
        if (i < n) {
            int g = 0;
            while (g + 1 < bounds.nranks && i >= bounds.b[g + 1]) ++g;
            const double *base = gathered + (int64_t)g * 2 * stride;
            w = base[i - bounds.b[g]];
            p = base[stride + i - bounds.b[g]];
        }
        wrow[i] = w;
        pfull[i] = p;
        if (i >= c0 && i < c0 + nloc) {
            const int64_t j = i - c0;
            num[j] = fma(-w, w, num[j]);
            ploc[j] = p;
            if (i == y) taken[j] = 1;
        }
    }
    // zero the padding of ploc (columns nloc .. ld)
    if (i >= nloc && i < ld) ploc[i] = 0.0;
}

// P[i][j] -= (p_i p_j) / p_y over the whole panel; row y and column y become exact zeros.
// (p_i p_j) is formed first so the update is bitwise symmetric in (i, j) and independent of the sharding.
template <int UNROLL>
__global__ void __launch_bounds__(256) downdate_kernel(double *__restrict__ prec, int64_t ld, int64_t n_rows,
                                                       const double *__restrict__ pfull,
                                                       const double *__restrict__ ploc, int64_t c0,
                                                       const vgp_candidate *cur, int rows_per_block) {
    const int64_t y = cur->index;
    if (y < 0) return;
    const int64_t col = (int64_t)blockIdx.x * 512 + 2 * threadIdx.x;
    if (col >= ld) return;
    const double inv = 1.0 / pfull[y];
    const double pj0 = ploc[col], pj1 = ploc[col + 1];
    const int64_t yl = y - c0;
    const bool z0 = col == yl, z1 = col + 1 == yl;
    for (int64_t rb = (int64_t)blockIdx.y * rows_per_block; rb < n_rows; rb += (int64_t)gridDim.y * rows_per_block) {
        const int64_t rend = min(rb + rows_per_block, n_rows);
        for (int64_t r = rb; r < rend; r += UNROLL) {
            double2 v[UNROLL];
            double pi[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (r + u < rend) {
                    v[u] = *reinterpret_cast<const double2 *>(prec + (r + u) * ld + col);
                    pi[u] = pfull[r + u];
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (r + u < rend) {
                    double2 o;
                    o.x = fma(-(pi[u] * pj0), inv, v[u].x);
                    o.y = fma(-(pi[u] * pj1), inv, v[u].y);
                    if (r + u == y) o = make_double2(0.0, 0.0);
                    if (z0) o.x = 0.0;
                    if (z1) o.y = 0.0;
                    *reinterpret_cast<double2 *>(prec + (r + u) * ld + col) = o;
                }
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------------------
// Peer-memory exchange: the per-selection exchanges are done by the step kernels themselves
// with stores into the peers' mailboxes over NVLink/NVSwitch and system-scope release/acquire flags -- no collective
// launch, no host involvement; k selections are enqueued back to back on every rank, TWO launches per selection:
//
//   peer_step_kernel      scores of the local candidates -> local winner -> its record and its conditioning history
//                         W[0..t)[y] stored into every rank's mailbox -> wait for all ranks' records -> global winner
//                         (same rule on every rank) -> conditioning row of the LOCAL candidates (needs only the
//                         winner's history, so the rows themselves never travel) -> numerators updated -> this rank's
//                         segment of the precision row P[y, J] stored into every rank's mailbox;
//   downdate_peer_kernel  waits for all segments, then the rank-1 downdate of the local panel reading p straight from
//                         the mailbox.
//
// (The first version had five launches -- score, publish, exchange, unpack, downdate -- and sent [w_J | p_J]: 0.17 ms
// of a 0.95 ms selection at 8 GPUs, profiles/r01_bench_n50k_g8_final.json.)
//
// Mailbox of one rank (written by its peers, read by its own kernels):
//   [0, 512)        rec_flag[64]   u64: exchange sequence number of the record stored by rank q
//   [512, 1024)     seg_flag[64]   u64: same for the precision-row segment of rank q
//   [1024, 1032)    error          int: set when a wait timed out
//   [2048, 6144)    recs[2][64]    vgp_candidate, double-buffered by sequence parity
//   [8192, ...)     hist[2][nranks][kmax] doubles, then (256-byte aligned) segs[2][nranks][stride] doubles
// ------------------------------------------------------------------------------------------------------------
constexpr size_t MB_REC_FLAG = 0, MB_SEG_FLAG = 512, MB_ERROR = 1024, MB_RECS = 2048, MB_HIST = 8192;
constexpr long long SPIN_LIMIT_CYCLES = 20000000000LL;      // ~10 s at 1.9 GHz: a dead peer fails the run, no hang
constexpr int HIST_SMEM = 2048;                             // history entries staged in shared memory

struct PeerTable {
    char *mb[MAX_RANKS];
    int nranks, rank;
    int64_t stride, kmax;
    size_t seg_off;                 // byte offset of segs[] in a mailbox
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ vgp_candidate *mb_rec(char *mb, unsigned long long seq, int r) {
    return reinterpret_cast<vgp_candidate *>(mb + MB_RECS) + (seq & 1) * MAX_RANKS + r;
}
__device__ __forceinline__ double *mb_hist(const PeerTable &pt, char *mb, unsigned long long seq, int r) {
    return reinterpret_cast<double *>(mb + MB_HIST) + ((int64_t)(seq & 1) * pt.nranks + r) * pt.kmax;
}
__device__ __forceinline__ double *mb_seg(const PeerTable &pt, char *mb, unsigned long long seq, int r) {
    return reinterpret_cast<double *>(mb + pt.seg_off) + ((int64_t)(seq & 1) * pt.nranks + r) * pt.stride;
}
// Thread q < nranks waits until rank q's flag in the local mailbox reaches `seq`.
__device__ __forceinline__ void wait_flags(char *mb, size_t flag_off, int nranks, unsigned long long seq) {
    if ((int)threadIdx.x < nranks) {
        const unsigned long long *f = reinterpret_cast<const unsigned long long *>(mb + flag_off) + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < seq) {
            if (clock64() - t0 > SPIN_LIMIT_CYCLES) {
                atomicExch(reinterpret_cast<int *>(mb + MB_ERROR), 1 + (int)threadIdx.x);
                break;
            }
        }
    }
    __syncthreads();
}

struct StepArgs {
    const double *cov, *prec;
    int64_t ld, c0, nloc, n_pad, t;
    double *num;
    int *taken;
    double small_, jitter;
    vgp_candidate *partials, *cur;
    unsigned *counter, *counter2;
    int64_t *sel;
    double *sel_score, *step_row, *wfull, *ploc;
};

// One selection up to the point where the precision row is on its way to every rank.  At most one CTA per SM
// (grid <= SM count, grid-stride over the local candidates): every CTA waits for flags, so all must be resident.
__global__ void __launch_bounds__(256) peer_step_kernel(PeerTable pt, unsigned long long seq, StepArgs a) {
    __shared__ double hist[HIST_SMEM];
    __shared__ vgp_candidate win;
    __shared__ bool last;
    char *me = pt.mb[pt.rank];
    const int64_t span = pt.stride > a.ld ? pt.stride : a.ld;
    // ---- 1. scores of the local candidates, this CTA's first strict maximum (placement_algorithm2.py:105-125)
    double s = NEG_INF, nm = 0.0, pd = 0.0;
    int64_t idx = INT64_MAX;
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < a.nloc; j += (int64_t)gridDim.x * 256) {
        const double pdj = a.prec[(a.c0 + j) * a.ld + j], nmj = a.num[j];
        if (!a.taken[j]) {
            const double den = 1.0 / pdj - a.jitter;        // sigma^2(y | Abar \ y)
            const double nom = nmj - a.jitter;              // sigma^2(y | A)
            double d = nom / den;
            if (fabs(den) < a.small_ || fabs(nom) < a.small_) d = 0.0;    // :116-119
            if (a.step_row) a.step_row[j] = d;
            if (d > -1.0 && better(s, idx, d, a.c0 + j)) {  // running best starts at -1, strict '<' (:106,:121)
                s = d;
                idx = a.c0 + j;
                nm = nmj;
                pd = pdj;
            }
        } else if (a.step_row) {
            a.step_row[j] = nan("");
        }
    }
    {
        double rs = s;
        int64_t ri = idx;
        int slot = 0;
        block_argmax(rs, ri, slot);
        if (threadIdx.x == 0) {
            win.index = ri;
            win.score = rs;
        }
    }
    __syncthreads();
    if (win.index == INT64_MAX) {
        if (threadIdx.x == 0) a.partials[blockIdx.x] = vgp_candidate{NEG_INF, -1, 0.0, 0.0};
    } else if (idx == win.index) {
        a.partials[blockIdx.x] = vgp_candidate{win.score, idx, nm, pd};
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicInc(a.counter, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        // ---- 2. local winner over the CTAs' partials; publish its record and its history to every rank
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
        __shared__ vgp_candidate mine;
        if (threadIdx.x == 0) {
            vgp_candidate w{NEG_INF, -1, 0.0, 0.0};
            if (slot >= 0) {
                w.score = rs;
                w.index = ri;
                w.num = __ldcg(&a.partials[slot].num);
                w.pdiag = __ldcg(&a.partials[slot].pdiag);
            }
            mine = w;
        }
        __syncthreads();
        for (int q = 0; q < pt.nranks; ++q) {
            if (mine.index >= 0) {
                double *dst = mb_hist(pt, pt.mb[q], seq, pt.rank);
                for (int64_t h = threadIdx.x; h < a.t; h += 256) dst[h] = a.wfull[h * a.n_pad + mine.index];
            }
            if (threadIdx.x == 0) *mb_rec(pt.mb[q], seq, pt.rank) = mine;
        }
        // one system-scope fence per releasing thread, after the CTA barrier (fences are cumulative: the other threads'
        // stores, ordered before the barrier, are covered) -- 256 x membar.sys per site cost tens of microseconds
        __syncthreads();
        if ((int)threadIdx.x < pt.nranks) {
            __threadfence_system();
            st_release_sys(reinterpret_cast<unsigned long long *>(pt.mb[threadIdx.x] + MB_REC_FLAG) + pt.rank, seq);
        }
    }
    // ---- 3. every CTA: all ranks' records have arrived (own rank's included) -> the same winner everywhere
    wait_flags(me, MB_REC_FLAG, pt.nranks, seq);
    __shared__ int win_rank;
    if (threadIdx.x == 0) {
        vgp_candidate w{NEG_INF, -1, 0.0, 0.0};
        int wr = 0;
        for (int r = 0; r < pt.nranks; ++r) {
            const vgp_candidate *src = mb_rec(me, seq, r);
            vgp_candidate c;
            c.score = __ldcg(&src->score);
            c.index = __ldcg((const long long *)&src->index);
            c.num = __ldcg(&src->num);
            c.pdiag = __ldcg(&src->pdiag);
            if (c.index < 0) continue;
            if (w.index < 0 || c.score > w.score || (c.score == w.score && c.index < w.index)) {
                w = c;
                wr = r;
            }
        }
        win = w;
        win_rank = wr;
        if (blockIdx.x == 0) {
            *a.cur = w;
            a.sel[a.t] = w.index;
            a.sel_score[a.t] = w.score;
        }
    }
    __syncthreads();
    const int64_t y = win.index;
    const double *whist = mb_hist(pt, me, seq, win_rank);
    if (y >= 0)
        for (int64_t h = threadIdx.x; h < a.t && h < HIST_SMEM; h += 256) hist[h] = __ldcg(whist + h);
    __syncthreads();
    // ---- 4. conditioning row of the local candidates, numerators, precision-row segment to every rank
    const double root = y >= 0 ? sqrt(win.num) : 1.0;
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < span; j += (int64_t)gridDim.x * 256) {
        double p = 0.0;
        if (y >= 0 && j < a.nloc) {
            double acc = a.cov[y * a.ld + j];
            if (a.c0 + j == y) acc += a.jitter;
            const int64_t ts = a.t < HIST_SMEM ? a.t : HIST_SMEM;
            for (int64_t h = 0; h < ts; ++h) acc = fma(-a.wfull[h * a.n_pad + a.c0 + j], hist[h], acc);
            for (int64_t h = ts; h < a.t; ++h) acc = fma(-a.wfull[h * a.n_pad + a.c0 + j], __ldcg(whist + h), acc);
            const double w = acc / root;
            a.wfull[a.t * a.n_pad + a.c0 + j] = w;
            a.num[j] = fma(-w, w, a.num[j]);
            if (a.c0 + j == y) a.taken[j] = 1;
            p = a.prec[y * a.ld + j];
        }
        if (j < a.ld) a.ploc[j] = p;
        if (j < pt.stride)
            for (int q = 0; q < pt.nranks; ++q) mb_seg(pt, pt.mb[q], seq, pt.rank)[j] = p;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();                 // this CTA's peer stores (all threads', via the barrier) before the count
        last = atomicInc(a.counter2, gridDim.x - 1) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && (int)threadIdx.x < pt.nranks) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long *>(pt.mb[threadIdx.x] + MB_SEG_FLAG) + pt.rank, seq);
    }
}

// downdate_kernel with the precision row read from the mailbox: p_i = segment of the rank that owns row i, staged per
// row block in shared memory (the rank lookup inside the streaming loop cost 2 % of the kernel: 3.21 vs 3.14 ms at
// n = 50 000 on 2 GPUs).
template <int UNROLL>
__global__ void __launch_bounds__(256) downdate_peer_kernel(double *__restrict__ prec, int64_t ld, int64_t n_rows,
                                                            int64_t n, PeerTable pt, unsigned long long seq,
                                                            Bounds bounds, const double *__restrict__ ploc, int64_t c0,
                                                            const vgp_candidate *cur, int rows_per_block) {
    __shared__ double sp[64];
    char *me = pt.mb[pt.rank];
    wait_flags(me, MB_SEG_FLAG, pt.nranks, seq);
    const int64_t y = cur->index;
    if (y < 0) return;
    const int64_t col = (int64_t)blockIdx.x * 512 + 2 * threadIdx.x;
    const bool active = col < ld;                        // inactive threads still take part in the barriers
    const double inv = 1.0 / cur->pdiag;                 // = P[y][y], the same bits on every rank
    const double pj0 = active ? ploc[col] : 0.0, pj1 = active ? ploc[col + 1] : 0.0;
    const int64_t yl = y - c0;
    const bool z0 = col == yl, z1 = col + 1 == yl;
    const double *segs = mb_seg(pt, me, seq, 0);
    for (int64_t rb = (int64_t)blockIdx.y * rows_per_block; rb < n_rows; rb += (int64_t)gridDim.y * rows_per_block) {
        const int64_t rend = min(rb + rows_per_block, n_rows);
        __syncthreads();
        if ((int64_t)threadIdx.x < rend - rb) {
            const int64_t row = rb + threadIdx.x;
            int g = 0;
            while (g + 1 < bounds.nranks && row >= bounds.b[g + 1]) ++g;
            sp[threadIdx.x] = row < n ? __ldcg(segs + (int64_t)g * pt.stride + (row - bounds.b[g])) : 0.0;
        }
        __syncthreads();
        if (!active) continue;
        for (int64_t r = rb; r < rend; r += UNROLL) {
            double2 v[UNROLL];
            double pi[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (r + u < rend) {
                    v[u] = *reinterpret_cast<const double2 *>(prec + (r + u) * ld + col);
                    pi[u] = sp[r + u - rb];
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (r + u < rend) {
                    double2 o;
                    o.x = fma(-(pi[u] * pj0), inv, v[u].x);
                    o.y = fma(-(pi[u] * pj1), inv, v[u].y);
                    if (r + u == y) o = make_double2(0.0, 0.0);
                    if (z0) o.x = 0.0;
                    if (z1) o.y = 0.0;
                    *reinterpret_cast<double2 *>(prec + (r + u) * ld + col) = o;
                }
            }
        }
    }
}

// num = diag(Sigma) (+ jitter); nothing taken
__global__ void __launch_bounds__(256) reset_kernel(const double *__restrict__ cov, int64_t ld, int64_t c0,
                                                    int64_t nloc, double jitter, double *num, int *taken,
                                                    double *ploc) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= ld) return;
    num[j] = j < nloc ? cov[(c0 + j) * ld + j] + jitter : 0.0;
    taken[j] = j < nloc ? 0 : 1;
    ploc[j] = 0.0;
}

// identity on the padding block of a square padded matrix: a[i][i] = 1 for i in [n, n_pad)
__global__ void pad_identity_kernel(double *a, int64_t ld, int64_t n, int64_t n_pad) {
    const int64_t i = n + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) a[i * ld + i] = 1.0;
}

int check_handle(vgp_greedy *h) {
    if (!h) {
        set_error("greedy handle is NULL");
        return VGP_ERR_INVALID;
    }
    return VGP_OK;
}

}  // namespace

#define H_LAUNCH_CHECK(h)      \
    do {                       \
        ++(h)->launches;       \
        VGP_LAUNCH_CHECK();    \
    } while (0)

// The dominant kernel: one streaming read-modify-write pass over the local precision panel.
static int launch_downdate(vgp_greedy *h, cudaStream_t s) {
    // One CTA per (8 rows x 512 columns), launched in address order, NO grid-stride loop: measured on the 20 GB panel
    // (tools/downdate_sweep.cu, profiles/r02_downdate_sweep.log) 5.73 ms = 6.99 TB/s against 6.31 ms = 6.35 TB/s for
    // 4 waves of CTAs striding over row blocks of 32 -- the block scheduler then walks the panel linearly and the
    // write of a DRAM page follows its read closely; CTAs that loop drift apart and keep more pages open.
    constexpr int rows_per_block = 8, unroll = 4;
    const unsigned gx = (unsigned)((h->ld + 511) / 512);
    const int64_t row_tiles = (h->n_pad + rows_per_block - 1) / rows_per_block;
    int64_t gy = row_tiles;
    if (gy > 65535) gy = 65535;                                   // beyond that the kernel's row loop takes over
    if (gy < 1) gy = 1;
    dim3 grid(gx, (unsigned)gy);
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    if (h->profile) {
        VGP_CUDA(cudaEventCreate(&pe0));
        VGP_CUDA(cudaEventCreate(&pe1));
        VGP_CUDA(cudaEventRecord(pe0, s));
    }
    downdate_kernel<unroll><<<grid, 256, 0, s>>>(h->prec, h->ld, h->n_pad, h->pfull, h->ploc, h->c0, h->cur, rows_per_block);
    H_LAUNCH_CHECK(h);
    if (h->profile) {
        VGP_CUDA(cudaEventRecord(pe1, s));
        h->prof_events.emplace_back(pe0, pe1);
    }
    return VGP_OK;
}

extern "C" {

int vgp_greedy_create(vgp_greedy **handle, int device, int64_t n, int64_t c0, int64_t nloc, int64_t kmax,
                      double small, double jitter) {
    VGP_REQUIRE(handle, "handle is NULL");
    *handle = nullptr;
    VGP_REQUIRE(n > 0 && nloc > 0 && c0 >= 0 && c0 + nloc <= n, "bad shard [%lld, %lld) of %lld", (long long)c0,
                (long long)(c0 + nloc), (long long)n);
    VGP_REQUIRE(kmax > 0 && kmax <= n, "kmax %lld outside [1, n]", (long long)kmax);
    VGP_ENTER(device);
    vgp_greedy *h = new (std::nothrow) vgp_greedy();
    VGP_REQUIRE(h, "out of host memory");
    h->device = device;
    h->n = n;
    h->c0 = c0;
    h->nloc = nloc;
    h->kmax = kmax;
    h->small_ = small;
    h->jitter = jitter;
    h->n_pad = round_up(n, TILE);
    h->ld = (c0 == 0 && nloc == n) ? h->n_pad : round_up(nloc, TILE);
    h->score_blocks = (int)((nloc + 255) / 256);
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) h->sm_count = prop.multiProcessorCount;
    const size_t panel = (size_t)h->n_pad * h->ld * sizeof(double);
    struct {
        void **p;
        size_t bytes;
    } allocs[] = {
        {(void **)&h->cov, panel},
        {(void **)&h->prec, panel},
        {(void **)&h->num, (size_t)h->ld * 8},
        {(void **)&h->ploc, (size_t)h->ld * 8},
        {(void **)&h->taken, (size_t)h->ld * 4},
        {(void **)&h->wfull, (size_t)kmax * h->n_pad * 8},
        {(void **)&h->pfull, (size_t)h->n_pad * 8},
        {(void **)&h->seg, (size_t)2 * h->n_pad * 8},
        {(void **)&h->partials, (size_t)(h->score_blocks + 256) * sizeof(vgp_candidate)},   // + the peer step's CTAs
        {(void **)&h->cur, sizeof(vgp_candidate)},
        {(void **)&h->best, sizeof(vgp_candidate)},
        {(void **)&h->counter, sizeof(unsigned)},
        {(void **)&h->sel, (size_t)kmax * 8},
        {(void **)&h->sel_score, (size_t)kmax * 8},
    };
    for (auto &a : allocs) {
        // the two panels come from (and go back to) the per-device workspace cache: see common.cuh
        const bool is_panel = a.p == (void **)&h->cov || a.p == (void **)&h->prec;
        e = is_panel ? cache_alloc(a.p, a.bytes) : device_malloc(a.p, a.bytes);
        if (e != cudaSuccess) {
            int rc = cuda_fail(e, "cudaMalloc (greedy state)", __FILE__, __LINE__);
            vgp_greedy_destroy(h);
            return rc;
        }
    }
    cudaMemset(h->cov, 0, panel);
    cudaMemset(h->prec, 0, panel);
    cudaMemset(h->counter, 0, sizeof(unsigned));
    cudaMemset(h->pfull, 0, (size_t)h->n_pad * 8);
    cudaMemset(h->wfull, 0, (size_t)kmax * h->n_pad * 8);
    cudaMemset(h->sel, 0xff, (size_t)kmax * 8);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        int rc = cuda_fail(e, "greedy state init", __FILE__, __LINE__);
        vgp_greedy_destroy(h);
        return rc;
    }
    *handle = h;
    return VGP_OK;
}

int vgp_greedy_destroy(vgp_greedy *h) {
    if (!h) return VGP_OK;
    VGP_ENTER(h->device);
    void *ptrs[] = {h->num, h->ploc, h->taken, h->wfull, h->pfull, h->seg,
                    h->partials, h->cur, h->best, h->counter, h->sel, h->sel_score, h->step_scores};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (void *p : {(void *)h->cov, (void *)h->prec, (void *)h->prec_saved}) cache_free(p);
    if (h->comm_ipc)
        for (int q = 0; q < h->comm_nranks; ++q)
            if (q != h->comm_rank && h->peer_mb[q]) cudaIpcCloseMemHandle(h->peer_mb[q]);
    if (h->mailbox) cudaFree(h->mailbox);
    if (h->counter2) cudaFree(h->counter2);
    h->ws.release();
    delete h;
    return VGP_OK;
}

int vgp_greedy_panels(vgp_greedy *h, double **cov_dev, double **prec_dev, int64_t *ld, int64_t *n_pad) {
    VGP_TRY(check_handle(h));
    if (cov_dev) *cov_dev = h->cov;
    if (prec_dev) *prec_dev = h->prec;
    if (ld) *ld = h->ld;
    if (n_pad) *n_pad = h->n_pad;
    return VGP_OK;
}

int vgp_greedy_reset(vgp_greedy *h, void *stream) {
    VGP_TRY(check_handle(h));
    VGP_ENTER(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    reset_kernel<<<(unsigned)((h->ld + 255) / 256), 256, 0, s>>>(h->cov, h->ld, h->c0, h->nloc, h->jitter, h->num,
                                                               h->taken, h->ploc);
    H_LAUNCH_CHECK(h);
    VGP_CUDA(cudaMemsetAsync(h->sel, 0xff, (size_t)h->kmax * 8, s));
    h->t = 0;
    return VGP_OK;
}

int vgp_greedy_factor(vgp_greedy *h, int *info_host, void *stream) {
    VGP_TRY(check_handle(h));
    VGP_REQUIRE(h->c0 == 0 && h->nloc == h->n, "vgp_greedy_factor needs a single shard holding all columns");
    VGP_ENTER(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t panel = (size_t)h->n_pad * h->ld * sizeof(double);
    VGP_CUDA(cudaMemcpyAsync(h->prec, h->cov, panel, cudaMemcpyDeviceToDevice, s));
    if (h->jitter != 0.0) {
        VGP_TRY(dense_add_diag(h->prec, h->ld, h->n, h->jitter, s));
        ++h->launches;
    }
    if (h->n_pad > h->n) {
        const int64_t extra = h->n_pad - h->n;
        pad_identity_kernel<<<(unsigned)((extra + 127) / 128), 128, 0, s>>>(h->prec, h->ld, h->n, h->n_pad);
        H_LAUNCH_CHECK(h);
    }
    const int64_t before = g_launches;
    int rc = dense_spd_inverse(h->prec, h->n_pad, h->ld, h->ws, info_host, s);
    h->launches += g_launches - before;
    VGP_TRY(rc);
    return vgp_greedy_reset(h, stream);
}

int vgp_greedy_save_precision(vgp_greedy *h, void *stream) {
    VGP_TRY(check_handle(h));
    VGP_ENTER(h->device);
    const size_t panel = (size_t)h->n_pad * h->ld * sizeof(double);
    if (!h->prec_saved) VGP_CUDA(cache_alloc((void **)&h->prec_saved, panel));
    VGP_CUDA(cudaMemcpyAsync(h->prec_saved, h->prec, panel, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return VGP_OK;
}

int vgp_greedy_restore_precision(vgp_greedy *h, void *stream) {
    VGP_TRY(check_handle(h));
    if (!h->prec_saved) {
        set_error("no saved precision panel");
        return VGP_ERR_STATE;
    }
    VGP_ENTER(h->device);
    const size_t panel = (size_t)h->n_pad * h->ld * sizeof(double);
    VGP_CUDA(cudaMemcpyAsync(h->prec, h->prec_saved, panel, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return vgp_greedy_reset(h, stream);
}

int vgp_greedy_local_best(vgp_greedy *h, vgp_candidate *best_dev, void *stream) {
    VGP_TRY(check_handle(h));
    VGP_REQUIRE(best_dev, "best_dev is NULL");
    if (h->t >= h->kmax) {
        set_error("greedy handle already holds kmax = %lld selections", (long long)h->kmax);
        return VGP_ERR_STATE;
    }
    VGP_ENTER(h->device);
    double *row = (h->record && h->step_scores) ? h->step_scores + h->t * h->nloc : nullptr;
    score_kernel<<<h->score_blocks, 256, 0, (cudaStream_t)stream>>>(h->prec, h->ld, h->c0, h->nloc, h->num, h->taken,
                                                                    h->small_, h->jitter, h->partials, h->counter,
                                                                    best_dev, row);
    H_LAUNCH_CHECK(h);
    return VGP_OK;
}

int vgp_greedy_select(vgp_greedy *h, const vgp_candidate *records_dev, int nrecords, void *stream) {
    VGP_TRY(check_handle(h));
    VGP_REQUIRE(records_dev && nrecords > 0, "no candidate records");
    VGP_ENTER(h->device);
    select_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(records_dev, nrecords, h->cur, h->sel, h->sel_score, h->t);
    H_LAUNCH_CHECK(h);
    return VGP_OK;
}

int vgp_greedy_segments(vgp_greedy *h, double *seg_dev, int64_t seg_stride, void *stream) {
    VGP_TRY(check_handle(h));
    VGP_REQUIRE(seg_dev && seg_stride >= h->nloc, "segment buffer too small");
    VGP_ENTER(h->device);
    segments_kernel<<<(unsigned)((seg_stride + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        h->cov, h->prec, h->ld, h->c0, h->nloc, h->wfull, h->n_pad, h->t, h->jitter, h->cur, seg_dev, seg_stride);
    H_LAUNCH_CHECK(h);
    return VGP_OK;
}

int vgp_greedy_apply(vgp_greedy *h, const double *gathered_dev, int64_t seg_stride, int nranks,
                     const int64_t *bounds_host, void *stream) {
    VGP_TRY(check_handle(h));
    VGP_REQUIRE(gathered_dev && bounds_host, "NULL argument");
    VGP_REQUIRE(nranks >= 1 && nranks <= MAX_RANKS, "nranks %d outside [1, %d]", nranks, MAX_RANKS);
    VGP_REQUIRE(bounds_host[0] == 0 && bounds_host[nranks] == h->n, "bounds do not cover [0, n)");
    for (int g = 0; g < nranks; ++g)
        VGP_REQUIRE(bounds_host[g + 1] - bounds_host[g] <= seg_stride && bounds_host[g + 1] >= bounds_host[g],
                    "rank %d segment does not fit seg_stride", g);
    VGP_ENTER(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    Bounds b;
    b.nranks = nranks;
    for (int g = 0; g <= nranks; ++g) b.b[g] = bounds_host[g];
    const int64_t span = h->n_pad > h->ld ? h->n_pad : h->ld;
    unpack_kernel<<<(unsigned)((span + 255) / 256), 256, 0, s>>>(gathered_dev, seg_stride, b, h->n, h->n_pad, h->c0,
                                                               h->nloc, h->ld, h->wfull + h->t * h->n_pad, h->pfull,
                                                               h->ploc, h->num, h->taken, h->cur);
    H_LAUNCH_CHECK(h);
    VGP_TRY(launch_downdate(h, s));
    ++h->t;
    return VGP_OK;
}

int vgp_greedy_profile(vgp_greedy *h, int enable) {
    VGP_TRY(check_handle(h));
    VGP_ENTER(h->device);
    for (auto &pr : h->prof_events) {
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    h->prof_events.clear();
    for (auto &pr : h->prof_step_events) cudaEventDestroy(pr.first);       // .second is a downdate event, freed above
    h->prof_step_events.clear();
    h->profile = enable ? 1 : 0;
    return VGP_OK;
}

/* Peer path: summed duration of the peer_step_kernel launches since profiling was enabled (includes its waits for the
 * other ranks' records); the downdate_peer_kernel figure of vgp_greedy_profile_read includes its wait for the segments. */
int vgp_greedy_profile_step_ms(vgp_greedy *h, double *total_ms) {
    VGP_TRY(check_handle(h));
    VGP_REQUIRE(total_ms, "NULL argument");
    VGP_ENTER(h->device);
    double sum = 0.0;
    for (auto &pr : h->prof_step_events) {
        float ms = 0.f;
        VGP_CUDA(cudaEventSynchronize(pr.second));
        VGP_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
        sum += ms;
    }
    *total_ms = sum;
    return VGP_OK;
}

int vgp_greedy_profile_read(vgp_greedy *h, double *total_ms, int64_t *launches) {
    VGP_TRY(check_handle(h));
    VGP_REQUIRE(total_ms && launches, "NULL argument");
    VGP_ENTER(h->device);
    double sum = 0.0;
    for (auto &pr : h->prof_events) {
        float ms = 0.f;
        VGP_CUDA(cudaEventSynchronize(pr.second));
        VGP_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
        sum += ms;
    }
    *total_ms = sum;
    *launches = (int64_t)h->prof_events.size();
    return VGP_OK;
}

int vgp_greedy_run(vgp_greedy *h, int64_t k, void *stream) {
    VGP_TRY(check_handle(h));
    VGP_REQUIRE(h->c0 == 0 && h->nloc == h->n, "vgp_greedy_run needs a single shard; drive shards step by step");
    VGP_REQUIRE(k >= 0 && h->t + k <= h->kmax, "k = %lld exceeds kmax = %lld (already %lld)", (long long)k,
                (long long)h->kmax, (long long)h->t);
    const int64_t bounds[2] = {0, h->n};
    for (int64_t i = 0; i < k; ++i) {
        VGP_TRY(vgp_greedy_local_best(h, h->best, stream));
        VGP_TRY(vgp_greedy_select(h, h->best, 1, stream));
        VGP_TRY(vgp_greedy_segments(h, h->seg, h->n_pad, stream));
        VGP_TRY(vgp_greedy_apply(h, h->seg, h->n_pad, 1, bounds, stream));
    }
    return VGP_OK;
}

// ---- peer-memory exchange --------------------------------------------------------------------------------
int vgp_greedy_comm_create(vgp_greedy *h, int rank, int nranks, const int64_t *bounds_host, void *ipc_handle_out,
                           void **mailbox_dev_out) {
    VGP_TRY(check_handle(h));
    VGP_REQUIRE(nranks >= 1 && nranks <= MAX_RANKS && rank >= 0 && rank < nranks, "bad rank %d of %d", rank, nranks);
    VGP_REQUIRE(bounds_host && bounds_host[0] == 0 && bounds_host[nranks] == h->n, "bounds do not cover [0, n)");
    VGP_REQUIRE(bounds_host[rank] == h->c0 && bounds_host[rank + 1] == h->c0 + h->nloc,
                "bounds[%d] do not match this handle's shard", rank);
    if (h->mailbox) {
        set_error("the handle already has a mailbox");
        return VGP_ERR_STATE;
    }
    VGP_ENTER(h->device);
    int64_t stride = 0;
    for (int g = 0; g < nranks; ++g) {
        VGP_REQUIRE(bounds_host[g + 1] >= bounds_host[g], "bounds not monotone");
        if (bounds_host[g + 1] - bounds_host[g] > stride) stride = bounds_host[g + 1] - bounds_host[g];
    }
    stride = round_up(stride, 2);
    h->comm_rank = rank;
    h->comm_nranks = nranks;
    h->comm_stride = stride;
    for (int g = 0; g <= nranks; ++g) h->comm_bounds[g] = bounds_host[g];
    h->comm_seg_off = (MB_HIST + (size_t)2 * nranks * h->kmax * 8 + 255) / 256 * 256;
    h->mailbox_bytes = h->comm_seg_off + (size_t)2 * nranks * stride * 8;
    VGP_CUDA(device_malloc((void **)&h->mailbox, h->mailbox_bytes));
    VGP_CUDA(cudaMemset(h->mailbox, 0, h->mailbox_bytes));
    if (!h->counter2) {
        VGP_CUDA(device_malloc((void **)&h->counter2, sizeof(unsigned)));
        VGP_CUDA(cudaMemset(h->counter2, 0, sizeof(unsigned)));
    }
    VGP_CUDA(cudaDeviceSynchronize());
    {   // Load every kernel of the exchange loop now: with lazy module loading a first launch can wait for the
        // device to drain, which must not happen while one of these kernels is spinning on a peer's flag.
        cudaFuncAttributes fa;
        VGP_CUDA(cudaFuncGetAttributes(&fa, peer_step_kernel));
        VGP_CUDA(cudaFuncGetAttributes(&fa, downdate_peer_kernel<4>));
    }
    if (ipc_handle_out) {
        cudaIpcMemHandle_t ipc;
        VGP_CUDA(cudaIpcGetMemHandle(&ipc, h->mailbox));
        memcpy(ipc_handle_out, &ipc, sizeof ipc);
    }
    if (mailbox_dev_out) *mailbox_dev_out = h->mailbox;
    return VGP_OK;
}

int vgp_greedy_comm_connect(vgp_greedy *h, const void *peers, int kind) {
    VGP_TRY(check_handle(h));
    VGP_REQUIRE(peers, "peers is NULL");
    if (!h->mailbox) {
        set_error("vgp_greedy_comm_create first");
        return VGP_ERR_STATE;
    }
    VGP_REQUIRE(kind == 0 || kind == 1, "kind must be 0 (device pointers) or 1 (IPC handles)");
    VGP_ENTER(h->device);
    for (int q = 0; q < h->comm_nranks; ++q) {
        if (q == h->comm_rank) {
            h->peer_mb[q] = h->mailbox;
        } else if (kind == 0) {
            h->peer_mb[q] = (char *)((void *const *)peers)[q];
            VGP_REQUIRE(h->peer_mb[q], "peer %d mailbox pointer is NULL", q);
        } else {
            cudaIpcMemHandle_t ipc;
            memcpy(&ipc, (const char *)peers + (size_t)q * sizeof ipc, sizeof ipc);
            void *ptr = nullptr;
            VGP_CUDA(cudaIpcOpenMemHandle(&ptr, ipc, cudaIpcMemLazyEnablePeerAccess));
            h->peer_mb[q] = (char *)ptr;
        }
    }
    h->comm_ipc = kind;
    return VGP_OK;
}

int vgp_greedy_run_peer(vgp_greedy *h, int64_t k, void *stream) {
    VGP_TRY(check_handle(h));
    if (!h->mailbox || !h->peer_mb[h->comm_rank]) {
        set_error("vgp_greedy_comm_create / vgp_greedy_comm_connect first");
        return VGP_ERR_STATE;
    }
    VGP_REQUIRE(k >= 0 && h->t + k <= h->kmax, "k = %lld exceeds kmax = %lld (already %lld)", (long long)k,
                (long long)h->kmax, (long long)h->t);
    VGP_ENTER(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    PeerTable pt;
    for (int q = 0; q < MAX_RANKS; ++q) pt.mb[q] = q < h->comm_nranks ? h->peer_mb[q] : nullptr;
    pt.nranks = h->comm_nranks;
    pt.rank = h->comm_rank;
    pt.stride = h->comm_stride;
    pt.kmax = h->kmax;
    pt.seg_off = h->comm_seg_off;
    Bounds b;
    b.nranks = h->comm_nranks;
    for (int g = 0; g <= h->comm_nranks; ++g) b.b[g] = h->comm_bounds[g];
    const int64_t span = h->comm_stride > h->ld ? h->comm_stride : h->ld;
    int64_t step_blocks = (span + 255) / 256;
    if (step_blocks > h->sm_count) step_blocks = h->sm_count;       // every CTA waits on flags: all must be resident
    if (step_blocks > 256) step_blocks = 256;                       // size of the partials buffer beyond score_blocks
    // one CTA per row block, no grid-stride loop (see launch_downdate: +7.5 % at 32 rows per block on the stand-alone
    // kernel); 32 rows per CTA keep the per-CTA flag check and the staging of p_i small against 128 KB of traffic
    constexpr int rows_per_block = 32;
    const unsigned gx = (unsigned)((h->ld + 511) / 512);
    const int64_t row_tiles = (h->n_pad + rows_per_block - 1) / rows_per_block;
    int64_t gy = row_tiles;
    if (gy > 65535) gy = 65535;
    if (gy < 1) gy = 1;
    for (int64_t i = 0; i < k; ++i) {
        const unsigned long long seq = ++h->epoch;
        StepArgs a;
        a.cov = h->cov;
        a.prec = h->prec;
        a.ld = h->ld;
        a.c0 = h->c0;
        a.nloc = h->nloc;
        a.n_pad = h->n_pad;
        a.t = h->t;
        a.num = h->num;
        a.taken = h->taken;
        a.small_ = h->small_;
        a.jitter = h->jitter;
        a.partials = h->partials;
        a.cur = h->cur;
        a.counter = h->counter;
        a.counter2 = h->counter2;
        a.sel = h->sel;
        a.sel_score = h->sel_score;
        a.step_row = (h->record && h->step_scores) ? h->step_scores + h->t * h->nloc : nullptr;
        a.wfull = h->wfull;
        a.ploc = h->ploc;
        cudaEvent_t pe0 = nullptr, pe1 = nullptr, se0 = nullptr;
        if (h->profile) {
            VGP_CUDA(cudaEventCreate(&se0));
            VGP_CUDA(cudaEventRecord(se0, s));
        }
        peer_step_kernel<<<(unsigned)step_blocks, 256, 0, s>>>(pt, seq, a);
        H_LAUNCH_CHECK(h);
        if (h->profile) {
            VGP_CUDA(cudaEventCreate(&pe0));
            VGP_CUDA(cudaEventCreate(&pe1));
            VGP_CUDA(cudaEventRecord(pe0, s));
            h->prof_step_events.emplace_back(se0, pe0);
        }
        downdate_peer_kernel<4><<<dim3(gx, (unsigned)gy), 256, 0, s>>>(h->prec, h->ld, h->n_pad, h->n, pt, seq, b, h->ploc,
                                                                       h->c0, h->cur, rows_per_block);
        H_LAUNCH_CHECK(h);
        if (h->profile) {
            VGP_CUDA(cudaEventRecord(pe1, s));
            h->prof_events.emplace_back(pe0, pe1);
        }
        ++h->t;
    }
    return VGP_OK;
}

int vgp_greedy_comm_status(vgp_greedy *h, int *error_host, void *stream) {
    VGP_TRY(check_handle(h));
    VGP_REQUIRE(error_host, "error_host is NULL");
    *error_host = 0;
    if (!h->mailbox) return VGP_OK;
    VGP_ENTER(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    VGP_CUDA(cudaMemcpyAsync(error_host, h->mailbox + MB_ERROR, sizeof(int), cudaMemcpyDeviceToHost, s));
    VGP_CUDA(cudaStreamSynchronize(s));
    if (*error_host != 0) {
        set_error("peer exchange timed out waiting for rank %d (a peer died or ran a different number of steps)",
                  *error_host - 1);
        return VGP_ERR_STATE;
    }
    return VGP_OK;
}

int vgp_greedy_results(vgp_greedy *h, int64_t *count, int64_t *selection_host, double *scores_host,
                       int64_t capacity, void *stream) {
    VGP_TRY(check_handle(h));
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

int vgp_greedy_record_scores(vgp_greedy *h, int enable) {
    VGP_TRY(check_handle(h));
    VGP_ENTER(h->device);
    if (enable && !h->step_scores)
        VGP_CUDA(device_malloc((void **)&h->step_scores, (size_t)h->kmax * h->nloc * 8));
    h->record = enable ? 1 : 0;
    return VGP_OK;
}

int vgp_greedy_step_scores(vgp_greedy *h, double *scores_host, int64_t capacity_rows, void *stream) {
    VGP_TRY(check_handle(h));
    VGP_REQUIRE(scores_host, "scores_host is NULL");
    if (!h->step_scores) {
        set_error("score recording was not enabled");
        return VGP_ERR_STATE;
    }
    VGP_ENTER(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t rows = h->t < capacity_rows ? h->t : capacity_rows;
    if (rows > 0)
        VGP_CUDA(cudaMemcpyAsync(scores_host, h->step_scores, (size_t)rows * h->nloc * 8, cudaMemcpyDeviceToHost, s));
    VGP_CUDA(cudaStreamSynchronize(s));
    return VGP_OK;
}

int vgp_greedy_launch_count(vgp_greedy *h, int64_t *launches) {
    VGP_TRY(check_handle(h));
    VGP_REQUIRE(launches, "launches is NULL");
    *launches = h->launches;
    return VGP_OK;
}

// formulation 0 of vgp_placement_host_ex (lazy.cu): the dense precision downdate
int vgp_placement_host_dense(int device, const double *cov_host, int64_t n, int64_t ld_host, int64_t k, double small,
                             double jitter, int64_t *selection_host, double *scores_host, double *step_scores_host,
                             double *seconds_host) {
    VGP_REQUIRE(cov_host && selection_host, "NULL argument");
    VGP_REQUIRE(n > 0 && ld_host >= n && k > 0 && k <= n, "bad sizes n=%lld ld=%lld k=%lld", (long long)n,
                (long long)ld_host, (long long)k);
    VGP_ENTER(device);
    vgp_greedy *h = nullptr;
    VGP_TRY(vgp_greedy_create(&h, device, n, 0, n, k, small, jitter));
    cudaStream_t s = nullptr;
    cudaEvent_t ev[4];
    for (auto &e : ev) cudaEventCreate(&e);
    int rc = VGP_OK;
    auto fail = [&](int code) {
        dense_set_pivot_floor(0.0);
        for (auto &e : ev) cudaEventDestroy(e);
        vgp_greedy_destroy(h);
        return code;
    };
    {   // numerical rank deficiency counts as "not positive definite": pivots below 1e-12 of the largest variance
        double scale = 0.0;
        for (int64_t i = 0; i < n; ++i) scale = std::max(scale, fabs(cov_host[i * ld_host + i]));
        dense_set_pivot_floor(1e-12 * scale);
    }
    cudaEventRecord(ev[0], s);
    cudaError_t ce = cudaMemcpy2DAsync(h->cov, (size_t)h->ld * 8, cov_host, (size_t)ld_host * 8, (size_t)n * 8,
                                       (size_t)n, cudaMemcpyHostToDevice, s);
    if (ce != cudaSuccess) return fail(cuda_fail(ce, "H2D of cov_vv", __FILE__, __LINE__));
    cudaEventRecord(ev[1], s);
    int info = 0;
    rc = vgp_greedy_factor(h, &info, s);
    if (rc != VGP_OK) return fail(rc);
    cudaEventRecord(ev[2], s);
    if (step_scores_host) {
        rc = vgp_greedy_record_scores(h, 1);
        if (rc != VGP_OK) return fail(rc);
    }
    rc = vgp_greedy_run(h, k, s);
    if (rc != VGP_OK) return fail(rc);
    int64_t count = 0;
    rc = vgp_greedy_results(h, &count, selection_host, scores_host, k, s);
    if (rc != VGP_OK) return fail(rc);
    if (step_scores_host) {
        rc = vgp_greedy_step_scores(h, step_scores_host, k, s);
        if (rc != VGP_OK) return fail(rc);
    }
    cudaEventRecord(ev[3], s);
    ce = cudaEventSynchronize(ev[3]);
    if (ce != cudaSuccess) return fail(cuda_fail(ce, "placement", __FILE__, __LINE__));
    if (seconds_host) {
        float ms;
        cudaEventElapsedTime(&ms, ev[0], ev[1]);
        seconds_host[0] = ms * 1e-3;
        cudaEventElapsedTime(&ms, ev[1], ev[2]);
        seconds_host[1] = ms * 1e-3;
        cudaEventElapsedTime(&ms, ev[2], ev[3]);
        seconds_host[2] = ms * 1e-3;
        cudaEventElapsedTime(&ms, ev[0], ev[3]);
        seconds_host[3] = ms * 1e-3;
    }
    return fail(VGP_OK);
}

}  // extern "C"
