// placement_algorithm_1/2 on positive SEMI-definite, rank-deficient covariances: the reference's pinv semantics.
//
// The reference takes np.linalg.pinv of Sigma_AA and of Sigma_{Abar \ y} (placement_algorithm2.py:371-413, call_pinv
// :399-405), and the cov_vv it really feeds is an empirical covariance M M^T / S of rank <= S over n > S locations
// (main_architecture_2.py:391-444, gp_functions.py:1019-1057).  On such a matrix the SPD formulation (greedy.cu,
// lazy.cu: sigma^2(y | Abar \ y) = 1 / (Sigma^-1)_yy) does not exist.  With Sigma = F F^T, F [n, r] of full column
// rank r (pivoted Cholesky, stopped at 1e-12 of the largest diagonal entry), every conditional variance is a squared
// distance in the r-dimensional factor space:
//
//   sigma^2(y | A)        = |f_y - proj_{span F_A} f_y|^2          (numerator;  pinv == projector on span F_A)
//   sigma^2(y | Abar \ y) = 0                      if f_y lies in span F_{Abar \ y}   <=>  leverage h_y < 1
//                         = 1 / (Sigma_AbarAbar^+)_yy   otherwise (y is essential for the span)    (denominator)
//   h_y = f_y^T G^-1 f_y,   (Sigma_AbarAbar^+)_yy = f_y^T G^-2 f_y,   G = sum_{i in Abar} f_i f_i^T   (r x r, SPD)
//
// so the whole step is O(n r^2): G by one product, its Cholesky, two triangular solves with all f_y as right-hand
// sides, column norms.  The numerator keeps the growing conditioning panel W of the SPD path (a selected point whose
// numerator is numerically zero adds no row: it is already in the span, which is what pinv of a singular Sigma_AA
// does).  When G loses rank (only possible once card Abar is down to about r) the factor is rebuilt on Abar.
// In the reference's own regime (r <= S << n) every denominator is zero, the guard (:116-119) sets every delta to 0
// and the first strict maximum above -1 is the lowest free index: the reference returns [0, 1, ..., k - 1] there
// (golden vectors from its unmodified code: tests/golden/greedy_golden.json, cases "lowrank_*").
#include <math.h>

#include <algorithm>
#include <vector>

#include "dense.cuh"

using namespace vgp;

namespace {

constexpr double RANK_TOL = 1e-12;      // relative to the largest diagonal entry: numerical rank of the factor
constexpr double LEVERAGE_TOL = 1e-6;   // h_y >= 1 - this: y is essential for the span of Abar

struct Best {
    double value;
    long long index;
};

// first index of the strict maximum of v over entries with skip[i] == 0 and v > floor_value; one block
__global__ void __launch_bounds__(1024) first_argmax_kernel(const double *v, const int *skip, int64_t n, double floor_value,
                                                            Best *out) {
    __shared__ double sv[32];
    __shared__ long long si[32];
    double bv = floor_value;
    long long bi = -1;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        if (skip && skip[i]) continue;
        const double x = v[i];
        if (x > bv) {               // strict: within a thread indices ascend, so the first maximal index wins
            bv = x;
            bi = i;
        }
    }
    auto better = [](double av, long long ai, double cv, long long ci) {
        if (ci < 0) return false;
        if (ai < 0) return true;
        return cv > av || (cv == av && ci < ai);
    };
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(bv, bi, ov, oi)) {
            bv = ov;
            bi = oi;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        sv[warp] = bv;
        si[warp] = bi;
    }
    __syncthreads();
    if (warp == 0) {
        bv = lane < (int)(blockDim.x >> 5) ? sv[lane] : floor_value;
        bi = lane < (int)(blockDim.x >> 5) ? si[lane] : -1;
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (better(bv, bi, ov, oi)) {
                bv = ov;
                bi = oi;
            }
        }
        if (lane == 0) {
            out->value = bv;
            out->index = bi;
        }
    }
}

// One step of the conditioning recurrence shared by the pivoted Cholesky and the numerator update:
//   row[i] = (cov[p][i] - sum_{s < j} panel[s][i] panel[s][p]) / sqrt(resid_p);   resid[i] -= row[i]^2
// resid_p = resid[p] comes by value: the thread that owns p overwrites resid[p] while the others still need it.
// zero_taken: entries of taken candidates are stored as 0 and their residual is left alone (factor on Abar only).
__global__ void __launch_bounds__(256) condition_row_kernel(const double *cov, int64_t ld, int64_t n, double *panel,
                                                            int64_t ldp, int64_t j, int64_t p, double resid_p,
                                                            double *resid, const int *taken, int zero_taken) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    double acc = cov[p * ld + i];
    for (int64_t s = 0; s < j; ++s) acc -= panel[s * ldp + i] * panel[s * ldp + p];
    double f = acc / sqrt(resid_p);
    if (zero_taken && taken[i]) f = 0.0;
    panel[j * ldp + i] = f;
    if (!(zero_taken && taken[i])) resid[i] = (i == p) ? 0.0 : resid[i] - f * f;
}

__global__ void __launch_bounds__(256) diag_init_kernel(const double *cov, int64_t ld, int64_t n, const int *taken,
                                                        double *resid, double *num) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const double d = cov[i * ld + i];
    if (resid) resid[i] = (taken && taken[i]) ? 0.0 : d;
    if (num) num[i] = d;
}

// out[i] = sum_{s < r} z[s][i]^2
__global__ void __launch_bounds__(256) column_norms_kernel(const double *z, int64_t ldz, int64_t r, int64_t n, double *out) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    double acc = 0.0;
    for (int64_t s = 0; s < r; ++s) {
        const double v = z[s * ldz + i];
        acc = fma(v, v, acc);
    }
    out[i] = acc;
}

// delta_i = 0 if |den| < small or |nom| < small else nom / den   (placement_algorithm2.py:116-119); NaN where taken
__global__ void __launch_bounds__(256) pinv_score_kernel(const double *num, const double *lev, const double *ginv2,
                                                         const int *taken, int64_t n, double small, double lev_tol,
                                                         double *delta) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    if (taken[i]) {
        delta[i] = nan("");
        return;
    }
    const double den = (lev[i] >= 1.0 - lev_tol && ginv2[i] > 0.0) ? 1.0 / ginv2[i] : 0.0;
    const double nom = num[i];
    delta[i] = (fabs(den) < small || fabs(nom) < small) ? 0.0 : nom / den;
}

__global__ void __launch_bounds__(256) take_kernel(double *ft, int64_t ldf, int64_t r, int64_t y, int *taken) {
    const int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (s < r) ft[s * ldf + y] = 0.0;
    if (s == 0) taken[y] = 1;
}

struct Buffers {
    std::vector<void *> ptrs;
    ~Buffers() {
        for (void *p : ptrs) cudaFree(p);
    }
    template <class T>
    int get(T **out, size_t count) {
        void *p = nullptr;
        VGP_CUDA(device_malloc(&p, (count ? count : 1) * sizeof(T)));
        ptrs.push_back(p);
        *out = (T *)p;
        return VGP_OK;
    }
};

struct State {
    int64_t n = 0, n_pad = 0, rcap = 0, r = 0;
    double *cov = nullptr, *ft = nullptr, *z = nullptr, *g = nullptr, *resid = nullptr, *num = nullptr, *w = nullptr;
    double *lev = nullptr, *ginv2 = nullptr, *delta = nullptr;
    int *taken = nullptr;
    Best *best = nullptr;
    double scale = 0.0;         // largest diagonal entry of cov_vv
    double rank_tol = RANK_TOL; // of the pivoted Cholesky; loosened when a factor turns out numerically singular
    DenseWorkspace ws;
    int64_t launches = 0, refactors = 0;
};

int read_best(State &st, cudaStream_t s, Best *host) {
    VGP_CUDA(cudaMemcpyAsync(host, st.best, sizeof(Best), cudaMemcpyDeviceToHost, s));
    VGP_CUDA(cudaStreamSynchronize(s));
    return VGP_OK;
}

// Pivoted Cholesky of Sigma restricted to the candidates not yet taken: Ft [r][n_pad] with Sigma_AbarAbar = F F^T.
int factorise(State &st, cudaStream_t s) {
    const unsigned vb = (unsigned)((st.n + 255) / 256);
    VGP_CUDA(cudaMemsetAsync(st.ft, 0, (size_t)st.rcap * st.n_pad * 8, s));
    diag_init_kernel<<<vb, 256, 0, s>>>(st.cov, st.n_pad, st.n, st.taken, st.resid, nullptr);
    VGP_LAUNCH_CHECK();
    st.r = 0;
    for (int64_t j = 0; j < st.rcap; ++j) {
        first_argmax_kernel<<<1, 1024, 0, s>>>(st.resid, st.taken, st.n, st.rank_tol * st.scale, st.best);
        VGP_LAUNCH_CHECK();
        Best b;
        VGP_TRY(read_best(st, s, &b));
        if (b.index < 0) break;                         // every remaining residual is below the rank tolerance
        condition_row_kernel<<<vb, 256, 0, s>>>(st.cov, st.n_pad, st.n, st.ft, st.n_pad, j, b.index, b.value, st.resid,
                                                st.taken, 1);
        VGP_LAUNCH_CHECK();
        st.r = j + 1;
    }
    if (st.r == st.rcap) {
        first_argmax_kernel<<<1, 1024, 0, s>>>(st.resid, st.taken, st.n, st.rank_tol * st.scale, st.best);
        VGP_LAUNCH_CHECK();
        Best b;
        VGP_TRY(read_best(st, s, &b));
        VGP_REQUIRE(b.index < 0, "numerical rank of cov_vv exceeds max_rank = %lld", (long long)st.rcap);
    }
    return VGP_OK;
}

// leverages and (Sigma_AbarAbar^+)_yy for all candidates; *rank_lost when G is numerically singular
int denominators(State &st, cudaStream_t s, bool *rank_lost) {
    *rank_lost = false;
    const int64_t rp = round_up(std::max<int64_t>(st.r, 1), TILE);
    const unsigned vb = (unsigned)((st.n + 255) / 256);
    if (st.r == 0) {        // nothing left to span: every conditional variance is zero
        VGP_CUDA(cudaMemsetAsync(st.lev, 0, (size_t)st.n_pad * 8, s));
        VGP_CUDA(cudaMemsetAsync(st.ginv2, 0, (size_t)st.n_pad * 8, s));
        return VGP_OK;
    }
    // G = Ft Ft^T over the first rp rows (rows r .. rp are zero: identity goes on the padding diagonal)
    VGP_TRY(dense_gemm(0, 1, rp, rp, st.n_pad, 1.0, st.ft, st.n_pad, st.ft, st.n_pad, 0.0, st.g, rp, GEMM_FULL, s));
    VGP_TRY(pad_identity(st.g, rp, st.r, rp, s));
    std::vector<double> diag((size_t)st.r);
    VGP_CUDA(cudaMemcpy2DAsync(diag.data(), 8, st.g, (size_t)(rp + 1) * 8, 8, (size_t)st.r, cudaMemcpyDeviceToHost, s));
    VGP_CUDA(cudaStreamSynchronize(s));
    const double gmax = *std::max_element(diag.begin(), diag.end());
    dense_set_pivot_floor(1e-14 * gmax);             // gross failures only: rank is decided by the pivoted Cholesky
    int info = 0;
    int rc = dense_potrf(st.g, rp, rp, st.ws, s);
    if (rc == VGP_OK) rc = dense_read_info(st.ws, &info, s);
    dense_set_pivot_floor(0.0);
    if (rc == VGP_ERR_NOT_PD) {
        *rank_lost = true;
        return VGP_OK;
    }
    VGP_TRY(rc);
    VGP_CUDA(cudaMemcpyAsync(st.z, st.ft, (size_t)rp * st.n_pad * 8, cudaMemcpyDeviceToDevice, s));
    VGP_TRY(dense_trsm(0, 0, rp, st.n_pad, 1.0, st.g, rp, st.z, st.n_pad, st.ws, true, s));       // Z = L^-1 F^T
    column_norms_kernel<<<vb, 256, 0, s>>>(st.z, st.n_pad, st.r, st.n, st.lev);
    VGP_LAUNCH_CHECK();
    VGP_TRY(dense_trsm(0, 1, rp, st.n_pad, 1.0, st.g, rp, st.z, st.n_pad, st.ws, true, s));       // Z = L^-T Z = G^-1 F^T
    column_norms_kernel<<<vb, 256, 0, s>>>(st.z, st.n_pad, st.r, st.n, st.ginv2);
    VGP_LAUNCH_CHECK();
    return VGP_OK;
}

}  // namespace

extern "C" {

int vgp_placement_host_pinv(int device, const double *cov_host, int64_t n, int64_t ld_host, int64_t k, double small,
                            int algorithm, int64_t max_rank, int64_t *selection_host, double *scores_host, double *step_scores_host,
                            int64_t *rank_host, double *seconds_host) {
    VGP_REQUIRE(cov_host && selection_host, "NULL argument");
    VGP_REQUIRE(n > 0 && ld_host >= n && k > 0 && k <= n, "bad sizes n=%lld ld=%lld k=%lld", (long long)n,
                (long long)ld_host, (long long)k);
    VGP_REQUIRE(algorithm == 1 || algorithm == 2, "algorithm must be 1 (naive) or 2 (lazy cache)");
    if (max_rank <= 0 || max_rank > n) max_rank = n;
    VGP_ENTER(device);
    cudaStream_t s = nullptr;
    State st;
    Buffers mem;
    st.n = n;
    st.n_pad = round_up(n, TILE);
    st.rcap = round_up(max_rank, TILE);
    VGP_REQUIRE((double)st.rcap * st.n_pad * 16 + (double)st.n_pad * st.n_pad * 8 < 150e9,
                "pseudo-inverse path: n = %lld with max_rank = %lld does not fit the device", (long long)n,
                (long long)max_rank);
    VGP_TRY(mem.get(&st.cov, (size_t)st.n_pad * st.n_pad));
    VGP_TRY(mem.get(&st.ft, (size_t)st.rcap * st.n_pad));
    VGP_TRY(mem.get(&st.z, (size_t)st.rcap * st.n_pad));
    VGP_TRY(mem.get(&st.g, (size_t)st.rcap * st.rcap));
    VGP_TRY(mem.get(&st.w, (size_t)k * st.n_pad));
    for (double **p : {&st.resid, &st.num, &st.lev, &st.ginv2, &st.delta}) VGP_TRY(mem.get(p, (size_t)st.n_pad));
    VGP_TRY(mem.get(&st.taken, (size_t)st.n_pad));
    VGP_TRY(mem.get(&st.best, 1));
    cudaEvent_t ev[3];
    for (auto &e : ev) cudaEventCreate(&e);
    struct EventGuard {
        cudaEvent_t *ev;
        ~EventGuard() {
            for (int i = 0; i < 3; ++i) cudaEventDestroy(ev[i]);
        }
    } event_guard{ev};
    struct WsGuard {
        DenseWorkspace &ws;
        ~WsGuard() { ws.release(); }
    } ws_guard{st.ws};
    const int64_t before = g_launches;
    cudaEventRecord(ev[0], s);
    // the whole symmetric matrix: the conditioning rows read cov[p][:]
    VGP_CUDA(cudaMemsetAsync(st.cov, 0, (size_t)st.n_pad * st.n_pad * 8, s));
    VGP_CUDA(cudaMemcpy2DAsync(st.cov, (size_t)st.n_pad * 8, cov_host, (size_t)ld_host * 8, (size_t)n * 8, (size_t)n,
                               cudaMemcpyHostToDevice, s));
    VGP_CUDA(cudaMemsetAsync(st.taken, 0, (size_t)st.n_pad * 4, s));
    for (int64_t i = 0; i < n; ++i) st.scale = std::max(st.scale, cov_host[i * ld_host + i]);
    VGP_REQUIRE(st.scale > 0.0, "cov_vv has no positive diagonal entry");
    const unsigned vb = (unsigned)((n + 255) / 256);
    diag_init_kernel<<<vb, 256, 0, s>>>(st.cov, st.n_pad, n, nullptr, nullptr, st.num);
    VGP_LAUNCH_CHECK();
    VGP_TRY(factorise(st, s));
    if (rank_host) *rank_host = st.r;
    {   // what the factor leaves over must be rounding noise: a clearly negative residual means cov_vv is indefinite
        std::vector<double> resid((size_t)n);
        VGP_CUDA(cudaMemcpyAsync(resid.data(), st.resid, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
        VGP_CUDA(cudaStreamSynchronize(s));
        const double lowest = *std::min_element(resid.begin(), resid.end());
        if (lowest < -1e-8 * st.scale) {
            set_error("cov_vv is not positive semi-definite: residual variance %.3e after a rank-%lld factor", lowest,
                      (long long)st.r);
            return VGP_ERR_NOT_PD;
        }
    }
    cudaEventRecord(ev[1], s);
    int64_t wrows = 0;
    // alg. 2 (placement_algorithm2.py:151-219): a cache of deltas, all stale at the start of a selection; the arg-max
    // entry of the cache is re-evaluated until an up-to-date one wins.  The deltas come from the device (one vector per
    // selection); the cache walk itself is an index scan on the host.  It is NOT alg. 1 on these inputs: a delta can
    // rise from 0 (guarded) to a positive value between selections, which a stale cache entry never shows.
    std::vector<double> cache, delta_host;
    std::vector<char> fresh, taken_host;
    if (algorithm == 2) {
        cache.assign((size_t)n, INFINITY);
        delta_host.resize((size_t)n);
        fresh.resize((size_t)n);
        taken_host.assign((size_t)n, 0);
    }
    for (int64_t t = 0; t < k; ++t) {
        bool lost = false;
        // Fewer candidates left than factor columns: G = sum of n - t rank-1 terms is singular for certain (an
        // unpivoted Cholesky does not show it reliably: the zero turns up as a pivot of 1e-11, not 1e-16) -> rebuild.
        if (n - t < st.r) {
            ++st.refactors;
            VGP_TRY(factorise(st, s));
        }
        VGP_TRY(denominators(st, s, &lost));
        // the remaining candidates no longer span the factor space (non-generic input): rebuild on Abar; if the fresh
        // factor is itself numerically singular, its weakest directions are rounding noise -- drop them
        for (int attempt = 0; lost && attempt < 4; ++attempt) {
            ++st.refactors;
            if (attempt > 0) st.rank_tol *= 1e3;
            VGP_TRY(factorise(st, s));
            VGP_TRY(denominators(st, s, &lost));
        }
        VGP_REQUIRE(!lost, "pseudo-inverse path: factor of the remaining candidates stays singular (rank tolerance %.1e)",
                    st.rank_tol);
        pinv_score_kernel<<<vb, 256, 0, s>>>(st.num, st.lev, st.ginv2, st.taken, n, small, LEVERAGE_TOL, st.delta);
        VGP_LAUNCH_CHECK();
        if (step_scores_host)
            VGP_CUDA(cudaMemcpyAsync(step_scores_host + t * n, st.delta, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
        Best b;
        if (algorithm == 1) {
            // first strict maximum above -1 over V \ A (placement_algorithm2.py:106-123)
            first_argmax_kernel<<<1, 1024, 0, s>>>(st.delta, st.taken, n, -1.0, st.best);
            VGP_LAUNCH_CHECK();
            VGP_TRY(read_best(st, s, &b));
        } else {
            VGP_CUDA(cudaMemcpyAsync(delta_host.data(), st.delta, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
            VGP_CUDA(cudaStreamSynchronize(s));
            std::fill(fresh.begin(), fresh.end(), 0);
            for (;;) {
                b.index = -1;
                b.value = -1.0;
                for (int64_t i = 0; i < n; ++i)             // argmax_cache_linear (:53-67): strict >, start at -1
                    if (!taken_host[(size_t)i] && b.value < cache[(size_t)i]) {
                        b.value = cache[(size_t)i];
                        b.index = i;
                    }
                if (b.index < 0 || fresh[(size_t)b.index]) break;
                cache[(size_t)b.index] = delta_host[(size_t)b.index];
                fresh[(size_t)b.index] = 1;
            }
            if (b.index >= 0) taken_host[(size_t)b.index] = 1;
        }
        if (b.index < 0) {
            set_error("list.remove(x): x not in list");             // the reference's failure mode (:144)
            return VGP_ERR_STATE;
        }
        selection_host[t] = b.index;
        if (scores_host) scores_host[t] = b.value;
        // condition the numerators on y unless it already lies in the span of the selected points
        double num_y = 0.0;
        VGP_CUDA(cudaMemcpyAsync(&num_y, st.num + b.index, 8, cudaMemcpyDeviceToHost, s));
        VGP_CUDA(cudaStreamSynchronize(s));
        if (num_y > RANK_TOL * st.scale) {
            condition_row_kernel<<<vb, 256, 0, s>>>(st.cov, st.n_pad, n, st.w, st.n_pad, wrows, b.index, num_y, st.num,
                                                    st.taken, 0);
            VGP_LAUNCH_CHECK();
            ++wrows;
        }
        take_kernel<<<(unsigned)((st.rcap + 255) / 256), 256, 0, s>>>(st.ft, st.n_pad, st.r, b.index, st.taken);
        VGP_LAUNCH_CHECK();
    }
    cudaEventRecord(ev[2], s);
    VGP_CUDA(cudaEventSynchronize(ev[2]));
    if (seconds_host) {
        float ms;
        cudaEventElapsedTime(&ms, ev[0], ev[1]);
        seconds_host[0] = 0.0;
        seconds_host[1] = ms * 1e-3;                   // H2D + pivoted Cholesky
        cudaEventElapsedTime(&ms, ev[1], ev[2]);
        seconds_host[2] = ms * 1e-3;
        cudaEventElapsedTime(&ms, ev[0], ev[2]);
        seconds_host[3] = ms * 1e-3;
    }
    (void)before;
    return VGP_OK;
}

}  // extern "C"
