/* TEST INFRASTRUCTURE ONLY -- C/OpenMP restatement of one selection step of the incremental greedy
 * (oracle/greedy_oracle.py IncrementalState), used as the CPU baseline in bench.py and to speed up the
 * oracle at n of a few thousand in tests.  Same arithmetic, element for element, as the NumPy form:
 *
 *   scores:   delta_j = guard(num_j - jitter, 1/P_jj - jitter)        placement_algorithm2.py:113-119
 *   segments: w_J = (Sigma[y,J] - sum_s W[s][J] W[s][y]) / sqrt(num_y), p_J = P[y,J]
 *   apply:    num_J -= w_J^2 ; P[:,J] -= (p p_J^T) / p_y ; row/column y zeroed
 *
 * Reference semantics: placement_algorithm2.py:105-145,371-413 (see the Python oracle for the derivation).
 * Built by oracle/build_oracle.py with gcc -O3 -fopenmp; no product code links or loads it.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>

/* explicit thread count (a launcher's OMP_NUM_THREADS=1 must not silently make the CPU arm single-threaded);
 * n <= 0 only queries.  Returns the count the next parallel region will use. */
int oracle_set_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
}

/* scores[j] for j < nloc (NaN where taken); returns the local first strict maximum above -1 or -1 */
int64_t oracle_scores(const double *prec, int64_t ld, int64_t c0, int64_t nloc, const double *num,
                      const unsigned char *taken, double small, double jitter, double *scores,
                      double *best_score) {
    int64_t best = -1;
    double bs = -1.0;
    for (int64_t j = 0; j < nloc; ++j) {
        if (taken[j]) {
            scores[j] = NAN;
            continue;
        }
        const double den = 1.0 / prec[(c0 + j) * ld + j] - jitter;
        const double nom = num[j] - jitter;
        double d = nom / den;
        if (fabs(den) < small || fabs(nom) < small) d = 0.0;
        scores[j] = d;
        if (bs < d) {
            bs = d;
            best = c0 + j;
        }
    }
    *best_score = bs;
    return best;
}

void oracle_segments(const double *cov, const double *prec, int64_t ld, int64_t c0, int64_t nloc,
                     const double *wfull, int64_t n, int64_t t, double jitter, int64_t y, double num_y,
                     double *w_seg, double *p_seg) {
    const double inv_sqrt = sqrt(num_y);
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < nloc; ++j) {
        double acc = cov[y * ld + j];
        if (c0 + j == y) acc += jitter;
        for (int64_t s = 0; s < t; ++s) acc -= wfull[s * n + c0 + j] * wfull[s * n + y];
        w_seg[j] = acc / inv_sqrt;
        p_seg[j] = prec[y * ld + j];
    }
}

void oracle_apply(double *prec, int64_t ld, int64_t n, int64_t c0, int64_t nloc, const double *w_seg,
                  const double *p_full, int64_t y, double *num, unsigned char *taken) {
    const double inv = 1.0 / p_full[y];
    const double *p_loc = p_full + c0;
    for (int64_t j = 0; j < nloc; ++j) num[j] -= w_seg[j] * w_seg[j];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double *row = prec + i * ld;
        const double pi = p_full[i];
        if (i == y) {
            for (int64_t j = 0; j < nloc; ++j) row[j] = 0.0;
        } else {
            for (int64_t j = 0; j < nloc; ++j) row[j] -= (pi * p_loc[j]) * inv;
        }
    }
    if (y >= c0 && y < c0 + nloc) {
        for (int64_t i = 0; i < n; ++i) prec[i * ld + (y - c0)] = 0.0;
        taken[y - c0] = 1;
    }
}
