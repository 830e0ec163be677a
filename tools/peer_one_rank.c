/* Stand-alone harness over the C-ABI (no Python): the peer-exchange greedy with ONE rank on one GPU -- the mailbox is
 * the rank's own -- so that peer_step_kernel / downdate_peer_kernel can be timed at n = 50 000 without a second GPU,
 * beside the single-shard path (vgp_greedy_run) on the same precision panel.  Selections of the two must agree.
 *   gcc -O2 -Iinclude -o tools/bin/peer_one_rank tools/peer_one_rank.c -Lvgposp_b200/lib -lvgposp -Wl,-rpath,'$ORIGIN/../../vgposp_b200/lib' -lm */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "vgposp.h"

#define OK(call)                                                                      \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_ != 0) {                                                               \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, vgp_last_error());          \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

int main(int argc, char **argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 50000;
    const int64_t k = 20, warm = 5;
    const double ls = 0.5 * cbrt(1000.0 / (double)n);
    double *xh = malloc((size_t)n * 3 * sizeof(double));
    uint64_t s = 88172645463325252ull;
    for (int64_t i = 0; i < 3 * n; ++i) {                   /* xorshift: uniform(-2, 2) */
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        xh[i] = -2.0 + 4.0 * (double)(s >> 11) / 9007199254740992.0;
    }
    void *xd = NULL;
    OK(vgp_malloc(0, (size_t)n * 24, &xd));
    OK(vgp_memcpy_h2d(0, xd, xh, (size_t)n * 24, NULL));
    vgp_greedy *h = NULL;
    OK(vgp_greedy_create(&h, 0, n, 0, n, k + warm, 1e-8, 0.0));
    double *cov = NULL, *prec = NULL;
    int64_t ld = 0, n_pad = 0;
    OK(vgp_greedy_panels(h, &cov, &prec, &ld, &n_pad));
    OK(vgp_expquad_matrix(0, xd, n, xd, n, 3, 1.0, ls, 1e-2, 0, cov, ld, NULL));
    int info = 0;
    OK(vgp_greedy_factor(h, &info, NULL));
    OK(vgp_greedy_save_precision(h, NULL));
    int64_t sel_a[64], sel_b[64], cnt = 0;
    double sc_a[64], sc_b[64], ms = 0.0, step_ms = 0.0;
    int64_t launches = 0;

    /* single-shard path */
    OK(vgp_greedy_run(h, warm, NULL));
    OK(vgp_greedy_restore_precision(h, NULL));
    OK(vgp_greedy_profile(h, 1));
    OK(vgp_greedy_run(h, k, NULL));
    OK(vgp_greedy_profile_read(h, &ms, &launches));
    OK(vgp_greedy_profile(h, 0));
    OK(vgp_greedy_results(h, &cnt, sel_a, sc_a, 64, NULL));
    printf("single shard : downdate_kernel      %.4f ms per launch (%lld launches), %.1f GB/s\n", ms / launches,
           (long long)launches, 16.0 * n * n / (ms / launches) * 1e-6);

    /* peer path with one rank */
    const int64_t bounds[2] = {0, n};
    void *mb = NULL;
    OK(vgp_greedy_comm_create(h, 0, 1, bounds, NULL, &mb));
    void *peers[1] = {mb};
    OK(vgp_greedy_comm_connect(h, peers, 0));
    OK(vgp_greedy_restore_precision(h, NULL));
    OK(vgp_greedy_run_peer(h, warm, NULL));
    OK(vgp_greedy_restore_precision(h, NULL));
    OK(vgp_greedy_profile(h, 1));
    OK(vgp_greedy_run_peer(h, k, NULL));
    OK(vgp_greedy_profile_read(h, &ms, &launches));
    OK(vgp_greedy_profile_step_ms(h, &step_ms));
    OK(vgp_greedy_profile(h, 0));
    int err = 0;
    OK(vgp_greedy_comm_status(h, &err, NULL));
    OK(vgp_greedy_results(h, &cnt, sel_b, sc_b, 64, NULL));
    printf("peer, 1 rank : downdate_peer_kernel %.4f ms per launch (%lld launches), %.1f GB/s; peer_step_kernel %.4f ms\n",
           ms / launches, (long long)launches, 16.0 * n * n / (ms / launches) * 1e-6, step_ms / launches);
    int same = cnt == k;
    for (int64_t i = 0; i < k && same; ++i) same = sel_a[i] == sel_b[i] && sc_a[i] == sc_b[i];
    printf("selections and scores of the two paths identical: %s  (first: %lld %lld %lld %lld)\n", same ? "yes" : "NO",
           (long long)sel_b[0], (long long)sel_b[1], (long long)sel_b[2], (long long)sel_b[3]);
    vgp_greedy_destroy(h);
    return same ? 0 : 2;
}
