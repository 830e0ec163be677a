// Microbenchmark (tools/, not part of the library): the precision downdate P[i][j] -= (p_i p_j) / p_y of csrc/greedy.cu
// as a stand-alone kernel, swept over CTA geometry, together with what the same buffer reaches under a plain in-place
// scale and a device-to-device copy -- to find out why the 20 GB panel of one GPU streams at 0.966 of the copy figure
// while the 2.5 - 10 GB panels of the sharded runs reach 0.99 with the same code.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/downdate_sweep tools/downdate_sweep.cu
//   tools/bin/downdate_sweep            (prints one line per variant: GB/s of 16 * rows * ld algorithmic bytes)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

// the library's kernel (greedy.cu, downdate_kernel): thread = 2 adjacent columns, CTA = 512 columns, row blocks of
// `rows_per_block` rows dealt to blockIdx.y with stride gridDim.y
template <int UNROLL>
__global__ void __launch_bounds__(256) dd_pairs(double *__restrict__ prec, int64_t ld, int64_t n_rows,
                                                const double *__restrict__ pfull, const double *__restrict__ ploc,
                                                int64_t y, int rows_per_block) {
    const int64_t col = (int64_t)blockIdx.x * 512 + 2 * threadIdx.x;
    if (col >= ld) return;
    const double inv = 1.0 / pfull[y];
    const double pj0 = ploc[col], pj1 = ploc[col + 1];
    const bool z0 = col == y, z1 = col + 1 == y;
    for (int64_t rb = (int64_t)blockIdx.y * rows_per_block; rb < n_rows; rb += (int64_t)gridDim.y * rows_per_block) {
        const int64_t rend = min(rb + rows_per_block, n_rows);
        for (int64_t r = rb; r < rend; r += UNROLL) {
            double2 v[UNROLL];
            double pi[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (r + u < rend) {
                    v[u] = *reinterpret_cast<const double2 *>(prec + (r + u) * ld + col);
                    pi[u] = pfull[r + u];
                }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
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

// thread = 4 adjacent columns (two 16-byte accesses 16 bytes apart -> one 32-byte sector per thread), CTA = 1024 columns
template <int UNROLL>
__global__ void __launch_bounds__(256) dd_quads(double *__restrict__ prec, int64_t ld, int64_t n_rows,
                                                const double *__restrict__ pfull, const double *__restrict__ ploc,
                                                int64_t y, int rows_per_block) {
    const int64_t col = (int64_t)blockIdx.x * 1024 + 4 * threadIdx.x;
    if (col >= ld) return;
    const double inv = 1.0 / pfull[y];
    double pj[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) pj[c] = col + c < ld ? ploc[col + c] : 0.0;
    for (int64_t rb = (int64_t)blockIdx.y * rows_per_block; rb < n_rows; rb += (int64_t)gridDim.y * rows_per_block) {
        const int64_t rend = min(rb + rows_per_block, n_rows);
        for (int64_t r = rb; r < rend; r += UNROLL) {
            double2 v[UNROLL][2];
            double pi[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (r + u < rend) {
                    const double2 *src = reinterpret_cast<const double2 *>(prec + (r + u) * ld + col);
                    v[u][0] = src[0];
                    v[u][1] = src[1];
                    pi[u] = pfull[r + u];
                }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (r + u < rend) {
                    double o[4] = {fma(-(pi[u] * pj[0]), inv, v[u][0].x), fma(-(pi[u] * pj[1]), inv, v[u][0].y),
                                   fma(-(pi[u] * pj[2]), inv, v[u][1].x), fma(-(pi[u] * pj[3]), inv, v[u][1].y)};
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (r + u == y || col + c == y) o[c] = 0.0;
                    double2 *dst = reinterpret_cast<double2 *>(prec + (r + u) * ld + col);
                    dst[0] = make_double2(o[0], o[1]);
                    dst[1] = make_double2(o[2], o[3]);
                }
        }
    }
}

// flat in-place scale over the same bytes (what a read-modify-write stream reaches with no indexing at all)
__global__ void __launch_bounds__(256) scale_flat(double2 *x, int64_t n2, double f) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n2; i += (int64_t)gridDim.x * 256) {
        double2 v = x[i];
        v.x *= f;
        v.y *= f;
        x[i] = v;
    }
}

__global__ void fill(double *x, int64_t n, double base, double step) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        x[i] = base + step * (double)(i % 1009);
}

__global__ void count_diff(const double *a, const double *b, int64_t n, unsigned long long *out) {
    unsigned long long c = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        c += __double_as_longlong(a[i]) != __double_as_longlong(b[i]);
    if (c) atomicAdd(out, c);
}

static cudaEvent_t e0, e1;
template <class F>
static void timeit(const char *name, double bytes, F launch) {
    for (int i = 0; i < 2; ++i) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f, sum = 0.f;
    const int reps = 5;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
        sum += ms;
    }
    CK(cudaGetLastError());
    printf("%-64s %9.4f ms best %9.4f ms mean  %8.1f GB/s best %8.1f GB/s mean\n", name, best, sum / reps,
           bytes / best * 1e-6, bytes / (sum / reps) * 1e-6);
    fflush(stdout);
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int64_t LD = 50048, ROWS = 50048, Y = 12345;
    double *panel, *other, *pfull, *ploc;
    CK(cudaMalloc(&panel, (size_t)ROWS * LD * 8));
    CK(cudaMalloc(&other, (size_t)ROWS * LD * 8));
    CK(cudaMalloc(&pfull, LD * 8));
    CK(cudaMalloc(&ploc, LD * 8));
    fill<<<sms * 8, 256>>>(panel, ROWS * LD, 1.0, 1e-6);
    fill<<<sms * 8, 256>>>(other, ROWS * LD, 1.0, 1e-6);
    fill<<<64, 256>>>(pfull, LD, 0.5, 1e-4);
    fill<<<64, 256>>>(ploc, LD, 0.5, 1e-4);
    CK(cudaDeviceSynchronize());
    printf("SMs %d; panel %lld x %lld doubles\n", sms, (long long)ROWS, (long long)LD);

    // ---- references at three footprints ------------------------------------------------------------------
    for (int64_t rows : {(int64_t)6256, (int64_t)25024, ROWS}) {
        const double bytes = 16.0 * rows * LD;
        char name[128];
        snprintf(name, sizeof name, "cudaMemcpy D2D  %5.1f GB -> other buffer", bytes / 2e9);
        timeit(name, bytes, [&] { CK(cudaMemcpyAsync(other, panel, (size_t)rows * LD * 8, cudaMemcpyDeviceToDevice)); });
        snprintf(name, sizeof name, "scale_flat in place, %5.1f GB read + written", bytes / 2e9);
        timeit(name, bytes, [&] { scale_flat<<<sms * 16, 256>>>((double2 *)panel, rows * LD / 2, 1.0000000001); });
    }

    // ---- the library's geometry at the panel sizes of the 1-, 2- and 8-GPU runs ---------------------------
    struct Geo { int64_t rows, ld; const char *what; };
    const Geo geos[] = {{ROWS, LD, "n=50k, 1 GPU (20 GB)"}, {ROWS, 25088, "n=50k, 2 GPUs (10 GB)"}, {ROWS, 6272, "n=50k, 8 GPUs (2.5 GB)"}};
    for (const Geo &g : geos) {
        const unsigned gx = (unsigned)((g.ld + 511) / 512);
        const int rpb = 32, waves = 4;
        int64_t gy = ((int64_t)sms * 8 * waves + gx - 1) / gx;
        const int64_t tiles = (g.rows + rpb - 1) / rpb;
        if (gy > tiles) gy = tiles;
        char name[128];
        snprintf(name, sizeof name, "dd_pairs<4> rpb 32 waves 4 (library), %s", g.what);
        timeit(name, 16.0 * g.rows * g.ld, [&] { dd_pairs<4><<<dim3(gx, (unsigned)gy), 256>>>(panel, g.ld, g.rows, pfull, ploc, Y, rpb); });
    }

    // ---- sweep on the 20 GB panel -------------------------------------------------------------------------
    const double bytes = 16.0 * ROWS * LD;
    for (int rpb : {4, 8, 16, 32, 64, 128})
        for (int waves : {1, 2, 4, 8, 0}) {
            const unsigned gx = (unsigned)((LD + 511) / 512);
            const int64_t tiles = (ROWS + rpb - 1) / rpb;
            int64_t gy = waves ? ((int64_t)sms * 8 * waves + gx - 1) / gx : tiles;
            if (gy > tiles) gy = tiles;
            if (gy > 65535) gy = 65535;
            char name[128];
            snprintf(name, sizeof name, "dd_pairs<4> rpb %3d waves %d grid %u x %lld", rpb, waves, gx, (long long)gy);
            timeit(name, bytes, [&] { dd_pairs<4><<<dim3(gx, (unsigned)gy), 256>>>(panel, LD, ROWS, pfull, ploc, Y, rpb); });
        }
    for (int rpb : {8, 32, 128}) {
        const unsigned gx = (unsigned)((LD + 511) / 512);
        const int64_t tiles = (ROWS + rpb - 1) / rpb;
        int64_t gy = ((int64_t)sms * 8 * 4 + gx - 1) / gx;
        if (gy > tiles) gy = tiles;
        char name[128];
        snprintf(name, sizeof name, "dd_pairs<8> rpb %3d waves 4", rpb);
        timeit(name, bytes, [&] { dd_pairs<8><<<dim3(gx, (unsigned)gy), 256>>>(panel, LD, ROWS, pfull, ploc, Y, rpb); });
        snprintf(name, sizeof name, "dd_pairs<2> rpb %3d waves 4", rpb);
        timeit(name, bytes, [&] { dd_pairs<2><<<dim3(gx, (unsigned)gy), 256>>>(panel, LD, ROWS, pfull, ploc, Y, rpb); });
    }
    for (int rpb : {8, 16, 32, 64})
        for (int waves : {2, 4, 8}) {
            const unsigned gx = (unsigned)((LD + 1023) / 1024);
            const int64_t tiles = (ROWS + rpb - 1) / rpb;
            int64_t gy = ((int64_t)sms * 8 * waves + gx - 1) / gx;
            if (gy > tiles) gy = tiles;
            char name[128];
            snprintf(name, sizeof name, "dd_quads<2> rpb %3d waves %d grid %u x %lld", rpb, waves, gx, (long long)gy);
            timeit(name, bytes, [&] { dd_quads<2><<<dim3(gx, (unsigned)gy), 256>>>(panel, LD, ROWS, pfull, ploc, Y, rpb); });
            snprintf(name, sizeof name, "dd_quads<4> rpb %3d waves %d", rpb, waves);
            timeit(name, bytes, [&] { dd_quads<4><<<dim3(gx, (unsigned)gy), 256>>>(panel, LD, ROWS, pfull, ploc, Y, rpb); });
        }

    // ---- the variants compute the same bits (4096-row slice, fresh data in both buffers) --------------------
    {
        const int64_t rows = 4096;
        unsigned long long *diff;
        CK(cudaMalloc(&diff, 8));
        fill<<<sms * 8, 256>>>(panel, rows * LD, 1.0, 1e-6);
        fill<<<sms * 8, 256>>>(other, rows * LD, 1.0, 1e-6);
        dd_pairs<4><<<dim3((unsigned)((LD + 511) / 512), 49), 256>>>(panel, LD, rows, pfull, ploc, Y % rows, 32);
        dd_quads<2><<<dim3((unsigned)((LD + 1023) / 1024), 64), 256>>>(other, LD, rows, pfull, ploc, Y % rows, 16);
        CK(cudaMemset(diff, 0, 8));
        count_diff<<<sms * 8, 256>>>(panel, other, rows * LD, diff);
        unsigned long long h = 0;
        CK(cudaMemcpy(&h, diff, 8, cudaMemcpyDeviceToHost));
        printf("dd_quads<2> vs dd_pairs<4> on %lld x %lld: %llu differing entries\n", (long long)rows, (long long)LD, h);
        fill<<<sms * 8, 256>>>(other, rows * LD, 1.0, 1e-6);
        dd_pairs<8><<<dim3((unsigned)((LD + 511) / 512), 200), 256>>>(other, LD, rows, pfull, ploc, Y % rows, 8);
        CK(cudaMemset(diff, 0, 8));
        count_diff<<<sms * 8, 256>>>(panel, other, rows * LD, diff);
        CK(cudaMemcpy(&h, diff, 8, cudaMemcpyDeviceToHost));
        printf("dd_pairs<8> rpb 8 vs dd_pairs<4> rpb 32: %llu differing entries\n", h);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
