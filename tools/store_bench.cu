// Microbenchmark: what does a pure float64 store stream reach on this GPU, for the store patterns of the
// kernel-matrix builder, and how does it degrade with FP64 work per element?  (tools/, not part of the library)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

template <int WORK>
__device__ __forceinline__ double work(double x, double c) {
    double p = c;
#pragma unroll
    for (int i = 0; i < WORK; ++i) p = fma(p, x, c);
    return p;
}

// pattern A: CTA tile 32 rows x 512 cols, thread = 2 adjacent columns, loop over rows (kernel_rect_kernel)
template <int WORK>
__global__ void __launch_bounds__(256) rect_pattern(double *out, long ld, long n1, long n2, double c) {
    const long row0 = (long)blockIdx.y * 32;
    const long col = (long)blockIdx.x * 512 + 2 * threadIdx.x;
    if (col >= n2) return;
    const double x0 = 1e-9 * col, x1 = 1e-9 * (col + 1);
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
        if (row0 + r >= n1) break;
        const double y = 1e-9 * (row0 + r);
        const double v0 = work<WORK>(x0 - y, c), v1 = work<WORK>(x1 - y, c);
        *reinterpret_cast<double2 *>(out + (row0 + r) * ld + col) = make_double2(v0, v1);
    }
}

// pattern B: CTA owns a contiguous 1 MB chunk of whole rows: 256 threads sweep each row left to right
template <int WORK>
__global__ void __launch_bounds__(256) row_pattern(double *out, long ld, long n1, long n2, double c) {
    for (long r = blockIdx.x; r < n1; r += gridDim.x) {
        const double y = 1e-9 * r;
        for (long col = 2 * threadIdx.x; col < n2; col += 512) {
            const double v0 = work<WORK>(1e-9 * col - y, c), v1 = work<WORK>(1e-9 * (col + 1) - y, c);
            *reinterpret_cast<double2 *>(out + r * ld + col) = make_double2(v0, v1);
        }
    }
}

template <class F>
float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    float best = 1e9;
    for (int i = 0; i < 5; ++i) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

template <int WORK>
void run(double *out, long n1, long n2) {
    const double gb = 8.0 * n1 * n2 / 1e9;
    dim3 grid((unsigned)((n2 + 511) / 512), (unsigned)((n1 + 31) / 32));
    float t = timeit([&] { rect_pattern<WORK><<<grid, 256>>>(out, n2, n1, n2, 0.5); });
    printf("rect_pattern work=%2d  %7.3f ms  %7.1f GB/s\n", WORK, t, gb / (t * 1e-3));
    t = timeit([&] { row_pattern<WORK><<<148 * 8, 256>>>(out, n2, n1, n2, 0.5); });
    printf("row_pattern  work=%2d  %7.3f ms  %7.1f GB/s\n", WORK, t, gb / (t * 1e-3));
}

int main(int argc, char **argv) {
    const long n1 = argc > 1 ? atol(argv[1]) : 50000, n2 = argc > 2 ? atol(argv[2]) : 25000;
    double *out;
    if (cudaMalloc(&out, (size_t)n1 * n2 * 8) != cudaSuccess) return 1;
    const double gb = 8.0 * n1 * n2 / 1e9;
    float t = timeit([&] { cudaMemsetAsync(out, 0, (size_t)n1 * n2 * 8); });
    printf("cudaMemset            %7.3f ms  %7.1f GB/s\n", t, gb / (t * 1e-3));
    run<0>(out, n1, n2);
    run<4>(out, n1, n2);
    run<8>(out, n1, n2);
    run<12>(out, n1, n2);
    run<16>(out, n1, n2);
    run<24>(out, n1, n2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
