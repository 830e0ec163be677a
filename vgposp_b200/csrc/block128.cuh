// Register-tiled elimination panels of one 128 x 128 block, shared by the dense layer's diagonal-block kernel
// (dense.cu, potf2_kernel) and the batched exact-GP likelihood kernel (gp.cu).
//
// 256 threads as a 16 x 16 grid; thread (ti, tc) owns the cyclic 8 x 8 sub-tile {(ti + 16 r, tc + 16 s)} in
// registers.  The column loop is split into 8 panels of 16 columns with the panel index a template parameter: which
// register holds column j and which rows / columns lie behind the pivot are compile-time facts, and the trailing
// update shrinks with the panel.
#pragma once
#include "common.cuh"

namespace vgp {

constexpr int NB = 128;

template <int SJ>
__device__ __forceinline__ void potf2_panel(double (&v)[8][8], double (*colbuf)[NB], int ti, int tc, int *info,
                                            int row_offset, double pivot_floor = 0.0) {
#pragma unroll 1
    for (int jj = 0; jj < 16; ++jj) {
        const int j = 16 * SJ + jj, jb = j & 1;
        if (tc == jj) {
#pragma unroll
            for (int r = SJ; r < 8; ++r) colbuf[jb][ti + 16 * r] = v[r][SJ];
        }
        __syncthreads();
        const double d = colbuf[jb][j];
        // pivot_floor > 0: a conditional variance this far below the matrix scale means numerical rank deficiency
        if (!(d > pivot_floor) && threadIdx.x == 0) atomicCAS(info, 0, row_offset + j + 1);
        const double piv = sqrt(d);
        const double inv = 1.0 / piv;
        double lr[8], lc[8];
#pragma unroll
        for (int r = SJ; r < 8; ++r) lr[r] = (r > SJ || ti > jj) ? colbuf[jb][ti + 16 * r] * inv : 0.0;     // i > j
#pragma unroll
        for (int s = SJ; s < 8; ++s) lc[s] = (s > SJ || tc > jj) ? colbuf[jb][tc + 16 * s] * inv : 0.0;     // c > j
#pragma unroll
        for (int r = SJ; r < 8; ++r)
#pragma unroll
            for (int s = SJ; s <= r; ++s) v[r][s] = fma(-lr[r], lc[s], v[r][s]);      // lower blocks only
        if (tc == jj) {        // column j is final: store L[i][j]
#pragma unroll
            for (int r = SJ; r < 8; ++r) {
                if (r > SJ || ti > jj) v[r][SJ] = lr[r];
                else if (ti == jj) v[r][SJ] = piv;
            }
        }
    }
}

}  // namespace vgp
