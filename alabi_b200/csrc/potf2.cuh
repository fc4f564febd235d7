// Diagonal-block kernel of the blocked Cholesky: factor a 128 x 128 SPD block
// and invert the factor, in ONE register-tiled sweep by a single CTA.
//
// 256 threads form a 16 x 16 grid; thread (ty, tx) owns the 8 x 8 cyclic
// sub-matrix rows {ty + 16 r}, columns {tx + 16 c} of both A (-> L) and B
// (identity -> L^-1) in registers.  Step j broadcasts column j of A and row j of
// B through shared memory (one __syncthreads per step, double buffered) and
// every thread applies the two rank-1 updates
//      A[i][k] -= A[i][j] A[k][j] / p_j          (j < k <= i)
//      B[i][c] -= A[i][j] B[j][c] / p_j          (c <= j < i)
// Columns/rows are scaled by 1/sqrt(p_j) only at the end, so the sweep has no
// second dependent phase.  The outer loop over j / 16 is unrolled so that all
// register-array indices are static and the triangular structure prunes work at
// compile time.
#pragma once
#include "common.cuh"

namespace abp {

constexpr int NB = AB_NB;

// 1/x to ~1 ulp: MUFU.RCP64H seed + two Newton steps (no slow-path call on the
// critical path of the column sweep; x is a positive pivot here)
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

template <bool FACTOR, int JR>
__device__ __forceinline__ void sweep16(double (&a)[8][8], double (&b)[8][8], double (*colA)[NB], double (*rowB)[NB],
                                        double* sdiag, double* sinv, int tx, int ty, int64_t o, int* info) {
    for (int jm = 0; jm < 16; jm++) {
        const int j = JR * 16 + jm, buf = jm & 1;
        if (tx == jm) {
#pragma unroll
            for (int r = JR; r < 8; r++) colA[buf][ty + 16 * r] = a[r][JR];
        }
        if (ty == jm) {
#pragma unroll
            for (int c = 0; c <= JR; c++) rowB[buf][tx + 16 * c] = b[JR][c];
        }
        __syncthreads();
        double p = colA[buf][j];
        if (FACTOR && !(p > 0.0)) {                 // LAPACK dpotrf: ajj <= 0 or NaN -> info = j + 1
            if (tx == 0 && ty == 0) atomicCAS(info, 0, (int)(o + j + 1));
            p = 1.0;
        }
        if (tx == 0 && ty == 0) sdiag[j] = p;          // sqrt / reciprocal of the pivots: after the sweep
        const double ip = fast_rcp(p);
        double ai[8];
#pragma unroll
        for (int r = JR; r < 8; r++) ai[r] = colA[buf][ty + 16 * r];
        if (FACTOR) {
#pragma unroll
            for (int c = JR; c < 8; c++) {
                const bool col_on = (c > JR) || (tx > jm);             // k > j
                const double ak = colA[buf][tx + 16 * c] * ip;
#pragma unroll
                for (int r = c; r < 8; r++) {
                    const bool on = col_on && ((r > c) || (ty >= tx));  // i >= k
                    if (on) a[r][c] = fma(-ai[r], ak, a[r][c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c <= JR; c++) {
            const bool col_on = (c < JR) || (tx <= jm);                 // c' <= j
            const double xk = rowB[buf][tx + 16 * c] * ip;
#pragma unroll
            for (int r = JR; r < 8; r++) {
                const bool on = col_on && ((r > JR) || (ty > jm));      // i > j
                if (on) b[r][c] = fma(-ai[r], xk, b[r][c]);
            }
        }
    }
}

// FACTOR = true : grid 1; factor the block at offset o, write L (upper part zeroed),
//                 D^-1 and the log-determinant part.
// FACTOR = false: grid = number of diagonal blocks; the matrix already holds L
//                 (imported factor); rebuild D^-1 and the log-determinant parts.
template <bool FACTOR>
__global__ void __launch_bounds__(256, 1)
potf2_inv_kernel(double* __restrict__ A, int64_t ld, int64_t o, double* __restrict__ Dinv,
                 double* __restrict__ logdet_part, int* __restrict__ info) {
    if (!FACTOR) {
        o = (int64_t)blockIdx.x * NB;
        Dinv += (int64_t)blockIdx.x * NB * NB;
        logdet_part += blockIdx.x;
    }
    __shared__ double colA[2][NB], rowB[2][NB], sdiag[NB], sinv[NB];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    double a[8][8], b[8][8];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 8; c++) {
            // only the lower block triangle (r >= c) is ever referenced: the rest
            // must not occupy registers
            if (r >= c) {
                const int i = ty + 16 * r, k = tx + 16 * c;
                a[r][c] = (k <= i) ? A[(o + i) * ld + o + k] : 0.0;
                b[r][c] = (k == i) ? 1.0 : 0.0;
            }
        }
    sweep16<FACTOR, 0>(a, b, colA, rowB, sdiag, sinv, tx, ty, o, info);
    sweep16<FACTOR, 1>(a, b, colA, rowB, sdiag, sinv, tx, ty, o, info);
    sweep16<FACTOR, 2>(a, b, colA, rowB, sdiag, sinv, tx, ty, o, info);
    sweep16<FACTOR, 3>(a, b, colA, rowB, sdiag, sinv, tx, ty, o, info);
    sweep16<FACTOR, 4>(a, b, colA, rowB, sdiag, sinv, tx, ty, o, info);
    sweep16<FACTOR, 5>(a, b, colA, rowB, sdiag, sinv, tx, ty, o, info);
    sweep16<FACTOR, 6>(a, b, colA, rowB, sdiag, sinv, tx, ty, o, info);
    sweep16<FACTOR, 7>(a, b, colA, rowB, sdiag, sinv, tx, ty, o, info);
    __syncthreads();
    if (tid < NB) {                                   // pivots -> diagonal of L and its reciprocal
        const double dj = FACTOR ? sqrt(sdiag[tid]) : sdiag[tid];
        sdiag[tid] = dj;
        sinv[tid] = 1.0 / dj;
    }
    __syncthreads();
    if (tid < 32) {
        double s = 0.0;
        for (int j = tid; j < NB; j += 32) s += 2.0 * log(sdiag[j]);
        s = ab_warp_sum(s);
        if (tid == 0) *logdet_part = s;
    }
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int i = ty + 16 * r;
        const double si = sinv[i];
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int k = tx + 16 * c;
            if (r >= c) {
                if (FACTOR) A[(o + i) * ld + o + k] = (k < i) ? a[r][c] * sinv[k] : ((k == i) ? sdiag[i] : 0.0);
                Dinv[i * NB + k] = (k <= i) ? b[r][c] * si : 0.0;
            } else {
                if (FACTOR) A[(o + i) * ld + o + k] = 0.0;
                Dinv[i * NB + k] = 0.0;
            }
        }
    }
}

}  // namespace abp
