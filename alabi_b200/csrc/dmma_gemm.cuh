// FP64 tensor-core GEMM core for sm_100a: 128x128 CTA tile, BK = 16, cp.async
// multi-stage pipeline, DMMA.8x8x4 (mma.sync.m8n8k4.f64) warp tiles of 64x32.
//
// On sm_100a every f64 mma shape lowers to DMMA.8x8x4 (checked with cuobjdump),
// tcgen05 has no f64 kind, so this warp-level path IS the FP64 tensor pipe.
//
// Each operand tile may be "K-major" (global layout [row][k], k contiguous) or
// "K-strided" (global layout [k][row], row contiguous).  Shared-memory leading
// dimensions are chosen == 4 (mod 16) doubles so that the 16 lanes of a
// half-warp (4 groupIDs x 4 threadID_in_group) hit 16 distinct 8-byte banks.
#pragma once
#include "common.cuh"

namespace abg {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4, THREADS = 256;
constexpr int LDK = BK + 4;            // K-major tile  [128][20]
constexpr int LDR = 128 + 4;           // K-strided tile [16][132]
constexpr int OPER_ELEMS = 128 * LDK;  // 2560 doubles >= 16 * 132
constexpr int STAGE_ELEMS = 2 * OPER_ELEMS;
constexpr int SMEM_BYTES = STAGES * STAGE_ELEMS * 8;   // 163840

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// Copy one 128 x 16 operand chunk global -> shared (all 256 threads, 4 x 16 B each).
template <bool KMAJOR>
__device__ __forceinline__ void load_operand(double* s, const double* g, int64_t ld, int tid) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int c = tid + i * THREADS;
        if (KMAJOR) {
            int row = c >> 3, c16 = c & 7;
            cp_async16(s + row * LDK + c16 * 2, g + (int64_t)row * ld + c16 * 2);
        } else {
            int k = c >> 6, c16 = c & 63;
            cp_async16(s + k * LDR + c16 * 2, g + (int64_t)k * ld + c16 * 2);
        }
    }
}

struct Acc {
    double v[8][4][2];   // [m-frag][n-frag][2]; 64x32 warp tile
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) v[i][j][0] = v[i][j][1] = 0.0;
    }
};

// acc += A_tile(128 x nk*16) * B_tile(128 x nk*16)^T.
//   AK: A is K-major, pointer at element [row0][k0]; else pointer at [k0][row0].
//   BKM: same for B.
// All 256 threads must call; smem must hold SMEM_BYTES.  Ends with the pipeline
// drained and a __syncthreads(), so smem can be reused by the caller.
//   TRI_A: the last 8 chunks (one 128-wide k block) multiply a LOWER-TRIANGULAR
//   128 x 128 A block (A[r][k] = 0 for k > r): m-fragments whose rows all lie
//   above the current k4-step are skipped (warp-uniform test).
template <bool AK, bool BKM, bool TRI_A = false>
__device__ __forceinline__ void mainloop(Acc& acc, const double* __restrict__ A, int64_t lda,
                                         const double* __restrict__ B, int64_t ldb, int nk,
                                         double* smem) {
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;          // 2 x 4 warps
    const int64_t a_step = AK ? (int64_t)BK : (int64_t)BK * lda;
    const int64_t b_step = BKM ? (int64_t)BK : (int64_t)BK * ldb;

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nk) {
            load_operand<AK>(smem + s * STAGE_ELEMS, A + s * a_step, lda, tid);
            load_operand<BKM>(smem + s * STAGE_ELEMS + OPER_ELEMS, B + s * b_step, ldb, tid);
        }
        cp_async_commit();
    }
    for (int kc = 0; kc < nk; kc++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nx = kc + STAGES - 1;
            if (nx < nk) {
                int st = nx % STAGES;
                load_operand<AK>(smem + st * STAGE_ELEMS, A + nx * a_step, lda, tid);
                load_operand<BKM>(smem + st * STAGE_ELEMS + OPER_ELEMS, B + nx * b_step, ldb, tid);
            }
            cp_async_commit();
        }
        const double* sA = smem + (kc % STAGES) * STAGE_ELEMS;
        const double* sB = sA + OPER_ELEMS;
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            double a[8], b[4];
#pragma unroll
            for (int f = 0; f < 8; f++)
                a[f] = AK ? sA[(wm * 64 + f * 8 + g) * LDK + kk * 4 + t]
                          : sA[(kk * 4 + t) * LDR + wm * 64 + f * 8 + g];
#pragma unroll
            for (int f = 0; f < 4; f++)
                b[f] = BKM ? sB[(wn * 32 + f * 8 + g) * LDK + kk * 4 + t]
                           : sB[(kk * 4 + t) * LDR + wn * 32 + f * 8 + g];
            const int ktri = TRI_A ? (kc - (nk - 8)) * BK + kk * 4 : -1;   // k offset inside the diagonal block
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (TRI_A && ktri > wm * 64 + i * 8 + 7) continue;
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc.v[i][j][0], acc.v[i][j][1], a[i], b[j]);
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
}

// Element (row, col) owned by acc.v[i][j][e] inside the 128 x 128 CTA tile.
__device__ __forceinline__ int acc_row(int i) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    return (warp >> 2) * 64 + i * 8 + (lane >> 2);
}
__device__ __forceinline__ int acc_col(int j) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    return (warp & 3) * 32 + j * 8 + (lane & 3) * 2;   // + e
}

// C[tile] = alpha * acc + beta * C[tile]   (C row-major, 16-byte vector accesses)
__device__ __forceinline__ void store_tile(const Acc& acc, double* __restrict__ C, int64_t ldc,
                                           double alpha, double beta) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int r = acc_row(i);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double2* p = reinterpret_cast<double2*>(C + (int64_t)r * ldc + acc_col(j));
            double2 o;
            if (beta != 0.0) {
                double2 c = *p;
                o.x = alpha * acc.v[i][j][0] + beta * c.x;
                o.y = alpha * acc.v[i][j][1] + beta * c.y;
            } else {
                o.x = alpha * acc.v[i][j][0];
                o.y = alpha * acc.v[i][j][1];
            }
            *p = o;
        }
    }
}

// linear index p -> (i, j) with 0 <= j <= i  (row-major lower triangle)
__device__ __forceinline__ void tri_decode(int p, int& i, int& j) {
    int r = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
    while ((r + 1) * (r + 2) / 2 <= p) r++;
    while (r * (r + 1) / 2 > p) r--;
    i = r;
    j = p - r * (r + 1) / 2;
}

}  // namespace abg
