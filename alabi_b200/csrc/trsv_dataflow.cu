// K2 solves, dataflow form: z = L^-1 r and alpha = L^-T z in ONE cooperative launch
// with per-block ready flags instead of a grid barrier per block step.
//
// Replaces george's `cho_solve` inside `GP.log_likelihood` / `_compute_alpha`
// (alabi/core.py:1248, 1261, 1441).  Block row k (128 rows) belongs to CTA k mod G:
//
//   forward   z_k     = D_k^-1 ( r_k - sum_{j<k} L_kj z_j )           j ascending
//   backward  alpha_k = D_k^-T ( z_k - sum_{j>k} L_jk^T alpha_j )     j descending
//
// A CTA pre-issues the 128 x 128 block L_kj into registers (coalesced rows, 64
// doubles per thread) BEFORE it waits for z_j, so the step on the critical path is:
// flag + 1 KB of z_j from L2, 64 FMAs per thread, one reduction, the D_k^-1 product
// from shared memory (D_k^-1 is staged there by cp.async when the task starts), a
// fence and the flag of z_k.  The summation order is fixed (block order j, lanes,
// then warps), so results do not depend on timing.
//
// Deadlock freedom: cooperative launch (all CTAs resident); every CTA walks its
// rows in dependency order, so the first unfinished row can always run.  Wait
// loops carry a watchdog that raises an abort flag instead of hanging.
#include "handle.h"
#include "dmma_gemm.cuh"

namespace {

constexpr int NB = AB_NB;

struct TrsvArgs {
    const double* L; int64_t ld; const double* Dinv; int T;
    const double* r; double* z; double* alpha;
    int* zflag; int* aflag; int* abort_flag;
    int backward;                 // 0: forward sweep only (z), 1: both
};

__device__ __forceinline__ int ld_acquire_i(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_i(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void wait_flag(const int* flag, int* abort_flag, bool& aborted) {
    if (aborted) return;
    long long t0 = 0;
    unsigned spins = 0;
    while (ld_acquire_i(flag) == 0) {
        if ((++spins & 1023u) == 0) {
            if (t0 == 0) t0 = clock64();
            if (*((volatile int*)abort_flag) != 0 || clock64() - t0 > 4000000000LL) {
                atomicExch(abort_flag, 1);
                aborted = true;
                return;
            }
        }
    }
}

// the 16 rows {warp + 8 i} of a 128 x 128 row-major block, columns {lane + 32 c}
__device__ __forceinline__ void load_block(double (&v)[16][4], const double* __restrict__ M, int64_t ldm, int warp, int lane) {
#pragma unroll
    for (int i = 0; i < 16; i++)
#pragma unroll
        for (int c = 0; c < 4; c++) v[i][c] = __ldcg(M + (int64_t)(warp + 8 * i) * ldm + lane + 32 * c);
}

__device__ __forceinline__ void stage_dinv(double* sD, const double* __restrict__ D, int tid) {
    for (int c = tid; c < NB * NB / 2; c += 256) abg::cp_async16(sD + 2 * c, D + 2 * c);
    abg::cp_async_commit();
}

__global__ void __launch_bounds__(256, 1)
trsv_dataflow_kernel(const __grid_constant__ TrsvArgs a) {
    extern __shared__ __align__(16) double sD[];          // D_k^-1, 128 x 128
    __shared__ double sw[NB], spart[8][NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, T = a.T;
    bool aborted = false;
    double v[16][4];

    // ---------------- forward: z = L^-1 r ----------------
    for (int k = blockIdx.x; k < T; k += G) {
        const int64_t o = (int64_t)k * NB;
        stage_dinv(sD, a.Dinv + (int64_t)k * NB * NB, tid);
        double part[16];
#pragma unroll
        for (int i = 0; i < 16; i++) part[i] = 0.0;
        if (k > 0) load_block(v, a.L + o * a.ld, a.ld, warp, lane);
        for (int j = 0; j < k; j++) {
            wait_flag(a.zflag + j, a.abort_flag, aborted);
            double x[4];
#pragma unroll
            for (int c = 0; c < 4; c++) x[c] = __ldcg(a.z + (int64_t)j * NB + lane + 32 * c);
#pragma unroll
            for (int i = 0; i < 16; i++) {
                double s = v[i][0] * x[0];
                s = fma(v[i][1], x[1], s);
                s = fma(v[i][2], x[2], s);
                s = fma(v[i][3], x[3], s);
                part[i] += s;
            }
            if (j + 1 < k) load_block(v, a.L + o * a.ld + (int64_t)(j + 1) * NB, a.ld, warp, lane);
        }
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const double s = ab_warp_sum(part[i]);
            if (lane == 0) sw[warp + 8 * i] = a.r[o + warp + 8 * i] - s;
        }
        abg::cp_async_wait<0>();
        __syncthreads();
        double x[4];
#pragma unroll
        for (int c = 0; c < 4; c++) x[c] = sw[lane + 32 * c];
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const double* Dr = sD + (warp + 8 * i) * NB + lane;
            double s = Dr[0] * x[0];
            s = fma(Dr[32], x[1], s);
            s = fma(Dr[64], x[2], s);
            s = fma(Dr[96], x[3], s);
            s = ab_warp_sum(s);
            if (lane == 0) a.z[o + warp + 8 * i] = s;
        }
        __threadfence();
        __syncthreads();                                  // also: sD / sw are free for the next task
        if (tid == 0) st_release_i(a.zflag + k, 1);
    }

    if (!a.backward) return;
    // ---------------- backward: alpha = L^-T z ----------------
    const int last = (T - 1) - (int)blockIdx.x;           // CTA c takes rows T-1-c, T-1-c-G, ...
    for (int k = last; k >= 0; k -= G) {
        const int64_t o = (int64_t)k * NB;
        stage_dinv(sD, a.Dinv + (int64_t)k * NB * NB, tid);
        double y[4] = {0.0, 0.0, 0.0, 0.0};
        if (k + 1 < T) load_block(v, a.L + (int64_t)(T - 1) * NB * a.ld + o, a.ld, warp, lane);
        for (int j = T - 1; j > k; j--) {
            wait_flag(a.aflag + j, a.abort_flag, aborted);
            double xr[16];
#pragma unroll
            for (int i = 0; i < 16; i++) xr[i] = __ldcg(a.alpha + (int64_t)j * NB + warp + 8 * i);
#pragma unroll
            for (int i = 0; i < 16; i++)
#pragma unroll
                for (int c = 0; c < 4; c++) y[c] = fma(v[i][c], xr[i], y[c]);
            if (j - 1 > k) load_block(v, a.L + (int64_t)(j - 1) * NB * a.ld + o, a.ld, warp, lane);
        }
#pragma unroll
        for (int c = 0; c < 4; c++) spart[warp][lane + 32 * c] = y[c];
        wait_flag(a.zflag + k, a.abort_flag, aborted);    // z_k (own forward row when G divides evenly)
        __syncthreads();
        if (tid < NB) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < 8; w++) s += spart[w][tid];
            sw[tid] = __ldcg(a.z + o + tid) - s;
        }
        abg::cp_async_wait<0>();
        __syncthreads();
        // alpha_k[c] = sum_i Dinv[i][c] w[i]
        double y2[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const double wi = sw[warp + 8 * i];
            const double* Dr = sD + (warp + 8 * i) * NB + lane;
#pragma unroll
            for (int c = 0; c < 4; c++) y2[c] = fma(Dr[32 * c], wi, y2[c]);
        }
        __syncthreads();                                  // spart is reused
#pragma unroll
        for (int c = 0; c < 4; c++) spart[warp][lane + 32 * c] = y2[c];
        __syncthreads();
        if (tid < NB) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < 8; w++) s += spart[w][tid];
            a.alpha[o + tid] = s;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release_i(a.aflag + k, 1);
    }
}

}  // namespace

// z = L^-1 r, alpha = L^-T z;  r (npad) is read only.  Returns 0 after enqueueing;
// the abort flag is copied to h->h_pinned + 10 (checked by the caller after its sync).
int ab_launch_trsv_dataflow(ab_gp* h, const double* r, int backward) {
    static unsigned long long configured = 0;   // per-device bit: cudaFuncSetAttribute is per device
    const int smem = NB * NB * (int)sizeof(double);
    if (h->device >= 64 || !((configured >> h->device) & 1ULL)) {
        AB_CUDA(cudaFuncSetAttribute(trsv_dataflow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (h->device < 64) configured |= 1ULL << h->device;
    }
    const int T = (int)(h->npad / NB);
    // control ints live behind the sums slot of scratch: [abort][zflag T][aflag T]
    int rc = ab_ensure_scratch(h, 256 + (size_t)(2 * T + 1) * sizeof(int));
    if (rc) return rc;
    int* ctrl = reinterpret_cast<int*>(h->scratch + 32);
    cudaStream_t s = h->stream;
    AB_CUDA(cudaMemsetAsync(ctrl, 0, (size_t)(2 * T + 1) * sizeof(int), s));
    TrsvArgs a;
    a.L = h->L; a.ld = h->npad; a.Dinv = h->Dinv; a.T = T;
    a.r = r; a.z = h->z; a.alpha = h->alpha;
    a.abort_flag = ctrl; a.zflag = ctrl + 1; a.aflag = ctrl + 1 + T;
    a.backward = backward;
    int grid = T < h->nsm ? T : h->nsm;
    void* args[] = {(void*)&a};
    AB_CUDA(cudaLaunchCooperativeKernel((void*)trsv_dataflow_kernel, dim3(grid), dim3(256), args, smem, s));
    ab_count_launches(1);
    AB_CUDA(cudaMemcpyAsync(h->h_pinned + 10, ctrl, sizeof(int), cudaMemcpyDeviceToHost, s));
    return 0;
}
