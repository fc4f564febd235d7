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
// z_j from L2, 64 FMAs per thread, one reduction, the D_k^-1 product from shared
// memory (D_k^-1 is staged there by cp.async when the task starts) and the stores of
// z_k.  The summation order is fixed (block order j, lanes, then warps), so results
// do not depend on timing.
//
// Signalling (round 2): the solution vectors travel between CTAs as 16-byte lines
// {lo, tag, hi, tag} per value (8-byte halves are single-copy atomic; a line whose two
// tags are set holds the value), written with relaxed gpu-scope vector stores and polled
// with relaxed vector loads: the data is its own ready flag.  That takes the flag poll
// (one L2 round trip before the dependent load of z_j), the __threadfence of 256 threads
// and the release store out of every one of the 2 N / 128 dependent block steps; z and
// alpha are also written in plain form for the kernels that follow.
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
    uint4* zline; uint4* aline; int* abort_flag;      // one line per value, zeroed before the launch
    int backward;                 // 0: forward sweep only (z), 1: both
};

__device__ __forceinline__ void st_line(uint4* p, double v) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)__double2loint(v)), "r"(1u),
                 "r"((unsigned)__double2hiint(v)), "r"(1u) : "memory");
}

// NV values x[c] = line[c * STRIDE]: all loads of a round are issued together; rounds repeat until
// every line carries its tags.  The watchdog raises the abort flag instead of hanging.
template <int NV, int STRIDE>
__device__ __forceinline__ void poll_lines(double (&x)[NV], const uint4* line, int* abort_flag, bool& aborted) {
    long long t0 = 0;
    unsigned spins = 0;
    for (;;) {
        bool ok = true;
#pragma unroll
        for (int c = 0; c < NV; c++) {
            unsigned x0, x1, x2, x3;
            asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "l"(line + c * STRIDE) : "memory");
            ok = ok && (x1 == 1u) && (x3 == 1u);
            x[c] = __hiloint2double((int)x2, (int)x0);
        }
        if (ok || aborted) return;
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

// Sums of 16 per-lane values over the 32 lanes of a warp with 16 exchanges instead of 16 x 5:
// every round halves the number of values a lane carries (it keeps the half its lane bit selects
// and adds the partner's copy of that half), so after offsets 16, 8, 4, 2 a lane holds one value
// and offset 1 completes it.  Returns the total of v[lane >> 1] (both lanes of a pair hold it).
// Fixed order, so results do not depend on timing.  (The butterfly per value cost 160 SHFL and 80
// DADD per warp twice per block step: the largest item on the dependent chain of the sweeps.)
__device__ __forceinline__ double warp_reduce16(const double (&v)[16], int lane) {
    double a8[8], a4[4], a2[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const double keep = b4 ? v[i + 8] : v[i], send = b4 ? v[i] : v[i + 8];
        a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const double keep = b3 ? a8[i + 4] : a8[i], send = b3 ? a8[i] : a8[i + 4];
        a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const double keep = b2 ? a4[i + 2] : a4[i], send = b2 ? a4[i] : a4[i + 2];
        a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const double keep = b1 ? a2[1] : a2[0], send = b1 ? a2[0] : a2[1];
    double r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
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
            double x[4];
            poll_lines<4, 32>(x, a.zline + (int64_t)j * NB + lane, a.abort_flag, aborted);
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
        {
            const double s = warp_reduce16(part, lane);          // row warp + 8 (lane >> 1)
            const int row = warp + 8 * (lane >> 1);
            if ((lane & 1) == 0) sw[row] = a.r[o + row] - s;
        }
        abg::cp_async_wait<0>();
        __syncthreads();
        double x[4];
#pragma unroll
        for (int c = 0; c < 4; c++) x[c] = sw[lane + 32 * c];
        double dz[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const double* Dr = sD + (warp + 8 * i) * NB + lane;
            double s = Dr[0] * x[0];
            s = fma(Dr[32], x[1], s);
            s = fma(Dr[64], x[2], s);
            s = fma(Dr[96], x[3], s);
            dz[i] = s;
        }
        {
            const double s = warp_reduce16(dz, lane);
            const int row = warp + 8 * (lane >> 1);
            if ((lane & 1) == 0) {
                st_line(a.zline + o + row, s);
                a.z[o + row] = s;
            }
        }
        __syncthreads();                                  // sD / sw are free for the next task
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
            double xr[16];
            poll_lines<16, 8>(xr, a.aline + (int64_t)j * NB + warp, a.abort_flag, aborted);
#pragma unroll
            for (int i = 0; i < 16; i++)
#pragma unroll
                for (int c = 0; c < 4; c++) y[c] = fma(v[i][c], xr[i], y[c]);
            if (j - 1 > k) load_block(v, a.L + (int64_t)(j - 1) * NB * a.ld + o, a.ld, warp, lane);
        }
#pragma unroll
        for (int c = 0; c < 4; c++) spart[warp][lane + 32 * c] = y[c];
        __syncthreads();
        if (tid < NB) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < 8; w++) s += spart[w][tid];
            double zk[1];                                 // z_k (own forward row when G divides evenly)
            poll_lines<1, 1>(zk, a.zline + o + tid, a.abort_flag, aborted);
            sw[tid] = zk[0] - s;
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
            st_line(a.aline + o + tid, s);
            a.alpha[o + tid] = s;
        }
        __syncthreads();
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
    // control area behind the sums slot of scratch: [abort flag, padded to 256 B][z lines npad][alpha lines npad]
    const size_t line_bytes = (size_t)2 * h->npad * sizeof(uint4);
    int rc = ab_ensure_scratch(h, 512 + line_bytes);
    if (rc) return rc;
    int* ctrl = reinterpret_cast<int*>(h->scratch + 32);
    cudaStream_t s = h->stream;
    AB_CUDA(cudaMemsetAsync(ctrl, 0, 256 + line_bytes, s));
    TrsvArgs a;
    a.L = h->L; a.ld = h->npad; a.Dinv = h->Dinv; a.T = T;
    a.r = r; a.z = h->z; a.alpha = h->alpha;
    a.abort_flag = ctrl;
    a.zline = reinterpret_cast<uint4*>(reinterpret_cast<char*>(ctrl) + 256);
    a.aline = a.zline + h->npad;
    a.backward = backward;
    int grid = T < h->nsm ? T : h->nsm;
    void* args[] = {(void*)&a};
    AB_CUDA(cudaLaunchCooperativeKernel((void*)trsv_dataflow_kernel, dim3(grid), dim3(256), args, smem, s));
    ab_count_launches(1);
    AB_CUDA(cudaMemcpyAsync(h->h_pinned + 10, ctrl, sizeof(int), cudaMemcpyDeviceToHost, s));
    return 0;
}
