// K2, dataflow form: the whole blocked Cholesky in ONE persistent kernel.
//
// Replaces george BasicSolver.compute (scipy cholesky) reached from
// alabi/gp_utils.py:243, alabi/core.py:1158,1430 — same result as chol.cu's
// multi-launch right-looking sweep, different schedule:
//
//   * one task per 128 x 128 tile (i, j), j <= i, of the lower triangle, handed
//     out in column-major order through an atomic counter (one CTA per SM);
//   * task (i, j) is LEFT-looking: it accumulates  S = sum_{k<j} L_ik L_jk^T  in one
//     long-K DMMA main loop (the tile of K is read once and written once, instead
//     of being re-read and re-written by every trailing update), then
//         i == j :  L_jj = chol(A_jj - S), D_j^-1 = L_jj^-1   (register-tiled sweep)
//         i  > j :  L_ij = (A_ij - S) D_j^-T                  (DMMA, A operand resident in smem)
//   * dependencies are per block row: prog[r] = number of final tiles in row r
//     (tiles of a row become final left to right).  The main loop waits on
//     prog[i] > k and prog[j] > k just before it prefetches k block k, so a task
//     runs ahead through every k block that is already final — look-ahead falls
//     out of the task order, there are no launch gaps and no wave quantisation.
//
// Deadlock freedom: tasks are popped in an order in which every dependency has
// a smaller index, and only resident CTAs hold tasks, so the lowest-index
// unfinished task can always run.  Every wait loop also carries a watchdog
// (about 2 s) that raises an abort flag instead of hanging the GPU.
#include <vector>
#include "handle.h"
#include "dmma_gemm.cuh"
#include "potf2.cuh"

namespace {

constexpr int NB = AB_NB;
using Core = abg::Main;
constexpr int LDC = NB + 4;                              // resident C tile [128][132]
constexpr int C_ELEMS = NB * LDC;
constexpr int BSTAGE = NB * Core::LDK;                   // one Dinv chunk [128][20]
constexpr int DF_SMEM_BYTES = (C_ELEMS + Core::STAGES * BSTAGE) * 8;     // 217088
static_assert(DF_SMEM_BYTES >= Core::SMEM_BYTES && DF_SMEM_BYTES <= 227 * 1024, "shared memory plan");

struct DfArgs {
    double* A; int64_t ld; int T;
    double* Dinv; double* logdet_parts; int* info;
    int* prog;                    // [T] final tiles per block row
    unsigned int* next_task;      // task counter
    int* abort_flag;
    const int2* tasks; int ntasks;
};

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// wait until prog[row] > k; `seen` caches the last value this thread observed
struct Waiter {
    const int* prog; int* abort_flag; bool aborted;
    __device__ __forceinline__ void wait(int row, int k, int& seen) {
        if (seen > k || aborted) return;
        long long t0 = 0;
        unsigned spins = 0;
        for (;;) {
            seen = ld_acquire(prog + row);
            if (seen > k) break;
            __nanosleep(40);
            if ((++spins & 255u) == 0) {
                if (t0 == 0) t0 = clock64();
                if (*((volatile int*)abort_flag) != 0 || clock64() - t0 > 4000000000LL) {
                    atomicExch(abort_flag, 1);
                    aborted = true;
                    return;
                }
            }
        }
    }
};

struct RowGate {
    Waiter* w; int i, j; int* seen_i; int* seen_j;
    __device__ __forceinline__ void operator()(int c) const {
        if ((c % Core::KB) == 0) {
            const int k = c / Core::KB;
            w->wait(i, k, *seen_i);
            w->wait(j, k, *seen_j);
        }
    }
};

__global__ void __launch_bounds__(Core::THREADS, 1)
chol_dataflow_kernel(const __grid_constant__ DfArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ abp::SweepShared sh;
    __shared__ int s_task;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3, wm = warp / Core::WN, wn = warp % Core::WN;
    Waiter w{a.prog, a.abort_flag, false};
    double* sC = smem;
    double* sBring = smem + C_ELEMS;
    double* const sdiag = sh.sdiag;
    double* const sinv = sh.sinv;
    unsigned sweep_phase = 0;
    if (tid == 0) abp::mbar_init((unsigned)__cvta_generic_to_shared(&sh.mbar), Core::THREADS);
    __syncthreads();

    for (;;) {
        if (tid == 0) s_task = (int)atomicAdd(a.next_task, 1u);
        __syncthreads();
        const int task = s_task;
        if (task >= a.ntasks) break;
        const int i = a.tasks[task].x, j = a.tasks[task].y;
        const int64_t row0 = (int64_t)i * NB, col0 = (int64_t)j * NB;
        int seen_i = 0, seen_j = 0;

        // ---- S = sum_{k<j} L_ik L_jk^T, gated per k block --------------------------
        abg::Acc acc;
        acc.zero();
        if (j > 0) {
            RowGate gate{&w, i, j, &seen_i, &seen_j};
            Core::mainloop<true, true, false, RowGate>(acc, a.A + row0 * a.ld, a.ld, a.A + col0 * a.ld, a.ld,
                                                       j * Core::KB, smem, gate);
        }
        // ---- C = A_ij - S, staged in shared memory ------------------------------------
        // (A_ij still holds the covariance values written before this launch)
#pragma unroll
        for (int f = 0; f < 8; f++) {
            const int r = Core::acc_row(f);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int c = Core::acc_col(q);
                const double2 v = *reinterpret_cast<const double2*>(a.A + (row0 + r) * a.ld + col0 + c);
                double2 o;
                o.x = v.x - acc.v[f][q][0];
                o.y = v.y - acc.v[f][q][1];
                *reinterpret_cast<double2*>(sC + r * LDC + c) = o;
            }
        }
        __syncthreads();

        if (i == j) {
            // ---- diagonal tile: Cholesky + inverse (potf2.cuh sweep) --------------------
            const int tx = tid & 15, ty = tid >> 4;
            double pa[8][8], pb[8][8];
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int c = 0; c < 8; c++)
                    if (r >= c) {
                        const int ii = ty + 16 * r, kk = tx + 16 * c;
                        pa[r][c] = (kk <= ii) ? sC[ii * LDC + kk] : 0.0;
                        pb[r][c] = (kk == ii) ? 1.0 : 0.0;
                    }
            abp::potf2_sweep<true>(pa, pb, sh, tx, ty, col0, a.info, sweep_phase);
            if (tid < 32) {
                double s = 0.0;
                for (int q = tid; q < NB; q += 32) s += 2.0 * log(sdiag[q]);
                s = ab_warp_sum(s);
                if (tid == 0) a.logdet_parts[j] = s;
            }
            double* Dj = a.Dinv + (int64_t)j * NB * NB;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int ii = ty + 16 * r;
                const double si = sinv[ii];
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const int kk = tx + 16 * c;
                    if (r >= c) {
                        a.A[(col0 + ii) * a.ld + col0 + kk] = (kk < ii) ? pa[r][c] * sinv[kk] : ((kk == ii) ? sdiag[ii] : 0.0);
                        Dj[ii * NB + kk] = (kk <= ii) ? pb[r][c] * si : 0.0;
                    } else {
                        a.A[(col0 + ii) * a.ld + col0 + kk] = 0.0;
                        Dj[ii * NB + kk] = 0.0;
                    }
                }
            }
        } else {
            // ---- off-diagonal tile: L_ij = C D_j^-T (A operand resident, D_j^-1 streamed) ----
            w.wait(j, j, seen_j);                                    // diagonal tile of column j is final
            const double* Dj = a.Dinv + (int64_t)j * NB * NB;
            abg::Acc out;
            out.zero();
            constexpr int NK = NB / Core::BK;
#pragma unroll
            for (int s = 0; s < Core::STAGES - 1; s++) {
                Core::load_operand<true, NB>(sBring + s * BSTAGE, Dj + s * Core::BK, NB, tid);
                abg::cp_async_commit();
            }
            for (int kc = 0; kc < NK; kc++) {
                abg::cp_async_wait<Core::STAGES - 2>();
                __syncthreads();
                const double* sB = sBring + (kc % Core::STAGES) * BSTAGE;
                const int nx = kc + Core::STAGES - 1;
                if (nx < NK) Core::load_operand<true, NB>(sBring + (nx % Core::STAGES) * BSTAGE, Dj + nx * Core::BK, NB, tid);
                abg::cp_async_commit();
                // D_j^-1 is lower triangular: column block c of the output only needs k <= c
#pragma unroll
                for (int kk = 0; kk < Core::BK / 4; kk++) {
                    const int k0 = kc * Core::BK + kk * 4;
                    if (k0 > wn * 32 + 31) continue;                 // warp-uniform
                    double fa[8], fb[4];
#pragma unroll
                    for (int f = 0; f < 8; f++) fa[f] = sC[(wm * 64 + f * 8 + g) * LDC + k0 + t];
#pragma unroll
                    for (int f = 0; f < 4; f++) fb[f] = sB[(wn * 32 + f * 8 + g) * Core::LDK + kk * 4 + t];
#pragma unroll
                    for (int f = 0; f < 8; f++)
#pragma unroll
                        for (int q = 0; q < 4; q++) abg::dmma884(out.v[f][q][0], out.v[f][q][1], fa[f], fb[q]);
                }
            }
            abg::cp_async_wait<0>();
            Core::store_tile(out, a.A + row0 * a.ld + col0, a.ld, 1.0, 0.0);
        }
        // ---- publish: the tile (and D_j^-1) is final -------------------------------------
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(a.prog + i, j + 1);
    }
}

}  // namespace

int ab_launch_factor_dataflow(ab_gp* h) {
    static bool configured = false;
    if (!configured) {
        AB_CUDA(cudaFuncSetAttribute(chol_dataflow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DF_SMEM_BYTES));
        configured = true;
    }
    const int T = (int)(h->npad / NB);
    const int ntasks = T * (T + 1) / 2;
    // control block in scratch: [0] next_task, [1] abort, [2..2+T) prog
    const size_t ctrl_ints = 2 + (size_t)T;
    int rc = ab_ensure_scratch(h, ctrl_ints * sizeof(int));
    if (rc) return rc;
    if (h->df_tasks_T != T) {                      // column-major task list for this T (host -> device once per size)
        std::vector<int2> tasks;
        tasks.reserve(ntasks);
        for (int j = 0; j < T; j++)
            for (int i = j; i < T; i++) tasks.push_back(make_int2(i, j));
        if (h->df_tasks) cudaFree(h->df_tasks);
        h->df_tasks = nullptr;
        AB_CUDA(cudaMalloc(&h->df_tasks, (size_t)ntasks * sizeof(int2)));
        AB_CUDA(cudaMemcpyAsync(h->df_tasks, tasks.data(), (size_t)ntasks * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
        AB_CUDA(cudaStreamSynchronize(h->stream));
        h->df_tasks_T = T;
    }
    cudaStream_t s = h->stream;
    int* ctrl = reinterpret_cast<int*>(h->scratch);
    AB_CUDA(cudaMemsetAsync(ctrl, 0, ctrl_ints * sizeof(int), s));
    AB_CUDA(cudaMemsetAsync(h->d_info, 0, sizeof(int), s));
    DfArgs a;
    a.A = h->L; a.ld = h->npad; a.T = T;
    a.Dinv = h->Dinv; a.logdet_parts = h->logdet_parts; a.info = h->d_info;
    a.next_task = reinterpret_cast<unsigned int*>(ctrl);
    a.abort_flag = ctrl + 1;
    a.prog = ctrl + 2;
    a.tasks = reinterpret_cast<const int2*>(h->df_tasks); a.ntasks = ntasks;
    const int grid = ntasks < h->nsm ? ntasks : h->nsm;
    ab_prof_begin(h, AB_PROF_FACTOR);
    chol_dataflow_kernel<<<grid, Core::THREADS, DF_SMEM_BYTES, s>>>(a);
    ab_prof_end(h, AB_PROF_FACTOR);
    ab_count_launches(1);
    AB_CHECK_LAUNCH();
    AB_CUDA(cudaMemcpyAsync(h->h_pinned + 9, ctrl + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
    return 0;
}
