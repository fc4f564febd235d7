// Development micro-benchmark (not part of the product): the FP64 DMMA main loop of
// alabi_b200/csrc/dmma_gemm.cuh fed by TMA instead of cp.async, so that the design decision
// "cp.async ring with padded pitch, not TMA"  rests on a measurement.
//
//   C(128 x 128) += A(128 x K, K-major) * B(K x 128, column index contiguous)     [predict_var's shape]
//
// TMA variant: per 16-deep k chunk ONE cp.async.bulk.tensor.2d for A (box 16 x 128 doubles, dense
// 128-byte rows, SWIZZLE_128B) and EIGHT for B (boxes of 16 columns x 16 k rows, 128-byte rows,
// SWIZZLE_128B), completion on an mbarrier per stage, slots released through a second mbarrier per
// stage (no CTA-wide barrier in the loop).  A dense 128-byte row pitch puts the 4 rows x 4 k-values
// a half-warp reads with one LDS.64 onto the same banks; the swizzle alone leaves a 2-way conflict
// (DESIGN.md), so the fragment rows / columns are PERMUTED inside their 8-row groups / 16-column
// boxes: lane group g reads row (2g mod 8) + g / 4 of A and column {0,1,8,9,2,3,10,11}[g] of B, which
// makes the 16 lanes of a half-warp hit 16 distinct 8-byte bank pairs.  (The permutation only
// relabels output rows / columns; an epilogue would undo it.)
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I alabi_b200/csrc -I include \
//             -o build/tma_gemm_bench tools/tma_gemm_bench.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "dmma_gemm.cuh"

namespace tma {

constexpr int BK = 16, STAGES = 4, THREADS = 256;
constexpr int A_BYTES = 128 * BK * 8, B_BYTES = BK * 128 * 8, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;     // + alignment slack + barriers

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// byte offset of element (row, k) inside a swizzled [rows][16 doubles] tile (128-byte rows)
__device__ __forceinline__ int swz(int row, int k) { return row * 128 + ((((k >> 1) ^ (row & 7)) << 4) | ((k & 1) << 3)); }

__global__ void __launch_bounds__(THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int nk, int reps,
            double* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + STAGES * STAGE_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3, wm = warp >> 2, wn = warp & 3;
    const int pg = ((2 * g) & 7) + (g >> 2);                         // permuted row inside an 8-row group
    const int pc = (g & 1) + ((g >> 1) & 1) * 8 + (g >> 2) * 2;      // permuted column inside a 16-column box
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(smem_u32(&bars[s]), 1);                        // full: one expect_tx arrival
            mbar_init(smem_u32(&bars[STAGES + s]), THREADS / 32);    // empty: one arrival per warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    abg::Acc acc;
    acc.zero();
    const long long total = (long long)nk * reps;
    auto issue = [&](long long c) {                                   // thread 0 only
        const int s = (int)(c % STAGES);
        const int kc = (int)(c % nk);
        const unsigned full = smem_u32(&bars[s]);
        unsigned char* st = base + s * STAGE_BYTES;
        mbar_expect_tx(full, STAGE_BYTES);
        tma_load_2d(smem_u32(st), &mapA, kc * BK, 0, full);
#pragma unroll
        for (int b = 0; b < 8; b++) tma_load_2d(smem_u32(st + A_BYTES + b * 2048), &mapB, b * 16, kc * BK, full);
    };
    if (tid == 0)
        for (long long c = 0; c < STAGES - 1 && c < total; c++) issue(c);
    for (long long c = 0; c < total; c++) {
        const int s = (int)(c % STAGES);
        mbar_wait(smem_u32(&bars[s]), (unsigned)((c / STAGES) & 1));
        const unsigned char* sA = base + s * STAGE_BYTES;
        const unsigned char* sB = sA + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / 4; kk++) {
            double a[8], b[4];
#pragma unroll
            for (int f = 0; f < 8; f++) a[f] = *reinterpret_cast<const double*>(sA + swz(wm * 64 + f * 8 + pg, kk * 4 + t));
#pragma unroll
            for (int f = 0; f < 4; f++) {
                // n-fragment f of warp column wn: box (wn * 2 + f / 2), half f % 2 of the box
                const int box = wn * 2 + (f >> 1), col = pc + (f & 1) * 4;
                b[f] = *reinterpret_cast<const double*>(sB + box * 2048 + swz(kk * 4 + t, col));
            }
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) abg::dmma884(acc.v[i][j][0], acc.v[i][j][1], a[i], b[j]);
            if (kk == 0 && tid == 0) {
                // refill the slot chunk c - 1 used, once every warp has released it
                const long long nx = c + STAGES - 1;
                if (nx < total) {
                    if (c > 0) mbar_wait(smem_u32(&bars[STAGES + (int)((c - 1) % STAGES)]), (unsigned)(((c - 1) / STAGES) & 1));
                    issue(nx);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[STAGES + s]));
    }
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) sum += acc.v[i][j][0] + acc.v[i][j][1];
    out[blockIdx.x * THREADS + threadIdx.x] = sum;
}

}  // namespace tma

// cp.async reference with the product configuration (same shape)
template <typename CFG>
__global__ void __launch_bounds__(CFG::THREADS, 1)
ref_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb, int nk, int reps,
           double* __restrict__ out) {
    extern __shared__ __align__(16) double smem[];
    typename CFG::Acc acc;
    acc.zero();
    for (int r = 0; r < reps; r++) CFG::template mainloop<true, false, false>(acc, A, lda, B, ldb, nk, smem);
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < CFG::MF; i++)
#pragma unroll
        for (int j = 0; j < CFG::NF; j++) s += acc.v[i][j][0] + acc.v[i][j][1];
    out[blockIdx.x * CFG::THREADS + threadIdx.x] = s;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
        printf("cuTensorMapEncodeTiled not available\n");
        return 1;
    }
    for (int K : {4096, 512, 128}) {
        const int reps = 32768 / K;
        double *A, *B, *out;
        cudaMalloc(&A, sizeof(double) * 128 * K); cudaMalloc(&B, sizeof(double) * (size_t)K * 128);
        cudaMalloc(&out, sizeof(double) * nsm * 256);
        // values chosen so that the two kernels can be compared: A[r][k] = small, B[k][c] = small
        std::vector<double> hA((size_t)128 * K), hB((size_t)K * 128);
        for (int r = 0; r < 128; r++) for (int k = 0; k < K; k++) hA[(size_t)r * K + k] = 1e-3 * ((r * 7 + k * 3) % 11 - 5);
        for (int k = 0; k < K; k++) for (int c = 0; c < 128; c++) hB[(size_t)k * 128 + c] = 1e-3 * ((k * 5 + c) % 13 - 6);
        cudaMemcpy(A, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(B, hB.data(), hB.size() * 8, cudaMemcpyHostToDevice);
        CUtensorMap mapA, mapB;
        {   // A: [128 rows][K] doubles, box {16 k, 128 rows}
            cuuint64_t dims[2] = {(cuuint64_t)K, 128}, strides[1] = {(cuuint64_t)K * 8};
            cuuint32_t box[2] = {16, 128}, es[2] = {1, 1};
            CUresult r = encode(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, A, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 1; }
        }
        {   // B: [K rows][128 cols] doubles, box {16 cols, 16 k rows}
            cuuint64_t dims[2] = {128, (cuuint64_t)K}, strides[1] = {128 * 8};
            cuuint32_t box[2] = {16, 16}, es[2] = {1, 1};
            CUresult r = encode(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, B, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 1; }
        }
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto time_it = [&](auto launch) {
            float best = 1e30f;
            for (int r = 0; r < 5; r++) {
                cudaEventRecord(e0);
                launch();
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (r > 0 && ms < best) best = ms;
            }
            return best;
        };
        const double fl = 2.0 * 128 * 128 * (double)K * reps * nsm;
        cudaFuncSetAttribute(tma::gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tma::SMEM_BYTES);
        float t_tma = time_it([&] { tma::gemm_kernel<<<nsm, tma::THREADS, tma::SMEM_BYTES>>>(mapA, mapB, K / 16, reps, out); });
        cudaError_t err = cudaDeviceSynchronize();
        std::vector<double> o1((size_t)nsm * 256), o2((size_t)nsm * 256);
        cudaMemcpy(o1.data(), out, o1.size() * 8, cudaMemcpyDeviceToHost);
        using C1 = abg::Core<16, 4, 1>;
        cudaFuncSetAttribute(ref_kernel<C1>, cudaFuncAttributeMaxDynamicSharedMemorySize, C1::SMEM_BYTES);
        float t_ref = time_it([&] { ref_kernel<C1><<<nsm, C1::THREADS, C1::SMEM_BYTES>>>(A, K, B, 128, K / 16, reps, out); });
        cudaMemcpy(o2.data(), out, o2.size() * 8, cudaMemcpyDeviceToHost);
        // both kernels hold the same 128 x 128 product, distributed differently over threads: compare the CTA totals
        double s1 = 0, s2 = 0;
        for (int i = 0; i < 256; i++) { s1 += o1[i]; s2 += o2[i]; }
        printf("K=%5d reps=%4d  TMA + mbarrier (swizzle-128B, permuted fragments): %7.3f ms %6.2f TFLOP/s | cp.async ring (product): %7.3f ms %6.2f TFLOP/s"
               " | CTA sums %.12e vs %.12e (%s)\n", K, reps, t_tma, fl / t_tma * 1e-9, t_ref, fl / t_ref * 1e-9, s1, s2,
               cudaGetErrorString(err));
        cudaFree(A); cudaFree(B); cudaFree(out);
    }
    return 0;
}
