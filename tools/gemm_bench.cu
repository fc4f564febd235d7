// Development micro-benchmark for the DMMA GEMM core (not part of the product):
// every CTA multiplies a 128 x K by a K x 128 operand that stays L2 resident, so
// the number is the main loop's own ceiling.  Variants: BK, stages, issue order.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I alabi_b200/csrc -I include -o build/gemm_bench tools/gemm_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "dmma_gemm.cuh"

template <typename CFG, bool AK, bool BKM>
__global__ void __launch_bounds__(CFG::THREADS, 1)
bench_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb, int nk, int reps,
             double* __restrict__ out) {
    extern __shared__ __align__(16) double smem[];
    typename CFG::Acc acc;
    acc.zero();
    for (int r = 0; r < reps; r++) CFG::template mainloop<AK, BKM, false>(acc, A, lda, B, ldb, nk, smem);
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < CFG::MF; i++)
#pragma unroll
        for (int j = 0; j < CFG::NF; j++) s += acc.v[i][j][0] + acc.v[i][j][1];
    out[blockIdx.x * CFG::THREADS + threadIdx.x] = s;
}

// predict_var-shaped: T row blocks, row block i multiplies (i + 1) k blocks, the last triangular
template <typename CFG>
__global__ void __launch_bounds__(CFG::THREADS, 1)
tri_kernel(const double* __restrict__ Linv, int64_t ld, int T, const double* __restrict__ P, int64_t ldp, double* __restrict__ out) {
    extern __shared__ __align__(16) double smem[];
    const double* Bp = P + (int64_t)blockIdx.x * 128;
    double ss = 0.0;
    for (int i = 0; i < T; i++) {
        typename CFG::Acc acc;
        acc.zero();
        CFG::template mainloop<true, false, true>(acc, Linv + (int64_t)i * 128 * ld, ld, Bp, ldp, (i + 1) * CFG::KB, smem);
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int j = 0; j < 4; j++) ss = fma(acc.v[a][j][0], acc.v[a][j][0], fma(acc.v[a][j][1], acc.v[a][j][1], ss));
    }
    out[blockIdx.x * CFG::THREADS + threadIdx.x] = ss;
}

template <typename CFG>
void run_tri(const char* name, int T, int waves) {
    const int64_t n = 128 * T, ldp = 148 * 128 * waves;
    double *Linv, *P, *out;
    cudaMalloc(&Linv, sizeof(double) * n * n); cudaMalloc(&P, sizeof(double) * n * ldp); cudaMalloc(&out, sizeof(double) * 148 * 256 * waves);
    cudaMemset(Linv, 0, sizeof(double) * n * n); cudaMemset(P, 0, sizeof(double) * n * ldp);
    auto kern = tri_kernel<CFG>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::SMEM_BYTES);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        kern<<<148 * waves, CFG::THREADS, CFG::SMEM_BYTES>>>(Linv, n, T, P, ldp, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    double fl = (double)n * n * ldp;     // algorithmic: N^2 flop per query
    printf("%-34s predict_var-shaped T=%3d waves=%d: %7.3f ms  %6.2f TFLOP/s algorithmic  (%s)\n", name, T, waves, best,
           fl / best * 1e-9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(Linv); cudaFree(P); cudaFree(out);
}

template <typename CFG, bool AK, bool BKM>
void run(const char* name, const double* A, const double* B, double* out, int K, int reps) {
    auto kern = bench_kernel<CFG, AK, BKM>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::SMEM_BYTES);
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int64_t lda = AK ? K : 128, ldb = BKM ? K : 128;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        kern<<<nsm, CFG::THREADS, CFG::SMEM_BYTES>>>(A, lda, B, ldb, K / CFG::BK, reps, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    double fl = 2.0 * 128 * 128 * (double)K * reps * nsm;
    printf("%-34s AK=%d BKM=%d K=%5d reps=%3d: %7.3f ms  %6.2f TFLOP/s  (%s)\n", name, AK, BKM, K, reps, best, fl / best * 1e-9,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int K = 4096;
    double *A, *B, *out;
    cudaMalloc(&A, sizeof(double) * 128 * K); cudaMalloc(&B, sizeof(double) * 128 * K);
    cudaMalloc(&out, sizeof(double) * 148 * 1024);
    cudaMemset(A, 0, sizeof(double) * 128 * K); cudaMemset(B, 0, sizeof(double) * 128 * K);
#define RUN(CFG, name)                                   \
    run<CFG, true, true>(name, A, B, out, K, 8);         \
    run<CFG, true, false>(name, A, B, out, K, 8);        \
    run<CFG, true, false>(name, A, B, out, 128, 256);    \
    run<CFG, true, false>(name, A, B, out, 512, 64);
    using C0 = abg::Core<16, 4, 0>; using C1 = abg::Core<16, 4, 1>; using C2 = abg::Core<32, 3, 0>;
    using C3 = abg::Core<32, 3, 1>; using C4 = abg::Core<16, 5, 1>;
    RUN(C0, "BK16 S4 prefetch-first (r01a)");
    RUN(C1, "BK16 S4 compute-first");
    RUN(C2, "BK32 S3 prefetch-first");
    RUN(C3, "BK32 S3 compute-first");
    RUN(C4, "BK16 S5 compute-first");
    run_tri<C0>("BK16 S4 prefetch-first", 8, 4); run_tri<C1>("BK16 S4 compute-first", 8, 4);
    run_tri<C3>("BK32 S3 compute-first", 8, 4);
    run_tri<C0>("BK16 S4 prefetch-first", 64, 1); run_tri<C1>("BK16 S4 compute-first", 64, 1);
    run_tri<C3>("BK32 S3 compute-first", 64, 1);
    return 0;
}
