// In-run FP64 yard-stick for bench.py: issue-rate peak of DMMA.8x8x4 (the FP64
// tensor pipe; MEASURED_PEAKS.json carries only HBM and bf16 figures).
#include "handle.h"
#include "alabi_b200.h"

namespace {
__global__ void __launch_bounds__(512) dmma_peak_kernel(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// FP64 FMA issue-rate peak (the roof of the kernel-evaluation loops: covariance build, predictive
// mean, sampler): 8 independent DFMA chains per thread.
__global__ void __launch_bounds__(512) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" int ab_fp64_fma_peak(int device, double* h_tflops) {
    if (!h_tflops) { ab_set_error("null argument"); return -1; }
    AB_CUDA(cudaSetDevice(device));
    int nsm = 0;
    AB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    const int threads = 512, blocks = nsm * 2, iters = 16384;
    double* out = nullptr;
    AB_CUDA(cudaMalloc(&out, sizeof(double) * threads * blocks));
    cudaEvent_t e0, e1;
    AB_CUDA(cudaEventCreate(&e0));
    AB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        AB_CUDA(cudaEventRecord(e0, 0));
        dfma_peak_kernel<<<blocks, threads>>>(out, iters, 0.9999999, 1e-9);
        AB_CUDA(cudaEventRecord(e1, 0));
        AB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        AB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best) best = ms;
    }
    AB_CHECK_LAUNCH();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *h_tflops = 2.0 * 8 * iters * (double)threads * blocks / (best * 1e-3) * 1e-12;
    return 0;
}

extern "C" int ab_fp64_tensor_peak(int device, double* h_tflops) {
    if (!h_tflops) { ab_set_error("null argument"); return -1; }
    AB_CUDA(cudaSetDevice(device));
    int nsm = 0;
    AB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    const int threads = 512, blocks = nsm * 2, iters = 8192;
    double* out = nullptr;
    AB_CUDA(cudaMalloc(&out, sizeof(double) * threads * blocks));
    cudaEvent_t e0, e1;
    AB_CUDA(cudaEventCreate(&e0));
    AB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        AB_CUDA(cudaEventRecord(e0, 0));
        dmma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        AB_CUDA(cudaEventRecord(e1, 0));
        AB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        AB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best) best = ms;
    }
    AB_CHECK_LAUNCH();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *h_tflops = 512.0 * 8 * iters * (threads / 32.0) * blocks / (best * 1e-3) * 1e-12;
    return 0;
}
