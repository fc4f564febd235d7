// Development micro-benchmark: dependent-issue latency of FP64 ops on the current GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_chain(double* out, long long* cyc, int iters, double a, double b) {
    double c = threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) c = fma(c, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = c;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void rcp_chain(double* out, long long* cyc, int iters) {
    double c = 1.0 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(c)); c = r; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = c;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lds_chain(double* out, long long* cyc, int iters) {
    __shared__ int idx[256];
    idx[threadIdx.x] = (threadIdx.x + 1) & 255;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) p = idx[p];
    }
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void bar_chain(double* out, long long* cyc, int iters) {
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
    long long h;
    const int iters = 1000;
    for (int threads : {32, 256}) {
        dfma_chain<<<1, threads>>>(out, cyc, iters, 1.0000001, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("threads=%3d DFMA dependent latency: %.1f cycles\n", threads, (double)h / (iters * 16));
        rcp_chain<<<1, threads>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("threads=%3d MUFU.RCP64H dependent latency: %.1f cycles\n", threads, (double)h / (iters * 16));
        lds_chain<<<1, threads>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("threads=%3d LDS dependent latency: %.1f cycles\n", threads, (double)h / (iters * 16));
        bar_chain<<<1, threads>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("threads=%3d __syncthreads: %.1f cycles\n", threads, (double)h / (iters * 16));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
