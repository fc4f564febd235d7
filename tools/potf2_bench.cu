// Development micro-benchmark: cycles per column of the register-tiled potf2 sweeps.
#include <cstdio>
#include <cuda_runtime.h>
#include "potf2.cuh"
void ab_set_error(const char*, ...) {}
template <int R>
__global__ void __launch_bounds__(256, 1) k(double* out, long long* cyc, int* info) {
    __shared__ abp::SweepShared sh;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    if (tid == 0) abp::mbar_init((unsigned)__cvta_generic_to_shared(&sh.mbar), 256);
    __syncthreads();
    double a[R][R], b[R][R];
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int c = 0; c < R; c++) if (r >= c) {
            const int i = ty + 16 * r, kk = tx + 16 * c;
            a[r][c] = (kk == i) ? 16.0 * R + 1.0 : ((kk < i) ? 0.01 * ((i * 7 + kk * 3) % 11) : 0.0);
            b[r][c] = (kk == i) ? 1.0 : 0.0;
        }
    unsigned phase = 0;
    long long t0 = clock64();
    abp::potf2_sweep<true, R>(a, b, sh, tx, ty, 0, info, phase);
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int c = 0; c < R; c++) if (r >= c) s += a[r][c] + b[r][c];
    out[tid] = s;
    if (tid == 0) cyc[0] = t1 - t0;
}
int main() {
    double* out; long long* cyc; int* info; cudaMalloc(&out, 8 * 256); cudaMalloc(&cyc, 8); cudaMalloc(&info, 4);
    long long h;
    for (int rep = 0; rep < 2; rep++) {
        k<8><<<1, 256>>>(out, cyc, info); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("R=8 (128 cols): %lld cycles, %.0f per column\n", h, h / 128.0);
        k<4><<<1, 256>>>(out, cyc, info); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("R=4 ( 64 cols): %lld cycles, %.0f per column\n", h, h / 64.0);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
