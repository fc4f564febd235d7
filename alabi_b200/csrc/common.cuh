// Shared device/host helpers for the alabi_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define AB_MAX_DIM 32     // maximum input dimension (per-dimension metric)
#define AB_NB 128         // block size of every blocked dense algorithm; matrices are padded to it

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL)
#error "alabi_b200 kernels are written for sm_100a (compile with -gencode arch=compute_100a,code=sm_100a)"
#endif

// Stationary kernel with an axis-aligned metric.  Coordinates are pre-scaled by
// 1/sqrt(M_k) so r^2 is a plain squared distance (george: r^2 = sum dx_k^2 / M_k;
// reference call site alabi/core.py:998-1014).
struct KernParams {
    int kind;                     // 0 ExpSquared, 1 Matern32, 2 Matern52
    int d;
    double amp;                   // exp(log_constant), 1 without an amplitude
    double diag_add;              // yerr^2 + exp(white_noise)
    double inv_len[AB_MAX_DIM];   // exp(-0.5 log_M_k)
};

template <int KIND>
__device__ __forceinline__ double ab_radial(double r2) {
    if (KIND == 0) return exp(-0.5 * r2);
    if (KIND == 1) { double r = sqrt(3.0 * r2); return (1.0 + r) * exp(-r); }
    double r = sqrt(5.0 * r2);
    return (1.0 + r + r * r / 3.0) * exp(-r);
}

// dk/d(r^2)
template <int KIND>
__device__ __forceinline__ double ab_radial_grad(double r2) {
    if (KIND == 0) return -0.5 * exp(-0.5 * r2);
    if (KIND == 1) { double r = sqrt(3.0 * r2); return -1.5 * exp(-r); }
    double r = sqrt(5.0 * r2);
    return -5.0 * (1.0 + r) * exp(-r) / 6.0;
}

#define AB_DISPATCH_KIND(kind, ...)                                   \
    do {                                                              \
        if ((kind) == 0) { constexpr int KIND = 0; __VA_ARGS__; }     \
        else if ((kind) == 1) { constexpr int KIND = 1; __VA_ARGS__; }\
        else { constexpr int KIND = 2; __VA_ARGS__; }                 \
    } while (0)

__device__ __forceinline__ double ab_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// error plumbing (api.cu owns the storage)
void ab_set_error(const char* fmt, ...);
#define AB_CUDA(call)                                                             \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) {                                                  \
            ab_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return -100 - (int)e_;                                                \
        }                                                                         \
    } while (0)
#define AB_CHECK_LAUNCH() AB_CUDA(cudaGetLastError())
