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

// Branch-free FP64 exp(x) for x <= 0 (the only arguments the kernels produce):
// n = rint(x log2 e), f = x - n ln2 (Cody-Waite), degree-13 Taylor on |f| <= 0.347
// (remainder 4e-18), scaling by 2^n through the exponent field; exp(x) < 1e-307
// flushes to 0.  Relative error ~2e-16, no slow-path call, no divergence.
__device__ __forceinline__ double ab_exp_neg(double x) {
    const double SHIFT = 6755399441055744.0;                 // 1.5 * 2^52
    double t = fma(x, 1.4426950408889634, SHIFT);
    int n = __double2loint(t);
    double nf = t - SHIFT;
    double f = fma(nf, -6.93147180369123816490e-01, x);
    f = fma(nf, -1.90821492927058770002e-10, f);
    double p = 1.6059043836821613e-10;                       // 1/13!
    p = fma(p, f, 2.0876756987868100e-09);                   // 1/12!
    p = fma(p, f, 2.5052108385441720e-08);                   // 1/11!
    p = fma(p, f, 2.7557319223985893e-07);                   // 1/10!
    p = fma(p, f, 2.7557319223985888e-06);                   // 1/9!
    p = fma(p, f, 2.4801587301587302e-05);                   // 1/8!
    p = fma(p, f, 1.9841269841269841e-04);                   // 1/7!
    p = fma(p, f, 1.3888888888888889e-03);                   // 1/6!
    p = fma(p, f, 8.3333333333333332e-03);                   // 1/5!
    p = fma(p, f, 4.1666666666666664e-02);                   // 1/4!
    p = fma(p, f, 1.6666666666666666e-01);                   // 1/3!
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    double r = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
    return (x > -707.0) ? r : 0.0;
}

// Branch-free sqrt(x) for finite x >= 0: MUFU.RSQ64H seed, two Newton steps on
// 1/sqrt, one correction of the root (<= 1 ulp); x < 1e-290 returns 0.
__device__ __forceinline__ double ab_sqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    double s = x * y;
    s = fma(0.5 * y, fma(-s, s, x), s);
    return (x > 1e-290) ? s : 0.0;
}

template <int KIND>
__device__ __forceinline__ double ab_radial(double r2) {
    if (KIND == 0) return ab_exp_neg(-0.5 * r2);
    if (KIND == 1) { double r = ab_sqrt_pos(3.0 * r2); return (1.0 + r) * ab_exp_neg(-r); }
    double r = ab_sqrt_pos(5.0 * r2);
    return (1.0 + r + r * r * 0.3333333333333333) * ab_exp_neg(-r);
}

// dk/d(r^2)
template <int KIND>
__device__ __forceinline__ double ab_radial_grad(double r2) {
    if (KIND == 0) return -0.5 * ab_exp_neg(-0.5 * r2);
    if (KIND == 1) { double r = ab_sqrt_pos(3.0 * r2); return -1.5 * ab_exp_neg(-r); }
    double r = ab_sqrt_pos(5.0 * r2);
    return -5.0 * (1.0 + r) * ab_exp_neg(-r) * 0.16666666666666666;
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

// Grid-wide barrier for kernels launched with cudaLaunchCooperativeKernel (all
// CTAs co-resident).  `counter` starts at 0 and only grows; `target` is the
// caller's running total.
__device__ __forceinline__ void ab_grid_barrier(unsigned long long* counter, unsigned long long& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(counter, 1ULL);
        while (*((volatile unsigned long long*)counter) < target) { }
        __threadfence();
    }
    __syncthreads();
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
