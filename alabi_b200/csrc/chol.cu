// K2: the multi-launch blocked right-looking FP64 Cholesky with DMMA trailing updates
// (factor schedules 0 / 1, kept as a cross-check of the dataflow kernel in
// chol_dataflow.cu, which is the default), plus the triangular inverse and K^-1 that
// the log-likelihood gradient and the predictive variance need, and small solve /
// reduction helpers.  The block triangular solves live in trsv_dataflow.cu.
//
// Replaces george BasicSolver.compute / apply_inverse / get_inverse (scipy
// cholesky + cho_solve) reached from alabi/gp_utils.py:243, alabi/core.py:1158,
// 1248, 1261, 1430.  Matrix layout: row-major, lower triangle, padded to a
// multiple of AB_NB = 128 with an identity block so every tile is full.
//
//   step k:  potf2_inv   one CTA factors the 128x128 diagonal block in shared
//                        memory and inverts it (D_k^-1)
//            trsm_panel  L_ik = A_ik D_k^-T      (DMMA GEMM, K = 128)
//            syrk        A_ij -= L_ik L_jk^T     (DMMA GEMM, K = 128)
//   With look-ahead the panel of step k+1 runs on a high-priority stream while
//   the rest of trailing update k is still in flight.
#include <vector>
#include "handle.h"
#include "dmma_gemm.cuh"
#include "potf2.cuh"

namespace {

constexpr int NB = AB_NB;

using abp::potf2_inv_kernel;

// ---------------------------------------------------------------------------
// panel solve and trailing update (DMMA GEMM core)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(abg::THREADS, 1)
trsm_panel_kernel(double* __restrict__ A, int64_t ld, int64_t o, const double* __restrict__ Dinv) {
    extern __shared__ __align__(16) double smem[];
    abg::Acc acc;
    acc.zero();
    double* Cp = A + (o + (int64_t)NB * (1 + blockIdx.x)) * ld + o;
    abg::mainloop<true, true>(acc, Cp, ld, Dinv, NB, NB / abg::BK, smem);
    abg::store_tile(acc, Cp, ld, 1.0, 0.0);
}

// mode 0: lower-triangular tiles of the trailing matrix starting at r0
// mode 1: tile (blockIdx.x, 0)  — first block column
// mode 2: first TWO block columns of an nt-row trailing matrix:
//         blockIdx.x < nt -> (blockIdx.x, 0), else (blockIdx.x - nt + 1, 1)
__global__ void __launch_bounds__(abg::THREADS, 1)
syrk_kernel(double* __restrict__ A, int64_t ld, int64_t o, int64_t r0, int kdim, int mode, int nt) {
    extern __shared__ __align__(16) double smem[];
    int ti, tj;
    if (mode == 0) abg::tri_decode(blockIdx.x, ti, tj);
    else if (mode == 1 || (int)blockIdx.x < nt) { ti = blockIdx.x; tj = 0; }
    else { ti = blockIdx.x - nt + 1; tj = 1; }
    abg::Acc acc;
    acc.zero();
    const double* Ap = A + (r0 + (int64_t)ti * NB) * ld + o;
    const double* Bp = A + (r0 + (int64_t)tj * NB) * ld + o;
    abg::mainloop<true, true>(acc, Ap, ld, Bp, ld, kdim / abg::BK, smem);
    abg::store_tile(acc, A + (r0 + (int64_t)ti * NB) * ld + r0 + (int64_t)tj * NB, ld, -1.0, 1.0);
}

// Small-tile variants for the panel chain (the critical path of the factorisation):
// a panel has only T-k-1 128-tiles, far fewer than 148 SMs for N <= 16384, and a
// 128 x 128 x 128 tile alone costs >= 17 us on one SM.  64-row tiles spread the same
// work over 2-4x more SMs.
using Half = abg::Core<16, 3, 1, 1, 4>;     // 64 x 128 tile, 128 threads, 90 KB (2 CTAs / SM)
using Small = abg::Core<16, 3, 1, 1, 2>;    // 64 x 64 tile, 64 threads, 60 KB (3 CTAs / SM)

// in place: the CTA reads and writes its own 64 rows x 128 columns only
__global__ void __launch_bounds__(Half::THREADS, 2)
trsm_half_kernel(double* __restrict__ A, int64_t ld, int64_t o, const double* __restrict__ Dinv) {
    extern __shared__ __align__(16) double smem[];
    abg::Acc acc;
    acc.zero();
    double* Cp = A + (o + (int64_t)NB * (1 + (blockIdx.x >> 1)) + 64 * (blockIdx.x & 1)) * ld + o;
    Half::mainloop<true, true>(acc, Cp, ld, Dinv, NB, NB / Half::BK, smem);
    Half::store_tile(acc, Cp, ld, 1.0, 0.0);
}

// modes 1 and 2 of syrk_kernel on 64 x 64 sub-tiles (4 CTAs per 128-tile; the
// strictly-upper sub-tile of a diagonal tile is skipped)
__global__ void __launch_bounds__(Small::THREADS, 3)
syrk_small_kernel(double* __restrict__ A, int64_t ld, int64_t o, int64_t r0, int kdim, int mode, int nt) {
    extern __shared__ __align__(16) double smem[];
    const int p = blockIdx.x >> 2, sm = (blockIdx.x >> 1) & 1, sn = blockIdx.x & 1;
    int ti, tj;
    if (mode == 1 || p < nt) { ti = p; tj = 0; }
    else { ti = p - nt + 1; tj = 1; }
    if (ti == tj && sn > sm) return;
    abg::Acc acc;
    acc.zero();
    const int64_t row = r0 + (int64_t)ti * NB + 64 * sm, col = r0 + (int64_t)tj * NB + 64 * sn;
    Small::mainloop<true, true>(acc, A + row * ld + o, ld, A + col * ld + o, ld, kdim / Small::BK, smem);
    Small::store_tile(acc, A + row * ld + col, ld, -1.0, 1.0);
}

// ---------------------------------------------------------------------------
// blocked triangular solves for z = L^-1 r and alpha = L^-T z
// ---------------------------------------------------------------------------
// forward step k: every CTA recomputes z_k = D_k^-1 r_k; CTA 0 stores it, CTA b >= 1
// applies r_{k+b} -= L_{k+b,k} z_k.
__global__ void __launch_bounds__(256)
trsv_fwd_kernel(const double* __restrict__ L, int64_t ld, const double* __restrict__ Dinv, int k,
                double* __restrict__ r, double* __restrict__ z) {
    __shared__ double sr[NB], sz[NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t o = (int64_t)k * NB;
    if (tid < NB) sr[tid] = r[o + tid];
    __syncthreads();
    const double* D = Dinv + (int64_t)k * NB * NB;
    for (int row = warp; row < NB; row += 8) {
        double s = 0.0;
#pragma unroll
        for (int c = lane; c < NB; c += 32) s = fma(D[row * NB + c], sr[c], s);
        s = ab_warp_sum(s);
        if (lane == 0) sz[row] = s;
    }
    __syncthreads();
    if (blockIdx.x == 0) {
        if (tid < NB) z[o + tid] = sz[tid];
        return;
    }
    const int64_t rb = o + (int64_t)blockIdx.x * NB;
    for (int row = warp; row < NB; row += 8) {
        const double* Lr = L + (rb + row) * ld + o;
        double s = 0.0;
#pragma unroll
        for (int c = lane; c < NB; c += 32) s = fma(Lr[c], sz[c], s);
        s = ab_warp_sum(s);
        if (lane == 0) r[rb + row] -= s;
    }
}

// backward step k: alpha_k = D_k^-T z_k (CTA 0 stores), CTA b >= 1 applies
// z_{b-1} -= L_{k,b-1}^T alpha_k.
__global__ void __launch_bounds__(NB)
trsv_bwd_kernel(const double* __restrict__ L, int64_t ld, const double* __restrict__ Dinv, int k,
                double* __restrict__ z, double* __restrict__ alpha) {
    __shared__ double sz[NB], sa[NB];
    const int tid = threadIdx.x;
    const int64_t o = (int64_t)k * NB;
    sz[tid] = z[o + tid];
    __syncthreads();
    const double* D = Dinv + (int64_t)k * NB * NB;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
    for (int c = 0; c < NB; c += 2) {
        s0 = fma(D[c * NB + tid], sz[c], s0);
        s1 = fma(D[(c + 1) * NB + tid], sz[c + 1], s1);
    }
    sa[tid] = s0 + s1;
    __syncthreads();
    if (blockIdx.x == 0) {
        alpha[o + tid] = sa[tid];
        return;
    }
    const int64_t cb = (int64_t)(blockIdx.x - 1) * NB;
    const double* Lp = L + o * ld + cb + tid;
    s0 = s1 = 0.0;
#pragma unroll 4
    for (int c = 0; c < NB; c += 2) {
        s0 = fma(Lp[(int64_t)c * ld], sa[c], s0);
        s1 = fma(Lp[(int64_t)(c + 1) * ld], sa[c + 1], s1);
    }
    z[cb + tid] -= s0 + s1;
}

__global__ void residual_kernel(const double* __restrict__ y, int64_t n, int64_t npad, double mean,
                                double* __restrict__ r) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npad) r[i] = (i < n) ? y[i] - mean : 0.0;
}

// single CTA, fixed summation order: out[0] = sum v^2, out[1] = sum parts
__global__ void __launch_bounds__(1024)
final_sums_kernel(const double* __restrict__ v, int64_t n, const double* __restrict__ parts, int nparts,
                  double* __restrict__ out) {
    __shared__ double sh[32];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) s = fma(v[i], v[i], s);
    s = ab_warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double q = 0.0;
        for (int w = 0; w < 32; w++) q += sh[w];
        double ld = 0.0;
        for (int p = 0; p < nparts; p++) ld += parts[p];
        out[0] = q;
        out[1] = ld;
    }
}

// ---------------------------------------------------------------------------
// triangular inverse by recursive halving and K^-1 = L^-T L^-1
// ---------------------------------------------------------------------------
struct InvNode { int r0, h, n2, tile_begin; };          // block units
struct InvLevel { int nnodes; InvNode nodes[64]; };

__global__ void copy_diag_blocks_kernel(const double* __restrict__ Dinv, double* __restrict__ Linv, int64_t ld) {
    const int64_t o = (int64_t)blockIdx.x * NB;
    const double* D = Dinv + (int64_t)blockIdx.x * NB * NB;
    for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) {
        int i = idx >> 7, j = idx & (NB - 1);
        Linv[(o + i) * ld + o + j] = D[idx];
    }
}

// PASS 1: W21 = L21 * Linv11 ; PASS 2: Linv21 = -Linv22 * W21
template <int PASS>
__global__ void __launch_bounds__(abg::THREADS, 1)
trinv_kernel(const double* __restrict__ L, double* __restrict__ Linv, double* __restrict__ W, int64_t ld,
             const __grid_constant__ InvLevel lv) {
    extern __shared__ __align__(16) double smem[];
    int nd = 0;
    for (int q = 1; q < lv.nnodes; q++)
        if ((int)blockIdx.x >= lv.nodes[q].tile_begin) nd = q;
    const InvNode node = lv.nodes[nd];
    const int p = blockIdx.x - node.tile_begin;
    const int i = p / node.h, j = p % node.h;
    const int64_t rowi = (int64_t)(node.r0 + node.h + i) * NB, colj = (int64_t)(node.r0 + j) * NB;
    abg::Acc acc;
    acc.zero();
    if (PASS == 1) {
        const double* Ap = L + rowi * ld + colj;                 // k starts at block j
        const double* Bp = Linv + colj * ld + colj;
        abg::mainloop<true, false>(acc, Ap, ld, Bp, ld, (node.h - j) * (NB / abg::BK), smem);
        abg::store_tile(acc, W + rowi * ld + colj, ld, 1.0, 0.0);
    } else {
        const int64_t k0 = (int64_t)(node.r0 + node.h) * NB;
        const double* Ap = Linv + rowi * ld + k0;
        const double* Bp = W + k0 * ld + colj;
        abg::mainloop<true, false>(acc, Ap, ld, Bp, ld, (i + 1) * (NB / abg::BK), smem);
        abg::store_tile(acc, Linv + rowi * ld + colj, ld, -1.0, 0.0);
    }
}

// Kinv(i, j) = sum_{k >= i} Linv(k, i)^T Linv(k, j), j <= i
__global__ void __launch_bounds__(abg::THREADS, 1)
kinv_kernel(const double* __restrict__ Linv, double* __restrict__ Kinv, int64_t ld, int T) {
    extern __shared__ __align__(16) double smem[];
    int i, j;
    abg::tri_decode(blockIdx.x, i, j);
    abg::Acc acc;
    acc.zero();
    const double* Ap = Linv + (int64_t)i * NB * ld + (int64_t)i * NB;
    const double* Bp = Linv + (int64_t)i * NB * ld + (int64_t)j * NB;
    abg::mainloop<false, false>(acc, Ap, ld, Bp, ld, (T - i) * (NB / abg::BK), smem);
    abg::store_tile(acc, Kinv + (int64_t)i * NB * ld + (int64_t)j * NB, ld, 1.0, 0.0);
}

__global__ void mirror_lower_kernel(double* __restrict__ A, int64_t ld, int64_t n) {
    __shared__ double t[32][33];
    int bi = blockIdx.y, bj = blockIdx.x;
    if (bj >= bi) return;                       // strictly-lower tiles -> upper
    int x = threadIdx.x, y = threadIdx.y;
    for (int yy = y; yy < 32; yy += 8) {
        int64_t r = (int64_t)bi * 32 + yy, c = (int64_t)bj * 32 + x;
        t[yy][x] = (r < n && c < n) ? A[r * ld + c] : 0.0;
    }
    __syncthreads();
    for (int yy = y; yy < 32; yy += 8) {
        int64_t r = (int64_t)bj * 32 + yy, c = (int64_t)bi * 32 + x;
        if (r < n && c < n) A[r * ld + c] = t[x][yy];
    }
}
__global__ void mirror_diag_tiles_kernel(double* __restrict__ A, int64_t ld, int64_t n) {
    int64_t o = (int64_t)blockIdx.x * 32;
    for (int idx = threadIdx.x; idx < 32 * 32; idx += blockDim.x) {
        int i = idx >> 5, j = idx & 31;
        if (j > i && o + j < n) A[(o + i) * ld + o + j] = A[(o + j) * ld + o + i];
    }
}

template <typename F>
int set_smem(F f, int bytes) {
    AB_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return 0;
}

// cudaFuncSetAttribute is per device: remember which devices are configured
int configure_once() {
    static unsigned long long done_mask = 0;
    int dev = 0;
    AB_CUDA(cudaGetDevice(&dev));
    const bool done = dev < 64 && ((done_mask >> dev) & 1ULL);
    if (done) return 0;
    int rc = 0;
    rc |= set_smem(trsm_panel_kernel, abg::SMEM_BYTES);
    rc |= set_smem(syrk_kernel, abg::SMEM_BYTES);
    rc |= set_smem(trsm_half_kernel, Half::SMEM_BYTES);
    rc |= set_smem(syrk_small_kernel, Small::SMEM_BYTES);
    rc |= set_smem(trinv_kernel<1>, abg::SMEM_BYTES);
    rc |= set_smem(trinv_kernel<2>, abg::SMEM_BYTES);
    rc |= set_smem(kinv_kernel, abg::SMEM_BYTES);
    if (rc) return rc;
    if (dev < 64) done_mask |= 1ULL << dev;
    return 0;
}

// returns the height of the subtree rooted at blocks [r0, r0 + nblk)
int inv_tree(int r0, int nblk, std::vector<std::vector<InvNode>>& levels) {
    if (nblk <= 1) return 0;
    int h = nblk / 2, n2 = nblk - h;
    int ha = inv_tree(r0, h, levels), hb = inv_tree(r0 + h, n2, levels);
    int ht = 1 + (ha > hb ? ha : hb);
    if ((int)levels.size() < ht + 1) levels.resize(ht + 1);
    levels[ht].push_back(InvNode{r0, h, n2, 0});
    return ht;
}
}  // namespace

// Right-looking factorisation with 128-wide panels whose trailing updates are
// applied two panels at a time (K = 256 DMMA GEMMs: half the C traffic and half
// the per-tile prologue/epilogue of a K = 128 sweep).  Pair (k, k+1):
//   potf2(k) trsm(k) | update column k+1 with panel k | potf2(k+1) trsm(k+1)
//   trailing(rows, cols >= k+2) -= [L_k L_k+1] [L_k L_k+1]^T
// Look-ahead: the first two block columns of the trailing update go first, then
// the next pair's panel chain runs on the high-priority stream while the rest
// of the trailing update is still in flight.
int ab_launch_factor(ab_gp* h) {
    int rc = configure_once();
    if (rc) return rc;
    const int T = (int)(h->npad / NB);
    const int64_t ld = h->npad;
    cudaStream_t ms = h->stream;
    AB_CUDA(cudaMemsetAsync(h->d_info, 0, sizeof(int), ms));
    ab_prof_begin(h, AB_PROF_FACTOR);
    ab_count_launches(3LL * T);
    const bool la = h->lookahead != 0 && T >= 4;
    cudaStream_t ps = la ? h->panel_stream : ms;
    const int G = abg::THREADS, SM = abg::SMEM_BYTES;
    auto Dk = [&](int k) { return h->Dinv + (int64_t)k * NB * NB; };
    // panel chain of the pair starting at block k (its columns are fully updated)
    // panel operations: 64-row tiles while a panel has fewer 128-tiles than SMs
    auto trsm = [&](int nblk, int64_t o, const double* D) {
        if (nblk <= 0) return;
        if (nblk < h->nsm) trsm_half_kernel<<<2 * nblk, Half::THREADS, Half::SMEM_BYTES, ps>>>(h->L, ld, o, D);
        else trsm_panel_kernel<<<nblk, G, SM, ps>>>(h->L, ld, o, D);
    };
    auto syrk_cols = [&](cudaStream_t st, int ntiles, int64_t o, int64_t r0, int kdim, int mode, int nt) {
        if (ntiles < h->nsm) syrk_small_kernel<<<4 * ntiles, Small::THREADS, Small::SMEM_BYTES, st>>>(h->L, ld, o, r0, kdim, mode, nt);
        else syrk_kernel<<<ntiles, G, SM, st>>>(h->L, ld, o, r0, kdim, mode, nt);
    };
    auto pair_panels = [&](int k) {
        const int64_t o = (int64_t)k * NB;
        potf2_inv_kernel<true><<<1, 256, 0, ps>>>(h->L, ld, o, Dk(k), h->logdet_parts + k, h->d_info);
        trsm(T - k - 1, o, Dk(k));
        if (k + 1 < T) {
            syrk_cols(ps, T - k - 1, o, o + NB, NB, 1, 0);                               // column k+1 -= L_k L_k^T
            potf2_inv_kernel<true><<<1, 256, 0, ps>>>(h->L, ld, o + NB, Dk(k + 1), h->logdet_parts + k + 1, h->d_info);
            trsm(T - k - 2, o + NB, Dk(k + 1));
        }
    };
    if (la) {
        AB_CUDA(cudaEventRecord(h->ev_fork, ms));
        AB_CUDA(cudaStreamWaitEvent(ps, h->ev_fork, 0));
    }
    pair_panels(0);
    if (la) AB_CUDA(cudaEventRecord(h->ev_panel, ps));
    for (int k = 0; k < T; k += 2) {
        const int64_t o = (int64_t)k * NB;
        if (la) AB_CUDA(cudaStreamWaitEvent(ms, h->ev_panel, 0));
        const int nt = T - k - 2;                    // block rows of the trailing matrix
        if (nt <= 0) break;
        const int kdim = 2 * NB;
        if (!la) {
            syrk_kernel<<<nt * (nt + 1) / 2, G, SM, ms>>>(h->L, ld, o, o + 2 * NB, kdim, 0, nt);
            pair_panels(k + 2);
            continue;
        }
        // (a) first two block columns of the trailing update, (b) next pair's panels, (c) the rest
        const int nfirst = nt >= 2 ? 2 * nt - 1 : nt;
        syrk_cols(ms, nfirst, o, o + 2 * NB, kdim, 2, nt);
        AB_CUDA(cudaEventRecord(h->ev_col, ms));
        AB_CUDA(cudaStreamWaitEvent(ps, h->ev_col, 0));
        pair_panels(k + 2);
        AB_CUDA(cudaEventRecord(h->ev_panel, ps));
        if (nt > 2)
            syrk_kernel<<<(nt - 2) * (nt - 1) / 2, G, SM, ms>>>(h->L, ld, o, o + 4 * NB, kdim, 0, nt - 2);
    }
    ab_prof_end(h, AB_PROF_FACTOR);
    AB_CHECK_LAUNCH();
    return 0;
}

int ab_launch_rebuild_dinv(ab_gp* h) {
    int rc = configure_once();
    if (rc) return rc;
    const int T = (int)(h->npad / NB);
    potf2_inv_kernel<false><<<T, 256, 0, h->stream>>>(h->L, h->npad, 0, h->Dinv, h->logdet_parts, h->d_info);
    AB_CHECK_LAUNCH();
    return 0;
}

// inverse and log-determinant part of ONE diagonal block of an existing factor
int ab_launch_rebuild_dinv_block(ab_gp* h, int kb) {
    int rc = configure_once();
    if (rc) return rc;
    const int64_t o = (int64_t)kb * NB;
    potf2_inv_kernel<false><<<1, 256, 0, h->stream>>>(h->L + o * h->npad + o, h->npad, 0, h->Dinv + o * NB,
                                                      h->logdet_parts + kb, h->d_info);
    AB_CHECK_LAUNCH();
    return 0;
}

// one-block factor: z = D_0^-1 r (r is not modified)
int ab_launch_trsv_single(ab_gp* h, const double* r) {
    AB_CUDA(cudaMemcpyAsync(h->alpha, r, NB * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));   // alpha: scratch here
    trsv_fwd_kernel<<<1, 256, 0, h->stream>>>(h->L, h->npad, h->Dinv, 0, h->alpha, h->z);
    AB_CHECK_LAUNCH();
    return 0;
}

// z = L^-1 (y - mean), alpha = L^-T z, quad = |z|^2, logdet = sum parts.
// Leaves {quad, logdet} in h->h_pinned[0..1] after a stream sync by the caller.
int ab_launch_solve_alpha(ab_gp* h, const double* y) {
    const int T = (int)(h->npad / NB);
    const int64_t ld = h->npad;
    cudaStream_t s = h->stream;
    residual_kernel<<<(unsigned)((h->npad + 255) / 256), 256, 0, s>>>(y, h->n, h->npad, h->mean, h->work);
    if (T >= 2) {
        int rc = ab_launch_trsv_dataflow(h, h->work, 1);          // both sweeps, one cooperative launch
        if (rc) return rc;
    } else {
        for (int k = 0; k < T; k++) trsv_fwd_kernel<<<T - k, 256, 0, s>>>(h->L, ld, h->Dinv, k, h->work, h->z);
        AB_CUDA(cudaMemcpyAsync(h->work, h->z, h->npad * sizeof(double), cudaMemcpyDeviceToDevice, s));
        for (int k = T - 1; k >= 0; k--) trsv_bwd_kernel<<<k + 1, NB, 0, s>>>(h->L, ld, h->Dinv, k, h->work, h->alpha);
    }
    final_sums_kernel<<<1, 1024, 0, s>>>(h->z, h->npad, h->logdet_parts, T, h->scratch);
    AB_CHECK_LAUNCH();
    AB_CUDA(cudaMemcpyAsync(h->h_pinned, h->scratch, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    return 0;
}

int ab_launch_build_linv(ab_gp* h) {
    int rc = configure_once();
    if (rc) return rc;
    const int T = (int)(h->npad / NB);
    const int64_t ld = h->npad;
    cudaStream_t s = h->stream;
    copy_diag_blocks_kernel<<<T, 256, 0, s>>>(h->Dinv, h->Linv, ld);
    std::vector<std::vector<InvNode>> levels;
    inv_tree(0, T, levels);
    for (size_t ht = 1; ht < levels.size(); ht++) {
        auto& nodes = levels[ht];
        for (size_t b = 0; b < nodes.size(); b += 64) {
            InvLevel lv;
            lv.nnodes = (int)((nodes.size() - b < 64) ? nodes.size() - b : 64);
            int tiles = 0;
            for (int q = 0; q < lv.nnodes; q++) {
                lv.nodes[q] = nodes[b + q];
                lv.nodes[q].tile_begin = tiles;
                tiles += lv.nodes[q].h * lv.nodes[q].n2;
            }
            trinv_kernel<1><<<tiles, abg::THREADS, abg::SMEM_BYTES, s>>>(h->L, h->Linv, h->Kinv, ld, lv);
            trinv_kernel<2><<<tiles, abg::THREADS, abg::SMEM_BYTES, s>>>(h->L, h->Linv, h->Kinv, ld, lv);
        }
    }
    AB_CHECK_LAUNCH();
    return 0;
}

int ab_launch_build_kinv(ab_gp* h) {
    int rc = configure_once();
    if (rc) return rc;
    const int T = (int)(h->npad / NB);
    kinv_kernel<<<T * (T + 1) / 2, abg::THREADS, abg::SMEM_BYTES, h->stream>>>(h->Linv, h->Kinv, h->npad, T);
    AB_CHECK_LAUNCH();
    return 0;
}

int ab_launch_mirror_lower(ab_gp* h, double* A, int64_t ld) {
    int64_t n = h->npad;
    unsigned nt = (unsigned)((n + 31) / 32);
    mirror_lower_kernel<<<dim3(nt, nt), dim3(32, 8), 0, h->stream>>>(A, ld, n);
    mirror_diag_tiles_kernel<<<nt, 256, 0, h->stream>>>(A, ld, n);
    AB_CHECK_LAUNCH();
    return 0;
}
