// K5 sampler kernels of kernel family 0 (ExpSquared): see ensemble_kernel.cuh
#include "ensemble_kernel.cuh"

int ab_ens_launch_k0(ab_gp* h, EnsArgs& A, int n_half, int p, int small_cta, int ws) {
    return launch_ens_kind<0>(h, A, n_half, p, small_cta, ws);
}
