#include "extra_terms.cuh"

#include <cfloat>

#include "collision_prox.cuh"
#include "tri_prox.cuh"

namespace aaadmm {

namespace {

__global__ void __launch_bounds__(128)
k_tri_prox(int variant, double *__restrict__ z, int64_t n, double lmin, double lmax) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    double F[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) F[k] = z[6 * i + k];
    double out[6];
    tri_prox_block(variant, F, lmin, lmax, out);
#pragma unroll
    for (int k = 0; k < 6; ++k) z[6 * i + k] = out[k];
}

__global__ void __launch_bounds__(128)
k_collision_prox(int n_objs, const int *__restrict__ types, const double *__restrict__ prm, double *__restrict__ z, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    double x[3] = {z[3 * i], z[3 * i + 1], z[3 * i + 2]};
    collision_prox_point(n_objs, types, prm, x);
#pragma unroll
    for (int k = 0; k < 3; ++k) z[3 * i + k] = x[k];
}

__global__ void k_spring_prox(double *__restrict__ z, const double *__restrict__ pins, const int *__restrict__ active, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !active[i]) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) z[3 * i + k] = pins[3 * i + k];
}

}  // namespace

void launch_tri_prox(int variant, double *z, int64_t n, double limit_min, double limit_max, cudaStream_t s) {
    if (n > 0) k_tri_prox<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(variant, z, n, limit_min, limit_max);
}
void launch_collision_prox(int n_objs, const int *types, const double *prm, double *z, int64_t n, cudaStream_t s) {
    if (n > 0) k_collision_prox<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(n_objs, types, prm, z, n);
}
void launch_spring_prox(double *z, const double *pins, const int *active, int64_t n, cudaStream_t s) {
    if (n > 0) k_spring_prox<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(z, pins, active, n);
}

}  // namespace aaadmm
