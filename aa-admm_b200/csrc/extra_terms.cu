#include "extra_terms.cuh"

#include <cfloat>

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

struct Payload {
    double dx, p[3];
};

__device__ __forceinline__ double norm3(const double *v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

// PassiveCollision::signed_distance of one object: keeps the smallest signed distance seen so far and
// the surface point that belongs to it (`if (dx > p.dx) return;`).
__device__ void signed_distance(int type, const double *q, const double *x, Payload &pl) {
    const double *c = q;
    const double rad = q[6];
    if (type == PASSIVE_FLOOR) {
        const double dx = x[1] - q[0];
        if (dx > pl.dx) return;
        pl.dx = dx;
        pl.p[0] = x[0];
        pl.p[1] = q[0];
        pl.p[2] = x[2];
    } else if (type == PASSIVE_SLIDE_FLOOR) {
        double nrm[3] = {q[3], q[4], q[5]};
        const double nn = norm3(nrm);  // the constructor normalises
        if (nn > 0.0) {
            nrm[0] /= nn;
            nrm[1] /= nn;
            nrm[2] /= nn;
        }
        const double dx = (x[0] - c[0]) * nrm[0] + (x[1] - c[1]) * nrm[1] + (x[2] - c[2]) * nrm[2];
        if (dx > pl.dx) return;
        pl.dx = dx;
#pragma unroll
        for (int k = 0; k < 3; ++k) pl.p[k] = x[k] - dx * nrm[k];
    } else if (type == PASSIVE_SPHERE) {
        double dir[3] = {x[0] - c[0], x[1] - c[1], x[2] - c[2]};
        const double len = norm3(dir);
        const double dx = len - rad;
        if (dx > pl.dx) return;
        pl.dx = dx;
#pragma unroll
        for (int k = 0; k < 3; ++k) pl.p[k] = c[k] + (len > 0.0 ? dir[k] / len : dir[k]) * rad;
    } else if (type == PASSIVE_PLANE_HALF_SPHERE) {
        const double px = x[0] - c[0], pz = x[2] - c[2];
        const double dc = sqrt(px * px + 0.0 * 0.0 + pz * pz) - rad;
        if (dc > 0.0) {
            const double dx = x[1] - c[1];
            if (dx > pl.dx) return;
            pl.dx = dx;
            pl.p[0] = x[0];
            pl.p[1] = c[1];
            pl.p[2] = x[2];
        } else {
            double dir[3] = {x[0] - c[0], x[1] - c[1], x[2] - c[2]};
            const double len = norm3(dir);
            const double dx = (x[1] - c[1] > 0.0) ? len + rad : rad - len;
            if (dx > pl.dx) return;
            pl.dx = dx;
#pragma unroll
            for (int k = 0; k < 3; ++k) pl.p[k] = c[k] + (len > 0.0 ? dir[k] / len : dir[k]) * rad;
        }
    } else {  // PASSIVE_CYLINDER: axis along z through `center`
        double dir[3] = {x[0] - c[0], x[1] - c[1], 0.0 - c[2]};
        const double len = norm3(dir);
        const double dx = len - rad;
        if (dx > pl.dx) return;
        pl.dx = dx;
#pragma unroll
        for (int k = 0; k < 3; ++k) pl.p[k] = c[k] + (len > 0.0 ? dir[k] / len : dir[k]) * rad + (k == 2 ? x[2] : 0.0);
    }
}

__global__ void __launch_bounds__(128)
k_collision_prox(int n_objs, const int *__restrict__ types, const double *__restrict__ prm, double *__restrict__ z, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    const double x[3] = {z[3 * i], z[3 * i + 1], z[3 * i + 2]};
    Payload pl;
    pl.dx = DBL_MAX;
    pl.p[0] = pl.p[1] = pl.p[2] = 0.0;
    for (int j = 0; j < n_objs; ++j) signed_distance(types[j], prm + 7 * j, x, pl);
    if (pl.dx < 0.0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) z[3 * i + k] = pl.p[k];
    }
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
