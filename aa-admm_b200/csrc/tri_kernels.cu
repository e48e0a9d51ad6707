// Triangle (cloth) terms of the hard_zxu loop: TriEnergyTerm (admm_anderson_hard_zxu/src/TriEnergyTerm.cpp:29-105)
// inside EnergyTerm::update_z / update_u (src/EnergyTerm.hpp:156-207), one thread per triangle.
//
// A triangle term has 6 rows: z_i = vec(F), F = [x1-x0, x2-x0] * rest_pose (3x2, column-major), scaled by the
// weight sqrt(K area). Layout as for the tets: idx int4[N] (x,y,z used), rest_pose 4 planes x N, u / z 6 planes x N
// (they follow the 9 T tet planes inside the same buffers, so the Anderson vector stays one contiguous range),
// contributions 9 doubles per triangle behind the tets' 12 T.
// These kernels run BEFORE the tet kernel of the same phase and leave their residual share in the control
// block (tri_prim2 / tri_comb); the tet kernel's finishing CTA adds it and takes the decisions.
#include <algorithm>

#include "collision_prox.cuh"
#include "tet_kernels.cuh"
#include "tri_prox.cuh"

namespace aaadmm {

// F(:,c) = d1 * R(0,c) + d2 * R(1,c), R column-major r[c*2+k]
__device__ __forceinline__ void tri_gradient(const double *__restrict__ pos, const int4 id, const double (&r)[4],
                                             double (&F)[6]) {
    double x0[3], d1[3], d2[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) x0[j] = pos[3 * (size_t)id.x + j];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        d1[j] = pos[3 * (size_t)id.y + j] - x0[j];
        d2[j] = pos[3 * (size_t)id.z + j] - x0[j];
    }
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 3; ++j) F[c * 3 + j] = d1[j] * r[c * 2 + 0] + d2[j] * r[c * 2 + 1];
}
__device__ __forceinline__ void tri_gradient_diff(const double *__restrict__ pa, const double *__restrict__ pb,
                                                  const int4 id, const double (&r)[4], double (&F)[6]) {
    double x0[3], d1[3], d2[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) x0[j] = pa[3 * (size_t)id.x + j] - pb[3 * (size_t)id.x + j];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        d1[j] = (pa[3 * (size_t)id.y + j] - pb[3 * (size_t)id.y + j]) - x0[j];
        d2[j] = (pa[3 * (size_t)id.z + j] - pb[3 * (size_t)id.z + j]) - x0[j];
    }
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 3; ++j) F[c * 3 + j] = d1[j] * r[c * 2 + 0] + d2[j] * r[c * 2 + 1];
}

// z = prox((D x - c + u)/w), |D x - W z - c|^2 and the corner contributions of rho dt^2 D^T W (W z - u).
template <int MODE>
__global__ void __launch_bounds__(TET_BLOCK)
k_tri_update_z_hard(TriArrays A, const double *__restrict__ pos, const double *__restrict__ u, double *__restrict__ z,
                    double *__restrict__ contrib, SolveState *st, double *partials) {
    if (st->done) return;
    if (MODE == MODE_REDO && !st->reject) return;
    const int N = A.n_tris;
    double acc[1] = {0.0};
    for (int t = blockIdx.x * TET_BLOCK + threadIdx.x; t < N; t += gridDim.x * TET_BLOCK) {
        const int4 id = A.idx[t];
        double r[4], F[6], zi[6], ui[6], zo[6];
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = A.rest_pose[(size_t)k * N + t];
        const double w = A.w[t], winv = 1.0 / w;
        tri_gradient(pos, id, r, F);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            ui[k] = u[(size_t)k * N + t];
            F[k] = w * F[k];
            zi[k] = (F[k] + ui[k]) * winv;
        }
        tri_prox_block(1, zi, A.limit_min[t], A.limit_max[t], zo);
        double y[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            z[(size_t)k * N + t] = zo[k];
            const double wz = w * zo[k];
            const double d = F[k] - wz;
            acc[0] += d * d;
            y[k] = wz - ui[k];
        }
        // corner a: sum_c Dc(a,c) y(:,c), Dc(1,c) = R(0,c), Dc(2,c) = R(1,c), Dc(0,c) = -(Dc(1,c) + Dc(2,c))
        double q[9];
#pragma unroll
        for (int a = 1; a < 3; ++a)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                q[a * 3 + j] = (A.rho_dt2 * (w * r[0 * 2 + a - 1])) * y[0 * 3 + j] +
                               (A.rho_dt2 * (w * r[1 * 2 + a - 1])) * y[1 * 3 + j];
#pragma unroll
        for (int j = 0; j < 3; ++j) q[j] = -(q[3 + j] + q[6 + j]);
        double *qo = contrib + (size_t)t * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) qo[k] = q[k];
    }
    double out[1];
    if (grid_reduce<1, TET_BLOCK>(acc, partials, &st->ticket, out)) {
        if (threadIdx.x == 0) st->tri_prim2 = out[0];
    }
}

// u += D x - W z - c; residual shares |D x - W z - c|^2 and |D (x - x_last)|^2.
template <int MODE>
__global__ void __launch_bounds__(TET_BLOCK)
k_tri_update_u_hard(TriArrays A, const double *__restrict__ pos_new, const double *__restrict__ pos_last,
                    const double *__restrict__ z, const double *__restrict__ u_in, double *__restrict__ u_out,
                    SolveState *st, double *partials) {
    if (st->done) return;
    const int N = A.n_tris;
    double acc[2] = {0.0, 0.0};
    for (int t = blockIdx.x * TET_BLOCK + threadIdx.x; t < N; t += gridDim.x * TET_BLOCK) {
        const int4 id = A.idx[t];
        double r[4], F[6], dFm[6];
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = A.rest_pose[(size_t)k * N + t];
        const double w = A.w[t];
        tri_gradient(pos_new, id, r, F);
        if (MODE == MODE_ITER) tri_gradient_diff(pos_new, pos_last, id, r, dFm);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double d = w * F[k] - w * z[(size_t)k * N + t];
            u_out[(size_t)k * N + t] = u_in[(size_t)k * N + t] + d;
            if (MODE == MODE_ITER) {
                acc[0] += d * d;
                const double e = w * dFm[k];
                acc[1] += e * e;
            }
        }
    }
    if (MODE != MODE_ITER) return;
    double out[2];
    if (grid_reduce<2, TET_BLOCK>(acc, partials, &st->ticket, out)) {
        if (threadIdx.x == 0) st->tri_comb = out[0] + out[1];
    }
}

// bconst += rho dt^2 D^T C_fix of the triangles that touch pinned vertices (per frame, after k_bconst).
__global__ void k_tri_bconst(TriArrays A, int slot0, const int64_t *__restrict__ inc_ptr, const int *__restrict__ inc,
                             const double *__restrict__ pos, double *__restrict__ bconst) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= A.n_free) return;
    const int N = A.n_tris;
    double s[3] = {0.0, 0.0, 0.0};
    const int64_t p1 = inc_ptr[v + 1];
    for (int64_t p = inc_ptr[v]; p < p1; ++p) {
        const int e = inc[p] - slot0;
        if (e < 0 || e >= 3 * N) continue;  // tet slot / collision-term slot
        const int t = e / 3, c = e - 3 * t;
        const int4 id = A.idx[t];
        const int ids[3] = {id.x, id.y, id.z};
        if (ids[0] < A.n_free && ids[1] < A.n_free && ids[2] < A.n_free) continue;
        double r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = A.rest_pose[(size_t)k * N + t];
        const double w = A.w[t];
        // Dc(a, col): a = 0 -> -(r[col*2] + r[col*2+1]), a >= 1 -> r[col*2 + a-1]
        double cf[6];  // C_fix block, cf[col*3+j] = -w sum_{a pinned} Dc(a,col) xpin_a[j]
#pragma unroll
        for (int col = 0; col < 2; ++col)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                double a = 0.0;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (ids[k] >= A.n_free) {
                        const double g = (k == 0) ? -(r[col * 2 + 0] + r[col * 2 + 1]) : r[col * 2 + k - 1];
                        a += (w * g) * pos[3 * (size_t)ids[k] + j];
                    }
                }
                cf[col * 3 + j] = -a;
            }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double a = 0.0;
#pragma unroll
            for (int col = 0; col < 2; ++col) {
                const double g = (c == 0) ? -(r[col * 2 + 0] + r[col * 2 + 1]) : r[col * 2 + c - 1];
                a += (A.rho_dt2 * (w * g)) * cf[col * 3 + j];
            }
            s[j] += a;
        }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) bconst[3 * (size_t)v + j] += s[j];
}

// ---------------------------------------------------------------------------------------------------------
// Collision terms (hard/src/CollisionEnergyTerm.hpp:40-91, created per vertex by Solver::initialize :386-392
// from set_collisions): 3 rows, D_i x = w x_idx, prox = projection onto the nearest penetrated passive object.
// u / z: 3 planes x P behind the triangles' planes; contributions 3 doubles per term.
// ---------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(TET_BLOCK)
k_pt_update_z_hard(PointArrays A, const double *__restrict__ pos, const double *__restrict__ u, double *__restrict__ z,
                   double *__restrict__ contrib, SolveState *st, double *partials) {
    if (st->done) return;
    if (MODE == MODE_REDO && !st->reject) return;
    const int P = A.n_pts;
    double acc[1] = {0.0};
    for (int t = blockIdx.x * TET_BLOCK + threadIdx.x; t < P; t += gridDim.x * TET_BLOCK) {
        const int v = A.vert[t];
        const double w = A.w[t], winv = 1.0 / w;
        double F[3], ui[3], zi[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            ui[k] = u[(size_t)k * P + t];
            F[k] = w * pos[3 * (size_t)v + k];
            zi[k] = (F[k] + ui[k]) * winv;
        }
        collision_prox_point(A.n_objs, A.types, A.prm, zi);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            z[(size_t)k * P + t] = zi[k];
            const double wz = w * zi[k];
            const double d = F[k] - wz;
            acc[0] += d * d;
            contrib[3 * (size_t)t + k] = (A.rho_dt2 * w) * (wz - ui[k]);
        }
    }
    double out[1];
    if (grid_reduce<1, TET_BLOCK>(acc, partials, &st->ticket, out)) {
        if (threadIdx.x == 0) st->pt_prim2 = out[0];
    }
}

template <int MODE>
__global__ void __launch_bounds__(TET_BLOCK)
k_pt_update_u_hard(PointArrays A, const double *__restrict__ pos_new, const double *__restrict__ pos_last,
                   const double *__restrict__ z, const double *__restrict__ u_in, double *__restrict__ u_out,
                   SolveState *st, double *partials) {
    if (st->done) return;
    const int P = A.n_pts;
    double acc[2] = {0.0, 0.0};
    for (int t = blockIdx.x * TET_BLOCK + threadIdx.x; t < P; t += gridDim.x * TET_BLOCK) {
        const int v = A.vert[t];
        const double w = A.w[t];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double xn = pos_new[3 * (size_t)v + k];
            const double d = w * xn - w * z[(size_t)k * P + t];
            u_out[(size_t)k * P + t] = u_in[(size_t)k * P + t] + d;
            if (MODE == MODE_ITER) {
                acc[0] += d * d;
                const double e = w * (xn - pos_last[3 * (size_t)v + k]);
                acc[1] += e * e;
            }
        }
    }
    if (MODE != MODE_ITER) return;
    double out[2];
    if (grid_reduce<2, TET_BLOCK>(acc, partials, &st->ticket, out)) {
        if (threadIdx.x == 0) st->pt_comb = out[0] + out[1];
    }
}

static int pt_grid(const PointArrays &A) {
    return std::max(1, std::min((A.n_pts + TET_BLOCK - 1) / TET_BLOCK, stream_grid(8)));
}
void launch_pt_update_z_hard(int mode, cudaStream_t s, const PointArrays &A, const double *pos, const double *u, double *z,
                             double *contrib, SolveState *st, double *partials) {
    const int g = pt_grid(A);
    if (mode == MODE_WARM)
        k_pt_update_z_hard<MODE_WARM><<<g, TET_BLOCK, 0, s>>>(A, pos, u, z, contrib, st, partials);
    else if (mode == MODE_ITER)
        k_pt_update_z_hard<MODE_ITER><<<g, TET_BLOCK, 0, s>>>(A, pos, u, z, contrib, st, partials);
    else
        k_pt_update_z_hard<MODE_REDO><<<g, TET_BLOCK, 0, s>>>(A, pos, u, z, contrib, st, partials);
}
void launch_pt_update_u_hard(int mode, cudaStream_t s, const PointArrays &A, const double *pos_new, const double *pos_last,
                             const double *z, const double *u_in, double *u_out, SolveState *st, double *partials) {
    const int g = pt_grid(A);
    if (mode == MODE_WARM)
        k_pt_update_u_hard<MODE_WARM><<<g, TET_BLOCK, 0, s>>>(A, pos_new, pos_last, z, u_in, u_out, st, partials);
    else
        k_pt_update_u_hard<MODE_ITER><<<g, TET_BLOCK, 0, s>>>(A, pos_new, pos_last, z, u_in, u_out, st, partials);
}

static int tri_grid(const TriArrays &A) {
    return std::max(1, std::min((A.n_tris + TET_BLOCK - 1) / TET_BLOCK, stream_grid(8)));
}

void launch_tri_update_z_hard(int mode, cudaStream_t s, const TriArrays &A, const double *pos, const double *u, double *z,
                              double *contrib, SolveState *st, double *partials) {
    const int g = tri_grid(A);
    if (mode == MODE_WARM)
        k_tri_update_z_hard<MODE_WARM><<<g, TET_BLOCK, 0, s>>>(A, pos, u, z, contrib, st, partials);
    else if (mode == MODE_ITER)
        k_tri_update_z_hard<MODE_ITER><<<g, TET_BLOCK, 0, s>>>(A, pos, u, z, contrib, st, partials);
    else
        k_tri_update_z_hard<MODE_REDO><<<g, TET_BLOCK, 0, s>>>(A, pos, u, z, contrib, st, partials);
}
void launch_tri_update_u_hard(int mode, cudaStream_t s, const TriArrays &A, const double *pos_new, const double *pos_last,
                              const double *z, const double *u_in, double *u_out, SolveState *st, double *partials) {
    const int g = tri_grid(A);
    if (mode == MODE_WARM)
        k_tri_update_u_hard<MODE_WARM><<<g, TET_BLOCK, 0, s>>>(A, pos_new, pos_last, z, u_in, u_out, st, partials);
    else
        k_tri_update_u_hard<MODE_ITER><<<g, TET_BLOCK, 0, s>>>(A, pos_new, pos_last, z, u_in, u_out, st, partials);
}
void launch_tri_bconst(cudaStream_t s, const TriArrays &A, int slot0, const int64_t *inc_ptr, const int *inc,
                       const double *pos, double *bconst) {
    k_tri_bconst<<<(A.n_free + 127) / 128, 128, 0, s>>>(A, slot0, inc_ptr, inc, pos, bconst);
}

}  // namespace aaadmm
