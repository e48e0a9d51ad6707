// Local step, rhs assembly and residuals of the tet ADMM loop (K1-K4 of SURVEY 2b).
//
// One thread per tet; SoA planes (u, z: 9 planes of n_tets doubles; B^-1: 9 planes) so that every
// warp-wide load/store is a contiguous 256-byte segment; vertex positions (3 doubles per vertex,
// free vertices first, pinned last) are gathered through L2.  D is never materialised:
//   D_i x - c_i = w_i * Ds(x_full) * B_i^-1      (EnergyTerm.hpp:167-178 with C_fix = m_C x_pin)
//   D^T(.)      = per-vertex gather over incident (tet, corner) pairs, fixed order, no atomics.
#include "tet_kernels.cuh"

#include <algorithm>

#include "lbfgs_prox.cuh"
#include "svd3.cuh"

namespace aaadmm {

// F = Ds * Binv, column-major F[r*3+j]; Ds columns are x_{k+1} - x_0.
__device__ __forceinline__ void deformation_gradient(const double *__restrict__ pos, const int4 id,
                                                     const double (&b)[9], double (&F)[9]) {
    double x0[3], d[9];
#pragma unroll
    for (int j = 0; j < 3; ++j) x0[j] = pos[3 * (size_t)id.x + j];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        d[0 * 3 + j] = pos[3 * (size_t)id.y + j] - x0[j];
        d[1 * 3 + j] = pos[3 * (size_t)id.z + j] - x0[j];
        d[2 * 3 + j] = pos[3 * (size_t)id.w + j] - x0[j];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            F[r * 3 + j] = d[0 * 3 + j] * b[r * 3 + 0] + d[1 * 3 + j] * b[r * 3 + 1] + d[2 * 3 + j] * b[r * 3 + 2];
}

// Same for a difference of two position sets (dual residual D (x - x_last)).
__device__ __forceinline__ void deformation_gradient_diff(const double *__restrict__ pa, const double *__restrict__ pb,
                                                          const int4 id, const double (&b)[9], double (&F)[9]) {
    double x0[3], d[9];
#pragma unroll
    for (int j = 0; j < 3; ++j) x0[j] = pa[3 * (size_t)id.x + j] - pb[3 * (size_t)id.x + j];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        d[0 * 3 + j] = (pa[3 * (size_t)id.y + j] - pb[3 * (size_t)id.y + j]) - x0[j];
        d[1 * 3 + j] = (pa[3 * (size_t)id.z + j] - pb[3 * (size_t)id.z + j]) - x0[j];
        d[2 * 3 + j] = (pa[3 * (size_t)id.w + j] - pb[3 * (size_t)id.w + j]) - x0[j];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            F[r * 3 + j] = d[0 * 3 + j] * b[r * 3 + 0] + d[1 * 3 + j] * b[r * 3 + 1] + d[2 * 3 + j] * b[r * 3 + 2];
}

// Corner contributions of D^T W (W z - u) scaled by rho dt^2: q[c*3+j], c = 0..3.
__device__ __forceinline__ void corner_contrib(const double (&b)[9], double w, double rho_dt2, const double (&y)[9],
                                               double (&q)[12]) {
#pragma unroll
    for (int c = 1; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            // G(r,c) = Binv(c-1, r) = b[r*3 + c-1]
            q[c * 3 + j] = (rho_dt2 * (w * b[0 * 3 + c - 1])) * y[0 * 3 + j] +
                           (rho_dt2 * (w * b[1 * 3 + c - 1])) * y[1 * 3 + j] +
                           (rho_dt2 * (w * b[2 * 3 + c - 1])) * y[2 * 3 + j];
        }
#pragma unroll
    for (int j = 0; j < 3; ++j) q[j] = -(q[3 + j] + q[6 + j] + q[9 + j]);
}

// hard_zxu local step (hard/src/Solver.cpp:133-140 + EnergyTerm::update_z + TetEnergyTerm::prox):
//   z = prox((D x - c + u)/w), prim^2 = |D x - W z - c|^2, and the per-corner D^T W(Wz - u)
//   contributions for the following x-update.  The finishing CTA takes the accept/reject
//   decision of hard/src/Solver.cpp:146 on the device.
template <int MODE>
__global__ void __launch_bounds__(TET_BLOCK, 4)
k_update_z_hard(TetArrays A, const double *__restrict__ pos, const double *__restrict__ u, double *__restrict__ z,
                double *__restrict__ contrib, SolveState *st, double *partials) {
    if (st->done) return;
    if (MODE == MODE_REDO && !st->reject) return;
    const int T = A.n_tets;
    double acc[1] = {0.0};
    // B^-1 and u are parked in shared memory while the Jacobi SVD runs: 36 fewer live registers
    __shared__ double s_park[18][TET_BLOCK];
    for (int t = blockIdx.x * TET_BLOCK + threadIdx.x; t < T; t += gridDim.x * TET_BLOCK) {
        if (A.material && A.material[t] != 0) continue;  // handled by k_hyper
        const int4 id = A.idx[t];
        double zi[9], F[9];
        const double w = A.w[t];
        {
            double b[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) b[k] = A.binv[(size_t)k * T + t];
            deformation_gradient(pos, id, b, F);
            const double winv = 1.0 / w;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const double ui = u[(size_t)k * T + t];
                F[k] = w * F[k];              // D_i x - c_i
                zi[k] = (F[k] + ui) * winv;   // W^-1 (D_i x + u_i - c_i)
                s_park[k][threadIdx.x] = b[k];
                s_park[9 + k][threadIdx.x] = ui;
            }
        }
        tet_prox_linear(zi);
        double b[9], y[9], q[12];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            z[(size_t)k * T + t] = zi[k];
            const double wz = w * zi[k];
            const double r = F[k] - wz;
            acc[0] += r * r;
            y[k] = wz - s_park[9 + k][threadIdx.x];
            b[k] = s_park[k][threadIdx.x];
        }
        corner_contrib(b, w, A.rho_dt2, y, q);
        double *qo = contrib + (size_t)t * 12;
#pragma unroll
        for (int k = 0; k < 12; k += 2) *reinterpret_cast<double2 *>(qo + k) = make_double2(q[k], q[k + 1]);
    }
    double out[1];
    if (grid_reduce<1, TET_BLOCK>(acc, partials, &st->ticket, out)) {
        if (threadIdx.x == 0) {
            const double tot = out[0] + st->hyper_prim2 + st->tri_prim2 + st->pt_prim2;
            const double prim = sqrt(tot);
            st->prim2 = tot;
            if (MODE == MODE_ITER) {
                if (st->accel && st->prev_prim < prim) {
                    st->reject = 1;
                } else {
                    st->reject = 0;
                    st->prev_prim = prim;
                }
            } else if (MODE == MODE_REDO) {
                st->prev_prim = prim;
                st->n_rejects += 1;
            }
        }
    }
}

// (u,x) <- default (u,x) and AndersonAcceleration::reset (hard/src/Solver.cpp:151-155).
__global__ void k_restore_if_reject(double *__restrict__ ucur, const double *__restrict__ gdef, int64_t n,
                                    SolveState *st) {
    if (st->done || !st->reject) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) ucur[i] = gdef[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->aa_iter = 0;
        st->aa_col = 0;
    }
}

// b = M x_bar + rho dt^2 D^T (W z + C_fix - u), written in the factor's elimination order.
__global__ void k_rhs_gather(int n_free, const int64_t *__restrict__ inc_ptr, const int *__restrict__ inc,
                             const double *__restrict__ contrib, const double *__restrict__ bconst,
                             const int *__restrict__ iperm, double *__restrict__ W, const SolveState *st, int when,
                             int dof_factor) {
    if (st->done || (when == 1 && !st->reject)) return;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_free) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    const int64_t p1 = inc_ptr[v + 1];
    int64_t p = inc_ptr[v];
    // RHS_BATCH slots at a time: their indices, then all their values, are loaded before the first add (the adds keep
    // the list order, so the sums are the same numbers as one slot at a time)
    constexpr int RHS_BATCH = 8;
    for (; p + RHS_BATCH <= p1; p += RHS_BATCH) {
        int e[RHS_BATCH];
#pragma unroll
        for (int k = 0; k < RHS_BATCH; ++k) e[k] = inc[p + k];  // contribution slot (tets: tet*4 + corner; triangles behind them)
        double q[RHS_BATCH][3];
#pragma unroll
        for (int k = 0; k < RHS_BATCH; ++k) {
            const double *src = contrib + 3 * (size_t)e[k];
            q[k][0] = src[0];
            q[k][1] = src[1];
            q[k][2] = src[2];
        }
#pragma unroll
        for (int k = 0; k < RHS_BATCH; ++k) {
            s0 += q[k][0];
            s1 += q[k][1];
            s2 += q[k][2];
        }
    }
    for (; p < p1; ++p) {
        const int e = inc[p];
        const double *q = contrib + 3 * (size_t)e;
        s0 += q[0];
        s1 += q[1];
        s2 += q[2];
    }
    if (dof_factor) {
        // factor of the full 3 n_free x 3 n_free system (e.g. Eigen's own, LinearSolver.hpp:79-84): one permutation
        // entry per degree of freedom
        W[iperm[3 * (size_t)v + 0]] = bconst[3 * (size_t)v + 0] + s0;
        W[iperm[3 * (size_t)v + 1]] = bconst[3 * (size_t)v + 1] + s1;
        W[iperm[3 * (size_t)v + 2]] = bconst[3 * (size_t)v + 2] + s2;
        return;
    }
    const size_t o = (size_t)iperm[v] * 3;
    W[o + 0] = bconst[3 * (size_t)v + 0] + s0;
    W[o + 1] = bconst[3 * (size_t)v + 1] + s1;
    W[o + 2] = bconst[3 * (size_t)v + 2] + s2;
}

// Per frame: bconst = M x_bar + rho dt^2 D^T C_fix, C_fix = m_C x_pin (hard/src/Solver.cpp:79,83).
__global__ void k_bconst(TetArrays A, const int64_t *__restrict__ inc_ptr, const int *__restrict__ inc,
                         const double *__restrict__ pos, const double *__restrict__ mass,
                         const double *__restrict__ xbar, double *__restrict__ bconst) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= A.n_free) return;
    const int T = A.n_tets;
    double s[3] = {0.0, 0.0, 0.0};
    const int64_t p1 = inc_ptr[v + 1];
    for (int64_t p = inc_ptr[v]; p < p1; ++p) {
        const int e = inc[p];
        if (e >= 4 * T) continue;  // triangle slot: k_tri_bconst
        const int t = e >> 2, c = e & 3;
        const int4 id = A.idx[t];
        const int ids[4] = {id.x, id.y, id.z, id.w};
        if (ids[0] < A.n_free && ids[1] < A.n_free && ids[2] < A.n_free && ids[3] < A.n_free) continue;
        double b[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) b[k] = A.binv[(size_t)k * T + t];
        const double w = A.w[t];
        // G(r,k): k=0 -> -(sum), k>=1 -> b[r*3+k-1]
        double cf[9];  // C_fix block, cf[r*3+j] = -w sum_{k pinned} G(r,k) xpin_k[j]
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                double a = 0.0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (ids[k] >= A.n_free) {
                        const double g = (k == 0) ? -(b[r * 3 + 0] + b[r * 3 + 1] + b[r * 3 + 2]) : b[r * 3 + k - 1];
                        a += (w * g) * pos[3 * (size_t)ids[k] + j];
                    }
                }
                cf[r * 3 + j] = -a;
            }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double a = 0.0;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double g = (c == 0) ? -(b[r * 3 + 0] + b[r * 3 + 1] + b[r * 3 + 2]) : b[r * 3 + c - 1];
                a += (A.rho_dt2 * (w * g)) * cf[r * 3 + j];
            }
            s[j] += a;
        }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) bconst[3 * (size_t)v + j] = mass[v] * xbar[3 * (size_t)v + j] + s[j];
}

// hard_zxu: u += D x - W z - c, comb = |D x - W z - c|^2 + |D (x - x_last)|^2
// (hard/src/Solver.cpp:183-196 + EnergyTerm::update_u).  The finishing CTA applies the break test
// and logs the iteration on the device.
template <int MODE>
__global__ void __launch_bounds__(TET_BLOCK)
k_update_u_hard(TetArrays A, const double *__restrict__ pos_new, const double *__restrict__ pos_last,
                const double *__restrict__ z, const double *__restrict__ u_in, double *__restrict__ u_out,
                SolveState *st, double *partials, double *hist_prim, double *hist_comb, int *hist_rej) {
    if (st->done) return;
    const int T = A.n_tets;
    double acc[2] = {0.0, 0.0};
    for (int t = blockIdx.x * TET_BLOCK + threadIdx.x; t < T; t += gridDim.x * TET_BLOCK) {
        const int4 id = A.idx[t];
        double b[9], F[9], dFm[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) b[k] = A.binv[(size_t)k * T + t];
        const double w = A.w[t];
        deformation_gradient(pos_new, id, b, F);
        if (MODE == MODE_ITER) deformation_gradient_diff(pos_new, pos_last, id, b, dFm);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const double r = w * F[k] - w * z[(size_t)k * T + t];
            u_out[(size_t)k * T + t] = u_in[(size_t)k * T + t] + r;
            if (MODE == MODE_ITER) {
                acc[0] += r * r;
                const double d = w * dFm[k];
                acc[1] += d * d;
            }
        }
    }
    if (MODE != MODE_ITER) return;
    double out[2];
    if (grid_reduce<2, TET_BLOCK>(acc, partials, &st->ticket, out)) {
        if (threadIdx.x == 0) {
            const double comb = out[0] + out[1] + st->tri_comb + st->pt_comb;
            st->comb = comb;
            if (comb < st->eps) {
                st->done = 1;
            } else {
                const int it = st->iter;
                hist_prim[it] = st->prev_prim;
                hist_comb[it] = comb;
                hist_comb[st->max_iters + it] = (double)(global_timer_ns() - st->t0);  // ns since the loop began
                hist_rej[it] = st->reject;
                st->iter = it + 1;
            }
            st->reject = 0;
        }
    }
}

__global__ void k_copy_if_not_done(double *__restrict__ dst, const double *__restrict__ src, int64_t n,
                                   const SolveState *st) {
    if (st->done) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}


// =========================================================================================
// xzu ordering (admm_anderson_xzu/src/Solver.cpp:78-257): kernels that differ from hard_zxu
// =========================================================================================

// u = W^-1 * get_all_gradient(z):  u_i = (1/w) K vol (z_i - U V^T)      (Solver.cpp:125-133)
__global__ void __launch_bounds__(TET_BLOCK, 4)
k_grad_u_xzu(TetArrays A, const double *__restrict__ z, double *__restrict__ u, const SolveState *st) {
    if (st->done) return;
    const int T = A.n_tets;
    for (int t = blockIdx.x * TET_BLOCK + threadIdx.x; t < T; t += gridDim.x * TET_BLOCK) {
        if (A.material && A.material[t] != 0) continue;  // handled by k_hyper
        double zi[9], g[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) zi[k] = z[(size_t)k * T + t];
        tet_grad_linear(zi, A.kvol[t], g);
        const double winv = 1.0 / A.w[t];
#pragma unroll
        for (int k = 0; k < 9; ++k) u[(size_t)k * T + t] = winv * g[k];
    }
}

// z0 = W^-1 (D x - C_fix): the unweighted deformation gradient                (Solver.cpp:80-82)
__global__ void k_z_from_x(TetArrays A, const double *__restrict__ pos, double *__restrict__ z) {
    const int T = A.n_tets;
    for (int t = blockIdx.x * TET_BLOCK + threadIdx.x; t < T; t += gridDim.x * TET_BLOCK) {
        const int4 id = A.idx[t];
        double b[9], F[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) b[k] = A.binv[(size_t)k * T + t];
        deformation_gradient(pos, id, b, F);
        const double w = A.w[t], winv = 1.0 / w;
#pragma unroll
        for (int k = 0; k < 9; ++k) z[(size_t)k * T + t] = (w * F[k]) * winv;
    }
}

// corner contributions of rho dt^2 D^T W (W z - u) for given z, u
// `when`: 0 always, 1 only if st->reject (the redo path of a rejected iterate)
__global__ void k_contrib(TetArrays A, const double *__restrict__ z, const double *__restrict__ u,
                          double *__restrict__ contrib, const SolveState *st, int when) {
    if (st->done || (when == 1 && !st->reject)) return;
    const int T = A.n_tets;
    for (int t = blockIdx.x * TET_BLOCK + threadIdx.x; t < T; t += gridDim.x * TET_BLOCK) {
        double b[9], y[9], q[12];
        const double w = A.w[t];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            b[k] = A.binv[(size_t)k * T + t];
            y[k] = w * z[(size_t)k * T + t] - u[(size_t)k * T + t];
        }
        corner_contrib(b, w, A.rho_dt2, y, q);
        double *qo = contrib + (size_t)t * 12;
#pragma unroll
        for (int k = 0; k < 12; k += 2) *reinterpret_cast<double2 *>(qo + k) = make_double2(q[k], q[k + 1]);
    }
}

// prim = |D x - W z - C_fix|; finishing CTA: accept/reject decision (Solver.cpp:154-159) or, on the
// redo path, just prev_prim = prim (:177-183)
template <int MODE>
__global__ void __launch_bounds__(TET_BLOCK)
k_prim_xzu(TetArrays A, const double *__restrict__ pos, const double *__restrict__ z, SolveState *st, double *partials) {
    if (st->done) return;
    if (MODE == MODE_REDO && !st->reject) return;
    const int T = A.n_tets;
    double acc[1] = {0.0};
    for (int t = blockIdx.x * TET_BLOCK + threadIdx.x; t < T; t += gridDim.x * TET_BLOCK) {
        const int4 id = A.idx[t];
        double b[9], F[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) b[k] = A.binv[(size_t)k * T + t];
        deformation_gradient(pos, id, b, F);
        const double w = A.w[t];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const double r = w * F[k] - w * z[(size_t)k * T + t];
            acc[0] += r * r;
        }
    }
    double out[1];
    if (grid_reduce<1, TET_BLOCK>(acc, partials, &st->ticket, out)) {
        if (threadIdx.x == 0) {
            const double prim = sqrt(out[0]);
            st->prim2 = out[0];
            if (MODE == MODE_ITER) {
                if (st->accel && st->prev_prim < prim) {
                    st->reject = 1;
                    st->n_rejects += 1;
                } else {
                    st->reject = 0;
                    st->prev_prim = prim;
                }
                st->skip_redo = !st->reject;
            } else {
                st->prev_prim = prim;
            }
        }
    }
}

// restore (u, x, z) = default_(u, x, z) and accelerator.replace(z) on a rejected iterate (:162-166)
__global__ void k_restore_xzu(double *__restrict__ u, const double *__restrict__ u_def, double *__restrict__ z,
                              const double *__restrict__ z_def, int64_t nz, double *__restrict__ x,
                              const double *__restrict__ x_def, int64_t nx, const SolveState *st) {
    if (st->done || !st->reject) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nz; i += stride) {
        u[i] = u_def[i];
        z[i] = z_def[i];
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += stride) x[i] = x_def[i];
}

// z_out = prox((D x - c + u)/w)  (EnergyTerm::update_z), no residual, no contributions
__global__ void __launch_bounds__(TET_BLOCK, 4)
k_update_z_plain(TetArrays A, const double *__restrict__ pos, const double *__restrict__ u, double *__restrict__ z_out,
                 const SolveState *st) {
    if (st && st->done) return;
    const int T = A.n_tets;
    for (int t = blockIdx.x * TET_BLOCK + threadIdx.x; t < T; t += gridDim.x * TET_BLOCK) {
        if (A.material && A.material[t] != 0) continue;  // handled by k_hyper
        const int4 id = A.idx[t];
        double b[9], F[9], zi[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) b[k] = A.binv[(size_t)k * T + t];
        deformation_gradient(pos, id, b, F);
        const double w = A.w[t], winv = 1.0 / w;
#pragma unroll
        for (int k = 0; k < 9; ++k) zi[k] = (w * F[k] + u[(size_t)k * T + t]) * winv;
        tet_prox_linear(zi);
#pragma unroll
        for (int k = 0; k < 9; ++k) z_out[(size_t)k * T + t] = zi[k];
    }
}

// u += D x - W z - c, predicated variants for the xzu loop (when: 0 always, 1 only if reject)
__global__ void __launch_bounds__(TET_BLOCK)
k_update_u_plain(TetArrays A, const double *__restrict__ pos, const double *__restrict__ z, double *__restrict__ u,
                 const SolveState *st, int when) {
    if (st->done || (when == 1 && !st->reject)) return;
    const int T = A.n_tets;
    for (int t = blockIdx.x * TET_BLOCK + threadIdx.x; t < T; t += gridDim.x * TET_BLOCK) {
        const int4 id = A.idx[t];
        double b[9], F[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) b[k] = A.binv[(size_t)k * T + t];
        deformation_gradient(pos, id, b, F);
        const double w = A.w[t];
#pragma unroll
        for (int k = 0; k < 9; ++k) u[(size_t)k * T + t] += w * F[k] - w * z[(size_t)k * T + t];
    }
}

// combined residual |W (z_a - z_b)|^2 + |D x - W z_a - C_fix|^2 (Solver.cpp:217-238); the finishing
// CTA logs the iteration and applies the break test (:240-250)
__global__ void __launch_bounds__(TET_BLOCK)
k_comb_xzu(TetArrays A, const double *__restrict__ pos, const double *__restrict__ za, const double *__restrict__ zb,
           SolveState *st, double *partials, double *hist_prim, double *hist_comb, int *hist_rej) {
    if (st->done) return;
    const int T = A.n_tets;
    double acc[2] = {0.0, 0.0};
    for (int t = blockIdx.x * TET_BLOCK + threadIdx.x; t < T; t += gridDim.x * TET_BLOCK) {
        const int4 id = A.idx[t];
        double b[9], F[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) b[k] = A.binv[(size_t)k * T + t];
        deformation_gradient(pos, id, b, F);
        const double w = A.w[t];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const double a = za[(size_t)k * T + t];
            const double d = w * (a - zb[(size_t)k * T + t]);
            const double r = w * F[k] - w * a;
            acc[0] += d * d;
            acc[1] += r * r;
        }
    }
    double out[2];
    if (grid_reduce<2, TET_BLOCK>(acc, partials, &st->ticket, out)) {
        if (threadIdx.x == 0) {
            const double comb = out[0] + out[1];
            st->comb = comb;
            const int it = st->iter;
            hist_prim[it] = st->prev_prim;
            hist_comb[it] = comb;
            hist_comb[st->max_iters + it] = (double)(global_timer_ns() - st->t0);  // ns since the loop began
            hist_rej[it] = st->reject;
            st->iter = it + 1;
            st->reject = 0;
            if (comb < st->eps) st->done = 1;
        }
    }
}

__global__ void k_copy2_if_not_done(double *__restrict__ d0, const double *__restrict__ s0, int64_t n0,
                                    double *__restrict__ d1, const double *__restrict__ s1, int64_t n1,
                                    const SolveState *st) {
    if (st->done) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n0; i += stride) d0[i] = s0[i];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += stride) d1[i] = s1[i];
}

// =========================================================================================
// Hyper-elastic tets (HyperElasticTet::prox / get_gradient, xzu/src/TetEnergyTerm.cpp:171-215):
// one thread per listed tet; runs BEFORE the linear kernel of the same step, which then adds
// st->hyper_prim2 to its own residual sum.
//   KIND 0: hard_zxu local step (z, contributions, residual)   `mode` as k_update_z_hard
//   KIND 1: z_out = prox((D x - c + u)/w)
//   KIND 2: u = W^-1 * vol * dPsi(z)
// =========================================================================================
template <int KIND>
__global__ void __launch_bounds__(TET_BLOCK)
k_hyper(TetArrays A, const double *__restrict__ pos, const double *__restrict__ uz_in, double *__restrict__ out,
        double *__restrict__ contrib, SolveState *st, double *partials, int mode) {
    if (st->done) return;
    if (KIND == 0 && mode == MODE_REDO && !st->reject) return;
    const int T = A.n_tets;
    double acc[1] = {0.0};
    for (int h = blockIdx.x * TET_BLOCK + threadIdx.x; h < A.n_hyper; h += gridDim.x * TET_BLOCK) {
        const int t = A.hyper_ids[h];
        HyperParams P;
        P.mu = A.mu[t];
        P.lambda = A.lambda[t];
        P.k = P.lambda + (2.0 / 3.0) * P.mu;
        P.vol = A.volume[t];
        P.material = A.material[t];
        const double w = A.w[t];
        if (KIND == 2) {
            double zi[9], g[9];
            for (int k = 0; k < 9; ++k) zi[k] = uz_in[(size_t)k * T + t];
            tet_grad_hyper(P, zi, g);
            const double winv = 1.0 / w;
            for (int k = 0; k < 9; ++k) out[(size_t)k * T + t] = winv * g[k];
            continue;
        }
        const int4 id = A.idx[t];
        double b[9], F[9], zi[9], ui[9];
        for (int k = 0; k < 9; ++k) b[k] = A.binv[(size_t)k * T + t];
        deformation_gradient(pos, id, b, F);
        const double winv = 1.0 / w;
        for (int k = 0; k < 9; ++k) {
            ui[k] = uz_in[(size_t)k * T + t];
            F[k] = w * F[k];
            zi[k] = (F[k] + ui[k]) * winv;
        }
        tet_prox_lbfgs(P, zi);
        for (int k = 0; k < 9; ++k) out[(size_t)k * T + t] = zi[k];
        if (KIND == 0) {
            double y[9], q[12];
            for (int k = 0; k < 9; ++k) {
                const double wz = w * zi[k];
                const double r = F[k] - wz;
                acc[0] += r * r;
                y[k] = wz - ui[k];
            }
            corner_contrib(b, w, A.rho_dt2, y, q);
            double *qo = contrib + (size_t)t * 12;
            for (int k = 0; k < 12; ++k) qo[k] = q[k];
        }
    }
    if (KIND != 0) return;
    double o[1];
    if (grid_reduce<1, TET_BLOCK>(acc, partials, &st->ticket, o)) {
        if (threadIdx.x == 0) st->hyper_prim2 = o[0];
    }
}

// ---- batched element kernels (unit parity) ----
__global__ void k_prox_batch(double *z, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double zi[9];
    for (int k = 0; k < 9; ++k) zi[k] = z[9 * i + k];
    tet_prox_linear(zi);
    for (int k = 0; k < 9; ++k) z[9 * i + k] = zi[k];
}
__global__ void k_fmuvt_batch(const double *z, double *out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double zi[9], g[9];
    for (int k = 0; k < 9; ++k) zi[k] = z[9 * i + k];
    tet_grad_linear(zi, 1.0, g);
    for (int k = 0; k < 9; ++k) out[9 * i + k] = g[k];
}

__global__ void k_prox_hyper_batch(HyperParams P, double *z, double *g, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double zi[9], gi[9];
    for (int k = 0; k < 9; ++k) zi[k] = z[9 * i + k];
    tet_grad_hyper(P, zi, gi);
    tet_prox_lbfgs(P, zi);
    for (int k = 0; k < 9; ++k) {
        z[9 * i + k] = zi[k];
        g[9 * i + k] = gi[k];
    }
}

// ---- host launchers (this translation unit is compiled with -fmad=false) ----
void launch_update_z_hard(int mode, int grid, cudaStream_t s, const TetArrays &A, const double *pos, const double *u,
                          double *z, double *contrib, SolveState *st, double *partials) {
    if (A.n_hyper > 0)
        k_hyper<0><<<(A.n_hyper + TET_BLOCK - 1) / TET_BLOCK, TET_BLOCK, 0, s>>>(A, pos, u, z, contrib, st, partials, mode);
    if (mode == MODE_WARM)
        k_update_z_hard<MODE_WARM><<<grid, TET_BLOCK, 0, s>>>(A, pos, u, z, contrib, st, partials);
    else if (mode == MODE_ITER)
        k_update_z_hard<MODE_ITER><<<grid, TET_BLOCK, 0, s>>>(A, pos, u, z, contrib, st, partials);
    else
        k_update_z_hard<MODE_REDO><<<grid, TET_BLOCK, 0, s>>>(A, pos, u, z, contrib, st, partials);
}
void launch_update_u_hard(int mode, int grid, cudaStream_t s, const TetArrays &A, const double *pos_new,
                          const double *pos_last, const double *z, const double *u_in, double *u_out, SolveState *st,
                          double *partials, double *hist_prim, double *hist_comb, int *hist_rej) {
    if (mode == MODE_WARM)
        k_update_u_hard<MODE_WARM><<<grid, TET_BLOCK, 0, s>>>(A, pos_new, pos_last, z, u_in, u_out, st, partials,
                                                              hist_prim, hist_comb, hist_rej);
    else
        k_update_u_hard<MODE_ITER><<<grid, TET_BLOCK, 0, s>>>(A, pos_new, pos_last, z, u_in, u_out, st, partials,
                                                              hist_prim, hist_comb, hist_rej);
}
void launch_restore_if_reject(int grid, cudaStream_t s, double *ucur, const double *gdef, int64_t n, SolveState *st) {
    k_restore_if_reject<<<grid, 256, 0, s>>>(ucur, gdef, n, st);
}
void launch_rhs_gather(cudaStream_t s, int n_free, const int64_t *inc_ptr, const int *inc, const double *contrib,
                       const double *bconst, const int *iperm, double *W, const SolveState *st, int when, int dof_factor) {
    k_rhs_gather<<<(n_free + 127) / 128, 128, 0, s>>>(n_free, inc_ptr, inc, contrib, bconst, iperm, W, st, when, dof_factor);
}
void launch_bconst(cudaStream_t s, const TetArrays &A, const int64_t *inc_ptr, const int *inc, const double *pos,
                   const double *mass, const double *xbar, double *bconst) {
    k_bconst<<<(A.n_free + 127) / 128, 128, 0, s>>>(A, inc_ptr, inc, pos, mass, xbar, bconst);
}
void launch_copy_if_not_done(cudaStream_t s, double *dst, const double *src, int64_t n, const SolveState *st) {
    k_copy_if_not_done<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, s>>>(dst, src, n, st);
}
void launch_prox_hyper_batch(int material, double mu, double lambda, double vol, double *d_z, double *d_g, int64_t n) {
    HyperParams P{mu, lambda, lambda + (2.0 / 3.0) * mu, vol, material};
    k_prox_hyper_batch<<<(unsigned)((n + 127) / 128), 128>>>(P, d_z, d_g, n);
}
void launch_prox_batch(double *d_z, int64_t n) { k_prox_batch<<<(unsigned)((n + 127) / 128), 128>>>(d_z, n); }
void launch_fmuvt_batch(const double *d_z, double *d_out, int64_t n) {
    k_fmuvt_batch<<<(unsigned)((n + 127) / 128), 128>>>(d_z, d_out, n);
}

void launch_grad_u_xzu(int grid, cudaStream_t s, const TetArrays &A, const double *z, double *u, const SolveState *st) {
    if (A.n_hyper > 0)
        k_hyper<2><<<(A.n_hyper + TET_BLOCK - 1) / TET_BLOCK, TET_BLOCK, 0, s>>>(A, nullptr, z, u, nullptr,
                                                                               const_cast<SolveState *>(st), nullptr, 0);
    k_grad_u_xzu<<<grid, TET_BLOCK, 0, s>>>(A, z, u, st);
}
void launch_z_from_x(int grid, cudaStream_t s, const TetArrays &A, const double *pos, double *z) {
    k_z_from_x<<<grid, TET_BLOCK, 0, s>>>(A, pos, z);
}
void launch_contrib(int grid, cudaStream_t s, const TetArrays &A, const double *z, const double *u, double *contrib,
                    const SolveState *st, int when) {
    k_contrib<<<grid, TET_BLOCK, 0, s>>>(A, z, u, contrib, st, when);
}
void launch_prim_xzu(int mode, int grid, cudaStream_t s, const TetArrays &A, const double *pos, const double *z,
                     SolveState *st, double *partials) {
    if (mode == MODE_ITER)
        k_prim_xzu<MODE_ITER><<<grid, TET_BLOCK, 0, s>>>(A, pos, z, st, partials);
    else
        k_prim_xzu<MODE_REDO><<<grid, TET_BLOCK, 0, s>>>(A, pos, z, st, partials);
}
void launch_restore_xzu(int grid, cudaStream_t s, double *u, const double *u_def, double *z, const double *z_def,
                        int64_t nz, double *x, const double *x_def, int64_t nx, const SolveState *st) {
    k_restore_xzu<<<grid, 256, 0, s>>>(u, u_def, z, z_def, nz, x, x_def, nx, st);
}
void launch_update_z_plain(int grid, cudaStream_t s, const TetArrays &A, const double *pos, const double *u,
                           double *z_out, const SolveState *st) {
    if (A.n_hyper > 0)
        k_hyper<1><<<(A.n_hyper + TET_BLOCK - 1) / TET_BLOCK, TET_BLOCK, 0, s>>>(A, pos, u, z_out, nullptr,
                                                                               const_cast<SolveState *>(st), nullptr, 0);
    k_update_z_plain<<<grid, TET_BLOCK, 0, s>>>(A, pos, u, z_out, st);
}
void launch_update_u_plain(int grid, cudaStream_t s, const TetArrays &A, const double *pos, const double *z, double *u,
                           const SolveState *st, int when) {
    k_update_u_plain<<<grid, TET_BLOCK, 0, s>>>(A, pos, z, u, st, when);
}
void launch_comb_xzu(int grid, cudaStream_t s, const TetArrays &A, const double *pos, const double *za,
                     const double *zb, SolveState *st, double *partials, double *hist_prim, double *hist_comb,
                     int *hist_rej) {
    k_comb_xzu<<<grid, TET_BLOCK, 0, s>>>(A, pos, za, zb, st, partials, hist_prim, hist_comb, hist_rej);
}
void launch_copy2_if_not_done(int grid, cudaStream_t s, double *d0, const double *s0, int64_t n0, double *d1,
                              const double *s1, int64_t n1, const SolveState *st) {
    k_copy2_if_not_done<<<grid, 256, 0, s>>>(d0, s0, n0, d1, s1, n1, st);
}

}  // namespace aaadmm
