// Anderson mixing on the device (K8-K11 of SURVEY 2b), two streaming passes per call.
//
// Restates AndersonAcceleration::compute_impl (hard/src/AndersonAcceleration.h:154-211; variant X:
// xzu/src/AndersonAcceleration.h:138-200) with the history kept RAW in HBM:
//   the reference normalises the newest dF column in place (dF(:,c) /= scale) and keeps the
//   normalised columns; here the columns stay unscaled and the scales live in the m x m Gram
//   matrix instead:  M(c,j) = <dF_c,dF_j>/(s_c s_j),  rhs_j = <dF_j,F>/s_j  - the same numbers
//   up to rounding, one full read+write pass over the Ne x 1 column cheaper.  As in the
//   reference only the row/column of the newest column is refreshed per call.
//   pass 1: F = G - u_cur (effective part); dF(:,c) += F; dG(:,c) += G; all 2 m_k dot products
//           in the same pass (warp shuffle -> CTA -> fixed-order last-CTA reduction); the last
//           CTA then solves the m_k x m_k system in one thread (cod_small.cuh).
//   pass 2: u_cur = G - dG(:,0:mk) (theta./scale); start the next column with -F / -G.
#pragma once
#include "cod_small.cuh"
#include "common.cuh"

namespace aaadmm {

constexpr int AA_BLOCK = 256;
constexpr int AA_ILP = 2;  // elements per thread and loop turn in the streaming passes

// Shared-memory scratch of the m_k x m_k solve between the two passes.
struct AaFinishWork {
    double dots[AA_MAX_M], rhsv[AA_MAX_M];  // <dF_c, dF_j> (dots[c] = |dF_c|^2) and <dF_j, F>, raw
    double g[AA_MAX_M], rhs[AA_MAX_M], theta[AA_MAX_M], hc[AA_MAX_M];
    double A[AA_MAX_M * AA_MAX_M];
    int trans[AA_MAX_M];
    CodWarpWork qr;
};

// Runs in warp 0 of the finishing CTA of pass 1 (all 32 lanes): hard/src/AndersonAcceleration.h:174-205 with the
// scaled Gram matrix (see the header comment). The scalings are one division per lane, the QR is cod_qr_warp, the rest
// of the complete orthogonal decomposition (cod_finish) runs on lane 0.
__device__ __forceinline__ void aa_finish_warp(SolveState *st, AaFinishWork &w) {
    const int lane = threadIdx.x & 31;
    const int iter = st->aa_iter, c = st->aa_col, m = st->aa_m;
    const double eps = 1e-14;
    const int mk = iter < m ? iter : m;
    const double nrm2 = w.dots[c];
    const double scale = fmax(eps, sqrt(nrm2));
    const double sc = (lane < mk && lane != c) ? st->aa_scale[lane] : scale;  // scale of column `lane`
    if (mk == 1) {
        if (lane == 0) {
            w.theta[0] = 0.0;
            const double sq = nrm2 / (scale * scale);
            st->aa_M[0] = sq;
            const double dF_norm = sqrt(sq);
            if (dF_norm > eps) w.theta[0] = ((w.rhsv[c] / scale) / dF_norm) / dF_norm;
        }
    } else {
        if (lane < mk) {
            const double g = (lane == c) ? nrm2 / (scale * scale) : w.dots[lane] / (scale * sc);
            st->aa_M[c * AA_MAX_M + lane] = g;
            st->aa_M[lane * AA_MAX_M + c] = g;
            w.g[lane] = g;
            w.rhs[lane] = w.rhsv[lane] / sc;
        }
        __syncwarp();
        for (int e = lane; e < mk * mk; e += 32) {
            const int j = e / mk, i = e - j * mk;
            w.A[e] = (j == c) ? w.g[i] : (i == c) ? w.g[j] : st->aa_M[j * AA_MAX_M + i];
        }
        __syncwarp();
        int nonzero_pivots;
        double maxpivot;
        cod_qr_warp(w.A, mk, w.hc, w.trans, nonzero_pivots, maxpivot, w.qr);
        if (lane == 0) cod_finish(w.A, mk, w.hc, w.trans, nonzero_pivots, maxpivot, w.rhs, w.theta);
    }
    __syncwarp();
    if (lane < mk) st->aa_coef[lane] = w.theta[lane] / sc;
    if (lane == 0) {
        st->aa_scale[c] = scale;
        st->aa_mk = mk;
        st->aa_col = (c + 1) % m;
        st->aa_iter = iter + 1;
    }
}

// G = [g_u (Ne) | g_x (Nt-Ne)]; when gx_dst != nullptr the x part is also copied there
// (the solver hands the fresh solve result in and keeps it as default_x).
template <int M>
__global__ void __launch_bounds__(AA_BLOCK)
k_aa_pass1(const double *__restrict__ g_u, const double *__restrict__ g_x, double *__restrict__ gx_dst,
           double *__restrict__ ucur, double *__restrict__ dF, double *__restrict__ dG, int64_t Ne, int64_t Nt,
           SolveState *st, double *partials, double *__restrict__ g_copy) {
    if (st->done || st->aa_skip) return;
    const int iter = st->aa_iter, c = st->aa_col, m = st->aa_m;
    const int64_t stride = (int64_t)gridDim.x * AA_BLOCK;
    const int64_t i0 = (int64_t)blockIdx.x * AA_BLOCK + threadIdx.x;
    if (iter == 0) {
        for (int64_t i = i0; i < Nt; i += stride) {
            double g;
            if (i < Ne) {
                g = g_u[i];
                dF[i] = -(g - ucur[i]);
            } else {
                g = g_x[i - Ne];
                if (gx_dst) gx_dst[i - Ne] = g;
            }
            if (g_copy) g_copy[i] = g;
            dG[i] = -g;
            ucur[i] = g;
        }
        // state changes only after every CTA has read it: last-CTA election as in grid_reduce
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int t = atomicAdd(&st->ticket, 1u);
            if (t == gridDim.x - 1) {
                st->ticket = 0u;
                st->aa_mk = 0;  // first call after init / reset: nothing to mix yet
                st->aa_iter = 1;
            }
        }
        return;
    }
    const int mk = iter < m ? iter : m;
    double acc[2 * M];
#pragma unroll
    for (int q = 0; q < 2 * M; ++q) acc[q] = 0.0;
    double *dFc = dF + (size_t)c * Ne;
    double *dGc = dG + (size_t)c * Nt;
    // AA_ILP elements per thread and turn: all loads are issued before the first store (the stores
    // to column c alias the loads of the other columns for the compiler, which would otherwise
    // serialise the turns and leave one element's worth of loads in flight).
    for (int64_t base = i0; base < Ne; base += stride * AA_ILP) {
        double g[AA_ILP], uc[AA_ILP], fc[AA_ILP], gc[AA_ILP], h[AA_ILP][M];
#pragma unroll
        for (int e = 0; e < AA_ILP; ++e) {
            const int64_t i = base + e * stride;
            if (i < Ne) {
                g[e] = g_u[i];
                uc[e] = ucur[i];
                fc[e] = dFc[i];
                gc[e] = dGc[i];
#pragma unroll
                for (int j = 0; j < M; ++j)
                    if (j < mk && j != c) h[e][j] = dF[(size_t)j * Ne + i];
            }
        }
#pragma unroll
        for (int e = 0; e < AA_ILP; ++e) {
            const int64_t i = base + e * stride;
            if (i < Ne) {
                const double F = g[e] - uc[e];
                const double a = fc[e] + F;
                dFc[i] = a;
                dGc[i] = gc[e] + g[e];
                if (g_copy) g_copy[i] = g[e];
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    if (j < mk) {
                        const double v = (j == c) ? a : h[e][j];
                        acc[j] += a * v;
                        acc[M + j] += v * F;
                    }
                }
            }
        }
    }
    for (int64_t i = Ne + i0; i < Nt; i += stride) {
        const double g = g_x[i - Ne];
        if (gx_dst) gx_dst[i - Ne] = g;
        if (g_copy) g_copy[i] = g;
        dGc[i] += g;
    }
    double out[2 * M];
    __shared__ AaFinishWork s_fin;
    if (grid_reduce<2 * M, AA_BLOCK>(acc, partials, &st->ticket, out)) {  // true in every thread of the last CTA
        if (threadIdx.x == 0) {
#pragma unroll
            for (int q = 0; q < M; ++q) {
                s_fin.dots[q] = out[q];
                s_fin.rhsv[q] = out[M + q];
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) aa_finish_warp(st, s_fin);
    }
}

template <int M>
__global__ void __launch_bounds__(AA_BLOCK)
k_aa_pass2(const double *__restrict__ g_u, const double *__restrict__ g_x, double *__restrict__ ucur,
           double *__restrict__ dF, double *__restrict__ dG, int64_t Ne, int64_t Nt, const SolveState *st) {
    if (st->done || st->aa_skip) return;
    const int mk = st->aa_mk;
    if (mk == 0) return;
    const int cn = st->aa_col;  // already advanced by pass 1
    double coef[M];
#pragma unroll
    for (int j = 0; j < M; ++j) coef[j] = (j < mk) ? st->aa_coef[j] : 0.0;
    const int64_t stride = (int64_t)gridDim.x * AA_BLOCK;
    for (int64_t base = (int64_t)blockIdx.x * AA_BLOCK + threadIdx.x; base < Nt; base += stride * AA_ILP) {
        double g[AA_ILP], uc[AA_ILP], s[AA_ILP];
#pragma unroll
        for (int e = 0; e < AA_ILP; ++e) {
            const int64_t i = base + e * stride;
            s[e] = 0.0;
            if (i < Nt) {
                g[e] = (i < Ne) ? g_u[i] : g_x[i - Ne];
                uc[e] = (i < Ne) ? ucur[i] : 0.0;
#pragma unroll
                for (int j = 0; j < M; ++j)
                    if (j < mk) s[e] += dG[(size_t)j * Nt + i] * coef[j];
            }
        }
#pragma unroll
        for (int e = 0; e < AA_ILP; ++e) {
            const int64_t i = base + e * stride;
            if (i < Nt) {
                if (i < Ne) dF[(size_t)cn * Ne + i] = -(g[e] - uc[e]);
                dG[(size_t)cn * Nt + i] = -g[e];
                ucur[i] = g[e] - s[e];
            }
        }
    }
}

// Host-side dispatch on the window size m (compile-time accumulator count).
int launch_aa_pass1(int m, int grid, cudaStream_t s, const double *g_u, const double *g_x, double *gx_dst,
                    double *ucur, double *dF, double *dG, int64_t Ne, int64_t Nt, SolveState *st, double *partials,
                    double *g_copy = nullptr);
int launch_aa_pass2(int m, int grid, cudaStream_t s, const double *g_u, const double *g_x, double *ucur,
                    double *dF, double *dG, int64_t Ne, int64_t Nt, const SolveState *st);

}  // namespace aaadmm
