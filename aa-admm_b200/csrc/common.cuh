// Shared device helpers: error handling, deterministic two-stage reductions, solver state.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

#ifndef AA_MAX_M
#define AA_MAX_M 16
#endif

namespace aaadmm {

void set_last_error(const std::string &msg);

#define AAADMM_CUDA_OK(expr)                                                                      \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ::aaadmm::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));         \
            return -1;                                                                            \
        }                                                                                         \
    } while (0)

// Number of CTAs used by the streaming kernels: a multiple of the SM count so that every SM
// holds the same number of resident CTAs (B200: 148 SMs).
int stream_grid(int ctas_per_sm);
int sm_count();

constexpr int RED_MAX_BLOCKS = 148 * 16;  // upper bound on the grid of a reducing kernel
constexpr int RED_MAX_Q = 2 * AA_MAX_M + 2;

// Device-resident control block of one ADMM solve: everything the safeguard needs, so that
// no per-iteration scalar ever travels to the host (hard/src/Solver.cpp:139-172,183-189).
struct SolveState {
    double prim2;      // squared primal residual of the latest local step
    double prev_prim;  // prev_prim_residual (norm)
    double comb;       // latest combined residual
    double eps;        // break threshold on comb (1e-20 in the reference)
    double hyper_prim2;  // residual share of the hyper-elastic tets (0 for linear scenes)
    int reject;        // decision of the current iteration
    int done;          // comb < eps reached: the remaining launches of this step are no-ops
    int iter;          // rows logged so far
    int n_rejects;
    int accel;         // Settings::ANDERSON
    int pad_;
    // Anderson state (hard/src/AndersonAcceleration.h:137-152)
    int aa_iter, aa_col, aa_m, aa_mk;
    double aa_scale[AA_MAX_M];
    double aa_M[AA_MAX_M * AA_MAX_M];  // scaled Gram matrix, column-major with ld = AA_MAX_M
    double aa_coef[AA_MAX_M];          // theta ./ scale of the current call
    unsigned int ticket;               // last-block election counter of the reducing kernels
    int loop_it, max_iters;            // device-side loop control of the graph WHILE node
    int skip_redo;                     // xzu: 1 unless the current iterate was rejected (guards the redo solve)
    int aa_skip;                       // geometry: 1 on a rejected turn (the Anderson passes do not run)
    unsigned long long t0;             // %globaltimer when the iteration loop began (per-iteration time stamps of the log)
    double tri_prim2, tri_comb;        // residual shares of the triangle terms (0 for tet-only scenes)
    double pt_prim2, pt_comb;          // residual shares of the collision terms
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
    return t;
}

// ---------------------------------------------------------------------------------------
// Deterministic grid reduction of NQ doubles per thread.
//   stage 1: warp shuffle tree -> shared -> one partial per CTA and quantity
//   stage 2: the LAST CTA to arrive (ticket) sums the partials of all CTAs in CTA order with a
//            fixed-shape tree, so the result does not depend on which CTA was last.
// Returns true in the finishing CTA (all its threads), with out[q] valid in thread 0.
// ---------------------------------------------------------------------------------------
template <int NQ, int BLOCK>
__device__ __forceinline__ bool grid_reduce(double (&v)[NQ], double *partials /*[NQ][gridDim.x]*/,
                                            unsigned int *ticket, double (&out)[NQ]) {
    __shared__ double s_red[NQ][BLOCK / 32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        double x = v[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) s_red[q][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x < NQ) {
        double x = 0.0;
#pragma unroll
        for (int w = 0; w < BLOCK / 32; ++w) x += s_red[threadIdx.x][w];
        partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = x;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    // stage 2
    // (all loads of all quantities are issued before the first use: one L2 round trip instead of NQ)
    const int nb = gridDim.x;
    double x[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) x[q] = 0.0;
    for (int b = threadIdx.x; b < nb; b += BLOCK) {
        double t[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) t[q] = __ldcg(&partials[(size_t)q * nb + b]);
#pragma unroll
        for (int q = 0; q < NQ; ++q) x[q] += t[q];
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x[q] += __shfl_down_sync(0xffffffffu, x[q], o);
        if (lane == 0) s_red[q][warp] = x[q];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            double x = 0.0;
#pragma unroll
            for (int w = 0; w < BLOCK / 32; ++w) x += s_red[q][w];
            out[q] = x;
        }
        *ticket = 0u;
    }
    return true;
}

}  // namespace aaadmm
