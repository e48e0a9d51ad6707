// Device-resident LDL^T apply (global step): x = P^T L^-T D^-1 L^-1 P b.
// Replaces m_cholesky->solve (xzu/src/LinearSolver.hpp:87-90 -> Eigen SimplicialCholesky.h:156-180)
// and Geometry's SimplicialLDLTSolver::solve (Geometry/SPDSolver.h:88-91).
#pragma once
#include <cstdint>
#include <vector>

#include "common.cuh"

namespace aaadmm {

struct LdltStats {
    int n = 0, n_blocks = 0, n_levels = 0, max_block = 0;
    int64_t nnz_L = 0, nnz_offdiag = 0, nnz_diag_dense = 0;
    int64_t dense_entries = 0;   // entries of the front matrices one sweep streams (with padding zeros)
    double bytes_per_solve = 0;  // algorithmic bytes of one apply: every factor value once per sweep + vectors
};

// One front = one supernode (elimination-tree chain of ns columns with a common set of k rows below).
struct FrontDesc {
    int first;      // first column (elimination order)
    int ns;         // columns
    int k;          // rows below the diagonal block
    int ld;         // leading dimension of the front matrix (>= ns + k, multiple of 4)
    int pad0, pad1;
    int64_t m_off;  // offset of the front matrix in `M`
    int64_t r_off;  // offset of the k row ids in `rows`
    int64_t u_off;  // offset of the update vector in `U`
    int64_t g_off;  // first entry of this front's ns + k rows in `gptr`
};

// One CTA's work in a sweep kernel: everything it needs in one 64-byte record (one dependent load).
struct alignas(16) SweepTask {
    int64_t m_off;  // front matrix
    int64_t g_off;  // forward: first gather row of the front; backward: offset of the row ids
    int64_t u_off;  // forward: update vector of the front
    int first, ns, k, ld;
    int start;      // forward: first row of the tile; backward: first column
    int shape;      // forward: log2(rows per tile); backward: columns of this CTA
    int wait_idx;   // dependency counter this task waits on ...
    int need;       // ... until it reaches this value (0: no dependency)
    int signal_idx; // counter to bump when the task is done (-1: none)
    int cw;         // backward: columns per warp
};

struct LdltDev {
    int n = 0, nrhs = 3, n_blocks = 0, n_levels = 0;
    int n_launches = 0;  // kernels of one apply
    int *perm = nullptr;   // perm[new] = old
    int *iperm = nullptr;  // iperm[old] = new
    double *dinv = nullptr;  // 1/D
    FrontDesc *fronts = nullptr;
    double *M = nullptr;     // front matrices [Linv ; Q], column-major ld x ns each (setup only)
    double *Mf = nullptr;    // the same entries tile by tile in the order the forward sweep streams them
    double *Mb = nullptr;    // ... and stage by stage in the order the backward sweep streams them
    int *rows = nullptr;     // row ids below each front
    int4 *gell = nullptr;    // per front row: up to 4 child update slots that add into it (-1 = none)
    int64_t *gptr = nullptr; // further slots (CSR), null when no row has more than 4
    int *gidx = nullptr;
    SweepTask *tasks = nullptr;  // forward tasks (fronts bottom-up), then backward tasks (top-down)
    int n_ftasks = 0, n_btasks = 0;
    int ws_cap = 0, v_cap = 0;   // staged columns (forward) / rows (backward) per chunk
    int *ctl = nullptr;          // [0], [1]: next forward / backward task; then one arrival counter per front and sweep
    int n_ctl = 0;
    unsigned long long *trace = nullptr;  // AAADMM_LDLT_TRACE: 4 time stamps per task
    std::vector<SweepTask> host_tasks;
    std::vector<int> host_level;
    // work vectors: W = permuted rhs, Yd = D^-1 L^-1 rhs, X = solution (elimination order), U = front updates
    double *W = nullptr, *Yd = nullptr, *X = nullptr, *U = nullptr;
    double *Va = nullptr;  // wide fronts: [D^-1 y ; -x(rows below) ; 0] assembled once per front (backward sweep)
    LdltStats stats;
    // ---- kept from the setup so that new VALUES can be loaded into the same structure (ldlt_dev_refactor) ----
    int64_t m_tot = 0, mf_tot = 0, mb_tot = 0;
    int2 *inv_tasks = nullptr;   // (front, first row) per 128 rows of a diagonal block
    int4 *q_tasks = nullptr;     // (front, first row, first column) per 64 x 64 tile of Q
    int n_inv_tasks = 0, n_q_tasks = 0;
    int2 *invd_tasks = nullptr, *invo_tasks = nullptr;  // blocked inverse: (front, diagonal block) / (front, block column) by distance
    int n_invd_tasks = 0;
    std::vector<int> invo_off;   // first task of block distance r + 1
    int64_t *tile_src = nullptr; // per task: offset of its front matrix in M
    double *A = nullptr;         // front matrices [T ; P] (only kept by factors created from a matrix)
    double *D = nullptr;         // pivots
    struct FactorPlan *plan = nullptr;  // device-side numeric factorisation (ldlt_factor.cu); null for factors given as (L, D)
    cudaStream_t setup_stream = nullptr;
};

// Builds the device structure from a strictly-lower CSC factor (rows sorted per column).
int ldlt_dev_create(LdltDev **out, int n, const int64_t *Lp, const int *Li, const double *Lx,
                    const double *D, const int *perm, int nrhs);
void ldlt_dev_destroy(LdltDev *f);
// The same structure from the PATTERN of L (Lp, Li) and of the matrix (Ap, Ai: lower CSC incl. diagonal, original
// numbering); the values Ax are factored on the device (ldlt_factor.cu). The structure stays allocated:
// ldlt_dev_refactor loads another matrix with the same pattern without any allocation.
int ldlt_dev_create_from_matrix(LdltDev **out, int n, const int64_t *Ap, const int *Ai, const double *Ax, const int64_t *Lp,
                                const int *Li, const int *perm, int nrhs);
// New values for a factor made by ldlt_dev_create_from_matrix (host array, lower CSC order of Ap / Ai). Enqueued on
// `stream` (stream-ordered with the applies the caller launches there afterwards); returns after the pivots were checked.
int ldlt_dev_refactor(LdltDev *f, const double *Ax, cudaStream_t stream);

// Developer aid (AAADMM_LDLT_TRACE=1 at creation): per-task time stamps of the last apply as CSV.
int ldlt_dev_dump_trace(LdltDev *f, const char *path);

// W (permuted rhs, n x nrhs interleaved) must already be filled. Result is written to
// x_out[perm[i]*nrhs + r] (original order) on `stream`.
// `skip` (may be null) points to a device flag; when it is non-zero every kernel returns at once
// (the solver's on-device break test).
int ldlt_dev_apply_permuted(LdltDev *f, double *x_out, cudaStream_t stream, const int *skip);
// b in original order -> x in original order (both device pointers).
int ldlt_dev_apply(LdltDev *f, const double *b, double *x_out, cudaStream_t stream, const int *skip);

}  // namespace aaadmm
