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
    double bytes_per_solve = 0;  // algorithmic bytes one apply streams (both sweeps)
};

struct LdltDev {
    int n = 0, nrhs = 3, n_blocks = 0, n_levels = 0;
    // permutation
    int *perm = nullptr;   // perm[new] = old
    int *iperm = nullptr;  // iperm[old] = new
    // block (supernode chain) partition and level schedule
    int *blk_of = nullptr;     // [n]
    int *blk_first = nullptr;  // [n_blocks+1]
    int64_t *linv_off = nullptr;  // [n_blocks+1] offsets of the dense ns x ns inverse blocks
    double *Linv = nullptr;       // row-major inverse of the unit-lower diagonal blocks
    double *LinvT = nullptr;      // its transpose (rows of L^-T)
    double *dinv = nullptr;       // 1/D
    // forward sweep: CSR of the off-block part; backward sweep: CSC of the same entries
    int64_t *fr_ptr = nullptr;
    int *fr_col = nullptr;
    double *fr_val = nullptr;
    int64_t *bc_ptr = nullptr;
    int *bc_row = nullptr;
    double *bc_val = nullptr;
    // per level and sweep kernel (fwd_off, fwd_diag, bwd_off, bwd_diag): LONG, SHORT and TINY rows;
    // list_ptr[(level*4 + kind)*3 + {0,1,2,3}] delimit them inside lev_rows
    std::vector<int> list_ptr;
    int *lev_rows = nullptr;
    // work vectors n x nrhs
    double *W = nullptr, *Y = nullptr, *X = nullptr;
    LdltStats stats;
};

// Builds the device structure from a strictly-lower CSC factor (rows sorted per column).
int ldlt_dev_create(LdltDev **out, int n, const int64_t *Lp, const int *Li, const double *Lx,
                    const double *D, const int *perm, int nrhs);
void ldlt_dev_destroy(LdltDev *f);

// W (permuted rhs, n x nrhs interleaved) must already be filled. Result is written to
// x_out[perm[i]*nrhs + r] (original order) on `stream`.
// `skip` (may be null) points to a device flag; when it is non-zero every kernel returns at once
// (the solver's on-device break test).
int ldlt_dev_apply_permuted(LdltDev *f, double *x_out, cudaStream_t stream, const int *skip);
// b in original order -> x in original order (both device pointers).
int ldlt_dev_apply(LdltDev *f, const double *b, double *x_out, cudaStream_t stream, const int *skip);

}  // namespace aaadmm
