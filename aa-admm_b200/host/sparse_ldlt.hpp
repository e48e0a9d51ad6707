// Host-side sparse LDL^T used at setup (the global step's system matrix is factored ONCE;
// the per-iteration apply runs on the GPU: csrc/ldlt_apply.cu).
//
// Role in the reference: admm::LDLTSolver::update_system -> Eigen::SimplicialLDLT::compute
// (xzu/src/LinearSolver.hpp:79-84) and Geometry's SimplicialLDLTSolver::initialize
// (Geometry/SPDSolver.h:73-86). The reference orders with AMD and factors column by column
// on one thread; this is a different algorithm for the same factorisation P A P^T = L D L^T:
// geometric nested dissection (wide, shallow elimination tree = few dependent levels for the
// GPU triangular solves) + supernodal multifrontal numeric phase on OpenMP tasks.
// The C ABI (include/aaadmm.h) accepts ANY (L, D, perm) in this CSC form, so a host that
// already has Eigen's factor (the reference) can pass that one instead.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace aaadmm {

// Symmetric matrix, lower triangle including the diagonal, CSC, rows sorted per column.
struct SymLower {
    int n = 0;
    std::vector<int64_t> p;
    std::vector<int> i;
    std::vector<double> x;
};

struct LdltFactor {
    int n = 0;
    std::vector<int> perm;    // perm[new] = old
    std::vector<int64_t> Lp;  // strictly-lower L, CSC, rows sorted, unit diagonal implied
    std::vector<int> Li;
    std::vector<double> Lx;
    std::vector<double> D;
    // statistics of the numeric phase
    int n_supernodes = 0;
    double flops = 0;
    double seconds_order = 0, seconds_symbolic = 0, seconds_numeric = 0;
    bool ok = false;  // false if a pivot was zero / non-finite
};

// On-disk factor cache (at 1M tets the setup dwarfs a short simulation). Little-endian binary:
//   magic "AAFACT01", uint64 key, int32 n, int64 nnz, perm[n] int32, Lp[n+1] int64, Li[nnz] int32, Lx[nnz] f64,
//   D[n] f64. `key` = matrix_key(A) of the matrix the factor belongs to (FNV-1a over n, pattern and values): a
//   cached factor is only used for a bit-identical matrix.
uint64_t matrix_key(const SymLower &A);
bool ldlt_save(const LdltFactor &F, uint64_t key, const std::string &path);
bool ldlt_load(const std::string &path, uint64_t key, LdltFactor &F);

// Builds SymLower from (row, col, value) triplets of the FULL or LOWER part; duplicates are
// summed; entries with row < col are mirrored into the lower triangle.
SymLower sym_from_triplets(int n, const std::vector<int> &r, const std::vector<int> &c,
                           const std::vector<double> &v, bool input_is_full);

// Nested-dissection ordering. coords: 3 doubles per node (may be null -> BFS bisection).
// Returns perm with perm[new] = old.
std::vector<int> nested_dissection(const SymLower &A, const double *coords, int leaf_size = 96);

// Numeric factorisation with the given ordering.
LdltFactor ldlt_factorize(const SymLower &A, const std::vector<int> &perm, int n_threads = 0);

// Symbolic phase alone: perm, Lp, Li (pattern of L) and the statistics; Lx and D stay empty. The numeric phase then
// runs on the device from the matrix values (aaadmm_ldlt_create_from_matrix, csrc/ldlt_factor.cu); the pattern is
// reusable for every matrix with the same sparsity (material sweeps over one mesh).
LdltFactor ldlt_symbolic(const SymLower &A, const std::vector<int> &perm, int n_threads = 0);

// Reference host solve (used by tests and by the setup self-check): x = A^-1 b for
// nrhs interleaved right-hand sides (b[i*nrhs + k]).
void ldlt_solve_host(const LdltFactor &F, const double *b, double *x, int nrhs);

}  // namespace aaadmm
