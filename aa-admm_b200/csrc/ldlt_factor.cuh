// Device-side numeric LDL^T factorisation into the per-front layout of ldlt_apply.cu (see ldlt_factor.cu).
#pragma once
#include <cstdint>
#include <vector>

#include "common.cuh"
#include "ldlt_apply.cuh"

namespace aaadmm {

struct FactorPlan;

// Pattern-only setup. fr / rows / parent / level / blk_of: the front structure ldlt_apply.cu derives from the pattern
// of L; perm[new] = old; Ap / Ai: lower CSC pattern of the matrix in its original numbering.
int factor_plan_build(FactorPlan **out, int n, const std::vector<FrontDesc> &fr, int nb, const std::vector<int> &rows,
                      const std::vector<int> &parent, const std::vector<int> &level, const std::vector<int> &blk_of,
                      const int *perm, const int64_t *Ap, const int *Ai, int64_t m_tot);
void factor_plan_destroy(FactorPlan *p);
int64_t factor_plan_nnz(const FactorPlan *p);
double *factor_plan_values(FactorPlan *p);  // device buffer the caller fills with the matrix values before a run
int factor_plan_launches(const FactorPlan *p);
// Factor: front matrices [T ; P] (unit diagonal explicit) into dA, pivots into D, reciprocals into dinv; on `s`.
int factor_plan_run(FactorPlan *p, double *dA, double *D, double *dinv, cudaStream_t s);
// Waits for `s`; -1 if a pivot was zero or non-finite.
int factor_plan_check(FactorPlan *p, cudaStream_t s);

}  // namespace aaadmm
