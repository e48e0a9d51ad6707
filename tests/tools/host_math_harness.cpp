// Developer/test harness: compiles the __host__ __device__ math helpers of csrc/ for the CPU so
// that logic errors are caught without GPU time. NOT part of the product libraries.
#include "../../aa-admm_b200/csrc/svd3.cuh"
#include "../../aa-admm_b200/csrc/cod_small.cuh"
#include "../../aa-admm_b200/csrc/lbfgs_prox.cuh"
extern "C" {
void harness_prox_hyper(int material, double mu, double lambda, double k, double vol, double *z, double *g, int n) {
    aaadmm::HyperParams P{mu, lambda, k, vol, material};
    for (int i = 0; i < n; ++i) { if (g) aaadmm::tet_grad_hyper(P, z + 9 * i, g + 9 * i); aaadmm::tet_prox_lbfgs(P, z + 9 * i); }
}
void harness_prox(double *z, int n) { for (int i = 0; i < n; ++i) aaadmm::tet_prox_linear(z + 9 * i); }
void harness_fmuvt(const double *z, double *o, int n) { for (int i = 0; i < n; ++i) aaadmm::tet_grad_linear(z + 9 * i, 1.0, o + 9 * i); }
int harness_cod(int m, const double *M, const double *rhs, double *x) { double A[256]; for (int i = 0; i < m * m; ++i) A[i] = M[i]; return aaadmm::cod_solve(A, m, rhs, x); }
}
