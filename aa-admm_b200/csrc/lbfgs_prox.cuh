// Hyper-elastic tet prox: per-tet 9-dimensional L-BFGS (one thread per tet, history in local memory).
//
// Restates HyperElasticTet::prox (xzu/src/TetEnergyTerm.cpp:171-183) with
//   NeoHookeanTet::NHProx   energy_density / gradient / U_gradient   :221-267
//   StVKTet::StVKProx       energy_density / gradient / U_gradient   :272-319
//   mcl::optlib::LBFGS<double,9>::minimize + LineSearch (Armijo backtracking)
//                           deps/mcloptlib/include/MCL/LBFGS.hpp:135-305
//                           (m = 6, epsilon = 1e-6, past = 1, delta = 1e-16, ftol = 1e-4, dec = 0.5,
//                            max_iters = 100, max_linesearch = 2000, min_step = 1e-20, max_step = 1e20)
// The minimised objective is vol * (Psi(z) + k/2 |z - v|^2), started at z = v.
// Where the reference throws (line-search step out of range) the device version stops the search.
#pragma once
#include <cmath>

#ifndef AAADMM_HD
#ifdef __CUDACC__
#define AAADMM_HD __host__ __device__ __forceinline__
#else
#define AAADMM_HD inline
#endif
#endif

namespace aaadmm {

struct HyperParams {
    double mu, lambda, k, vol;
    int material;  // 1 Neo-Hookean, 2 St Venant-Kirchhoff
};

namespace hyper {

AAADMM_HD double det3c(const double *m) {
    return m[0] * (m[4] * m[8] - m[7] * m[5]) - m[3] * (m[1] * m[8] - m[7] * m[2]) + m[6] * (m[1] * m[5] - m[4] * m[2]);
}
// Eigen's 3x3 inverse through cofactors (Eigen/src/LU/InverseImpl.h:120-170), column-major
AAADMM_HD void inv3c(const double *m, double *inv) {
#define HM(r, c) m[(c) * 3 + (r)]
#define HCOF(i, j) (HM(((i) + 1) % 3, ((j) + 1) % 3) * HM(((i) + 2) % 3, ((j) + 2) % 3) - HM(((i) + 1) % 3, ((j) + 2) % 3) * HM(((i) + 2) % 3, ((j) + 1) % 3))
    const double c0 = HCOF(0, 0), c1 = HCOF(1, 0), c2 = HCOF(2, 0);
    const double det = (c0 * HM(0, 0) + c1 * HM(1, 0)) + c2 * HM(2, 0);
    const double id = 1.0 / det;
    inv[0] = c0 * id;
    inv[3] = c1 * id;
    inv[6] = c2 * id;
    inv[1] = HCOF(0, 1) * id;
    inv[4] = HCOF(1, 1) * id;
    inv[7] = HCOF(2, 1) * id;
    inv[2] = HCOF(0, 2) * id;
    inv[5] = HCOF(1, 2) * id;
    inv[8] = HCOF(2, 2) * id;
#undef HCOF
#undef HM
}
// C = F^T F
AAADMM_HD void ftf(const double *F, double *C) {
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) C[j * 3 + i] = F[i * 3 + 0] * F[j * 3 + 0] + F[i * 3 + 1] * F[j * 3 + 1] + F[i * 3 + 2] * F[j * 3 + 2];
}

AAADMM_HD double energy_density(const HyperParams &P, const double *F) {
    double C[9];
    ftf(F, C);
    if (P.material == 1) {
        const double J = det3c(F);
        const double I1 = C[0] + C[4] + C[8], I3 = J * J;
        const double l3 = log(I3);
        return 0.5 * P.mu * (I1 - l3 - 3.0) + 0.125 * P.lambda * l3 * l3;
    }
    double E[9];
    for (int k = 0; k < 9; ++k) E[k] = 0.5 * (C[k] - ((k % 4 == 0) ? 1.0 : 0.0));
    const double tr = E[0] + E[4] + E[8];
    double ee = 0.0;  // trace(E^T E)
    for (int j = 0; j < 3; ++j) ee += E[j * 3 + 0] * E[j * 3 + 0] + E[j * 3 + 1] * E[j * 3 + 1] + E[j * 3 + 2] * E[j * 3 + 2];
    return P.mu * ee + 0.5 * P.lambda * tr * tr;
}

AAADMM_HD void u_gradient(const HyperParams &P, const double *F, double *G) {
    if (P.material == 1) {
        double Fi[9];
        inv3c(F, Fi);
        const double lj = P.lambda * log(det3c(F));
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) {
                const double fit = Fi[r * 3 + c];  // (F^-1)^T (r,c)
                G[c * 3 + r] = P.mu * (F[c * 3 + r] - fit) + lj * fit;
            }
        return;
    }
    double C[9], S[9];
    ftf(F, C);
    double tr = 0.0;
    for (int k = 0; k < 9; ++k) C[k] = 0.5 * (C[k] - ((k % 4 == 0) ? 1.0 : 0.0));
    tr = C[0] + C[4] + C[8];
    for (int k = 0; k < 9; ++k) S[k] = 2.0 * P.mu * C[k] + ((k % 4 == 0) ? P.lambda * tr : 0.0);
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) G[c * 3 + r] = F[0 * 3 + r] * S[c * 3 + 0] + F[1 * 3 + r] * S[c * 3 + 1] + F[2 * 3 + r] * S[c * 3 + 2];
}

// fx = vol * (Psi(x) + k/2 |v - x|^2), grad = vol * (dPsi(x) + k (x - v))
AAADMM_HD double value_grad(const HyperParams &P, const double *v, const double *x, double *grad) {
    u_gradient(P, x, grad);
    double q = 0.0;
    for (int i = 0; i < 9; ++i) {
        const double d = v[i] - x[i];
        q += d * d;
        grad[i] = P.vol * (grad[i] + P.k * (x[i] - v[i]));
    }
    return P.vol * (energy_density(P, x) + 0.5 * P.k * q);
}

AAADMM_HD double dot9(const double *a, const double *b) {
    double s = 0.0;
    for (int i = 0; i < 9; ++i) s += a[i] * b[i];
    return s;
}

}  // namespace hyper

// z (in: v, out: minimiser)
AAADMM_HD void tet_prox_lbfgs(const HyperParams &P, double *z) {
    using namespace hyper;
    const int M = 6;
    const double epsilon = 1e-6, delta = 1e-16, ftol = 1e-4, min_step = 1e-20, max_step = 1e+20;
    double v[9], x[9], xp[9], grad[9], gradp[9], drt[9], s[M][9], y[M][9], ys[M], alpha[M];
    for (int i = 0; i < 9; ++i) v[i] = x[i] = z[i];
    double fx = value_grad(P, v, x, grad);
    double xnorm = sqrt(dot9(x, x)), gnorm = sqrt(dot9(grad, grad));
    double fx_prev = fx;
    if (gnorm <= epsilon * fmax(xnorm, 1.0)) return;
    for (int i = 0; i < 9; ++i) drt[i] = -grad[i];
    double step = 1.0 / sqrt(dot9(drt, drt));
    int k = 1, end = 0;
    for (;;) {
        for (int i = 0; i < 9; ++i) {
            xp[i] = x[i];
            gradp[i] = grad[i];
        }
        {  // LineSearch: backtracking, Armijo
            const double fx_init = fx, dg_test = ftol * dot9(grad, drt);
            for (int it = 0; it < 2000; ++it) {
                for (int i = 0; i < 9; ++i) x[i] = xp[i] + step * drt[i];
                fx = value_grad(P, v, x, grad);
                if (!(fx > fx_init + step * dg_test)) break;
                if (step < min_step || step > max_step) break;  // the reference throws here
                step *= 0.5;
            }
        }
        xnorm = sqrt(dot9(x, x));
        gnorm = sqrt(dot9(grad, grad));
        if (gnorm <= epsilon * fmax(xnorm, 1.0)) break;
        if (fabs(fx_prev - fx) < delta) break;
        fx_prev = fx;
        if (k >= 100) break;
        for (int i = 0; i < 9; ++i) {
            s[end][i] = x[i] - xp[i];
            y[end][i] = grad[i] - gradp[i];
        }
        const double ysv = dot9(y[end], s[end]), yy = dot9(y[end], y[end]);
        ys[end] = ysv;
        for (int i = 0; i < 9; ++i) drt[i] = -grad[i];
        const int bound = k < M ? k : M;
        end = (end + 1) % M;
        int j = end;
        for (int i = 0; i < bound; ++i) {
            j = (j + M - 1) % M;
            alpha[j] = dot9(s[j], drt) / ys[j];
            for (int q = 0; q < 9; ++q) drt[q] -= alpha[j] * y[j][q];
        }
        for (int q = 0; q < 9; ++q) drt[q] *= (ysv / yy);
        for (int i = 0; i < bound; ++i) {
            const double beta = dot9(y[j], drt) / ys[j];
            for (int q = 0; q < 9; ++q) drt[q] += (alpha[j] - beta) * s[j][q];
            j = (j + 1) % M;
        }
        step = 1.0;
        ++k;
    }
    for (int i = 0; i < 9; ++i) z[i] = x[i];
}

// HyperElasticTet::get_gradient: g = vol * dPsi(F)
AAADMM_HD void tet_grad_hyper(const HyperParams &P, const double *F, double *g) {
    hyper::u_gradient(P, F, g);
    for (int i = 0; i < 9; ++i) g[i] = P.vol * g[i];
}

}  // namespace aaadmm
