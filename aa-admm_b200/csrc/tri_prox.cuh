// TriEnergyTerm::prox on one column-major 3x2 block (xzu/src/TriEnergyTerm.cpp:77-107, hard/src/TriEnergyTerm.cpp:74-105),
// shared by the unit batch kernel (extra_terms.cu) and the fused local step of the hard_zxu loop (tri_kernels.cu).
// Include only from translation units compiled with -fmad=false.
#pragma once
#include <cfloat>

namespace aaadmm {

// Thin SVD of a 3x2 block by one one-sided (Hestenes) Jacobi rotation of its two columns: F J = [g1 g2]
// with g1 . g2 = 0, sigma_i = |g_i|, u_i = g_i / sigma_i, V = J. The reference takes U, sigma, V from
// Eigen's JacobiSVD with a full-pivoting QR preconditioner; U f(Sigma) V^T does not depend on how the
// factors were obtained as long as the singular values are distinct from zero.
struct Svd32 {
    double u1[3], u2[3], s1, s2, v11, v12, v21, v22;  // V = [v11 v12; v21 v22], columns v1, v2
};

__device__ __forceinline__ Svd32 svd32(const double *F) {
    const double *f1 = F, *f2 = F + 3;
    const double a = f1[0] * f1[0] + f1[1] * f1[1] + f1[2] * f1[2];
    const double c = f2[0] * f2[0] + f2[1] * f2[1] + f2[2] * f2[2];
    const double b = f1[0] * f2[0] + f1[1] * f2[1] + f1[2] * f2[2];
    double cs = 1.0, sn = 0.0;
    if (fabs(b) > DBL_MIN) {
        const double zeta = (c - a) / (2.0 * b);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        cs = 1.0 / sqrt(1.0 + t * t);
        sn = cs * t;
    }
    Svd32 r;
    double g1[3], g2[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        g1[k] = cs * f1[k] - sn * f2[k];
        g2[k] = sn * f1[k] + cs * f2[k];
    }
    r.v11 = cs;
    r.v21 = -sn;
    r.v12 = sn;
    r.v22 = cs;
    r.s1 = sqrt(g1[0] * g1[0] + g1[1] * g1[1] + g1[2] * g1[2]);
    r.s2 = sqrt(g2[0] * g2[0] + g2[1] * g2[1] + g2[2] * g2[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        r.u1[k] = r.s1 > 0.0 ? g1[k] / r.s1 : (k == 0 ? 1.0 : 0.0);
        r.u2[k] = r.s2 > 0.0 ? g2[k] / r.s2 : 0.0;
    }
    if (!(r.s2 > 0.0)) {  // rank deficient: any unit vector orthogonal to u1
        const int m = fabs(r.u1[0]) <= fabs(r.u1[1]) ? (fabs(r.u1[0]) <= fabs(r.u1[2]) ? 0 : 2) : (fabs(r.u1[1]) <= fabs(r.u1[2]) ? 1 : 2);
        double e[3] = {0.0, 0.0, 0.0};
        e[m] = 1.0;
        const double d = r.u1[m];
        double n2 = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            r.u2[k] = e[k] - d * r.u1[k];
            n2 += r.u2[k] * r.u2[k];
        }
        n2 = sqrt(n2);
#pragma unroll
        for (int k = 0; k < 3; ++k) r.u2[k] /= n2;
    }
    return r;
}

// variant 0 = xzu, 1 = hard_zxu; lmin / lmax = Lame::limit_min / limit_max of the term.
__device__ __forceinline__ void tri_prox_block(int variant, const double (&F)[6], double lmin, double lmax, double (&out)[6]) {
    const Svd32 s = svd32(F);
    const bool check = lmin > 0.0 || lmax < 99.0;
    if (variant == 0) {
        // p = U [I;0] V^T, z = (p + z) / 2, then the column norms are clamped into [lmin, lmax]
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            out[k] = 0.5 * ((s.u1[k] * s.v11 + s.u2[k] * s.v12) + F[k]);
            out[3 + k] = 0.5 * ((s.u1[k] * s.v21 + s.u2[k] * s.v22) + F[3 + k]);
        }
        if (check) {
            const double l0 = sqrt(out[0] * out[0] + out[1] * out[1] + out[2] * out[2]);
            const double l1 = sqrt(out[3] * out[3] + out[4] * out[4] + out[5] * out[5]);
            // the reference applies the four tests one after the other with the ORIGINAL norms
            double f0 = 1.0, f1 = 1.0;
            if (l0 < lmin) f0 *= lmin / l0;
            if (l1 < lmin) f1 *= lmin / l1;
            if (l0 > lmax) f0 *= lmax / l0;
            if (l1 > lmax) f1 *= lmax / l1;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                out[k] *= f0;
                out[3 + k] *= f1;
            }
        }
    } else {
        // Sigma' = ((1, 1) + sigma) / 2, clamped into [lmin, lmax]; z = U Sigma' V^T
        double a = (1.0 + s.s1) / 2.0, b = (1.0 + s.s2) / 2.0;
        if (check) {
            const double l0 = a, l1 = b;
            if (l0 < lmin) a = lmin;
            if (l1 < lmin) b = lmin;
            if (l0 > lmax) a = lmax;
            if (l1 > lmax) b = lmax;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            out[k] = a * s.u1[k] * s.v11 + b * s.u2[k] * s.v12;
            out[3 + k] = a * s.u1[k] * s.v21 + b * s.u2[k] * s.v22;
        }
    }
}

}  // namespace aaadmm
