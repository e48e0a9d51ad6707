// Per-element FP64 math of the tet local step (one thread per tet).
//
// Restates, operation by operation, what the reference's LINEAR tet does through Eigen 3.3.4:
//   TetEnergyTerm::prox          xzu/src/TetEnergyTerm.cpp:101-123 (hard: :74-96)
//   TetEnergyTerm::get_gradient  xzu/src/TetEnergyTerm.cpp:156-165
// which call JacobiSVD<Matrix3d> (two-sided Jacobi, no QR preconditioner for square input):
//   Eigen/src/SVD/JacobiSVD.h:660-770, Eigen/src/misc/RealSvd2x2.h:17-52,
//   Eigen/src/Jacobi/Jacobi.h:80-113 (makeJacobi), :300-441 (apply_rotation_in_the_plane).
// Matrices are column-major 3x3: a[c*3+r].
#pragma once
#include <cfloat>
#include <cmath>

#ifdef __CUDACC__
#define AAADMM_HD __host__ __device__ __forceinline__
#else
#define AAADMM_HD inline
#endif

namespace aaadmm {

struct Rot2 {
    double c, s;
};

// rows p,q of a: row_p' = c row_p + s row_q ; row_q' = -s row_p + c row_q
AAADMM_HD void rot_rows(double *a, int p, int q, Rot2 j) {
    if (j.c == 1.0 && j.s == 0.0) return;
#pragma unroll
    for (int col = 0; col < 3; ++col) {
        const double x = a[col * 3 + p], y = a[col * 3 + q];
        a[col * 3 + p] = j.c * x + j.s * y;
        a[col * 3 + q] = j.c * y - j.s * x;
    }
}
// columns p,q of a: col_p' = c col_p + s col_q ; col_q' = -s col_p + c col_q
AAADMM_HD void rot_cols(double *a, int p, int q, Rot2 j) {
    if (j.c == 1.0 && j.s == 0.0) return;
#pragma unroll
    for (int row = 0; row < 3; ++row) {
        const double x = a[p * 3 + row], y = a[q * 3 + row];
        a[p * 3 + row] = j.c * x + j.s * y;
        a[q * 3 + row] = j.c * y - j.s * x;
    }
}

// JacobiRotation::makeJacobi(x, y, z) for the symmetric 2x2 [[x,y],[y,z]].
AAADMM_HD Rot2 make_jacobi(double x, double y, double z) {
    Rot2 r;
    const double deno = 2.0 * fabs(y);
    if (deno < DBL_MIN) {
        r.c = 1.0;
        r.s = 0.0;
        return r;
    }
    const double tau = (x - z) / deno;
    const double w = sqrt(tau * tau + 1.0);
    const double t = (tau > 0.0) ? 1.0 / (tau + w) : 1.0 / (tau - w);
    const double sign_t = t > 0.0 ? 1.0 : -1.0;
    const double n = 1.0 / sqrt(t * t + 1.0);
    // y / |y| of the reference is exactly +-1 for the finite non-zero y that reach this line: no division needed
    r.s = -sign_t * (y > 0.0 ? 1.0 : -1.0) * fabs(t) * n;
    r.c = n;
    return r;
}

// real_2x2_jacobi_svd on the (p,q) block of w.
AAADMM_HD void svd2x2(const double *w, int p, int q, Rot2 &jl, Rot2 &jr) {
    double m00 = w[p * 3 + p], m01 = w[q * 3 + p], m10 = w[p * 3 + q], m11 = w[q * 3 + q];
    Rot2 rot1;
    const double t = m00 + m11;
    const double d = m10 - m01;
    if (fabs(d) < DBL_MIN) {
        rot1.s = 0.0;
        rot1.c = 1.0;
    } else {
        const double u = t / d;
        const double tmp = sqrt(1.0 + u * u);
        rot1.s = 1.0 / tmp;
        rot1.c = u / tmp;
    }
    if (!(rot1.c == 1.0 && rot1.s == 0.0)) {
        const double a00 = rot1.c * m00 + rot1.s * m10, a01 = rot1.c * m01 + rot1.s * m11;
        const double a10 = rot1.c * m10 - rot1.s * m00, a11 = rot1.c * m11 - rot1.s * m01;
        m00 = a00;
        m01 = a01;
        m10 = a10;
        m11 = a11;
    }
    (void)m10;
    jr = make_jacobi(m00, m01, m11);
    // j_left = rot1 * j_right.transpose()
    const double tc = jr.c, ts = -jr.s;
    jl.c = rot1.c * tc - rot1.s * ts;
    jl.s = rot1.c * ts + rot1.s * tc;
}

// JacobiSVD<Matrix3d>(F, ComputeFullU|ComputeFullV): F = U diag(sv) V^T, sv sorted descending.
AAADMM_HD void jacobi_svd3(const double *F, double *U, double *sv, double *V) {
    double scale = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) scale = fmax(scale, fabs(F[k]));
    if (scale == 0.0) scale = 1.0;
    double w[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        w[k] = F[k] / scale;
        U[k] = (k % 4 == 0) ? 1.0 : 0.0;
        V[k] = (k % 4 == 0) ? 1.0 : 0.0;
    }
    const double precision = 2.0 * DBL_EPSILON;
    double maxDiag = fmax(fabs(w[0]), fmax(fabs(w[4]), fabs(w[8])));
    bool finished = false;
    int guard = 0;
    while (!finished && guard++ < 64) {
        finished = true;
#pragma unroll
        for (int p = 1; p < 3; ++p) {
#pragma unroll
            for (int q = 0; q < p; ++q) {
                const double threshold = fmax(DBL_MIN, precision * maxDiag);
                if (fabs(w[q * 3 + p]) > threshold || fabs(w[p * 3 + q]) > threshold) {
                    finished = false;
                    Rot2 jl, jr;
                    svd2x2(w, p, q, jl, jr);
                    rot_rows(w, p, q, jl);  // applyOnTheLeft(p,q,j_left)
                    rot_cols(U, p, q, jl);  // U.applyOnTheRight(p,q,j_left.transpose())
                    Rot2 jrt;
                    jrt.c = jr.c;
                    jrt.s = -jr.s;
                    rot_cols(w, p, q, jrt);  // applyOnTheRight(p,q,j_right)
                    rot_cols(V, p, q, jrt);
                    maxDiag = fmax(maxDiag, fmax(fabs(w[p * 3 + p]), fabs(w[q * 3 + q])));
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double a = w[i * 3 + i];
        sv[i] = fabs(a);
        if (a < 0.0) {
            U[i * 3 + 0] = -U[i * 3 + 0];
            U[i * 3 + 1] = -U[i * 3 + 1];
            U[i * 3 + 2] = -U[i * 3 + 2];
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) sv[i] *= scale;
    // sort descending (first maximum wins), swapping columns of U and V
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int pos = i;
        double mx = sv[i];
#pragma unroll
        for (int k = i + 1; k < 3; ++k)
            if (sv[k] > mx) {
                mx = sv[k];
                pos = k;
            }
        if (mx == 0.0) break;
        if (pos != i) {
            double tsv = sv[i];
            sv[i] = sv[pos];
            sv[pos] = tsv;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                double tu = U[i * 3 + r];
                U[i * 3 + r] = U[pos * 3 + r];
                U[pos * 3 + r] = tu;
                double tv = V[i * 3 + r];
                V[i * 3 + r] = V[pos * 3 + r];
                V[pos * 3 + r] = tv;
            }
        }
    }
}

// Matrix3d::determinant() (Eigen/src/LU/Determinant.h bruteforce_det3_helper).
AAADMM_HD double det3(const double *m) {
    const double h0 = m[0 * 3 + 0] * (m[1 * 3 + 1] * m[2 * 3 + 2] - m[2 * 3 + 1] * m[1 * 3 + 2]);
    const double h1 = m[1 * 3 + 0] * (m[0 * 3 + 1] * m[2 * 3 + 2] - m[2 * 3 + 1] * m[0 * 3 + 2]);
    const double h2 = m[2 * 3 + 0] * (m[0 * 3 + 1] * m[1 * 3 + 2] - m[1 * 3 + 1] * m[0 * 3 + 2]);
    return h0 - h1 + h2;
}

// R = U diag(1,1,s3) V^T, summed over k in order 0,1,2.
AAADMM_HD void usvt(const double *U, double s3, const double *V, double *R) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i)
            R[j * 3 + i] = U[0 * 3 + i] * V[0 * 3 + j] + U[1 * 3 + i] * V[1 * 3 + j] + (U[2 * 3 + i] * s3) * V[2 * 3 + j];
}

// TetEnergyTerm::prox: z <- 0.5 (U diag(1,1,+-1) V^T + z), flip iff det(z) < 1e-16.
AAADMM_HD void tet_prox_linear(double *z) {
    double U[9], V[9], sv[3], R[9];
    jacobi_svd3(z, U, sv, V);
    const double s3 = (det3(z) < 1e-16) ? -1.0 : 1.0;
    usvt(U, s3, V, R);
#pragma unroll
    for (int k = 0; k < 9; ++k) z[k] = 0.5 * (R[k] + z[k]);
}

// TetEnergyTerm::get_gradient: g = (K vol) (F - U V^T)   (no inversion flip)
AAADMM_HD void tet_grad_linear(const double *F, double kvol, double *g) {
    double U[9], V[9], sv[3], R[9];
    jacobi_svd3(F, U, sv, V);
    usvt(U, 1.0, V, R);
#pragma unroll
    for (int k = 0; k < 9; ++k) g[k] = kvol * (F[k] - R[k]);
}

}  // namespace aaadmm
