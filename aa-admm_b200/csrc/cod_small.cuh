// m x m rank-revealing least-squares solve used by Anderson mixing (one thread, m <= AA_MAX_M).
//
// Restates Eigen 3.3.4's CompleteOrthogonalDecomposition<MatrixXd>::compute + solve as
// AndersonAcceleration::compute_impl uses it (hard/src/AndersonAcceleration.h:193-196):
//   column-pivoted Householder QR with LAPACK-style norm downdating
//       Eigen/src/QR/ColPivHouseholderQR.h:480-577
//   rank = #{ i < nonzero_pivots : |R_ii| > eps * m * max_j |R_jj| }      :378-385 / :255-263
//   when rank < m, Householder reflectors Z from the right reduce [R11 R12] to [T11 0] and the
//   minimum-norm solution is returned  Eigen/src/QR/CompleteOrthogonalDecomposition.h:409-524
//   Householder vectors: Eigen/src/Householder/Householder.h:65-96.
// Summations run sequentially (Eigen's packet order differs in the last bits only).
#pragma once
#include <cfloat>
#include <cmath>

#ifndef AAADMM_HD
#ifdef __CUDACC__
#define AAADMM_HD __host__ __device__ __forceinline__
#else
#define AAADMM_HD inline
#endif
#endif

#ifndef AA_MAX_M
#define AA_MAX_M 16
#endif

namespace aaadmm {

// makeHouseholder on x[0..n-1] (stride inc): on exit x[0] unchanged by this routine
// (caller stores beta), x[1..] = essential part. Returns tau, beta.
AAADMM_HD void make_householder(double *x, int n, int inc, double &tau, double &beta) {
    double tailSq = 0.0;
    for (int i = 1; i < n; ++i) tailSq += x[i * inc] * x[i * inc];
    const double c0 = x[0];
    if (tailSq <= DBL_MIN) {
        tau = 0.0;
        beta = c0;
        for (int i = 1; i < n; ++i) x[i * inc] = 0.0;
    } else {
        beta = sqrt(c0 * c0 + tailSq);
        if (c0 >= 0.0) beta = -beta;
        const double den = c0 - beta;
        for (int i = 1; i < n; ++i) x[i * inc] = x[i * inc] / den;
        tau = (beta - c0) / beta;
    }
}

// Column-pivoted Householder QR of the m x m matrix A (column-major, overwritten by R and the essential parts of the
// reflectors) with LAPACK-style norm downdating: Eigen/src/QR/ColPivHouseholderQR.h:480-577. One thread.
AAADMM_HD void cod_qr(double *A, int m, double *hc, int *trans, int &nonzero_pivots, double &maxpivot) {
    double normU[AA_MAX_M], normD[AA_MAX_M];
    const int rows = m, cols = m, size = m;
#define A_(r, cc) A[(cc) * m + (r)]
    double maxnorm = 0.0;
    for (int k = 0; k < cols; ++k) {
        double s = 0.0;
        for (int i = 0; i < rows; ++i) s += A_(i, k) * A_(i, k);
        normD[k] = sqrt(s);
        normU[k] = normD[k];
        if (normU[k] > maxnorm) maxnorm = normU[k];
    }
    const double th0 = maxnorm * DBL_EPSILON;
    const double threshold_helper = (th0 * th0) / (double)rows;
    const double downdate_threshold = sqrt(DBL_EPSILON);
    nonzero_pivots = size;
    maxpivot = 0.0;
    for (int k = 0; k < size; ++k) {
        int big = k;
        double bn = normU[k];
        for (int j = k + 1; j < cols; ++j)
            if (normU[j] > bn) {
                bn = normU[j];
                big = j;
            }
        const double big_sq = bn * bn;
        if (nonzero_pivots == size && big_sq < threshold_helper * (double)(rows - k)) nonzero_pivots = k;
        trans[k] = big;
        if (k != big) {
            for (int i = 0; i < rows; ++i) {
                double t = A_(i, k);
                A_(i, k) = A_(i, big);
                A_(i, big) = t;
            }
            double t = normU[k];
            normU[k] = normU[big];
            normU[big] = t;
            t = normD[k];
            normD[k] = normD[big];
            normD[big] = t;
        }
        double beta;
        make_householder(&A_(k, k), rows - k, 1, hc[k], beta);
        A_(k, k) = beta;
        if (fabs(beta) > maxpivot) maxpivot = fabs(beta);
        // apply H_k = I - tau v v^T (v = [1; essential]) to the trailing columns
        const int nr = rows - k;
        for (int j = k + 1; j < cols; ++j) {
            if (nr == 1) {
                A_(k, j) *= (1.0 - hc[k]);
            } else if (hc[k] != 0.0) {
                double tmp = 0.0;
                for (int i = k + 1; i < rows; ++i) tmp += A_(i, k) * A_(i, j);
                tmp += A_(k, j);
                A_(k, j) -= hc[k] * tmp;
                for (int i = k + 1; i < rows; ++i) A_(i, j) -= hc[k] * A_(i, k) * tmp;
            }
        }
        for (int j = k + 1; j < cols; ++j) {
            if (normU[j] != 0.0) {
                double temp = fabs(A_(k, j)) / normU[j];
                temp = (1.0 + temp) * (1.0 - temp);
                temp = temp < 0.0 ? 0.0 : temp;
                const double r = normU[j] / normD[j];
                const double temp2 = temp * (r * r);
                if (temp2 <= downdate_threshold) {
                    double s = 0.0;
                    for (int i = k + 1; i < rows; ++i) s += A_(i, j) * A_(i, j);
                    normD[j] = sqrt(s);
                    normU[j] = normD[j];
                } else {
                    normU[j] *= sqrt(temp);
                }
            }
        }
    }
#undef A_
}

// Everything after the QR: rank decision, the Z reflectors of the complete orthogonal decomposition when rank < m,
// c = Q^T b, back substitution, Z^* y and the column permutation. One thread. Returns the rank.
AAADMM_HD int cod_finish(double *A, int m, const double *hc, const int *trans, int nonzero_pivots, double maxpivot,
                         const double *b, double *x) {
    double zc[AA_MAX_M], c[AA_MAX_M], y[AA_MAX_M];
    const int rows = m, cols = m, size = m;
#define A_(r, cc) A[(cc) * m + (r)]
    // rank
    const double premult = fabs(maxpivot) * (DBL_EPSILON * (double)size);
    int rank = 0;
    for (int i = 0; i < nonzero_pivots; ++i) rank += (fabs(A_(i, i)) > premult);
    if (rank == 0) {
        for (int i = 0; i < m; ++i) x[i] = 0.0;
        return 0;
    }
    // COD: zero out R12 with reflectors from the right
    if (rank < cols) {
        const int nt = cols - rank + 1;
        for (int k = rank - 1; k >= 0; --k) {
            if (k != rank - 1)
                for (int i = 0; i <= k; ++i) {
                    double t = A_(i, k);
                    A_(i, k) = A_(i, rank - 1);
                    A_(i, rank - 1) = t;
                }
            double beta;
            make_householder(&A_(k, rank - 1), nt, m, zc[k], beta);
            A_(k, rank - 1) = beta;
            if (k > 0 && zc[k] != 0.0) {
                // applyHouseholderOnTheRight on rows 0..k-1, columns rank-1..cols-1
                for (int i = 0; i < k; ++i) {
                    if (nt == 1) {
                        A_(i, rank - 1) *= (1.0 - zc[k]);
                    } else {
                        double tmp = 0.0;
                        for (int j = 1; j < nt; ++j) tmp += A_(i, rank - 1 + j) * A_(k, rank - 1 + j);
                        tmp += A_(i, rank - 1);
                        A_(i, rank - 1) -= zc[k] * tmp;
                        for (int j = 1; j < nt; ++j) A_(i, rank - 1 + j) -= zc[k] * tmp * A_(k, rank - 1 + j);
                    }
                }
            }
            if (k != rank - 1)
                for (int i = 0; i <= k; ++i) {
                    double t = A_(i, k);
                    A_(i, k) = A_(i, rank - 1);
                    A_(i, rank - 1) = t;
                }
        }
    }
    // c = Q^T b : apply H_0, H_1, ..., H_{rank-1} in order
    for (int i = 0; i < m; ++i) c[i] = b[i];
    for (int k = 0; k < rank; ++k) {
        const int nr = rows - k;
        if (nr == 1) {
            c[k] *= (1.0 - hc[k]);
        } else if (hc[k] != 0.0) {
            double tmp = 0.0;
            for (int i = k + 1; i < rows; ++i) tmp += A_(i, k) * c[i];
            tmp += c[k];
            c[k] -= hc[k] * tmp;
            for (int i = k + 1; i < rows; ++i) c[i] -= hc[k] * A_(i, k) * tmp;
        }
    }
    // solve T z = c(0:rank): upper-triangular back substitution
    for (int i = rank - 1; i >= 0; --i) {
        double s = c[i];
        for (int j = i + 1; j < rank; ++j) s -= A_(i, j) * y[j];
        y[i] = s / A_(i, i);
    }
    for (int i = rank; i < m; ++i) y[i] = 0.0;
    if (rank < cols) {
        // y <- Z^* y
        const int nt = cols - rank + 1;
        for (int k = 0; k < rank; ++k) {
            if (k != rank - 1) {
                double t = y[k];
                y[k] = y[rank - 1];
                y[rank - 1] = t;
            }
            if (nt > 1 && zc[k] != 0.0) {
                double tmp = 0.0;
                for (int j = 1; j < nt; ++j) tmp += A_(k, rank - 1 + j) * y[rank - 1 + j];
                tmp += y[rank - 1];
                y[rank - 1] -= zc[k] * tmp;
                for (int j = 1; j < nt; ++j) y[rank - 1 + j] -= zc[k] * A_(k, rank - 1 + j) * tmp;
            }
            if (k != rank - 1) {
                double t = y[k];
                y[k] = y[rank - 1];
                y[rank - 1] = t;
            }
        }
    }
    // x = P y with P built from the transpositions (applyTranspositionOnTheRight in order)
    int pidx[AA_MAX_M];
    for (int i = 0; i < m; ++i) pidx[i] = i;
    for (int k = 0; k < size; ++k) {
        int t = pidx[k];
        pidx[k] = pidx[trans[k]];
        pidx[trans[k]] = t;
    }
    for (int i = 0; i < m; ++i) x[pidx[i]] = y[i];
#undef A_
    return rank;
}

// Solves min ||A x - b|| (min-norm x when rank deficient). A is m x m column-major (destroyed),
// b length m, x length m. Returns the detected rank.
AAADMM_HD int cod_solve(double *A, int m, const double *b, double *x) {
    double hc[AA_MAX_M];
    int trans[AA_MAX_M], nonzero_pivots;
    double maxpivot;
    cod_qr(A, m, hc, trans, nonzero_pivots, maxpivot);
    return cod_finish(A, m, hc, trans, nonzero_pivots, maxpivot, b, x);
}

#ifdef __CUDACC__
// Warp-cooperative cod_qr: exactly the operations of cod_qr on the same operands and in the same order per result,
// with the independent ones on different lanes. cod_qr in one thread is a chain of about 60 FP64 division /
// square-root sequences (35 dependent instructions each) plus the dot products: 22 us for m = 5 while every other SM
// waits. Per elimination step the column norms' downdates, the reflector applications (one lane per trailing column)
// and the divisions of the Householder vector (one lane per row, the tau division on one more lane) are independent;
// quantities every lane needs (pivot, tail sum, beta) are recomputed redundantly on all lanes instead of broadcast.
// A, hc, trans, w.* live in shared memory; all 32 lanes of ONE converged warp call this. m <= AA_MAX_M <= 32.
struct CodWarpWork {
    double normU[AA_MAX_M], normD[AA_MAX_M];
};
__device__ __forceinline__ void cod_qr_warp(double *A, int m, double *hc, int *trans, int &nonzero_pivots, double &maxpivot,
                                            CodWarpWork &w) {
    const int lane = threadIdx.x & 31;
#define A_(r, cc) A[(cc) * m + (r)]
    if (lane < m) {
        double s = 0.0;
        for (int i = 0; i < m; ++i) s += A_(i, lane) * A_(i, lane);
        w.normD[lane] = sqrt(s);
        w.normU[lane] = w.normD[lane];
    }
    __syncwarp();
    double maxnorm = 0.0;
    for (int k = 0; k < m; ++k)
        if (w.normU[k] > maxnorm) maxnorm = w.normU[k];
    const double th0 = maxnorm * DBL_EPSILON;
    const double threshold_helper = (th0 * th0) / (double)m;
    const double downdate_threshold = sqrt(DBL_EPSILON);
    nonzero_pivots = m;
    maxpivot = 0.0;
    for (int k = 0; k < m; ++k) {
        // pivot column: every lane, from the shared norms (uniform result)
        int big = k;
        double bn = w.normU[k];
        for (int j = k + 1; j < m; ++j)
            if (w.normU[j] > bn) {
                bn = w.normU[j];
                big = j;
            }
        const double big_sq = bn * bn;
        if (nonzero_pivots == m && big_sq < threshold_helper * (double)(m - k)) nonzero_pivots = k;
        __syncwarp();  // all lanes have read the norms
        if (lane == 0) trans[k] = big;
        if (k != big) {
            if (lane < m) {
                const double t = A_(lane, k);
                A_(lane, k) = A_(lane, big);
                A_(lane, big) = t;
            }
            if (lane == 0) {
                double t = w.normU[k];
                w.normU[k] = w.normU[big];
                w.normU[big] = t;
                t = w.normD[k];
                w.normD[k] = w.normD[big];
                w.normD[big] = t;
            }
        }
        __syncwarp();
        // make_householder on A(k:m-1, k): tail sum and beta on every lane, one division per lane
        double tailSq = 0.0;
        for (int i = k + 1; i < m; ++i) tailSq += A_(i, k) * A_(i, k);
        const double c0 = A_(k, k);
        double beta, tau;
        __syncwarp();  // column k has been read by everybody before it is scaled
        if (tailSq <= DBL_MIN) {
            tau = 0.0;
            beta = c0;
            if (lane > k && lane < m) A_(lane, k) = 0.0;
        } else {
            beta = sqrt(c0 * c0 + tailSq);
            if (c0 >= 0.0) beta = -beta;
            const double den = c0 - beta;
            // lanes k+1..m-1: essential part x_i / den; all other lanes: tau = (beta - c0) / beta (same instruction stream)
            const bool mine = lane > k && lane < m;
            const double num = mine ? A_(lane, k) : (beta - c0);
            const double q = num / (mine ? den : beta);
            if (mine) A_(lane, k) = q;
            tau = __shfl_sync(0xffffffffu, q, 0);  // lane 0 is never one of the row lanes (0 > k is false)
        }
        if (lane == 0) {
            hc[k] = tau;
            A_(k, k) = beta;
        }
        if (fabs(beta) > maxpivot) maxpivot = fabs(beta);
        __syncwarp();
        // reflector on the trailing columns and their norm downdates: one lane per column j
        const int j = k + 1 + lane;
        if (j < m) {
            if (m - k == 1) {
                A_(k, j) *= (1.0 - tau);
            } else if (tau != 0.0) {
                double tmp = 0.0;
                for (int i = k + 1; i < m; ++i) tmp += A_(i, k) * A_(i, j);
                tmp += A_(k, j);
                A_(k, j) -= tau * tmp;
                for (int i = k + 1; i < m; ++i) A_(i, j) -= tau * A_(i, k) * tmp;
            }
            if (w.normU[j] != 0.0) {
                double temp = fabs(A_(k, j)) / w.normU[j];
                temp = (1.0 + temp) * (1.0 - temp);
                temp = temp < 0.0 ? 0.0 : temp;
                const double r = w.normU[j] / w.normD[j];
                const double temp2 = temp * (r * r);
                if (temp2 <= downdate_threshold) {
                    double s = 0.0;
                    for (int i = k + 1; i < m; ++i) s += A_(i, j) * A_(i, j);
                    w.normD[j] = sqrt(s);
                    w.normU[j] = w.normD[j];
                } else {
                    w.normU[j] *= sqrt(temp);
                }
            }
        }
        __syncwarp();
    }
#undef A_
}
#endif  // __CUDACC__

}  // namespace aaadmm
