/* TEST INFRASTRUCTURE ONLY - CPU restatement ("port") of the reference algorithm for the AA-ADMM
 * hot path, in plain C99.  Not part of the product and never linked or called by it; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load liboracle_port.so.
 *
 * Pinned against the unmodified reference compiled into oracle/_ref (tests/test_oracle_*.py) and
 * against the golden trajectories under tests/golden/ that oracle/_ref produced
 * (tests/golden/make_golden.py).
 *
 * What is restated, with the reference lines it follows:
 *   svd3 / prox / gradient   Eigen 3.3.4 JacobiSVD<Matrix3d> (Eigen/src/SVD/JacobiSVD.h:660-770,
 *                            misc/RealSvd2x2.h:17-52, Jacobi/Jacobi.h:80-113) as used by
 *                            TetEnergyTerm::prox / get_gradient (xzu/src/TetEnergyTerm.cpp:101-123,156-165)
 *   cod_solve                Eigen CompleteOrthogonalDecomposition solve (QR/ColPivHouseholderQR.h:480-577,
 *                            QR/CompleteOrthogonalDecomposition.h:409-524)
 *   anderson                 AndersonAcceleration::compute_impl, variant H (hard/src/AndersonAcceleration.h:154-211);
 *                            variant X (xzu/src/AndersonAcceleration.h:138-200) is the same with total == effective
 *   scene setup              TetEnergyTerm ctor + get_reduction (xzu/src/TetEnergyTerm.cpp:32-88),
 *                            Solver::initialize (hard/src/Solver.cpp:361-491): D, W, A = M + rho dt^2 D^T W^2 D
 *   ldl                      simplicial up-looking LDL^T (the algorithm of Eigen::SimplicialLDLT,
 *                            Eigen/src/SparseCholesky/SimplicialCholesky_impl.h), natural ordering
 *   step_hard                Solver::step, hard_zxu ordering (hard/src/Solver.cpp:34-234)
 *   step_xzu                 Solver::step, xzu ordering (xzu/src/Solver.cpp:34-263)
 * The port factors the scalar system Ahat (A = Ahat (x) I3) in the mesh's natural order.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* 3x3 Jacobi SVD, column-major a[c*3+r]                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct { double c, s; } rot_t;

static void rot_rows(double *a, int p, int q, rot_t j) {
    if (j.c == 1.0 && j.s == 0.0) return;
    for (int col = 0; col < 3; ++col) {
        double x = a[col * 3 + p], y = a[col * 3 + q];
        a[col * 3 + p] = j.c * x + j.s * y;
        a[col * 3 + q] = j.c * y - j.s * x;
    }
}
static void rot_cols(double *a, int p, int q, rot_t j) {
    if (j.c == 1.0 && j.s == 0.0) return;
    for (int row = 0; row < 3; ++row) {
        double x = a[p * 3 + row], y = a[q * 3 + row];
        a[p * 3 + row] = j.c * x + j.s * y;
        a[q * 3 + row] = j.c * y - j.s * x;
    }
}
/* JacobiRotation::makeJacobi(x,y,z) */
static rot_t make_jacobi(double x, double y, double z) {
    rot_t r;
    double deno = 2.0 * fabs(y);
    if (deno < DBL_MIN) { r.c = 1.0; r.s = 0.0; return r; }
    double tau = (x - z) / deno;
    double w = sqrt(tau * tau + 1.0);
    double t = (tau > 0.0) ? 1.0 / (tau + w) : 1.0 / (tau - w);
    double sign_t = t > 0.0 ? 1.0 : -1.0;
    double n = 1.0 / sqrt(t * t + 1.0);
    r.s = -sign_t * (y / fabs(y)) * fabs(t) * n;
    r.c = n;
    return r;
}
/* real_2x2_jacobi_svd */
static void svd2(const double *w, int p, int q, rot_t *jl, rot_t *jr) {
    double m00 = w[p * 3 + p], m01 = w[q * 3 + p], m10 = w[p * 3 + q], m11 = w[q * 3 + q];
    rot_t r1;
    double t = m00 + m11, d = m10 - m01;
    if (fabs(d) < DBL_MIN) { r1.s = 0.0; r1.c = 1.0; }
    else { double u = t / d; double tmp = sqrt(1.0 + u * u); r1.s = 1.0 / tmp; r1.c = u / tmp; }
    if (!(r1.c == 1.0 && r1.s == 0.0)) {
        double a00 = r1.c * m00 + r1.s * m10, a01 = r1.c * m01 + r1.s * m11;
        double a11 = r1.c * m11 - r1.s * m01;
        m00 = a00; m01 = a01; m11 = a11;
    }
    *jr = make_jacobi(m00, m01, m11);
    jl->c = r1.c * jr->c - r1.s * (-jr->s);
    jl->s = r1.c * (-jr->s) + r1.s * jr->c;
}
static void svd3(const double *F, double *U, double *sv, double *V) {
    double scale = 0.0, w[9];
    for (int k = 0; k < 9; ++k) if (fabs(F[k]) > scale) scale = fabs(F[k]);
    if (scale == 0.0) scale = 1.0;
    for (int k = 0; k < 9; ++k) { w[k] = F[k] / scale; U[k] = V[k] = (k % 4 == 0) ? 1.0 : 0.0; }
    double maxd = fmax(fabs(w[0]), fmax(fabs(w[4]), fabs(w[8])));
    int finished = 0, guard = 0;
    while (!finished && guard++ < 64) {
        finished = 1;
        for (int p = 1; p < 3; ++p) for (int q = 0; q < p; ++q) {
            double th = fmax(DBL_MIN, 2.0 * DBL_EPSILON * maxd);
            if (fabs(w[q * 3 + p]) > th || fabs(w[p * 3 + q]) > th) {
                finished = 0;
                rot_t jl, jr, jrt;
                svd2(w, p, q, &jl, &jr);
                rot_rows(w, p, q, jl);
                rot_cols(U, p, q, jl);
                jrt.c = jr.c; jrt.s = -jr.s;
                rot_cols(w, p, q, jrt);
                rot_cols(V, p, q, jrt);
                maxd = fmax(maxd, fmax(fabs(w[p * 3 + p]), fabs(w[q * 3 + q])));
            }
        }
    }
    for (int i = 0; i < 3; ++i) {
        double a = w[i * 3 + i];
        sv[i] = fabs(a);
        if (a < 0.0) for (int r = 0; r < 3; ++r) U[i * 3 + r] = -U[i * 3 + r];
    }
    for (int i = 0; i < 3; ++i) sv[i] *= scale;
    for (int i = 0; i < 3; ++i) {
        int pos = i; double mx = sv[i];
        for (int k = i + 1; k < 3; ++k) if (sv[k] > mx) { mx = sv[k]; pos = k; }
        if (mx == 0.0) break;
        if (pos != i) {
            double t = sv[i]; sv[i] = sv[pos]; sv[pos] = t;
            for (int r = 0; r < 3; ++r) {
                t = U[i * 3 + r]; U[i * 3 + r] = U[pos * 3 + r]; U[pos * 3 + r] = t;
                t = V[i * 3 + r]; V[i * 3 + r] = V[pos * 3 + r]; V[pos * 3 + r] = t;
            }
        }
    }
}
static double det3(const double *m) {
    double h0 = m[0] * (m[4] * m[8] - m[7] * m[5]);
    double h1 = m[3] * (m[1] * m[8] - m[7] * m[2]);
    double h2 = m[6] * (m[1] * m[5] - m[4] * m[2]);
    return h0 - h1 + h2;
}
static void usvt(const double *U, double s3, const double *V, double *R) {
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i)
        R[j * 3 + i] = U[0 * 3 + i] * V[0 * 3 + j] + U[1 * 3 + i] * V[1 * 3 + j] + (U[2 * 3 + i] * s3) * V[2 * 3 + j];
}
/* TetEnergyTerm::prox */
static void tet_prox(double *z) {
    double U[9], V[9], sv[3], R[9];
    svd3(z, U, sv, V);
    usvt(U, det3(z) < 1e-16 ? -1.0 : 1.0, V, R);
    for (int k = 0; k < 9; ++k) z[k] = 0.5 * (R[k] + z[k]);
}
/* TetEnergyTerm::get_gradient */
static void tet_grad(const double *F, double kvol, double *g) {
    double U[9], V[9], sv[3], R[9];
    svd3(F, U, sv, V);
    usvt(U, 1.0, V, R);
    for (int k = 0; k < 9; ++k) g[k] = kvol * (F[k] - R[k]);
}
void port_tet_prox(double *z, int n) { for (int i = 0; i < n; ++i) tet_prox(z + 9 * i); }
void port_tet_fmuvt(const double *z, double *out, int n) { for (int i = 0; i < n; ++i) tet_grad(z + 9 * i, 1.0, out + 9 * i); }

/* ------------------------------------------------------------------------------------------ */
/* Eigen COD solve, m x m column-major                                                          */
/* ------------------------------------------------------------------------------------------ */
#define MAXM 32
static void householder(double *x, int n, int inc, double *tau, double *beta) {
    double tail = 0.0, c0 = x[0];
    for (int i = 1; i < n; ++i) tail += x[i * inc] * x[i * inc];
    if (tail <= DBL_MIN) { *tau = 0.0; *beta = c0; for (int i = 1; i < n; ++i) x[i * inc] = 0.0; }
    else {
        double b = sqrt(c0 * c0 + tail);
        if (c0 >= 0.0) b = -b;
        for (int i = 1; i < n; ++i) x[i * inc] /= (c0 - b);
        *tau = (b - c0) / b; *beta = b;
    }
}
int port_cod_solve(int m, const double *Min, const double *b, double *x) {
    double A[MAXM * MAXM], hc[MAXM], zc[MAXM], nu[MAXM], nd[MAXM], c[MAXM], y[MAXM];
    int tr[MAXM], pidx[MAXM];
#define AA(r, cc) A[(cc) * m + (r)]
    memcpy(A, Min, sizeof(double) * m * m);
    double maxn = 0.0;
    for (int k = 0; k < m; ++k) {
        double s = 0.0;
        for (int i = 0; i < m; ++i) s += AA(i, k) * AA(i, k);
        nd[k] = nu[k] = sqrt(s);
        if (nu[k] > maxn) maxn = nu[k];
    }
    double th = (maxn * DBL_EPSILON) * (maxn * DBL_EPSILON) / (double)m, down = sqrt(DBL_EPSILON);
    int nz = m; double maxpiv = 0.0;
    for (int k = 0; k < m; ++k) {
        int big = k; double bn = nu[k];
        for (int j = k + 1; j < m; ++j) if (nu[j] > bn) { bn = nu[j]; big = j; }
        if (nz == m && bn * bn < th * (double)(m - k)) nz = k;
        tr[k] = big;
        if (k != big) {
            for (int i = 0; i < m; ++i) { double t = AA(i, k); AA(i, k) = AA(i, big); AA(i, big) = t; }
            double t = nu[k]; nu[k] = nu[big]; nu[big] = t;
            t = nd[k]; nd[k] = nd[big]; nd[big] = t;
        }
        double beta;
        householder(&AA(k, k), m - k, 1, &hc[k], &beta);
        AA(k, k) = beta;
        if (fabs(beta) > maxpiv) maxpiv = fabs(beta);
        for (int j = k + 1; j < m; ++j) {
            if (m - k == 1) AA(k, j) *= (1.0 - hc[k]);
            else if (hc[k] != 0.0) {
                double tmp = 0.0;
                for (int i = k + 1; i < m; ++i) tmp += AA(i, k) * AA(i, j);
                tmp += AA(k, j);
                AA(k, j) -= hc[k] * tmp;
                for (int i = k + 1; i < m; ++i) AA(i, j) -= hc[k] * AA(i, k) * tmp;
            }
        }
        for (int j = k + 1; j < m; ++j) if (nu[j] != 0.0) {
            double t = fabs(AA(k, j)) / nu[j];
            t = (1.0 + t) * (1.0 - t); if (t < 0.0) t = 0.0;
            double r = nu[j] / nd[j];
            if (t * r * r <= down) {
                double s = 0.0;
                for (int i = k + 1; i < m; ++i) s += AA(i, j) * AA(i, j);
                nd[j] = nu[j] = sqrt(s);
            } else nu[j] *= sqrt(t);
        }
    }
    double pre = fabs(maxpiv) * (DBL_EPSILON * (double)m);
    int rank = 0;
    for (int i = 0; i < nz; ++i) rank += fabs(AA(i, i)) > pre;
    if (rank == 0) { for (int i = 0; i < m; ++i) x[i] = 0.0; return 0; }
    int nt = m - rank + 1;
    if (rank < m) for (int k = rank - 1; k >= 0; --k) {
        if (k != rank - 1) for (int i = 0; i <= k; ++i) { double t = AA(i, k); AA(i, k) = AA(i, rank - 1); AA(i, rank - 1) = t; }
        double beta;
        householder(&AA(k, rank - 1), nt, m, &zc[k], &beta);
        AA(k, rank - 1) = beta;
        if (k > 0 && zc[k] != 0.0) for (int i = 0; i < k; ++i) {
            double tmp = 0.0;
            for (int j = 1; j < nt; ++j) tmp += AA(i, rank - 1 + j) * AA(k, rank - 1 + j);
            tmp += AA(i, rank - 1);
            AA(i, rank - 1) -= zc[k] * tmp;
            for (int j = 1; j < nt; ++j) AA(i, rank - 1 + j) -= zc[k] * tmp * AA(k, rank - 1 + j);
        }
        if (k != rank - 1) for (int i = 0; i <= k; ++i) { double t = AA(i, k); AA(i, k) = AA(i, rank - 1); AA(i, rank - 1) = t; }
    }
    for (int i = 0; i < m; ++i) c[i] = b[i];
    for (int k = 0; k < rank; ++k) {
        if (m - k == 1) c[k] *= (1.0 - hc[k]);
        else if (hc[k] != 0.0) {
            double tmp = 0.0;
            for (int i = k + 1; i < m; ++i) tmp += AA(i, k) * c[i];
            tmp += c[k];
            c[k] -= hc[k] * tmp;
            for (int i = k + 1; i < m; ++i) c[i] -= hc[k] * AA(i, k) * tmp;
        }
    }
    for (int i = rank - 1; i >= 0; --i) {
        double s = c[i];
        for (int j = i + 1; j < rank; ++j) s -= AA(i, j) * y[j];
        y[i] = s / AA(i, i);
    }
    for (int i = rank; i < m; ++i) y[i] = 0.0;
    if (rank < m) for (int k = 0; k < rank; ++k) {
        if (k != rank - 1) { double t = y[k]; y[k] = y[rank - 1]; y[rank - 1] = t; }
        if (nt > 1 && zc[k] != 0.0) {
            double tmp = 0.0;
            for (int j = 1; j < nt; ++j) tmp += AA(k, rank - 1 + j) * y[rank - 1 + j];
            tmp += y[rank - 1];
            y[rank - 1] -= zc[k] * tmp;
            for (int j = 1; j < nt; ++j) y[rank - 1 + j] -= zc[k] * AA(k, rank - 1 + j) * tmp;
        }
        if (k != rank - 1) { double t = y[k]; y[k] = y[rank - 1]; y[rank - 1] = t; }
    }
    for (int i = 0; i < m; ++i) pidx[i] = i;
    for (int k = 0; k < m; ++k) { int t = pidx[k]; pidx[k] = pidx[tr[k]]; pidx[tr[k]] = t; }
    for (int i = 0; i < m; ++i) x[pidx[i]] = y[i];
#undef AA
    return rank;
}

/* ------------------------------------------------------------------------------------------ */
/* Anderson acceleration (variant H)                                                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int m, nt, ne, iter, col;
    double *u, *F, *dF, *scale, *G, *dG, *M, *theta;
} aa_t;

void *port_aa_new(int m, int total_dim, int effective_dim) {
    aa_t *a = (aa_t *)calloc(1, sizeof(aa_t));
    a->m = m; a->nt = total_dim; a->ne = effective_dim; a->iter = -1; a->col = -1;
    a->u = calloc(total_dim, sizeof(double)); a->F = calloc(effective_dim, sizeof(double));
    a->dF = calloc((size_t)effective_dim * m, sizeof(double)); a->scale = calloc(m, sizeof(double));
    a->G = calloc(total_dim, sizeof(double)); a->dG = calloc((size_t)total_dim * m, sizeof(double));
    a->M = calloc(m * m, sizeof(double)); a->theta = calloc(m, sizeof(double));
    return a;
}
void port_aa_free(void *p) {
    aa_t *a = (aa_t *)p;
    free(a->u); free(a->F); free(a->dF); free(a->scale); free(a->G); free(a->dG); free(a->M); free(a->theta); free(a);
}
void port_aa_replace(void *p, const double *u) { aa_t *a = p; memcpy(a->u, u, sizeof(double) * a->nt); }
void port_aa_reset(void *p, const double *u) { aa_t *a = p; memcpy(a->u, u, sizeof(double) * a->nt); a->iter = 0; a->col = 0; }
void port_aa_init(void *p, const double *u) { port_aa_reset(p, u); }
void port_aa_compute(void *p, const double *g, double *out) {
    aa_t *a = p;
    const int m = a->m, ne = a->ne, nt = a->nt;
    memcpy(a->G, g, sizeof(double) * nt);
    for (int i = 0; i < ne; ++i) a->F[i] = a->G[i] - a->u[i];
    if (a->iter == 0) {
        for (int i = 0; i < ne; ++i) a->dF[i] = -a->F[i];
        for (int i = 0; i < nt; ++i) a->dG[i] = -a->G[i];
        memcpy(a->u, a->G, sizeof(double) * nt);
    } else {
        double *dFc = a->dF + (size_t)a->col * ne, *dGc = a->dG + (size_t)a->col * nt;
        for (int i = 0; i < ne; ++i) dFc[i] += a->F[i];
        for (int i = 0; i < nt; ++i) dGc[i] += a->G[i];
        const double eps = 1e-14;
        double s = 0.0;
        for (int i = 0; i < ne; ++i) s += dFc[i] * dFc[i];
        double scale = fmax(eps, sqrt(s));
        a->scale[a->col] = scale;
        for (int i = 0; i < ne; ++i) dFc[i] /= scale;
        int mk = a->iter < m ? a->iter : m;
        if (mk == 1) {
            a->theta[0] = 0.0;
            double sq = 0.0;
            for (int i = 0; i < ne; ++i) sq += dFc[i] * dFc[i];
            a->M[0] = sq;
            double nrm = sqrt(sq);
            if (nrm > eps) {
                double d = 0.0;
                for (int i = 0; i < ne; ++i) d += (dFc[i] / nrm) * (a->F[i] / nrm);
                a->theta[0] = d;
            }
        } else {
            double Mk[MAXM * MAXM], rhs[MAXM];
            for (int j = 0; j < mk; ++j) {
                const double *dFj = a->dF + (size_t)j * ne;
                double d = 0.0;
                for (int i = 0; i < ne; ++i) d += dFc[i] * dFj[i];
                a->M[a->col + j * m] = d;   /* row col */
                a->M[j + a->col * m] = d;   /* column col */
            }
            for (int j = 0; j < mk; ++j) {
                const double *dFj = a->dF + (size_t)j * ne;
                double d = 0.0;
                for (int i = 0; i < ne; ++i) d += dFj[i] * a->F[i];
                rhs[j] = d;
                for (int i = 0; i < mk; ++i) Mk[j * mk + i] = a->M[i + j * m];
            }
            port_cod_solve(mk, Mk, rhs, a->theta);
        }
        for (int i = 0; i < nt; ++i) {
            double acc = 0.0;
            for (int j = 0; j < mk; ++j) acc += a->dG[(size_t)j * nt + i] * (a->theta[j] / a->scale[j]);
            a->u[i] = a->G[i] - acc;
        }
        a->col = (a->col + 1) % m;
        dFc = a->dF + (size_t)a->col * ne; dGc = a->dG + (size_t)a->col * nt;
        for (int i = 0; i < ne; ++i) dFc[i] = -a->F[i];
        for (int i = 0; i < nt; ++i) dGc[i] = -a->G[i];
    }
    a->iter++;
    memcpy(out, a->u, sizeof(double) * nt);
}

/* ------------------------------------------------------------------------------------------ */
/* Tet scene: setup + the two step orderings                                                    */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int nv, nt, nf, np;
    int *tet;       /* 4 per tet, vertex ids */
    int *v2f;       /* vertex -> free index or -1 */
    int *pin_rank;  /* vertex -> rank among pinned (ascending id) or -1 */
    double *binv, *w, *kvol, *mass; /* mass per free vertex */
    double dt, rho_dt2;
    int ordering;   /* 0 hard_zxu, 1 xzu */
    /* scalar system, lower CSC + factor */
    int *Lp, *Li, *parent; double *Lx, *D;
} scene_t;

static void inverse3(const double *m, double *inv, double *det_out) {
#define MM(r, c) m[(c) * 3 + (r)]
#define COF(i, j) (MM(((i) + 1) % 3, ((j) + 1) % 3) * MM(((i) + 2) % 3, ((j) + 2) % 3) - MM(((i) + 1) % 3, ((j) + 2) % 3) * MM(((i) + 2) % 3, ((j) + 1) % 3))
    double c0 = COF(0, 0), c1 = COF(1, 0), c2 = COF(2, 0);
    double det = (c0 * MM(0, 0) + c1 * MM(1, 0)) + c2 * MM(2, 0);
    double id = 1.0 / det;
    inv[0] = c0 * id; inv[3] = c1 * id; inv[6] = c2 * id;
    inv[1] = COF(0, 1) * id; inv[4] = COF(1, 1) * id; inv[7] = COF(2, 1) * id;
    inv[2] = COF(0, 2) * id; inv[5] = COF(1, 2) * id; inv[8] = COF(2, 2) * id;
    *det_out = det3(m);
#undef COF
#undef MM
}

/* G(r,c): reduction coefficient of corner c for deformation-gradient column r */
static double Gc(const double *b, int r, int c) {
    return c == 0 ? -(b[r * 3 + 0] + b[r * 3 + 1] + b[r * 3 + 2]) : b[r * 3 + c - 1];
}

static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }

void port_scene_free(void *p) {
    scene_t *s = p;
    if (!s) return;
    free(s->tet); free(s->v2f); free(s->pin_rank); free(s->binv); free(s->w); free(s->kvol); free(s->mass);
    free(s->Lp); free(s->Li); free(s->parent); free(s->Lx); free(s->D); free(s);
}

/* rest: 3 doubles per vertex; masses: per vertex; pins: vertex ids */
void *port_scene_new(int nv, const double *rest, int nt, const int *tets, const double *masses, double youngs,
                     double poisson, int npins, const int *pins, double dt, double penalty, int ordering) {
    scene_t *s = calloc(1, sizeof(scene_t));
    s->nv = nv; s->nt = nt; s->np = npins; s->dt = dt; s->ordering = ordering;
    s->rho_dt2 = (ordering == 0 ? penalty : 1.0) * dt * dt;
    s->tet = malloc(sizeof(int) * 4 * nt); memcpy(s->tet, tets, sizeof(int) * 4 * nt);
    s->v2f = malloc(sizeof(int) * nv); s->pin_rank = malloc(sizeof(int) * nv);
    int *sorted = malloc(sizeof(int) * (npins + 1));
    memcpy(sorted, pins, sizeof(int) * npins);
    qsort(sorted, npins, sizeof(int), cmp_int);
    for (int v = 0; v < nv; ++v) { s->v2f[v] = 0; s->pin_rank[v] = -1; }
    for (int k = 0; k < npins; ++k) { s->v2f[sorted[k]] = -1; s->pin_rank[sorted[k]] = k; }
    free(sorted);
    int nf = 0;
    for (int v = 0; v < nv; ++v) if (s->v2f[v] == 0) s->v2f[v] = nf++;
    s->nf = nf;
    s->mass = malloc(sizeof(double) * nf);
    for (int v = 0; v < nv; ++v) if (s->v2f[v] >= 0) s->mass[s->v2f[v]] = masses[v];
    s->binv = malloc(sizeof(double) * 9 * nt); s->w = malloc(sizeof(double) * nt); s->kvol = malloc(sizeof(double) * nt);
    double mu = youngs / (2.0 * (1.0 + poisson)), lambda = youngs * poisson / ((1.0 + poisson) * (1.0 - 2.0 * poisson));
    double K = lambda + (2.0 / 3.0) * mu;
    for (int t = 0; t < nt; ++t) {
        const int *tv = tets + 4 * t;
        double e[9], det;
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) e[c * 3 + r] = rest[3 * tv[c + 1] + r] - rest[3 * tv[0] + r];
        inverse3(e, s->binv + 9 * t, &det);
        double vol = det / 6.0;
        if (vol < 0) { port_scene_free(s); return NULL; }
        s->w[t] = sqrt(K * vol);
        s->kvol[t] = K * vol;
    }
    /* dense-row assembly of the scalar system through per-column sorted lists */
    int *cnt = calloc(nf + 1, sizeof(int));
    for (int t = 0; t < nt; ++t) for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) {
        int va = s->v2f[tets[4 * t + a]], vb = s->v2f[tets[4 * t + b]];
        if (va >= 0 && vb >= 0 && va >= vb) cnt[vb + 1]++;
    }
    for (int j = 0; j < nf; ++j) cnt[j + 1] += cnt[j];
    int ntr = cnt[nf];
    int *ti = malloc(sizeof(int) * (ntr + 1)); double *tx = malloc(sizeof(double) * (ntr + 1));
    int *pos = malloc(sizeof(int) * (nf + 1)); memcpy(pos, cnt, sizeof(int) * (nf + 1));
    for (int t = 0; t < nt; ++t) {
        const double *b = s->binv + 9 * t; double w = s->w[t];
        for (int a = 0; a < 4; ++a) for (int c = 0; c < 4; ++c) {
            int va = s->v2f[tets[4 * t + a]], vb = s->v2f[tets[4 * t + c]];
            if (va >= 0 && vb >= 0 && va >= vb) {
                double v = 0.0;
                for (int r = 0; r < 3; ++r) v += (s->rho_dt2 * (w * Gc(b, r, a))) * (w * Gc(b, r, c));
                ti[pos[vb]] = va; tx[pos[vb]] = v; pos[vb]++;
            }
        }
    }
    /* compress each column (sum duplicates), diagonal gets the mass */
    int *Ap = calloc(nf + 1, sizeof(int)); int *Ai = malloc(sizeof(int) * (ntr + nf + 1)); double *Ax = malloc(sizeof(double) * (ntr + nf + 1));
    double *acc = calloc(nf, sizeof(double)); int *mark = malloc(sizeof(int) * nf); int *list = malloc(sizeof(int) * nf);
    for (int i = 0; i < nf; ++i) mark[i] = -1;
    int nnz = 0;
    for (int j = 0; j < nf; ++j) {
        int nl = 0;
        mark[j] = j; list[nl++] = j; acc[j] = s->mass[j];
        for (int p = cnt[j]; p < cnt[j + 1]; ++p) {
            int i = ti[p];
            if (mark[i] != j) { mark[i] = j; list[nl++] = i; acc[i] = 0.0; }
            acc[i] += tx[p];
        }
        qsort(list, nl, sizeof(int), cmp_int);
        for (int k = 0; k < nl; ++k) { Ai[nnz] = list[k]; Ax[nnz] = acc[list[k]]; nnz++; }
        Ap[j + 1] = nnz;
    }
    free(cnt); free(ti); free(tx); free(pos); free(acc); free(mark); free(list);
    /* up-looking LDL^T (symbolic: etree + column counts; numeric: row by row) */
    int n = nf;
    /* build upper-by-row access: for row k we need A(i,k) for i<k, i.e. entries of column i at row k:
       transpose the strict lower part */
    int *Tp = calloc(n + 1, sizeof(int));
    for (int j = 0; j < n; ++j) for (int p = Ap[j]; p < Ap[j + 1]; ++p) if (Ai[p] > j) Tp[Ai[p] + 1]++;
    for (int j = 0; j < n; ++j) Tp[j + 1] += Tp[j];
    int *Ti = malloc(sizeof(int) * (Tp[n] + 1)); double *Tx = malloc(sizeof(double) * (Tp[n] + 1));
    int *tp = malloc(sizeof(int) * (n + 1)); memcpy(tp, Tp, sizeof(int) * (n + 1));
    double *diag = malloc(sizeof(double) * n);
    for (int j = 0; j < n; ++j) for (int p = Ap[j]; p < Ap[j + 1]; ++p) {
        if (Ai[p] == j) diag[j] = Ax[p];
        else { Ti[tp[Ai[p]]] = j; Tx[tp[Ai[p]]] = Ax[p]; tp[Ai[p]]++; }
    }
    s->parent = malloc(sizeof(int) * n); s->Lp = calloc(n + 1, sizeof(int)); s->D = malloc(sizeof(double) * n);
    int *Lnz = calloc(n, sizeof(int)), *flag = malloc(sizeof(int) * n), *pattern = malloc(sizeof(int) * n);
    for (int k = 0; k < n; ++k) {
        s->parent[k] = -1; flag[k] = k;
        for (int p = Tp[k]; p < Tp[k + 1]; ++p) {
            int i = Ti[p];
            for (; flag[i] != k; i = s->parent[i]) {
                if (s->parent[i] == -1) s->parent[i] = k;
                Lnz[i]++; flag[i] = k;
            }
        }
    }
    for (int k = 0; k < n; ++k) s->Lp[k + 1] = s->Lp[k] + Lnz[k];
    s->Li = malloc(sizeof(int) * (s->Lp[n] + 1)); s->Lx = malloc(sizeof(double) * (s->Lp[n] + 1));
    double *Y = calloc(n, sizeof(double));
    for (int k = 0; k < n; ++k) Lnz[k] = 0;
    for (int k = 0; k < n; ++k) {
        int top = n;
        Y[k] = 0.0; flag[k] = k;
        for (int p = Tp[k]; p < Tp[k + 1]; ++p) {
            int i = Ti[p], len = 0;
            Y[i] += Tx[p];
            for (; flag[i] != k; i = s->parent[i]) { pattern[len++] = i; flag[i] = k; }
            while (len > 0) pattern[--top] = pattern[--len];
        }
        double dk = diag[k];
        for (; top < n; ++top) {
            int i = pattern[top];
            double yi = Y[i]; Y[i] = 0.0;
            int p2 = s->Lp[i] + Lnz[i];
            for (int p = s->Lp[i]; p < p2; ++p) Y[s->Li[p]] -= s->Lx[p] * yi;
            double lki = yi / s->D[i];
            dk -= lki * yi;
            s->Li[p2] = k; s->Lx[p2] = lki; Lnz[i]++;
        }
        s->D[k] = dk;
    }
    free(Ap); free(Ai); free(Ax); free(Tp); free(Ti); free(Tx); free(tp); free(diag); free(Lnz); free(flag); free(pattern); free(Y);
    return s;
}

int port_scene_nfree(void *p) { return ((scene_t *)p)->nf; }

/* x (3 per free vertex) <- A^-1 b, three right-hand sides at once */
static void solve3(const scene_t *s, const double *b, double *x) {
    int n = s->nf;
    memcpy(x, b, sizeof(double) * 3 * n);
    for (int j = 0; j < n; ++j) for (int p = s->Lp[j]; p < s->Lp[j + 1]; ++p)
        for (int r = 0; r < 3; ++r) x[3 * s->Li[p] + r] -= s->Lx[p] * x[3 * j + r];
    for (int j = 0; j < n; ++j) for (int r = 0; r < 3; ++r) x[3 * j + r] /= s->D[j];
    for (int j = n - 1; j >= 0; --j) for (int p = s->Lp[j]; p < s->Lp[j + 1]; ++p)
        for (int r = 0; r < 3; ++r) x[3 * j + r] -= s->Lx[p] * x[3 * s->Li[p] + r];
}

/* D_i x - c_i = w * Ds * Binv with free positions from xf and pinned ones from xp */
static void wF(const scene_t *s, int t, const double *xf, const double *xp, double *out) {
    const int *tv = s->tet + 4 * t; const double *b = s->binv + 9 * t; double w = s->w[t];
    double X[4][3];
    for (int c = 0; c < 4; ++c) for (int j = 0; j < 3; ++j)
        X[c][j] = s->v2f[tv[c]] >= 0 ? xf[3 * s->v2f[tv[c]] + j] : xp[3 * s->pin_rank[tv[c]] + j];
    for (int r = 0; r < 3; ++r) for (int j = 0; j < 3; ++j) {
        double f = (X[1][j] - X[0][j]) * b[r * 3 + 0] + (X[2][j] - X[0][j]) * b[r * 3 + 1] + (X[3][j] - X[0][j]) * b[r * 3 + 2];
        out[r * 3 + j] = w * f;
    }
}
/* update_z for all tets: z = prox((Dx - c + u)/w) */
static void update_z(const scene_t *s, const double *xf, const double *xp, const double *u, double *z) {
    for (int t = 0; t < s->nt; ++t) {
        double d[9], zi[9], winv = 1.0 / s->w[t];
        wF(s, t, xf, xp, d);
        for (int k = 0; k < 9; ++k) zi[k] = (d[k] + u[9 * t + k]) * winv;
        tet_prox(zi);
        memcpy(z + 9 * t, zi, sizeof(zi));
    }
}
static void update_u(const scene_t *s, const double *xf, const double *xp, const double *z, double *u) {
    for (int t = 0; t < s->nt; ++t) {
        double d[9];
        wF(s, t, xf, xp, d);
        for (int k = 0; k < 9; ++k) u[9 * t + k] += d[k] - s->w[t] * z[9 * t + k];
    }
}
static double prim2(const scene_t *s, const double *xf, const double *xp, const double *z) {
    double acc = 0.0;
    for (int t = 0; t < s->nt; ++t) {
        double d[9];
        wF(s, t, xf, xp, d);
        for (int k = 0; k < 9; ++k) { double r = d[k] - s->w[t] * z[9 * t + k]; acc += r * r; }
    }
    return acc;
}
/* x = A^-1 (M xbar + rho dt^2 D^T (W z + C_fix - u)) */
static void solve_x(const scene_t *s, const double *xbar, const double *xp, const double *z, const double *u, double *x,
                    double *rhs) {
    for (int v = 0; v < s->nf; ++v) for (int j = 0; j < 3; ++j) rhs[3 * v + j] = s->mass[v] * xbar[3 * v + j];
    for (int t = 0; t < s->nt; ++t) {
        const int *tv = s->tet + 4 * t; const double *b = s->binv + 9 * t; double w = s->w[t];
        double y[9];
        for (int r = 0; r < 3; ++r) for (int j = 0; j < 3; ++j) {
            double cf = 0.0;   /* C_fix block */
            for (int c = 0; c < 4; ++c) if (s->v2f[tv[c]] < 0) cf -= (w * Gc(b, r, c)) * xp[3 * s->pin_rank[tv[c]] + j];
            y[r * 3 + j] = w * z[9 * t + r * 3 + j] + cf - u[9 * t + r * 3 + j];
        }
        for (int c = 0; c < 4; ++c) if (s->v2f[tv[c]] >= 0) for (int j = 0; j < 3; ++j) {
            double a = 0.0;
            for (int r = 0; r < 3; ++r) a += (s->rho_dt2 * (w * Gc(b, r, c))) * y[r * 3 + j];
            rhs[3 * s->v2f[tv[c]] + j] += a;
        }
    }
    solve3(s, rhs, x);
}

/* One Solver::step(). x, v: 3 per vertex (in/out). pin_pts: 3 per pin in ascending pinned-vertex order.
 * hist_*: admm_iters entries. Returns the number of logged rows. */
int port_scene_step(void *p, double *x, double *v, const double *pin_pts, int iters, int m, int accel, double gravity,
                    double *hist_prim, double *hist_comb, int *hist_rej) {
    scene_t *s = p;
    const int nf = s->nf, Z = 9 * s->nt, nv = s->nv;
    const double dt = s->dt, eps = 1e-20;
    double *xbar = malloc(sizeof(double) * 3 * nf), *cx = malloc(sizeof(double) * 3 * nf), *rhs = malloc(sizeof(double) * 3 * nf);
    double *z = calloc(Z, sizeof(double)), *u = calloc(Z, sizeof(double));
    double *du = malloc(sizeof(double) * Z), *dx = malloc(sizeof(double) * 3 * nf), *dz = malloc(sizeof(double) * Z);
    double *lastx = malloc(sizeof(double) * 3 * nf), *buf = malloc(sizeof(double) * (Z + 3 * nf)), *out = malloc(sizeof(double) * (Z + 3 * nf));
    if (fabs(gravity) > 0) for (int i = 0; i < nv; ++i) if (s->v2f[i] >= 0) v[3 * i + 1] += dt * gravity;
    for (int i = 0; i < nv; ++i) if (s->v2f[i] >= 0) for (int j = 0; j < 3; ++j) xbar[3 * s->v2f[i] + j] = x[3 * i + j] + dt * v[3 * i + j];
    memcpy(cx, xbar, sizeof(double) * 3 * nf);
    int rows = 0;
    double prev_prim = 1e+20, prim = 0.0, comb = 0.0;
    const double *final_x = cx;
    if (s->ordering == 0) {
        /* ---- hard_zxu (hard/src/Solver.cpp:74-226) ---- */
        update_z(s, cx, pin_pts, u, z);
        solve_x(s, xbar, pin_pts, z, u, cx, rhs);
        update_u(s, cx, pin_pts, z, u);
        memcpy(du, u, sizeof(double) * Z); memcpy(dx, cx, sizeof(double) * 3 * nf);
        void *aa = NULL;
        if (accel && m > 0) { aa = port_aa_new(m, Z + 3 * nf, Z); memcpy(buf, u, sizeof(double) * Z); memcpy(buf + Z, cx, sizeof(double) * 3 * nf); port_aa_init(aa, buf); }
        for (int it = 0; it < iters; ++it) {
            update_z(s, cx, pin_pts, u, z);
            prim = sqrt(prim2(s, cx, pin_pts, z));
            int rej = 0;
            if (accel && prev_prim < prim) {
                memcpy(u, du, sizeof(double) * Z); memcpy(cx, dx, sizeof(double) * 3 * nf);
                memcpy(buf, u, sizeof(double) * Z); memcpy(buf + Z, cx, sizeof(double) * 3 * nf); port_aa_reset(aa, buf);
                update_z(s, cx, pin_pts, u, z);
                prim = sqrt(prim2(s, cx, pin_pts, z));
                rej = 1;
            }
            memcpy(lastx, cx, sizeof(double) * 3 * nf);
            prev_prim = prim;
            solve_x(s, xbar, pin_pts, z, u, cx, rhs);
            double dual = 0.0;
            for (int t = 0; t < s->nt; ++t) {   /* |D (x - last_x)|^2: pinned parts cancel */
                const int *tv = s->tet + 4 * t; const double *b = s->binv + 9 * t; double w = s->w[t];
                double X[4][3];
                for (int c = 0; c < 4; ++c) for (int j = 0; j < 3; ++j) { int fidx = s->v2f[tv[c]]; X[c][j] = fidx >= 0 ? cx[3 * fidx + j] - lastx[3 * fidx + j] : 0.0; }
                for (int r = 0; r < 3; ++r) for (int j = 0; j < 3; ++j) {
                    double f = (X[1][j] - X[0][j]) * b[r * 3 + 0] + (X[2][j] - X[0][j]) * b[r * 3 + 1] + (X[3][j] - X[0][j]) * b[r * 3 + 2];
                    dual += (w * f) * (w * f);
                }
            }
            comb = prim2(s, cx, pin_pts, z) + dual;
            if (comb < eps) break;
            update_u(s, cx, pin_pts, z, u);
            if (accel) {
                memcpy(du, u, sizeof(double) * Z); memcpy(dx, cx, sizeof(double) * 3 * nf);
                memcpy(buf, du, sizeof(double) * Z); memcpy(buf + Z, dx, sizeof(double) * 3 * nf);
                port_aa_compute(aa, buf, out);
                memcpy(u, out, sizeof(double) * Z); memcpy(cx, out + Z, sizeof(double) * 3 * nf);
            }
            hist_prim[rows] = prim; hist_comb[rows] = comb; hist_rej[rows] = rej; rows++;
        }
        if (aa) port_aa_free(aa);
        final_x = accel ? dx : cx;
    } else {
        /* ---- xzu (xzu/src/Solver.cpp:78-257) ---- */
        double *g = malloc(sizeof(double) * Z), *defz = malloc(sizeof(double) * Z), *combz = malloc(sizeof(double) * Z), *combx = malloc(sizeof(double) * 3 * nf);
        /* curr_z = W^-1 (D xbar - C_fix) */
        for (int t = 0; t < s->nt; ++t) { double d[9]; wF(s, t, xbar, pin_pts, d); for (int k = 0; k < 9; ++k) z[9 * t + k] = d[k] * (1.0 / s->w[t]); }
        solve_x(s, xbar, pin_pts, z, u, cx, rhs);
        update_z(s, cx, pin_pts, u, z);
        memcpy(dz, z, sizeof(double) * Z); memcpy(dx, cx, sizeof(double) * 3 * nf); memcpy(du, u, sizeof(double) * Z);
        void *aa = port_aa_new(m > 0 ? m : 1, Z, Z);
        port_aa_init(aa, z);
        for (int it = 0; it < iters; ++it) {
            if (accel) {
                for (int t = 0; t < s->nt; ++t) { tet_grad(z + 9 * t, s->kvol[t], g + 9 * t); for (int k = 0; k < 9; ++k) u[9 * t + k] = (1.0 / s->w[t]) * g[9 * t + k]; }
            } else update_u(s, cx, pin_pts, z, u);
            solve_x(s, xbar, pin_pts, z, u, cx, rhs);
            prim = sqrt(prim2(s, cx, pin_pts, z));
            if (accel && prev_prim < prim) {
                memcpy(u, du, sizeof(double) * Z); memcpy(cx, dx, sizeof(double) * 3 * nf); memcpy(z, dz, sizeof(double) * Z);
                port_aa_replace(aa, z);
                update_u(s, cx, pin_pts, z, u);
                solve_x(s, xbar, pin_pts, z, u, cx, rhs);
                prim = sqrt(prim2(s, cx, pin_pts, z));
            }
            prev_prim = prim;
            if (accel) {
                memcpy(dx, cx, sizeof(double) * 3 * nf); memcpy(du, u, sizeof(double) * Z);
                update_z(s, cx, pin_pts, u, dz);
                port_aa_compute(aa, dz, z);
                /* combined residual "for drawing figures": extra solve + local step on copies */
                solve_x(s, xbar, pin_pts, dz, u, combx, rhs);
                update_z(s, combx, pin_pts, u, combz);
                double dual = 0.0;
                for (int t = 0; t < s->nt; ++t) for (int k = 0; k < 9; ++k) { double d = s->w[t] * (combz[9 * t + k] - dz[9 * t + k]); dual += d * d; }
                comb = dual + prim2(s, combx, pin_pts, combz);
            } else {
                memcpy(combz, z, sizeof(double) * Z);   /* last_z */
                update_z(s, cx, pin_pts, u, z);
                double dual = 0.0;
                for (int t = 0; t < s->nt; ++t) for (int k = 0; k < 9; ++k) { double d = s->w[t] * (z[9 * t + k] - combz[9 * t + k]); dual += d * d; }
                comb = dual + prim2(s, cx, pin_pts, z);
            }
            hist_prim[rows] = prim; hist_comb[rows] = comb; hist_rej[rows] = 0; rows++;
            if (comb < eps) break;
        }
        port_aa_free(aa);
        free(g); free(defz); free(combz); free(combx);
        final_x = cx;
    }
    for (int i = 0; i < nv; ++i) for (int j = 0; j < 3; ++j) {
        double nx = s->v2f[i] >= 0 ? final_x[3 * s->v2f[i] + j] : pin_pts[3 * s->pin_rank[i] + j];
        v[3 * i + j] = (nx - x[3 * i + j]) * (1.0 / dt);
        x[3 * i + j] = nx;
    }
    free(xbar); free(cx); free(rhs); free(z); free(u); free(du); free(dx); free(dz); free(lastx); free(buf); free(out);
    return rows;
}
