/* TEST INFRASTRUCTURE ONLY - CPU restatement (plain C) of the Geometry path and of the remaining element
 * types of bldeng/AA-ADMM, used by tests/ as a checker next to the compiled reference (oracle/_ref).
 * Nothing in the product may call this file. Pinned by tests/test_oracle_cpu.py against oracle/_ref
 * (the unmodified reference classes) on the same inputs.
 *
 *   Constraint<3>::apply_transform / add_constraint        Geometry/Constraint.h:73-94, 132-159
 *   EdgeLengthConstraint / AngleConstraint / PlaneConstraint::project_impl   :211-214, :243-291, :406-413
 *   closest point on a triangle mesh (brute force over the triangles; the reference walks an igl::AABB tree
 *   and tests triangles with the same region logic)          Geometry/external/igl/point_simplex_squared_distance.cpp:44-110
 *   LinearRegularization<3>                                 Geometry/LinearRegularization.h:47-153
 *   ALMGeometrySolver<3>::setup_ADMM / solve_ADMM           Geometry/ALMGeometrySolver.h:81-161, 163-283, 404-461
 *   GeometrySolver<3>::setup_ADMM / solve_ADMM              Geometry/GeometrySolver.h:85-155, 156-263, 383-459
 *   TriEnergyTerm::prox (xzu / hard)                        xzu/src/TriEnergyTerm.cpp:77-107, hard/src/TriEnergyTerm.cpp:74-105
 *   Collision::prox + analytic passive objects              hard/src/CollisionEnergyTerm.hpp:79-91, PassiveObject.hpp:32-136
 * The global system is small in the tests and is solved densely (LDL^T without pivoting).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* Anderson acceleration of aaadmm_port.c (same shared library) */
void *port_aa_new(int m, int total_dim, int effective_dim);
void port_aa_free(void *p);
void port_aa_init(void *p, const double *u);
void port_aa_reset(void *p, const double *u);
void port_aa_replace(void *p, const double *u);
void port_aa_compute(void *p, const double *g, double *out);

enum { GEO_PLANE = 0, GEO_EDGE = 1, GEO_ANGLE = 2 };
#define MAXK 16

typedef struct {
    int use_alm;
    /* hard constraints */
    int nh, cap_h, n_idx, cap_idx;
    int *type, *ptr, *idx, *col0;
    double *param; /* 4 per constraint */
    /* soft closest-point constraint(s): one point each */
    int ns, cap_s;
    int *soft_pt;
    double soft_w;
    int nv, nt;
    double *V;
    int *T;
    /* regularisation rows */
    int nr, cap_r, n_ri, cap_ri;
    int *rptr, *ridx;
    double *rcoef, *rtarget; /* coefficients (already * sqrt(w)), 3 targets per row (already * sqrt(w)) */
    /* setup */
    int P, zc, zc_all;
    double rho;
    double *A; /* P x P, LDL^T in place: strictly lower = L, diagonal = D */
    double *rhs_fixed; /* P x 3 */
} geo_t;

static void *grow(void *p, int *cap, int need, size_t elem) {
    if (need <= *cap) return p;
    int c = *cap ? *cap : 16;
    while (c < need) c *= 2;
    *cap = c;
    return realloc(p, (size_t)c * elem);
}

void *port_geo_new(int use_alm) {
    geo_t *g = calloc(1, sizeof(geo_t));
    g->use_alm = use_alm;
    return g;
}
void port_geo_free(void *p) {
    geo_t *g = p;
    free(g->type); free(g->ptr); free(g->idx); free(g->col0); free(g->param); free(g->soft_pt); free(g->V); free(g->T);
    free(g->rptr); free(g->ridx); free(g->rcoef); free(g->rtarget); free(g->A); free(g->rhs_fixed);
    free(g);
}

static void add_hard(geo_t *g, int type, const int *idx, int k, const double *prm) {
    if (g->nh + 2 > g->cap_h) {
        g->cap_h = g->cap_h ? 2 * g->cap_h : 64;
        g->type = realloc(g->type, sizeof(int) * g->cap_h);
        g->ptr = realloc(g->ptr, sizeof(int) * (g->cap_h + 1));
        g->param = realloc(g->param, sizeof(double) * 4 * g->cap_h);
    }
    g->idx = grow(g->idx, &g->cap_idx, g->n_idx + k, sizeof(int));
    if (g->nh == 0) g->ptr[0] = 0;
    g->type[g->nh] = type;
    memcpy(g->param + 4 * g->nh, prm, 4 * sizeof(double));
    memcpy(g->idx + g->n_idx, idx, k * sizeof(int));
    g->n_idx += k;
    g->ptr[g->nh + 1] = g->n_idx;
    g->nh++;
}
void port_geo_add_plane(void *p, const int *idx, int k, double w) {
    (void)w;
    double prm[4] = {0, 0, 0, 0};
    add_hard(p, GEO_PLANE, idx, k, prm);
}
void port_geo_add_edge(void *p, int i0, int i1, double w, double len) {
    (void)w;
    int idx[2] = {i0, i1};
    double prm[4] = {len, 0, 0, 0};
    add_hard(p, GEO_EDGE, idx, 2, prm);
}
static double clampd(double v, double lo, double hi) { return fmin(fmax(lo, v), hi); }
/* AngleConstraint constructor (Constraint.h:228-237) */
void port_geo_add_angle(void *p, int tip, int s1, int s2, double w, double amin, double amax) {
    (void)w;
    const double pi = 3.14159265358979323846;
    int idx[3] = {tip, s1, s2};
    const double mn = fmax(0.0, amin), mx = fmin(pi, amax);
    double prm[4] = {mn, mx, clampd(cos(mn), -1.0, 1.0), clampd(cos(mx), -1.0, 1.0)};
    add_hard(p, GEO_ANGLE, idx, 3, prm);
}
/* ReferenceSurfceConstraint over points 0..n_points-1 (Constraint.h:351-394) */
void port_geo_add_ref_surface(void *p, int n_points, double w, const double *V, int nv, const int *F, int nf) {
    geo_t *g = p;
    g->soft_pt = grow(g->soft_pt, &g->cap_s, g->ns + n_points, sizeof(int));
    for (int i = 0; i < n_points; ++i) g->soft_pt[g->ns + i] = i;
    g->ns += n_points;
    g->soft_w = w;
    g->nv = nv; g->nt = nf;
    g->V = realloc(g->V, sizeof(double) * 3 * nv);
    g->T = realloc(g->T, sizeof(int) * 3 * nf);
    memcpy(g->V, V, sizeof(double) * 3 * nv);
    memcpy(g->T, F, sizeof(int) * 3 * nf);
}
/* LinearRegularization rows: coefficients and targets are scaled by sqrt(weight) (LinearRegularization.h:47-117) */
static void add_reg(geo_t *g, const int *idx, const double *coef, int n, double w, const double *target3) {
    const double sw = sqrt(w);
    if (g->nr + 2 > g->cap_r) {
        g->cap_r = g->cap_r ? 2 * g->cap_r : 64;
        g->rptr = realloc(g->rptr, sizeof(int) * (g->cap_r + 1));
        g->rtarget = realloc(g->rtarget, sizeof(double) * 3 * g->cap_r);
    }
    if (g->n_ri + n > g->cap_ri) {
        g->cap_ri = 2 * (g->n_ri + n) + 64;
        g->ridx = realloc(g->ridx, sizeof(int) * g->cap_ri);
        g->rcoef = realloc(g->rcoef, sizeof(double) * g->cap_ri);
    }
    if (g->nr == 0) g->rptr[0] = 0;
    for (int i = 0; i < n; ++i) {
        g->ridx[g->n_ri + i] = idx[i];
        g->rcoef[g->n_ri + i] = coef[i] * sw;
    }
    g->n_ri += n;
    for (int r = 0; r < 3; ++r) g->rtarget[3 * g->nr + r] = target3[r] * sw;
    g->rptr[g->nr + 1] = g->n_ri;
    g->nr++;
}
void port_geo_add_uniform_laplacian(void *p, const int *idx, int n, double w, const double *ref_pts /* may be NULL */) {
    double coef[64], t[3] = {0, 0, 0};
    coef[0] = 1.0;
    for (int i = 1; i < n; ++i) coef[i] = -1.0 / (double)(n - 1);
    if (ref_pts)
        for (int i = 0; i < n; ++i)
            for (int r = 0; r < 3; ++r) t[r] += ref_pts[3 * idx[i] + r] * coef[i];
    add_reg(p, idx, coef, n, w, t);
}
void port_geo_add_closeness(void *p, int idx, double w, const double *target3) {
    double one = 1.0;
    add_reg(p, &idx, &one, 1, w, target3);
}

/* ---- per-constraint operators ---------------------------------------------------------------- */
static int n_cols(int type, int k) { return type == GEO_PLANE ? k : k - 1; }

/* Constraint::apply_transform: out[j*3 + r], returns the number of columns */
static int transform(int type, const int *ids, int k, const double *x, double *out) {
    if (type == GEO_PLANE) {
        double mean[3] = {0, 0, 0};
        for (int j = 0; j < k; ++j)
            for (int r = 0; r < 3; ++r) {
                out[3 * j + r] = x[3 * ids[j] + r];
                mean[r] += out[3 * j + r];
            }
        for (int r = 0; r < 3; ++r) mean[r] /= (double)k;
        for (int j = 0; j < k; ++j)
            for (int r = 0; r < 3; ++r) out[3 * j + r] -= mean[r];
        return k;
    }
    for (int j = 1; j < k; ++j)
        for (int r = 0; r < 3; ++r) out[3 * (j - 1) + r] = x[3 * ids[j] + r] - x[3 * ids[0] + r];
    return k - 1;
}

static double norm3(const double *v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
static void normalized(const double *v, double *o) {
    const double n = norm3(v);
    for (int r = 0; r < 3; ++r) o[r] = n > 0.0 ? v[r] / n : v[r];
}

/* PlaneConstraint::project_impl: remove the component along the direction of least variance of the k
 * mean-centred points = eigenvector of the smallest eigenvalue of A A^T (cyclic Jacobi on the 3x3 matrix). */
static void project_plane(const double *v, int k, double *out) {
    double C[3][3] = {{0}}, E[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int j = 0; j < k; ++j)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) C[a][b] += v[3 * j + a] * v[3 * j + b];
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = fabs(C[0][1]) + fabs(C[0][2]) + fabs(C[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(C[p][q]) < DBL_MIN) continue;
                const double th = (C[q][q] - C[p][p]) / (2.0 * C[p][q]);
                const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(1.0 + th * th));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int a = 0; a < 3; ++a) {
                    const double cp = C[a][p], cq = C[a][q];
                    C[a][p] = c * cp - s * cq;
                    C[a][q] = s * cp + c * cq;
                }
                for (int a = 0; a < 3; ++a) {
                    const double cp = C[p][a], cq = C[q][a];
                    C[p][a] = c * cp - s * cq;
                    C[q][a] = s * cp + c * cq;
                }
                for (int a = 0; a < 3; ++a) {
                    const double ep = E[a][p], eq = E[a][q];
                    E[a][p] = c * ep - s * eq;
                    E[a][q] = s * ep + c * eq;
                }
            }
    }
    int m = 0;
    if (C[1][1] < C[m][m]) m = 1;
    if (C[2][2] < C[m][m]) m = 2;
    double n[3] = {E[0][m], E[1][m], E[2][m]}, nn[3];
    normalized(n, nn);
    for (int j = 0; j < k; ++j) {
        const double d = nn[0] * v[3 * j] + nn[1] * v[3 * j + 1] + nn[2] * v[3 * j + 2];
        for (int r = 0; r < 3; ++r) out[3 * j + r] = v[3 * j + r] - nn[r] * d;
    }
}

/* EdgeLengthConstraint::project_impl */
static void project_edge(const double *v, double len, double *out) {
    double u[3];
    normalized(v, u);
    for (int r = 0; r < 3; ++r) out[r] = u[r] * len;
}

/* AngleConstraint::project_impl */
static void project_angle(const double *v, const double *prm, double *out) {
    const double min_angle = prm[0], max_angle = prm[1], min_cos = prm[2], max_cos = prm[3];
    memcpy(out, v, 6 * sizeof(double));
    const double *v1 = v, *v2 = v + 3;
    const double v1s = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2], v2s = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
    const double v1n = sqrt(v1s), v2n = sqrt(v2s);
    double u1[3], u2[3];
    normalized(v1, u1);
    normalized(v2, u2);
    const double cg = clampd(u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2], -1.0, 1.0);
    if ((1.0 - fabs(cg) > 1e-14) && (cg > min_cos || cg < max_cos)) {
        const double gamma = acos(cg);
        double eta = cg > min_cos ? (min_angle - gamma) : (gamma - max_angle);
        eta = fmax(eta, 0.0);
        double theta = 0.5 * atan2(v2s * sin(2 * eta), v1s + v2s * cos(2 * eta));
        theta = fmax(0.0, fmin(eta, theta));
        const double phi = eta - theta;
        double t3[3], t4[3], u3[3], u4[3];
        for (int r = 0; r < 3; ++r) {
            t3[r] = u2[r] - u1[r] * cg;
            t4[r] = u1[r] - u2[r] * cg;
        }
        normalized(t3, u3);
        normalized(t4, u4);
        if (cg > min_cos)
            for (int r = 0; r < 3; ++r) {
                u3[r] *= -1.0;
                u4[r] *= -1.0;
            }
        for (int r = 0; r < 3; ++r) {
            out[r] = (u1[r] * cos(theta) + u3[r] * sin(theta)) * (v1n * cos(theta));
            out[3 + r] = (u2[r] * cos(phi) + u4[r] * sin(phi)) * (v2n * cos(phi));
        }
    }
}

static void project(int type, const double *v, int kc, const double *prm, double *out) {
    if (type == GEO_PLANE)
        project_plane(v, kc, out);
    else if (type == GEO_EDGE)
        project_edge(v, prm[0], out);
    else
        project_angle(v, prm, out);
}

/* unit entry: n constraints of one type with k transformed columns each */
void port_geo_project(int type, int n, int k, const double *cols, const double *param4, double *out) {
    for (int i = 0; i < n; ++i) project(type, cols + 3 * (size_t)k * i, k, param4 + 4 * i, out + 3 * (size_t)k * i);
}

/* closest point on triangle (a, b, c) to p: Ericson's regions, as igl's point_simplex_squared_distance */
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double closest_on_triangle(const double *p, const double *a, const double *b, const double *c, double *out) {
    double ab[3], ac[3], ap[3], bp[3], cp[3];
    for (int r = 0; r < 3; ++r) {
        ab[r] = b[r] - a[r];
        ac[r] = c[r] - a[r];
        ap[r] = p[r] - a[r];
    }
    const double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    int done = 0;
    if (d1 <= 0.0 && d2 <= 0.0) {
        memcpy(out, a, 3 * sizeof(double));
        done = 1;
    }
    double d3 = 0, d4 = 0, d5 = 0, d6 = 0, vc = 0, vb = 0;
    if (!done) {
        for (int r = 0; r < 3; ++r) bp[r] = p[r] - b[r];
        d3 = dot3(ab, bp);
        d4 = dot3(ac, bp);
        if (d3 >= 0.0 && d4 <= d3) {
            memcpy(out, b, 3 * sizeof(double));
            done = 1;
        }
    }
    if (!done) {
        vc = d1 * d4 - d3 * d2;
        const int a_ne_b = !(a[0] == b[0] && a[1] == b[1] && a[2] == b[2]);
        if (a_ne_b && vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
            const double v = d1 / (d1 - d3);
            for (int r = 0; r < 3; ++r) out[r] = a[r] + v * ab[r];
            done = 1;
        }
    }
    if (!done) {
        for (int r = 0; r < 3; ++r) cp[r] = p[r] - c[r];
        d5 = dot3(ab, cp);
        d6 = dot3(ac, cp);
        if (d6 >= 0.0 && d5 <= d6) {
            memcpy(out, c, 3 * sizeof(double));
            done = 1;
        }
    }
    if (!done) {
        vb = d5 * d2 - d1 * d6;
        if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
            const double w = d2 / (d2 - d6);
            for (int r = 0; r < 3; ++r) out[r] = a[r] + w * ac[r];
            done = 1;
        }
    }
    if (!done) {
        const double va = d3 * d6 - d5 * d4;
        if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
            const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
            for (int r = 0; r < 3; ++r) out[r] = b[r] + w * (c[r] - b[r]);
        } else {
            const double denom = 1.0 / (va + vb + vc);
            const double v = vb * denom, w = vc * denom;
            for (int r = 0; r < 3; ++r) out[r] = a[r] + ab[r] * v + ac[r] * w;
        }
    }
    const double dx = p[0] - out[0], dy = p[1] - out[1], dz = p[2] - out[2];
    return dx * dx + dy * dy + dz * dz;
}
static void closest_point(const double *V, const int *T, int nt, const double *p, double *out) {
    double best = DBL_MAX;
    for (int t = 0; t < nt; ++t) {
        double c[3];
        const double d = closest_on_triangle(p, V + 3 * T[3 * t], V + 3 * T[3 * t + 1], V + 3 * T[3 * t + 2], c);
        if (d < best) {
            best = d;
            memcpy(out, c, 3 * sizeof(double));
        }
    }
}
void port_geo_closest_points(const double *V, int nv, const int *T, int nt, const double *q, int nq, double *out) {
    (void)nv;
    for (int i = 0; i < nq; ++i) closest_point(V, T, nt, q + 3 * i, out + 3 * i);
}

/* ---- setup ------------------------------------------------------------------------------------ */
/* D row r of the hard constraints as (point, coefficient) pairs (Constraint::add_constraint, unweighted) */
static int hard_row(const geo_t *g, int c, int j, int *pts, double *coef) {
    const int *ids = g->idx + g->ptr[c];
    const int k = g->ptr[c + 1] - g->ptr[c];
    if (g->type[c] == GEO_PLANE) {
        for (int i = 0; i < k; ++i) {
            pts[i] = ids[i];
            coef[i] = (i == j) ? 1.0 - 1.0 / k : -1.0 / k;
        }
        return k;
    }
    pts[0] = ids[0];
    coef[0] = -1.0;
    pts[1] = ids[j + 1];
    coef[1] = 1.0;
    return 2;
}

int port_geo_setup(void *p, int n_points, double rho) {
    geo_t *g = p;
    g->P = n_points;
    g->rho = rho;
    free(g->col0);
    g->col0 = malloc(sizeof(int) * (g->nh + 1));
    int zc = 0;
    for (int c = 0; c < g->nh; ++c) {
        g->col0[c] = zc;
        zc += n_cols(g->type[c], g->ptr[c + 1] - g->ptr[c]);
    }
    g->zc = zc;
    g->zc_all = zc + (g->use_alm ? 0 : g->ns);
    const int P = n_points;
    free(g->A);
    free(g->rhs_fixed);
    g->A = calloc((size_t)P * P, sizeof(double));
    g->rhs_fixed = calloc((size_t)P * 3, sizeof(double));
    double *A = g->A;
    int pts[MAXK];
    double coef[MAXK];
    /* rho D_hard^T D_hard */
    for (int c = 0; c < g->nh; ++c) {
        const int kc = n_cols(g->type[c], g->ptr[c + 1] - g->ptr[c]);
        for (int j = 0; j < kc; ++j) {
            const int n = hard_row(g, c, j, pts, coef);
            for (int a = 0; a < n; ++a)
                for (int b = 0; b < n; ++b) A[(size_t)pts[a] * P + pts[b]] += rho * coef[a] * coef[b];
        }
    }
    /* soft rows: ALM weighted by sqrt(w) outside the penalty term; GS unweighted inside it */
    for (int i = 0; i < g->ns; ++i) A[(size_t)g->soft_pt[i] * P + g->soft_pt[i]] += g->use_alm ? g->soft_w : rho;
    /* L^T L and L^T targets */
    for (int r = 0; r < g->nr; ++r) {
        const int n = g->rptr[r + 1] - g->rptr[r];
        const int *ri = g->ridx + g->rptr[r];
        const double *rc = g->rcoef + g->rptr[r];
        for (int a = 0; a < n; ++a) {
            for (int b = 0; b < n; ++b) A[(size_t)ri[a] * P + ri[b]] += rc[a] * rc[b];
            for (int k = 0; k < 3; ++k) g->rhs_fixed[3 * ri[a] + k] += rc[a] * g->rtarget[3 * r + k];
        }
    }
    /* dense LDL^T in place (lower) */
    for (int j = 0; j < P; ++j) {
        double d = A[(size_t)j * P + j];
        for (int k = 0; k < j; ++k) d -= A[(size_t)j * P + k] * A[(size_t)j * P + k] * A[(size_t)k * P + k];
        if (!(d > 0.0)) return -1;
        A[(size_t)j * P + j] = d;
        for (int i = j + 1; i < P; ++i) {
            double s = A[(size_t)i * P + j];
            for (int k = 0; k < j; ++k) s -= A[(size_t)i * P + k] * A[(size_t)j * P + k] * A[(size_t)k * P + k];
            A[(size_t)i * P + j] = s / d;
        }
    }
    return 0;
}

/* x (P x 3 interleaved) = A^-1 b */
static void solve3(const geo_t *g, const double *b, double *x) {
    const int P = g->P;
    const double *A = g->A;
    memcpy(x, b, sizeof(double) * 3 * P);
    for (int i = 0; i < P; ++i)
        for (int k = 0; k < i; ++k)
            for (int r = 0; r < 3; ++r) x[3 * i + r] -= A[(size_t)i * P + k] * x[3 * k + r];
    for (int i = 0; i < P; ++i)
        for (int r = 0; r < 3; ++r) x[3 * i + r] /= A[(size_t)i * P + i];
    for (int i = P - 1; i >= 0; --i)
        for (int k = i + 1; k < P; ++k)
            for (int r = 0; r < 3; ++r) x[3 * i + r] -= A[(size_t)k * P + i] * x[3 * k + r];
}

/* Dx of all hard rows (and, for GS, the soft rows = the points themselves) */
static void compute_dx(const geo_t *g, const double *x, double *dx) {
    for (int c = 0; c < g->nh; ++c)
        transform(g->type[c], g->idx + g->ptr[c], g->ptr[c + 1] - g->ptr[c], x, dx + 3 * g->col0[c]);
    if (!g->use_alm)
        for (int i = 0; i < g->ns; ++i) memcpy(dx + 3 * (g->zc + i), x + 3 * g->soft_pt[i], 3 * sizeof(double));
}

/* rhs = rhs_fixed + rho D^T (z - u) [+ w * closest for ALM] ; then solve */
static void x_update(const geo_t *g, const double *z, const double *u, const double *cp, double *rhs, double *x) {
    const int P = g->P;
    memcpy(rhs, g->rhs_fixed, sizeof(double) * 3 * P);
    int pts[MAXK];
    double coef[MAXK];
    for (int c = 0; c < g->nh; ++c) {
        const int kc = n_cols(g->type[c], g->ptr[c + 1] - g->ptr[c]);
        for (int j = 0; j < kc; ++j) {
            const int n = hard_row(g, c, j, pts, coef);
            const int col = g->col0[c] + j;
            for (int a = 0; a < n; ++a)
                for (int r = 0; r < 3; ++r) rhs[3 * pts[a] + r] += g->rho * coef[a] * (z[3 * col + r] - u[3 * col + r]);
        }
    }
    for (int i = 0; i < g->ns; ++i) {
        const int pt = g->soft_pt[i];
        if (g->use_alm) {
            for (int r = 0; r < 3; ++r) rhs[3 * pt + r] += g->soft_w * cp[3 * i + r];
        } else {
            const int col = g->zc + i;
            for (int r = 0; r < 3; ++r) rhs[3 * pt + r] += g->rho * (z[3 * col + r] - u[3 * col + r]);
        }
    }
    solve3(g, rhs, x);
}

/* ALMGeometrySolver<3>::solve_ADMM */
static int solve_alm(geo_t *g, const double *init_x, int max_iter, int m, double *x_out, double *hist) {
    const int P = g->P, NU = 3 * g->zc, N = NU + 3 * P;
    double *cur = calloc(N, sizeof(double)), *def = calloc(N, sizeof(double)), *nw = calloc(N, sizeof(double));
    double *dx = calloc(NU + 3, sizeof(double)), *prev = calloc(NU + 3, sizeof(double)), *z = calloc(NU + 3, sizeof(double));
    double *cp = calloc(3 * (g->ns + 1), sizeof(double)), *rhs = calloc(3 * P, sizeof(double)), *v = calloc(3 * MAXK, sizeof(double));
    memcpy(cur + NU, init_x, sizeof(double) * 3 * P);
    memcpy(def + NU, init_x, sizeof(double) * 3 * P);
    void *aa = NULL;
    if (m > 0) {
        aa = port_aa_new(m, N, N);
        port_aa_init(aa, cur);
    }
    int iter = 0, reset = 0;
    double prev_res = DBL_MAX;
    for (int turn = 0; iter < max_iter && turn < 8 * max_iter + 16; ++turn) {
        const double *cu = cur, *cx = cur + NU;
        compute_dx(g, cx, dx);
        memcpy(prev, dx, sizeof(double) * NU);
        /* z update: hard projections of Dx + u, soft closest points of the points themselves */
        for (int c = 0; c < g->nh; ++c) {
            const int kc = n_cols(g->type[c], g->ptr[c + 1] - g->ptr[c]), o = 3 * g->col0[c];
            for (int j = 0; j < 3 * kc; ++j) v[j] = dx[o + j] + cu[o + j];
            project(g->type[c], v, kc, g->param + 4 * c, z + o);
        }
        for (int i = 0; i < g->ns; ++i) closest_point(g->V, g->T, g->nt, cx + 3 * g->soft_pt[i], cp + 3 * i);
        x_update(g, z, cu, cp, rhs, nw + NU);
        compute_dx(g, nw + NU, dx);
        double r1 = 0, r2 = 0;
        for (int k = 0; k < NU; ++k) {
            nw[k] = cu[k] + dx[k] - z[k];
            r1 += (dx[k] - z[k]) * (dx[k] - z[k]);
            r2 += (dx[k] - prev[k]) * (dx[k] - prev[k]);
        }
        const double res = r1 + r2;
        const int accept = (m <= 0) || reset || res < prev_res;
        if (accept) {
            memcpy(def, nw, sizeof(double) * N);
            hist[iter++] = res;
            prev_res = res;
            reset = 0;
            if (aa)
                port_aa_compute(aa, nw, cur);
            else
                memcpy(cur, nw, sizeof(double) * N);
        } else {
            memcpy(cur, def, sizeof(double) * N);
            reset = 1;
            if (aa) port_aa_reset(aa, cur);
        }
    }
    memcpy(x_out, def + NU, sizeof(double) * 3 * P);
    if (aa) port_aa_free(aa);
    free(cur); free(def); free(nw); free(dx); free(prev); free(z); free(cp); free(rhs); free(v);
    return iter;
}

/* GeometrySolver<3>::solve_ADMM (older variant) */
static void gs_z_update(const geo_t *g, const double *dx, const double *u, double *z, double *v) {
    for (int c = 0; c < g->nh; ++c) {
        const int kc = n_cols(g->type[c], g->ptr[c + 1] - g->ptr[c]), o = 3 * g->col0[c];
        for (int j = 0; j < 3 * kc; ++j) v[j] = dx[o + j] + u[o + j];
        project(g->type[c], v, kc, g->param + 4 * c, z + o);
    }
    const double a = g->rho / (g->soft_w + g->rho);
    for (int i = 0; i < g->ns; ++i) {
        const int o = 3 * (g->zc + i);
        double q[3], c[3];
        for (int r = 0; r < 3; ++r) q[r] = dx[o + r] + u[o + r];
        closest_point(g->V, g->T, g->nt, q, c);
        for (int r = 0; r < 3; ++r) z[o + r] = q[r] * a + c[r] * (1 - a);
    }
}
static double gs_residual(const double *dx, const double *z, int n) {
    double s = 0;
    for (int k = 0; k < n; ++k) s += (dx[k] - z[k]) * (dx[k] - z[k]);
    return sqrt(s);
}
static int solve_gs(geo_t *g, const double *init_x, int max_iter, int m, double *x_out, double *hist) {
    const int P = g->P, NU = 3 * g->zc_all, N = NU + 3 * P;
    double *cur = calloc(N, sizeof(double)), *def = calloc(N, sizeof(double));
    double *dx = calloc(NU + 3, sizeof(double)), *z = calloc(NU + 3, sizeof(double)), *rhs = calloc(3 * P, sizeof(double));
    double *v = calloc(3 * MAXK, sizeof(double));
    memcpy(cur + NU, init_x, sizeof(double) * 3 * P);
    memcpy(def + NU, init_x, sizeof(double) * 3 * P);
    /* ADMM_init_variables: one warm-up turn */
    compute_dx(g, cur + NU, dx);
    gs_z_update(g, dx, cur, z, v);
    x_update(g, z, cur, NULL, rhs, def + NU);
    compute_dx(g, def + NU, dx);
    for (int k = 0; k < NU; ++k) def[k] = cur[k] + dx[k] - z[k];
    memcpy(cur, def, sizeof(double) * N);
    void *aa = NULL;
    if (m > 0) {
        aa = port_aa_new(m, N, NU);
        port_aa_init(aa, cur);
    }
    int iter = 0;
    double prev_res = DBL_MAX;
    while (iter < max_iter) {
        gs_z_update(g, dx, cur, z, v);
        double res = gs_residual(dx, z, NU);
        if (m > 0 && res > prev_res) { /* need_reset: back to the un-accelerated iterate */
            double *t = cur; cur = def; def = t;
            port_aa_replace(aa, cur);
            compute_dx(g, cur + NU, dx);
            gs_z_update(g, dx, cur, z, v);
            res = gs_residual(dx, z, NU);
        }
        hist[iter++] = res;
        if (iter >= max_iter) break;
        prev_res = res;
        x_update(g, z, cur, NULL, rhs, def + NU);
        compute_dx(g, def + NU, dx);
        for (int k = 0; k < NU; ++k) def[k] = cur[k] + dx[k] - z[k];
        if (aa) {
            double *acc = malloc(sizeof(double) * N);
            port_aa_compute(aa, def, acc);
            memcpy(cur, acc, sizeof(double) * N);
            free(acc);
        } else {
            double *t = cur; cur = def; def = t;
        }
        compute_dx(g, cur + NU, dx);
    }
    memcpy(x_out, cur + NU, sizeof(double) * 3 * P);
    if (aa) port_aa_free(aa);
    free(cur); free(def); free(dx); free(z); free(rhs); free(v);
    return iter;
}

int port_geo_solve(void *p, const double *init_x, int max_iter, int m, double *x_out, double *hist) {
    geo_t *g = p;
    return g->use_alm ? solve_alm(g, init_x, max_iter, m, x_out, hist) : solve_gs(g, init_x, max_iter, m, x_out, hist);
}

/* ---- remaining element types ---------------------------------------------- */
/* TriEnergyTerm::prox; variant 0 = xzu, 1 = hard. Thin SVD of the 3x2 block by one Hestenes rotation. */
void port_tri_prox(int variant, double *z, int n, double lmin, double lmax) {
    for (int i = 0; i < n; ++i) {
        double *F = z + 6 * i, *f1 = F, *f2 = F + 3;
        const double a = dot3(f1, f1), c = dot3(f2, f2), b = dot3(f1, f2);
        double cs = 1.0, sn = 0.0;
        if (fabs(b) > DBL_MIN) {
            const double zeta = (c - a) / (2.0 * b);
            const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            cs = 1.0 / sqrt(1.0 + t * t);
            sn = cs * t;
        }
        double g1[3], g2[3], u1[3], u2[3];
        for (int k = 0; k < 3; ++k) {
            g1[k] = cs * f1[k] - sn * f2[k];
            g2[k] = sn * f1[k] + cs * f2[k];
        }
        const double s1 = norm3(g1), s2 = norm3(g2);
        for (int k = 0; k < 3; ++k) {
            u1[k] = g1[k] / s1;
            u2[k] = g2[k] / s2;
        }
        const double v11 = cs, v21 = -sn, v12 = sn, v22 = cs;
        const int check = lmin > 0.0 || lmax < 99.0;
        double out[6];
        if (variant == 0) {
            for (int k = 0; k < 3; ++k) {
                out[k] = 0.5 * ((u1[k] * v11 + u2[k] * v12) + F[k]);
                out[3 + k] = 0.5 * ((u1[k] * v21 + u2[k] * v22) + F[3 + k]);
            }
            if (check) {
                const double l0 = norm3(out), l1 = norm3(out + 3);
                double f0 = 1.0, f1s = 1.0;
                if (l0 < lmin) f0 *= lmin / l0;
                if (l1 < lmin) f1s *= lmin / l1;
                if (l0 > lmax) f0 *= lmax / l0;
                if (l1 > lmax) f1s *= lmax / l1;
                for (int k = 0; k < 3; ++k) {
                    out[k] *= f0;
                    out[3 + k] *= f1s;
                }
            }
        } else {
            double sa = (1.0 + s1) / 2.0, sb = (1.0 + s2) / 2.0;
            if (check) {
                const double l0 = sa, l1 = sb;
                if (l0 < lmin) sa = lmin;
                if (l1 < lmin) sb = lmin;
                if (l0 > lmax) sa = lmax;
                if (l1 > lmax) sb = lmax;
            }
            for (int k = 0; k < 3; ++k) {
                out[k] = sa * u1[k] * v11 + sb * u2[k] * v12;
                out[3 + k] = sa * u1[k] * v21 + sb * u2[k] * v22;
            }
        }
        memcpy(F, out, sizeof(out));
    }
}

/* Collision::prox over analytic passive objects; types 0 Floor, 1 SlideFloor, 2 Sphere, 3 PlaneAndHalfSphere,
 * 4 Cylinder; 7 parameters per object {cx, cy, cz, nx, ny, nz, radius} (Floor: cx = height) */
void port_collision_prox(int n_objs, const int *types, const double *prm, double *z, int n) {
    for (int i = 0; i < n; ++i) {
        double *x = z + 3 * i, best = DBL_MAX, pt[3] = {0, 0, 0};
        for (int j = 0; j < n_objs; ++j) {
            const double *q = prm + 7 * j, *c = q, rad = q[6];
            double dx, cand[3];
            if (types[j] == 0) {
                dx = x[1] - q[0];
                cand[0] = x[0]; cand[1] = q[0]; cand[2] = x[2];
            } else if (types[j] == 1) {
                double nrm[3];
                normalized(q + 3, nrm);
                dx = (x[0] - c[0]) * nrm[0] + (x[1] - c[1]) * nrm[1] + (x[2] - c[2]) * nrm[2];
                for (int k = 0; k < 3; ++k) cand[k] = x[k] - dx * nrm[k];
            } else if (types[j] == 2) {
                double dir[3] = {x[0] - c[0], x[1] - c[1], x[2] - c[2]}, u[3];
                dx = norm3(dir) - rad;
                normalized(dir, u);
                for (int k = 0; k < 3; ++k) cand[k] = c[k] + u[k] * rad;
            } else if (types[j] == 3) {
                const double px = x[0] - c[0], pz = x[2] - c[2];
                if (sqrt(px * px + pz * pz) - rad > 0.0) {
                    dx = x[1] - c[1];
                    cand[0] = x[0]; cand[1] = c[1]; cand[2] = x[2];
                } else {
                    double dir[3] = {x[0] - c[0], x[1] - c[1], x[2] - c[2]}, u[3];
                    dx = (x[1] - c[1] > 0.0) ? norm3(dir) + rad : rad - norm3(dir);
                    normalized(dir, u);
                    for (int k = 0; k < 3; ++k) cand[k] = c[k] + u[k] * rad;
                }
            } else {
                double dir[3] = {x[0] - c[0], x[1] - c[1], 0.0 - c[2]}, u[3];
                dx = norm3(dir) - rad;
                normalized(dir, u);
                for (int k = 0; k < 3; ++k) cand[k] = c[k] + u[k] * rad + (k == 2 ? x[2] : 0.0);
            }
            if (dx > best) continue; /* `if (dx > p.dx) return;` */
            best = dx;
            memcpy(pt, cand, sizeof(pt));
        }
        if (best < 0.0) memcpy(x, pt, sizeof(pt));
    }
}

/* ---- hyper-elastic tet prox ---------------------------------------------------
 * HyperElasticTet::prox (xzu/src/TetEnergyTerm.cpp:171-183): argmin_z vol * (Psi(z) + k/2 |z - v|^2) from z0 = v by
 * mcl::optlib::LBFGS<double,9> (deps/mcloptlib/include/MCL/LBFGS.hpp:205-305: m = 6, epsilon 1e-6, delta 1e-16,
 * Armijo backtracking :135-203 with ftol 1e-4, dec 0.5, at most 100 iterations).
 * Psi: NeoHookeanTet::NHProx :221-267 (material 1), StVKTet::StVKProx :272-319 (material 2). F column-major. */
typedef struct { double mu, lambda, k, vol; int material; } hyper_t;

static double det3c(const double *m) {
    return m[0] * (m[4] * m[8] - m[7] * m[5]) - m[3] * (m[1] * m[8] - m[7] * m[2]) + m[6] * (m[1] * m[5] - m[4] * m[2]);
}
static void inv3c(const double *m, double *inv) { /* Eigen's cofactor inverse */
#define HM(r, c) m[(c) * 3 + (r)]
#define HCOF(i, j) (HM(((i) + 1) % 3, ((j) + 1) % 3) * HM(((i) + 2) % 3, ((j) + 2) % 3) - HM(((i) + 1) % 3, ((j) + 2) % 3) * HM(((i) + 2) % 3, ((j) + 1) % 3))
    const double c0 = HCOF(0, 0), c1 = HCOF(1, 0), c2 = HCOF(2, 0);
    const double id = 1.0 / ((c0 * HM(0, 0) + c1 * HM(1, 0)) + c2 * HM(2, 0));
    inv[0] = c0 * id; inv[3] = c1 * id; inv[6] = c2 * id;
    inv[1] = HCOF(0, 1) * id; inv[4] = HCOF(1, 1) * id; inv[7] = HCOF(2, 1) * id;
    inv[2] = HCOF(0, 2) * id; inv[5] = HCOF(1, 2) * id; inv[8] = HCOF(2, 2) * id;
#undef HCOF
#undef HM
}
static void ftf(const double *F, double *C) {
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) C[j * 3 + i] = F[i * 3] * F[j * 3] + F[i * 3 + 1] * F[j * 3 + 1] + F[i * 3 + 2] * F[j * 3 + 2];
}
static double psi(const hyper_t *P, const double *F) {
    double C[9];
    ftf(F, C);
    if (P->material == 1) {
        const double J = det3c(F), I1 = C[0] + C[4] + C[8], l3 = log(J * J);
        return 0.5 * P->mu * (I1 - l3 - 3.0) + 0.125 * P->lambda * l3 * l3;
    }
    double E[9], ee = 0.0;
    for (int k = 0; k < 9; ++k) E[k] = 0.5 * (C[k] - ((k % 4 == 0) ? 1.0 : 0.0));
    const double tr = E[0] + E[4] + E[8];
    for (int j = 0; j < 3; ++j) ee += E[j * 3] * E[j * 3] + E[j * 3 + 1] * E[j * 3 + 1] + E[j * 3 + 2] * E[j * 3 + 2];
    return P->mu * ee + 0.5 * P->lambda * tr * tr;
}
static void dpsi(const hyper_t *P, const double *F, double *G) {
    if (P->material == 1) {
        double Fi[9];
        inv3c(F, Fi);
        const double lj = P->lambda * log(det3c(F));
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) {
                const double fit = Fi[r * 3 + c];
                G[c * 3 + r] = P->mu * (F[c * 3 + r] - fit) + lj * fit;
            }
        return;
    }
    double C[9], S[9];
    ftf(F, C);
    for (int k = 0; k < 9; ++k) C[k] = 0.5 * (C[k] - ((k % 4 == 0) ? 1.0 : 0.0));
    const double tr = C[0] + C[4] + C[8];
    for (int k = 0; k < 9; ++k) S[k] = 2.0 * P->mu * C[k] + ((k % 4 == 0) ? P->lambda * tr : 0.0);
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) G[c * 3 + r] = F[r] * S[c * 3] + F[3 + r] * S[c * 3 + 1] + F[6 + r] * S[c * 3 + 2];
}
static double value_grad(const hyper_t *P, const double *v, const double *x, double *grad) {
    dpsi(P, x, grad);
    double q = 0.0;
    for (int i = 0; i < 9; ++i) {
        const double d = v[i] - x[i];
        q += d * d;
        grad[i] = P->vol * (grad[i] + P->k * (x[i] - v[i]));
    }
    return P->vol * (psi(P, x) + 0.5 * P->k * q);
}
static double dot9(const double *a, const double *b) {
    double s = 0.0;
    for (int i = 0; i < 9; ++i) s += a[i] * b[i];
    return s;
}
static void prox_lbfgs(const hyper_t *P, double *z) {
    enum { M = 6 };
    const double epsilon = 1e-6, delta = 1e-16, ftol = 1e-4, min_step = 1e-20, max_step = 1e+20;
    double v[9], x[9], xp[9], grad[9], gradp[9], drt[9], s[M][9], y[M][9], ys[M], alpha[M];
    for (int i = 0; i < 9; ++i) v[i] = x[i] = z[i];
    double fx = value_grad(P, v, x, grad), fx_prev = fx;
    if (sqrt(dot9(grad, grad)) <= epsilon * fmax(sqrt(dot9(x, x)), 1.0)) return;
    for (int i = 0; i < 9; ++i) drt[i] = -grad[i];
    double step = 1.0 / sqrt(dot9(drt, drt));
    int k = 1, end = 0;
    for (;;) {
        memcpy(xp, x, sizeof(x));
        memcpy(gradp, grad, sizeof(grad));
        const double fx_init = fx, dg_test = ftol * dot9(grad, drt);
        for (int it = 0; it < 2000; ++it) {
            for (int i = 0; i < 9; ++i) x[i] = xp[i] + step * drt[i];
            fx = value_grad(P, v, x, grad);
            if (!(fx > fx_init + step * dg_test)) break;
            if (step < min_step || step > max_step) break;
            step *= 0.5;
        }
        if (sqrt(dot9(grad, grad)) <= epsilon * fmax(sqrt(dot9(x, x)), 1.0)) break;
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
    memcpy(z, x, sizeof(x));
}
/* n column-major 3x3 blocks in place; grad (may be NULL) receives vol * dPsi/dF of the INPUT blocks; k = bulk modulus */
void port_tet_prox_hyper(int material, double mu, double lambda, double vol, double *z, double *grad, int n) {
    hyper_t P = {mu, lambda, lambda + (2.0 / 3.0) * mu, vol, material};
    for (int i = 0; i < n; ++i) {
        if (grad) {
            dpsi(&P, z + 9 * i, grad + 9 * i);
            for (int q = 0; q < 9; ++q) grad[9 * i + q] *= vol;
        }
        prox_lbfgs(&P, z + 9 * i);
    }
}
