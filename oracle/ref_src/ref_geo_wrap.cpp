// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product; never linked by it.
//
// C wrapper around the UNMODIFIED reference Geometry solver classes (Geometry/ALMGeometrySolver.h,
// GeometrySolver.h, Constraint.h, LinearRegularization.h, SPDSolver.h, igl::AABB), compiled by
// oracle/Makefile into oracle/_ref/libref_geo.so.  Constraints are built from plain arrays (the
// recipes of PlanarityOpt.cpp:147-246 / WireMeshOpt.cpp:253-289 restated by the caller) so that the
// product and the reference see the same constraint lists.
#include <cstring>
#include <memory>
#include <vector>

#include "ALMGeometrySolver.h"
#include "GeometrySolver.h"

namespace {
struct Handle {
    ALMGeometrySolver<3> alm;
    GeometrySolver<3> gs;
    bool use_alm = true;
};
struct ProbeEdge : EdgeLengthConstraint<3> { using EdgeLengthConstraint<3>::EdgeLengthConstraint; };
}  // namespace

extern "C" {

void *ref_geo_new(int use_alm) {
    Handle *h = new Handle();
    h->use_alm = use_alm != 0;
    return h;
}
void ref_geo_free(void *p) { delete static_cast<Handle *>(p); }

static void add_c(Handle *h, Constraint<3> *c, bool soft) {
    if (h->use_alm) {
        if (soft) h->alm.add_soft_constraint(c); else h->alm.add_hard_constraint(c);
    } else {
        if (soft) h->gs.add_soft_constraint(c); else h->gs.add_hard_constraint(c);
    }
}
void ref_geo_add_plane(void *p, const int *idx, int k, double weight, int soft) {
    add_c(static_cast<Handle *>(p), new PlaneConstraint(std::vector<int>(idx, idx + k), weight), soft);
}
void ref_geo_add_edge(void *p, int i0, int i1, double weight, double len, int soft) {
    add_c(static_cast<Handle *>(p), new EdgeLengthConstraint<3>(i0, i1, weight, len), soft);
}
void ref_geo_add_angle(void *p, int tip, int s1, int s2, double weight, double amin, double amax, int soft) {
    add_c(static_cast<Handle *>(p), new AngleConstraint<3>(tip, s1, s2, weight, amin, amax), soft);
}
// ReferenceSurfceConstraint over points 0..n_points-1 (V: nv x 3 row-major, F: nf x 3 row-major)
void ref_geo_add_ref_surface(void *p, int n_points, double weight, const double *V, int nv, const int *F, int nf, int soft) {
    Matrix3X Vm(3, nv);
    Eigen::Matrix3Xi Fm(3, nf);
    for (int i = 0; i < nv; ++i) for (int k = 0; k < 3; ++k) Vm(k, i) = V[3 * i + k];
    for (int i = 0; i < nf; ++i) for (int k = 0; k < 3; ++k) Fm(k, i) = F[3 * i + k];
    add_c(static_cast<Handle *>(p), new ReferenceSurfceConstraint(n_points, weight, Vm, Fm), soft);
}
void ref_geo_add_relative_uniform_laplacian(void *p, const int *idx, int n, double weight, const double *ref_pts, int n_pts) {
    Handle *h = static_cast<Handle *>(p);
    Matrix3X R = Eigen::Map<const Matrix3X>(ref_pts, 3, n_pts);
    std::vector<int> ids(idx, idx + n);
    if (h->use_alm) h->alm.add_relative_uniform_laplacian(ids, weight, R); else h->gs.add_relative_uniform_laplacian(ids, weight, R);
}
void ref_geo_add_uniform_laplacian(void *p, const int *idx, int n, double weight) {
    Handle *h = static_cast<Handle *>(p);
    std::vector<int> ids(idx, idx + n);
    if (h->use_alm) h->alm.add_uniform_laplacian(ids, weight); else h->gs.add_uniform_laplacian(ids, weight);
}
void ref_geo_add_closeness(void *p, int idx, double weight, const double *target3) {
    Handle *h = static_cast<Handle *>(p);
    Vector3 t(target3[0], target3[1], target3[2]);
    if (h->use_alm) h->alm.add_closeness(idx, weight, t); else h->gs.add_closeness(idx, weight, t);
}
int ref_geo_setup(void *p, int n_points, double rho) {
    Handle *h = static_cast<Handle *>(p);
    return (h->use_alm ? h->alm.setup_ADMM(n_points, rho) : h->gs.setup_ADMM(n_points, rho)) ? 0 : -1;
}
// init_x: 3 x n_points column-major (xyz per point). Returns the number of logged iterations.
int ref_geo_solve(void *p, const double *init_x, int n_points, int max_iter, int anderson_m) {
    Handle *h = static_cast<Handle *>(p);
    Matrix3X x0 = Eigen::Map<const Matrix3X>(init_x, 3, n_points);
    if (h->use_alm) {
        h->alm.solve_ADMM(x0, 1e-8, max_iter, anderson_m);
        return (int)h->alm.function_values_.size();
    }
    h->gs.solve_ADMM(x0, 1e-8, max_iter, anderson_m);
    return (int)h->gs.function_values_.size();
}
void ref_geo_history(void *p, double *values, double *secs) {
    Handle *h = static_cast<Handle *>(p);
    const std::vector<double> &f = h->use_alm ? h->alm.function_values_ : h->gs.function_values_;
    const std::vector<double> &t = h->use_alm ? h->alm.elapsed_time_ : h->gs.elapsed_time_;
    memcpy(values, f.data(), f.size() * sizeof(double));
    if (secs) memcpy(secs, t.data(), t.size() * sizeof(double));
}
void ref_geo_solution(void *p, double *x, int n_points) {
    Handle *h = static_cast<Handle *>(p);
    const Matrix3X &s = h->use_alm ? h->alm.get_solution() : h->gs.get_solution();
    memcpy(x, s.data(), sizeof(double) * 3 * n_points);
}

// ---- unit-level access -------------------------------------------------------------------------
// project one constraint on `pts` (3 x k column-major, already transformed): kind 0 plane, 1 edge, 2 angle
void ref_geo_project(int kind, int k, const double *pts, double *out, double a, double b) {
    Matrix3X in = Eigen::Map<const Matrix3X>(pts, 3, k), o = in;
    std::vector<Triplet> trip;
    int idO = 0;
    if (kind == 0) {
        std::vector<int> ids(k); for (int i = 0; i < k; ++i) ids[i] = i;
        PlaneConstraint c(ids, 1.0); c.add_constraint(false, trip, idO); c.project(in, o);
    } else if (kind == 1) {
        EdgeLengthConstraint<3> c(0, 1, 1.0, a); c.add_constraint(false, trip, idO); c.project(in, o);
    } else {
        AngleConstraint<3> c(0, 1, 2, 1.0, a, b); c.add_constraint(false, trip, idO); c.project(in, o);
    }
    memcpy(out, o.data(), sizeof(double) * 3 * k);
}
// closest points on a triangle mesh for n queries (igl::AABB::squared_distance, the batch form the
// ReferenceSurfceConstraint uses)
void ref_geo_closest_points(const double *V, int nv, const int *F, int nf, const double *Q, int nq, double *C, int *I, double *sqrD) {
    MatrixX3 Vm(nv, 3), Qm(nq, 3), Cm;
    Eigen::MatrixX3i Fm(nf, 3);
    for (int i = 0; i < nv; ++i) for (int k = 0; k < 3; ++k) Vm(i, k) = V[3 * i + k];
    for (int i = 0; i < nf; ++i) for (int k = 0; k < 3; ++k) Fm(i, k) = F[3 * i + k];
    for (int i = 0; i < nq; ++i) for (int k = 0; k < 3; ++k) Qm(i, k) = Q[3 * i + k];
    igl::AABB<MatrixX3, 3> tree;
    tree.init(Vm, Fm);
    VectorX d; Eigen::VectorXi idx;
    tree.squared_distance(Vm, Fm, Qm, d, idx, Cm);
    for (int i = 0; i < nq; ++i) { for (int k = 0; k < 3; ++k) C[3 * i + k] = Cm(i, k); I[i] = idx(i); sqrD[i] = d(i); }
}

}  // extern "C"
