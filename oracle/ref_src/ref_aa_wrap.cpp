// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product; never linked by it.
//
// C wrapper around the UNMODIFIED reference AndersonAcceleration classes, compiled
// twice by oracle/Makefile into oracle/_ref/libref_aa.so:
//   -DREF_AA_H -I admm_anderson_hard_zxu/src   (variant H, hard/src/AndersonAcceleration.h:38-212;
//                                                byte-identical to Geometry/AndersonAcceleration.h)
//   -DREF_AA_X -I admm_anderson_xzu/src        (variant X, xzu/src/AndersonAcceleration.h:39-295)
// The class is renamed per translation unit so both can live in one library.
#ifdef REF_AA_H
#define AndersonAcceleration RefAndersonH
#else
#define AndersonAcceleration RefAndersonX
#endif
#include "AndersonAcceleration.h"
#include <cstring>

extern "C" {

#ifdef REF_AA_H
void *ref_aa_h_new(int m, int total_dim, int effective_dim) {
    return new AndersonAcceleration(m, total_dim, effective_dim);
}
void ref_aa_h_free(void *p) { delete static_cast<AndersonAcceleration *>(p); }
void ref_aa_h_init(void *p, const double *u, int n) {
    VectorX v = Eigen::Map<const VectorX>(u, n);
    static_cast<AndersonAcceleration *>(p)->init(v);
}
void ref_aa_h_reset(void *p, const double *u, int n) {
    VectorX v = Eigen::Map<const VectorX>(u, n);
    static_cast<AndersonAcceleration *>(p)->reset(v);
}
void ref_aa_h_replace(void *p, const double *u, int n) {
    VectorX v = Eigen::Map<const VectorX>(u, n);
    static_cast<AndersonAcceleration *>(p)->replace(v);
}
void ref_aa_h_compute(void *p, const double *g, double *accel, int n) {
    VectorX gv = Eigen::Map<const VectorX>(g, n), out(n);
    static_cast<AndersonAcceleration *>(p)->compute(gv, out);
    memcpy(accel, out.data(), n * sizeof(double));
}
// two-block forms (effective block first)
void ref_aa_h_init2(void *p, const double *u1, int n1, const double *u2, int n2) {
    VectorX a = Eigen::Map<const VectorX>(u1, n1), b = Eigen::Map<const VectorX>(u2, n2);
    static_cast<AndersonAcceleration *>(p)->init(a, b);
}
void ref_aa_h_compute2(void *p, const double *g1, int n1, const double *g2, int n2, double *o1, double *o2) {
    VectorX a = Eigen::Map<const VectorX>(g1, n1), b = Eigen::Map<const VectorX>(g2, n2), oa(n1), ob(n2);
    static_cast<AndersonAcceleration *>(p)->compute(a, b, oa, ob);
    memcpy(o1, oa.data(), n1 * sizeof(double));
    memcpy(o2, ob.data(), n2 * sizeof(double));
}

// Eigen::CompleteOrthogonalDecomposition<MatrixXX>::solve as compute_impl uses it
// (hard/src/AndersonAcceleration.h:193-196). M is m x m column-major.
void ref_cod_solve(int m, const double *M, const double *rhs, double *out) {
    MatrixXX A = Eigen::Map<const MatrixXX>(M, m, m);
    VectorX b = Eigen::Map<const VectorX>(rhs, m);
    Eigen::CompleteOrthogonalDecomposition<MatrixXX> cod;
    cod.compute(A);
    VectorX x = cod.solve(b);
    memcpy(out, x.data(), m * sizeof(double));
}
int ref_cod_rank(int m, const double *M) {
    MatrixXX A = Eigen::Map<const MatrixXX>(M, m, m);
    Eigen::CompleteOrthogonalDecomposition<MatrixXX> cod;
    cod.compute(A);
    return (int)cod.rank();
}
#else
void *ref_aa_x_new() { return new AndersonAcceleration(); }
void ref_aa_x_free(void *p) { delete static_cast<AndersonAcceleration *>(p); }
void ref_aa_x_init(void *p, int m, int d, const double *g0) {
    VectorX v = Eigen::Map<const VectorX>(g0, d);
    static_cast<AndersonAcceleration *>(p)->init(m, d, v);
}
void ref_aa_x_replace(void *p, const double *g, int d) {
    VectorX v = Eigen::Map<const VectorX>(g, d);
    static_cast<AndersonAcceleration *>(p)->replace(v);
}
void ref_aa_x_compute(void *p, double *curr_g, const double *g, int d) {
    VectorX gv = Eigen::Map<const VectorX>(g, d), out(d);
    static_cast<AndersonAcceleration *>(p)->compute(out, gv);
    memcpy(curr_g, out.data(), d * sizeof(double));
}
#endif

}  // extern "C"
