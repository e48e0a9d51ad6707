// Geometry ADMM (planar-quad / wire-mesh optimisation): local projections, soft closest-point
// constraint, rhs assembly and residual + accept/reject logic of
//   ALMGeometrySolver<3>::solve_ADMM   Geometry/ALMGeometrySolver.h:163-283 (ops :404-461)
//   GeometrySolver<3>::solve_ADMM      Geometry/GeometrySolver.h:156-263   (ops :383-459)
//   Constraint<3> and subclasses       Geometry/Constraint.h:48-414
//   igl::AABB closest point            Geometry/external/igl/AABB.cpp, point_simplex_squared_distance.cpp:44-110
// One thread per constraint / per point. Points are 3 doubles (xyz) per column, as Eigen's Matrix3X.
#pragma once
#include "common.cuh"

namespace aaadmm {

enum { GEO_PLANE = 0, GEO_EDGE = 1, GEO_ANGLE = 2 };
constexpr int GEO_MAX_K = 16;   // max points of one plane constraint
constexpr int GEO_BLOCK = 128;

struct GeoConstraints {
    int n;                 // number of hard constraints
    const int *type;       // [n]
    const int *idx_ptr;    // [n+1] into idx
    const int *idx;        // point ids
    const int *col0;       // [n] first output column (idO_)
    const double *param;   // [n][4]: edge: target length; angle: min, max, cos(min), cos(max)
};

struct BvhNode {
    double lo[3], hi[3];
    int left, right;  // inner: child ids; leaf: left = -(first_tri + 1), right = tri count
};

struct GeoSoft {
    int n;                  // soft closest-point constraints (one point each)
    const int *point;       // [n]
    double weight;          // constraint weight (the reference stores its square root)
    const BvhNode *nodes;
    const int *tri_order;   // leaf triangle ids
    const double *tri;      // [n_tris][9] corner coordinates a, b, c
    int *last_tri;          // [n] closest triangle of the previous call (search bound), -1 initially
    const int *order;       // [n] thread j handles constraint order[j] (spatially sorted: coherent BVH descents), or null
};

// Geometry loop state that rides in SolveState's generic fields:
//   prev_prim = prev_residual, reject = reset flag, iter = accepted iterations, done = loop finished

// current x -> Dx (kept as prev_Dx), v = Dx + u, z = project(v)       (ALMGeometrySolver.h:200-205,425-435)
// zmu (may be null): also z - u, for launch_geo_rhs(z = zmu, u = null) of the same turn
void launch_geo_local(cudaStream_t s, const GeoConstraints &C, const double *x, const double *u, double *prev_dx,
                      double *z, double *zmu, const SolveState *st);
// closest point on the reference surface for every soft point          (:436-439, Constraint.h:340-345,377-383)
void launch_geo_soft(cudaStream_t s, const GeoSoft &S, const double *x, double *cp, const SolveState *st);
// rhs = rhs_fixed + rho D_hard^T (z - u) + D_soft^T z_soft, in elimination order   (:442-450)
void launch_geo_rhs(cudaStream_t s, int n_points, const int64_t *dt_ptr, const int *dt_col, const double *dt_val,
                    const double *z, const double *u, const double *rhs_fixed, const int *soft_of_point,
                    double soft_weight, const double *cp, const int *iperm, double *W, const SolveState *st);
// new_u = u + D new_x - z; r = |D new_x - z|^2 + |D new_x - prev_Dx|^2; accept / reject / count   (:211-263)
void launch_geo_u_resid(cudaStream_t s, const GeoConstraints &C, const double *x_new, const double *u, const double *z,
                        const double *prev_dx, double *u_new, SolveState *st, double *partials, double *hist,
                        int accel);
// on accept without acceleration: current = default = new; on reject: current = default
void launch_geo_select(cudaStream_t s, double *cur, double *def, const double *nw, int64_t n, const SolveState *st,
                       int accel);
// ---- GeometrySolver<3> (older variant, Geometry/GeometrySolver.h:156-263, ops :383-459) ----------
// Dx for hard rows (transform) and soft rows (the point itself), rows [0, n_hard cols) then one per soft point
void launch_gs_dx(cudaStream_t s, const GeoConstraints &C, const GeoSoft &S, int zc_hard, const double *x, double *dx,
                  const SolveState *st, int when);
// z = project(Dx + u) (hard) / a (Dx+u) + (1-a) closest(Dx+u) (soft, a = rho/(w+rho)); residual |Dx - z|;
// mode 0: warm-up turn (no residual); mode 1: decide need_reset = accel && residual > prev (GeometrySolver.h:186-190); mode 2: redo after the swap
void launch_gs_z(cudaStream_t s, int mode, const GeoConstraints &C, const GeoSoft &S, int zc_hard, double rho,
                 const double *dx, const double *u, double *z, double *cp_scratch, SolveState *st, double *partials,
                 double *hist);
// current = default: when = 1 only after a rejected iterate (:192-197), when = 0 always (:243-246)
void launch_gs_take_default(cudaStream_t s, double *cur, const double *def, int64_t n, const SolveState *st, int when);
// default_u = current_u + Dx - z (:450-455)
void launch_gs_u(cudaStream_t s, const double *u_cur, const double *dx, const double *z, double *u_def, int64_t n,
                 const SolveState *st);

// unit-parity entry: project n_c constraints given already transformed columns (device arrays)
void launch_geo_project_only(const GeoConstraints &C, const double *v, double *z);
void launch_geo_closest_only(const GeoSoft &S, const double *q, double *cp, int *tri_out);

}  // namespace aaadmm
