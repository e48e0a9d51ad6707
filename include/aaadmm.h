/* aaadmm.h - C ABI of libaaadmm_b200.so: the B200 (sm_100a) implementation of AA-ADMM's
 * per-iteration hot path (local projections, pre-factored global solve, Anderson mixing,
 * safeguard).  Plain pointers and sizes only; every entry point returns 0 on success and a
 * negative value on failure (aaadmm_last_error() holds the message).  Nothing throws across
 * this boundary; the C++ mirror classes in aa-admm_b200/host translate to the reference's
 * throw / `return false` conventions.  All device state is owned by the handles.  One CUDA
 * stream per handle, no internal host threads, re-entrant across handles.
 *
 * The reference has no FFI layer (header-only / static library); each group below
 * names the reference interface it replaces.
 */
#ifndef AAADMM_H_
#define AAADMM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AAADMM_MAX_M 16

const char *aaadmm_last_error(void);
int aaadmm_device_count(void);          /* number of CUDA devices, <=0 if none / no driver   */
int aaadmm_set_device(int device);      /* cudaSetDevice for the calling host thread          */
int aaadmm_device_sms(void);

/* ------------------------------------------------------------------------------------------
 * Anderson acceleration.
 * Replaces class AndersonAcceleration:
 *   variant H  admm_anderson_hard_zxu/src/AndersonAcceleration.h:38-212 (== Geometry/AndersonAcceleration.h)
 *              ctor(m,total_dim,effective_dim) :40-56, replace :58-71, reset :73-91,
 *              compute :93-114, init :116-135, compute_impl :154-211
 *   variant X  admm_anderson_xzu/src/AndersonAcceleration.h: init(m,d,g0) :279-295,
 *              compute(curr_g,g) :138-200, replace(g) :51-54      (= variant H with total==effective)
 * Vectors are flat FP64; "host" calls copy in/out, "dev" calls take device pointers.
 * ------------------------------------------------------------------------------------------ */
typedef struct aaadmm_aa aaadmm_aa;
int aaadmm_aa_create(aaadmm_aa **out, int m, int64_t total_dim, int64_t effective_dim);
int aaadmm_aa_destroy(aaadmm_aa *aa);
int aaadmm_aa_init(aaadmm_aa *aa, const double *u, int64_t n);          /* host u */
int aaadmm_aa_reset(aaadmm_aa *aa, const double *u, int64_t n);
int aaadmm_aa_replace(aaadmm_aa *aa, const double *u, int64_t n);
int aaadmm_aa_compute(aaadmm_aa *aa, const double *g, double *accel_u, int64_t n);
int aaadmm_aa_init_dev(aaadmm_aa *aa, const double *d_u, int64_t n);    /* device u */
int aaadmm_aa_reset_dev(aaadmm_aa *aa, const double *d_u, int64_t n);
int aaadmm_aa_replace_dev(aaadmm_aa *aa, const double *d_u, int64_t n);
int aaadmm_aa_compute_dev(aaadmm_aa *aa, const double *d_g, double *d_accel_u, int64_t n);
/* iteration count since init/reset and next history column (iter_, col_idx_) */
int aaadmm_aa_state(aaadmm_aa *aa, int *iter, int *col);

/* ------------------------------------------------------------------------------------------
 * Sparse LDL^T apply.
 * Replaces LDLTSolver::solve (admm_anderson_xzu/src/LinearSolver.hpp:87-90) and
 * SimplicialLDLTSolver::solve (Geometry/SPDSolver.h:88-91).  The factor is computed ONCE on
 * the host (by the caller: Eigen::SimplicialLDLT matrixL()/vectorD()/permutationP() in the
 * reference, aa-admm_b200/host/sparse_ldlt in this repo) and handed over as:
 *   L  strictly lower triangle, CSC, row indices ascending per column, unit diagonal implied
 *   D  n pivots;  perm[new] = old  (P A P^T = L D L^T)
 * nrhs is 1 or 3 right-hand sides, interleaved (b[i*nrhs + r]).
 * ------------------------------------------------------------------------------------------ */
typedef struct aaadmm_ldlt aaadmm_ldlt;
int aaadmm_ldlt_create(aaadmm_ldlt **out, int n, const int64_t *Lp, const int *Li, const double *Lx,
                       const double *D, const int *perm, int nrhs);
/* The same object from the MATRIX instead of a finished factor: A (lower CSC incl. diagonal, rows ascending, original
 * numbering) with the pattern of L (Lp, Li as above, from a symbolic analysis: aa-admm_b200/host/sparse_ldlt:
 * ldlt_symbolic) and the ordering. The numeric factorisation runs on the device (multifrontal, FP64, no pivoting) and
 * replaces the numeric half of Eigen::SimplicialLDLT::compute (LinearSolver.hpp:79-84). aaadmm_ldlt_refactor loads
 * another matrix with the SAME pattern (a material sweep over one mesh) into the same object: no allocation, no
 * analysis, stream-ordered on the object's stream. Both fail (-1) on a zero or non-finite pivot. */
int aaadmm_ldlt_create_from_matrix(aaadmm_ldlt **out, int n, const int64_t *Ap, const int *Ai, const double *Ax,
                                   const int64_t *Lp, const int *Li, const int *perm, int nrhs);
int aaadmm_ldlt_refactor(aaadmm_ldlt *f, const double *Ax);
int aaadmm_ldlt_destroy(aaadmm_ldlt *f);
int aaadmm_ldlt_solve(aaadmm_ldlt *f, const double *b, double *x);          /* host vectors */
int aaadmm_ldlt_solve_dev(aaadmm_ldlt *f, const double *d_b, double *d_x);  /* device vectors */
/* stats[0..7] = n, blocks, levels, max_block, nnz_L, nnz_offblock, dense_diag_entries, bytes_per_solve */
/* Developer aid: with AAADMM_LDLT_TRACE=1 in the environment at creation, writes the per-task time stamps
 * (globaltimer ns: start, ring primed, dependencies met, done) of the last apply as CSV. */
int aaadmm_ldlt_dump_trace(aaadmm_ldlt *h, const char *path);
int aaadmm_ldlt_stats(aaadmm_ldlt *f, double *stats8);

/* ------------------------------------------------------------------------------------------
 * Tet-mesh ADMM step.
 * Replaces the loop of admm::Solver::step():
 *   hard_zxu ordering  admm_anderson_hard_zxu/src/Solver.cpp:74-226  (z -> x -> u, AA on (u,x))
 *   xzu ordering       admm_anderson_xzu/src/Solver.cpp:78-257       (x -> z -> u, AA on z)
 * with EnergyTerm::update_z/update_u/get_all_gradient (src/EnergyTerm.hpp:156-207) and
 * TetEnergyTerm::prox/get_gradient (src/TetEnergyTerm.cpp:101-123,156-165) as batched kernels.
 * Vertices are numbered free-first (reference free order), pinned vertices last.
 * ------------------------------------------------------------------------------------------ */
typedef struct aaadmm_tetscene aaadmm_tetscene;

typedef struct {
    int n_verts, n_free, n_tets;
    const int *tet;          /* 4*n_tets vertex ids in the free-first numbering              */
    const double *binv;      /* 9*n_tets column-major inverse rest edge matrices              */
    const double *weight;    /* n_tets ADMM weights sqrt(K vol)                                */
    const double *kvol;      /* n_tets K*vol (xzu gradient)                                    */
    const int *material;     /* n_tets: 0 linear, 1 neo-hookean, 2 stvk (may be NULL = linear) */
    const double *mu;        /* n_tets (hyper-elastic only, may be NULL)                       */
    const double *lambda;    /* n_tets (hyper-elastic only, may be NULL)                       */
    const double *mass_free; /* n_free scalar lumped masses                                    */
    const int64_t *inc_ptr;  /* n_free+1: incidence CSR over free vertices                     */
    const int *inc;          /* entries = contribution slots: tet*4+corner, triangles 4*n_tets+tri*3+corner,
                                collision terms 4*n_tets+3*n_tris+term                                */
    double rho_dt2;          /* penalty * dt^2 (hard) or dt^2 (xzu)                            */
    const double *volume;    /* n_tets rest volumes (hyper-elastic only, may be NULL)          */
    /* Triangle (cloth) terms of the same scene: TriEnergyTerm, admm_anderson_hard_zxu/src/TriEnergyTerm.cpp:29-105
     * (6 rows per triangle, F = [x1-x0, x2-x0] * rest_pose). hard_zxu ordering only. n_tris may be 0; a scene
     * may also consist of triangles only (n_tets = 0). */
    int n_tris;
    const int *tri;              /* 3*n_tris vertex ids in the free-first numbering                   */
    const double *tri_rest_pose; /* 4*n_tris column-major inverse rest matrices (2x2)                 */
    const double *tri_weight;    /* n_tris ADMM weights sqrt(K area)                                  */
    const double *tri_limit_min; /* n_tris strain limits Lame::limit_min (NULL = -100, no limiting)   */
    const double *tri_limit_max; /* n_tris strain limits Lame::limit_max (NULL = +100, no limiting)   */
    /* Collision terms: one Collision energy term per listed FREE vertex (Solver::set_collisions +
     * hard/src/Solver.cpp:386-392, CollisionEnergyTerm.hpp:40-91: 3 rows, D_i x = w x_idx) against the analytic
     * passive objects added with Solver::add_obstacle (PassiveObject.hpp:32-136). hard_zxu ordering only. */
    int n_collisions;
    const int *collision_vert;      /* n_collisions free-first vertex ids (< n_free)                  */
    const double *collision_weight; /* n_collisions weights (reference: sqrt(K_soft_rubber * 2))      */
    int n_obstacles;
    const int *obstacle_type;       /* n_obstacles AAADMM_PASSIVE_*                                   */
    const double *obstacle_prm;     /* n_obstacles x 7: {cx, cy, cz, nx, ny, nz, radius} (Floor: cx = y) */
} aaadmm_tetscene_desc;

#define AAADMM_ORDER_HARD_ZXU 0
#define AAADMM_ORDER_XZU 1

typedef struct {
    int ordering;       /* AAADMM_ORDER_*                                   */
    int admm_iters;     /* Settings::admm_iters                             */
    int anderson_m;     /* Settings::Anderson_m                             */
    int accel;          /* 0 NOACC, 1 ANDERSON                              */
    double eps;         /* break threshold on the combined residual (1e-20) */
    int log_comb_xzu;   /* xzu: also run the extra solve+local step that the reference uses only to
                           log the combined residual (xzu/src/Solver.cpp:217-233); 1 = as reference */
} aaadmm_step_opts;

typedef struct {
    int iters_logged;   /* rows written to the history arrays                          */
    int rejects;        /* rejected (reset) accelerated iterates                       */
    int broke_early;    /* combined residual fell below eps                            */
    float loop_ms;      /* device time of the iteration loop (CUDA events)             */
    float step_ms;      /* device time including uploads/downloads of this call        */
    int kernel_launches;
} aaadmm_step_result;

/* `factor` is borrowed (must outlive the scene). Either n == n_free with nrhs == 3 (A = Ahat (x) I3, one scalar factor
 * for the three axes) or n == 3*n_free with nrhs == 1: a factor of the full system in the reference's degree-of-freedom
 * order 3*vertex + axis, e.g. Eigen::SimplicialLDLT's own matrixL / vectorD / permutationP (LinearSolver.hpp:79-84). */
int aaadmm_tetscene_create(aaadmm_tetscene **out, const aaadmm_tetscene_desc *desc, aaadmm_ldlt *factor);
int aaadmm_tetscene_destroy(aaadmm_tetscene *s);
/* Another material on the same scene (a member of a material sweep; the factor object is refreshed separately with
 * aaadmm_ldlt_refactor): per-tet weight / K vol / mu / lambda (n_tets each; mu, lambda may be NULL for linear scenes),
 * per-triangle weight and strain limits (NULL when the scene has no triangles) and rho_dt2. Nothing is allocated;
 * element connectivity, rest shapes, incidence lists and all state buffers stay. */
int aaadmm_tetscene_update_material(aaadmm_tetscene *s, const double *weight, const double *kvol, const double *mu,
                                    const double *lambda, const double *tri_weight, const double *tri_limit_min,
                                    const double *tri_limit_max, double rho_dt2);
/* One time step of the ADMM loop.
 *   x_bar   3*n_free  predicted free positions x + dt v (host)
 *   x_pin   3*(n_verts-n_free) pinned positions (host)
 *   x_out   3*n_free  resulting free positions (host)
 *   hist_prim/hist_comb/hist_reject: admm_iters entries each (host, may be NULL) */
int aaadmm_tetscene_step(aaadmm_tetscene *s, const aaadmm_step_opts *opts, const double *x_bar,
                         const double *x_pin, double *x_out, double *hist_prim, double *hist_comb,
                         int *hist_reject, aaadmm_step_result *result);
/* Same loop with inputs already resident (the x_bar / x_pin of the last aaadmm_tetscene_step
 * call are reused) and nothing copied back: used to time the HBM-resident rate. */
int aaadmm_tetscene_step_resident(aaadmm_tetscene *s, const aaadmm_step_opts *opts, aaadmm_step_result *result);
/* Device time stamps of the last step's logged iterations: ms[i] = milliseconds from the start of the iteration loop
 * (after the warm start) to the moment iteration i was logged (%globaltimer in the CTA that writes the log), the
 * cumulative times the reference writes to ./result/residual-*.txt (hard/src/Solver.cpp:210-212). n <= rows logged. */
int aaadmm_tetscene_iteration_times(aaadmm_tetscene *s, double *ms, int n);
/* Debug/parity access: copies z (9*n_tets, reference layout: 9 consecutive doubles per tet)
 * and u of the last step to the host. Both are locals of the reference's Solver::step (hard/src/Solver.cpp:84-86) and
 * do not survive it there. One difference on a frame that ends through the break test `comb < 1e-20`
 * (hard/src/Solver.cpp:188-189): the reference leaves the loop before its last update of u, here the kernel that
 * evaluates the test has already written that update; x is the same. */
int aaadmm_tetscene_read_zu(aaadmm_tetscene *s, double *z, double *u);
/* Per-kernel device timings of the last profiled step (see aaadmm_tetscene_profile). */
#define AAADMM_NPROF 8
/* Runs `iters` iterations of the hard_zxu loop from the current state with CUDA events around
 * every phase; ms[AAADMM_NPROF] receives the average per-iteration time of
 * {update_z, rhs, ldlt_apply, update_u_resid, aa_pass1, aa_pass2, safeguard, total}. */
int aaadmm_tetscene_profile(aaadmm_tetscene *s, const aaadmm_step_opts *opts, int iters, float *ms);
/* Algorithmic bytes per launch of the same phases, for the roofline report. */
int aaadmm_tetscene_algo_bytes(aaadmm_tetscene *s, int anderson_m, double *bytes);

/* ------------------------------------------------------------------------------------------
 * Geometry ADMM (planar-quad and wire-mesh optimisation).
 * Replaces ALMGeometrySolver<3>::solve_ADMM (Geometry/ALMGeometrySolver.h:163-283) with the shipped
 * Constraint<3> subclasses as device batches (Geometry/Constraint.h): PlaneConstraint :396-414,
 * EdgeLengthConstraint :194-218, AngleConstraint :220-296, and the closest-point soft constraints
 * PointToRefSurfaceConstraint :328-349 / ReferenceSurfceConstraint :351-394 (igl::AABB).
 * setup_ADMM (:81-161) stays on the host: it builds rho*D_hard^T, the constant rhs and the P x P
 * system matrix, factors it once and hands the factor over (aaadmm_ldlt, nrhs = 3).
 * ------------------------------------------------------------------------------------------ */
typedef struct aaadmm_geo aaadmm_geo;

#define AAADMM_GEO_PLANE 0 /* MEAN_CENTERING, k output columns      */
#define AAADMM_GEO_EDGE 1  /* SUBTRACT_FIRST, 1 output column       */
#define AAADMM_GEO_ANGLE 2 /* SUBTRACT_FIRST, 2 output columns      */

/* Solver variant. ALM: ALMGeometrySolver<3> (soft rows outside z/u, residual |Dx-z|^2 + |Dx-Dx_prev|^2,
 * reject and re-run, solution = default_x). GS: the older GeometrySolver<3>
 * (Geometry/GeometrySolver.h:85-263): the soft closest-point rows are part of z/u (columns n_zcols ..
 * n_zcols+n_soft, unweighted, combined as a v + (1-a) closest(v), a = rho/(w+rho)), dt_* then also
 * carries those rows, the residual is |Dx-z|, a growing residual swaps back to the un-accelerated
 * iterate, every turn counts as an iteration and the solution is current_x. */
#define AAADMM_GEO_ALM 0
#define AAADMM_GEO_GS 1

typedef struct {
    int n_points;
    int n_hard;             /* hard constraints in insertion order (their output columns follow it)  */
    const int *type;        /* n_hard: AAADMM_GEO_*                                                  */
    const int *idx_ptr;     /* n_hard+1 offsets into idx                                             */
    const int *idx;         /* point ids (idI_)                                                      */
    const double *param;    /* 4 per constraint: edge {length}; angle {min, max, cos min, cos max}   */
    int n_zcols;            /* columns of z_hard / u                                                 */
    const int64_t *dt_ptr;  /* n_points+1: CSR of rho * D_hard^T (rows = points, cols = z columns)   */
    const int *dt_col;
    const double *dt_val;
    int n_soft;             /* closest-point soft constraints, one point each                        */
    const int *soft_point;  /* n_soft point ids                                                      */
    double soft_weight;     /* their weight (D_soft^T z_soft = weight * closest point)               */
    int n_ref_verts;        /* reference triangle mesh                                               */
    const double *ref_verts; /* 3 per vertex                                                         */
    int n_ref_tris;
    const int *ref_tris;    /* 3 per triangle                                                        */
    const double *rhs_fixed; /* 3 per point: L^T * regularisation targets                            */
    int variant;            /* AAADMM_GEO_ALM or AAADMM_GEO_GS                                        */
    double rho;             /* penalty parameter (GS: the soft combination needs it; dt_val has it folded in) */
} aaadmm_geo_desc;

/* `factor` (nrhs = 3, n = n_points) is borrowed. */
int aaadmm_geo_create(aaadmm_geo **out, const aaadmm_geo_desc *desc, aaadmm_ldlt *factor);
int aaadmm_geo_destroy(aaadmm_geo *g);
/* init_x / x_out: 3 per point; hist: max_iter combined residuals of the accepted iterations.
 * anderson_m <= 0 runs plain ADMM. result->iters_logged = accepted iterations, ->rejects = resets. */
int aaadmm_geo_solve(aaadmm_geo *g, const double *init_x, int max_iter, int anderson_m, double *x_out, double *hist,
                     aaadmm_step_result *result);
/* Per logged iteration of the last aaadmm_geo_solve: 1 if it follows a reset of the accelerator (the iterate before it
 * was rejected), the flag the reference's solvers keep in Anderson_reset_ (Geometry/ALMGeometrySolver.h:391). */
int aaadmm_geo_reset_flags(aaadmm_geo *g, int *flags, int n);
/* Device time stamps of the last aaadmm_geo_solve: ms[i] = milliseconds from the start of the loop to the moment logged
 * iteration i was accepted (%globaltimer in the CTA that writes the log): the elapsed_time_ column of the reference's
 * ./result/residual-*.txt (Geometry/ALMGeometrySolver.h:262-266). n <= iterations logged. */
int aaadmm_geo_iteration_times(aaadmm_geo *g, double *ms, int n);
/* unit parity: project `n` constraints of one type on already transformed columns (3 per column, host) */
int aaadmm_geo_project(int type, int n, int k, const double *cols, const double *param4, double *out);
/* unit parity: nearest points on a triangle mesh for nq host queries */
int aaadmm_geo_closest_points(const double *verts, int nv, const int *tris, int nt, const double *q, int nq,
                              double *closest, int *tri);

/* ------------------------------------------------------------------------------------------
 * Batched element kernels exposed for unit parity (same device code the step uses).
 *   prox_linear : TetEnergyTerm::prox on n column-major 3x3 blocks (in place, host array)
 *   grad_linear : out = F - U V^T  (TetEnergyTerm::get_gradient / (K vol))
 *   cod_solve   : Eigen COD least-squares solve of an m x m column-major system
 * ------------------------------------------------------------------------------------------ */
int aaadmm_tet_prox_linear(double *z, int64_t n);
int aaadmm_tet_f_minus_uvt(const double *z, double *out, int64_t n);
/* HyperElasticTet::prox (xzu/src/TetEnergyTerm.cpp:171-183; material 1 NeoHookean, 2 StVK; L-BFGS of
 * deps/mcloptlib/include/MCL/LBFGS.hpp:205-305) in place on n blocks; grad (may be NULL) receives
 * get_gradient = vol * dPsi/dF of the INPUT blocks. */
int aaadmm_tet_prox_hyper(int material, double mu, double lambda, double vol, double *z, double *grad, int64_t n);
int aaadmm_cod_solve(int m, const double *M, const double *rhs, double *x, int *rank);

/* ------------------------------------------------------------------------------------------
 * Remaining element types, as batches on host arrays (in place):
 *   tri_prox       TriEnergyTerm::prox on n column-major 3x2 blocks; variant AAADMM_ORDER_XZU:
 *                  xzu/src/TriEnergyTerm.cpp:77-107, AAADMM_ORDER_HARD_ZXU: hard/src/TriEnergyTerm.cpp:74-105;
 *                  limit_min / limit_max = Lame::limit_min / limit_max (-100 / 100: no strain limiting)
 *   collision_prox Collision::prox (hard/src/CollisionEnergyTerm.hpp:79-91) of n points against n_objs analytic
 *                  passive objects (hard/src/PassiveObject.hpp:32-136); 7 parameters per object:
 *                  {cx, cy, cz, nx, ny, nz, radius}; Floor uses cx as its height
 *   spring_prox    SpringPin::prox (hard/src/SpringEnergyTerm.hpp:66-70): z = pin where active
 * ------------------------------------------------------------------------------------------ */
#define AAADMM_PASSIVE_FLOOR 0
#define AAADMM_PASSIVE_SLIDE_FLOOR 1
#define AAADMM_PASSIVE_SPHERE 2
#define AAADMM_PASSIVE_PLANE_HALF_SPHERE 3
#define AAADMM_PASSIVE_CYLINDER 4
int aaadmm_tri_prox(int variant, double *z, int64_t n, double limit_min, double limit_max);
int aaadmm_collision_prox(int n_objs, const int *types, const double *params7, double *z, int64_t n);
int aaadmm_spring_prox(double *z, const double *pins, const int *active, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* AAADMM_H_ */
