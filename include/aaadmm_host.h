/* aaadmm_host.h - C entry points of libaaadmm_host.so: the host-side mirror of the reference's
 * scene/solver classes (aa-admm_b200/host), exposed for ctypes (tests, bench.py, smoke()).
 * The compute boundary is include/aaadmm.h; nothing in this library computes the hot path.
 * Every function returns 0 (or a handle) on success, <0 / NULL on failure
 * (aaadmm_host_last_error()).
 */
#ifndef AAADMM_HOST_H_
#define AAADMM_HOST_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char *aaadmm_host_last_error(void);

/* Beam scenes: mcl::factory::make_tet_blocks + beams.cpp centre/scale/offset + find_pins /
 * stretch_beams (admm_anderson_hard_zxu/samples/Asia2019/beams.cpp:66-160) restated in O(n). */
void *aaadmm_host_beam_new(void);
void aaadmm_host_beam_free(void *h);
int aaadmm_host_beam_add(void *h, int cx, int cy, int cz, float y_shift, float density);
int aaadmm_host_beam_counts(void *h, int *n_verts, int *n_tets, int *n_pins);
int aaadmm_host_beam_copy(void *h, float *verts, int *tets, float *masses, int *pin_idx, double *pin_pts, int *pin_side);
int aaadmm_host_beam_stretch(void *h, double dt);

/* Setup factorisation (role of LDLTSolver::update_system, LinearSolver.hpp:79-84):
 * A lower CSC incl. diagonal; coords 3 per node or NULL. */
void *aaadmm_host_factor_new(int n, const int64_t *Ap, const int *Ai, const double *Ax, const double *coords,
                             int leaf_size, int n_threads);
/* On-disk factor cache (host/sparse_ldlt.hpp: ldlt_save / ldlt_load), keyed by a hash of the matrix. */
int aaadmm_host_factor_save(void *h, int n, const int64_t *Ap, const int *Ai, const double *Ax, const char *path);
void *aaadmm_host_factor_load(int n, const int64_t *Ap, const int *Ai, const double *Ax, const char *path);
void aaadmm_host_factor_free(void *h);
int64_t aaadmm_host_factor_nnz(void *h);
int aaadmm_host_factor_copy(void *h, int64_t *Lp, int *Li, double *Lx, double *D, int *perm);
int aaadmm_host_factor_solve(void *h, const double *b, double *x, int nrhs);
int aaadmm_host_factor_stats(void *h, double *s6);

/* Mesh files of the reference's samples (host/MeshIO.hpp = mcl::meshio of deps/mclscene/include/MCL/MeshIO.hpp:55-345):
 * kind 0 = TetGen ASCII `path`.ele + `path`.node (0- or 1-based, inverted tets re-ordered, float32 vertices),
 * kind 1 = Wavefront .obj (v / f records, triangles). copy: vertices (3 float per vertex), elements (4 or 3 ints) and the
 * lumped float32 masses binding::add_tetmesh (1522 kg/m^3) / add_trimesh (1 kg/m^2) give the nodes. */
void *aaadmm_host_mesh_load(const char *path, int kind);
void aaadmm_host_mesh_free(void *h);
int aaadmm_host_mesh_counts(void *h, int *n_verts, int *n_elems);
int aaadmm_host_mesh_copy(void *h, float *verts, int *elems, float *masses);
int aaadmm_host_mesh_save(void *h, const char *path);

/* Operator setup alone, no device involved (role of Solver::initialize, hard/src/Solver.cpp:361-491): the scalar
 * system matrix Ahat of A = M + rho dt^2 D^T W^2 D = Ahat (x) I3 for a scene of tets and triangles with the pinned
 * vertices eliminated; lower CSC incl. diagonal over the free vertices (dev_to_vert maps them back).
 * collision_verts: vertices that carry a Collision energy term (Solver::set_collisions), may be NULL / 0. */
void *aaadmm_host_system_new(const float *verts, int n_verts, const int *tets, int n_tets, const int *tris, int n_tris,
                             const float *masses, double youngs, double poisson, const int *pins, int n_pins,
                             double rho_dt2, const int *collision_verts, int n_collisions);
/* another uniform material / rho dt^2 on the same mesh: values only (update_tet_system_materials) */
int aaadmm_host_system_update(void *h, double youngs, double poisson, double rho_dt2);
void aaadmm_host_system_free(void *h);
int aaadmm_host_system_counts(void *h, int *n_free, int64_t *nnz);
int aaadmm_host_system_copy(void *h, int64_t *Ap, int *Ai, double *Ax, int *dev_to_vert);

/* admm::Solver mirror (hard_zxu/src/Solver.hpp:38-261): add_tetmesh = binding::add_tetmesh
 * (samples/utils/AddMeshes.hpp:97-177), set_pins, initialize, step. ordering 0 = hard_zxu, 1 = xzu. */
void *aaadmm_host_solver_new(void);
void aaadmm_host_solver_free(void *h);
int aaadmm_host_solver_add_tetmesh(void *h, const float *verts, int n_verts, const int *tets, int n_tets,
                                   const float *masses, double youngs, double poisson, int material);
/* binding::add_trimesh (samples/utils/AddMeshes.hpp:180-230) + create_tris_from_mesh (hard/src/TriEnergyTerm.hpp:32-47);
 * limit_min / limit_max = Lame::limit_min / limit_max (strain limiting). hard_zxu ordering only. */
int aaadmm_host_solver_add_trimesh(void *h, const float *verts, int n_verts, const int *tris, int n_tris,
                                   const float *masses, double youngs, double poisson, double limit_min,
                                   double limit_max);
/* WindForce (src/ExplicitForce.hpp:39-46) over n_tris triangles (global vertex ids) pushed into Solver::ext_forces */
int aaadmm_host_solver_add_wind(void *h, const int *tris, int n_tris, const double *dir3);
/* Solver::set_collisions (vertices constrained in place, hard/src/Solver.cpp:318-344) and Solver::add_obstacle with an
 * analytic obstacle (PassiveObject.hpp:32-136): type = AAADMM_PASSIVE_*, prm7 = {cx, cy, cz, nx, ny, nz, radius}
 * (Floor: prm7[0] = y). Both before initialize; hard_zxu ordering only. */
int aaadmm_host_solver_set_collisions(void *h, const int *idx, int n);
int aaadmm_host_solver_add_obstacle(void *h, int type, const double *prm7);
/* WindForce::project on caller arrays (x, v: 3 per vertex; v updated in place) */
int aaadmm_host_wind_project(const int *tris, int n_tris, const double *dir3, double dt, const double *x, double *v,
                             int n_verts);
int aaadmm_host_solver_set_pins(void *h, const int *idx, const double *pts, int n);
int aaadmm_host_solver_initialize(void *h, double dt, int iters, double gravity, int anderson_m, int accel,
                                  double penalty, int ordering, int nd_leaf);
/* Solver::set_external_factor: the next initialize() uses this factor (n = free vertices: factor of Ahat; n = 3 x free
 * vertices: factor of the full system in the reference's dof order, e.g. Eigen's own) instead of factoring itself. */
int aaadmm_host_solver_set_factor(void *h, int n, const int64_t *Lp, const int *Li, const double *Lx, const double *D,
                                  const int *perm);
/* Parameter sweeps on one Solver: set_material gives every energy term the Lame parameters (youngs, poisson), set_x
 * overwrites the node positions (3 per node), and the next aaadmm_host_solver_initialize with unchanged mesh / pins /
 * terms keeps the analysis and all device buffers and redoes the numeric part only (was_incremental() = 1). */
int aaadmm_host_solver_set_material(void *h, double youngs, double poisson);
int aaadmm_host_solver_set_x(void *h, const double *x);
int aaadmm_host_solver_was_incremental(void *h);
int aaadmm_host_solver_step(void *h);
int aaadmm_host_solver_set_iters(void *h, int iters, int anderson_m, int accel);
int aaadmm_host_solver_n_dof(void *h);
int aaadmm_host_solver_get_x(void *h, double *x);
int aaadmm_host_solver_get_v(void *h, double *v);
int aaadmm_host_solver_hist_rows(void *h);
int aaadmm_host_solver_hist_copy(void *h, double *prim, double *comb, int *rej);
int aaadmm_host_solver_info(void *h, double *out8);
int aaadmm_host_solver_factor_info(void *h, double *out6);
void *aaadmm_host_solver_device_scene(void *h);
void *aaadmm_host_solver_device_factor(void *h);
/* TriEnergyTerm constructor (hard/src/TriEnergyTerm.cpp:34-56): rest pose inverse (column-major 2x2), area, weight;
 * -1 for an inverted rest triangle */
int aaadmm_host_tri_constants(const double *rest9, double youngs, double poisson, double *rest_pose4, double *area, double *weight);
int aaadmm_host_tet_constants(const double *rest12, double youngs, double poisson, double *binv9, double *vol,
                              double *weight);

/* ALMGeometrySolver<3> mirror (Geometry/ALMGeometrySolver.h:52-287) with the shipped constraints
 * (Geometry/Constraint.h): planes/edges/angles are hard constraints, the reference-surface closest-point
 * constraint is soft, Laplacian / closeness rows are regularisation. */
void *aaadmm_host_geo_new(void);
/* variant: AAADMM_GEO_ALM (ALMGeometrySolver<3>) or AAADMM_GEO_GS (GeometrySolver<3>, Geometry/GeometrySolver.h:52-267) */
void *aaadmm_host_geo_new_variant(int variant);
void aaadmm_host_geo_free(void *h);
int aaadmm_host_geo_add_plane(void *h, const int *idx, int k, double weight);
int aaadmm_host_geo_add_edge(void *h, int i0, int i1, double weight, double len);
int aaadmm_host_geo_add_angle(void *h, int tip, int s1, int s2, double weight, double amin, double amax);
int aaadmm_host_geo_add_ref_surface(void *h, int n_points, double weight, const double *V, int nv, const int *F, int nf);
int aaadmm_host_geo_add_relative_uniform_laplacian(void *h, const int *idx, int n, double weight, const double *ref_pts, int n_pts);
int aaadmm_host_geo_add_uniform_laplacian(void *h, const int *idx, int n, double weight);
int aaadmm_host_geo_add_closeness(void *h, int idx, double weight, const double *target3);
int aaadmm_host_geo_setup(void *h, int n_points, double rho);
int aaadmm_host_geo_solve(void *h, const double *init_x, int n_points, int max_iter, int anderson_m);
int aaadmm_host_geo_history(void *h, double *values);
int aaadmm_host_geo_elapsed(void *h, double *secs); /* elapsed_time_ of the same iterations (cumulative seconds) */
int aaadmm_host_geo_solution(void *h, double *x, int n_points);
int aaadmm_host_geo_info(void *h, double *out4);

/* Geometry front-end (aa-admm_b200/host/GeometryApps.hpp): polygon meshes from Wavefront .obj numbered as OpenMesh numbers
 * them, subdivide_and_smooth_mesh (Geometry/MeshTypes.h:214-342) and the constraint recipes of the two applications
 * (Geometry/PlanarityOpt.cpp:134-246: app 0, prm = {penalty, closeness_w, laplacian_w, relative_laplacian_w};
 *  Geometry/WireMeshOpt.cpp:226-289: app 1, prm = {penalty, min_angle, max_angle, edge_length, closeness_w, laplacian_w}). */
void *aaadmm_host_polymesh_load(const char *path);
void *aaadmm_host_polymesh_new(const double *verts, int n_verts, const int *face_ptr, const int *face_idx, int n_faces);
void aaadmm_host_polymesh_free(void *h);
int aaadmm_host_polymesh_save(void *h, const char *path);
int aaadmm_host_polymesh_counts(void *h, int *counts4, double *avg_edge_length); /* vertices, faces, corners, edges */
int aaadmm_host_polymesh_copy(void *h, double *verts, int *face_ptr, int *face_idx, int *edges);
void *aaadmm_host_polymesh_subdivide_and_smooth(void *h);
int aaadmm_host_geoapp_optimize(int app, void *mesh, void *ref_mesh, int max_iter, int anderson_m, const double *prm,
                                double *hist, int *n_hist, double *solution, double *info3);

/* The same recipes with the setup kept (constraints, system matrix, factorisation, device upload once; any number of
 * solves from the mesh's own positions): info4 = {device loop ms, resets, kernel launches, wall ms of solve_ADMM}. */
void *aaadmm_host_geoapp_new(int kind, void *mesh, void *ref_mesh, const double *prm);
void aaadmm_host_geoapp_free(void *h);
/* out8 = points, hard constraints, z / u columns, soft constraints, nnz(L), fronts, tree levels, bytes of one apply */
int aaadmm_host_geoapp_stats(void *h, double *out8);
void *aaadmm_host_geoapp_device_factor(void *h); /* aaadmm_ldlt* of the application's solver (developer aid: trace dump) */
int aaadmm_host_geoapp_solve(void *h, int max_iter, int anderson_m, double *hist, int *n_hist, double *solution, double *info4);

#ifdef __cplusplus
}
#endif
#endif
