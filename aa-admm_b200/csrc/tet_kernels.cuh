// Local step, rhs assembly and residuals of the tet ADMM loop (K1-K4 of SURVEY 2b).
//
// One thread per tet; SoA planes (u, z: 9 planes of n_tets doubles; B^-1: 9 planes) so that every
// warp-wide load/store is a contiguous 256-byte segment; vertex positions (3 doubles per vertex,
// free vertices first, pinned last) are gathered through L2.  D is never materialised:
//   D_i x - c_i = w_i * Ds(x_full) * B_i^-1      (EnergyTerm.hpp:167-178 with C_fix = m_C x_pin)
//   D^T(.)      = per-vertex gather over incident (tet, corner) pairs, fixed order, no atomics.
#pragma once
#include "common.cuh"

namespace aaadmm {

constexpr int TET_BLOCK = 128;

struct TetArrays {
    int n_tets, n_free, n_verts;
    const int4 *idx;      // [T] vertex ids (free-first numbering)
    const double *binv;   // [9][T]
    const double *w;      // [T]
    const double *kvol;   // [T]
    double rho_dt2;
    // hyper-elastic tets (Neo-Hookean / StVK prox by per-tet L-BFGS); all null / 0 for linear scenes
    const int *material;   // [T] 0 linear, 1 NH, 2 StVK
    const double *mu, *lambda, *volume;  // [T]
    const int *hyper_ids;  // tets with material != 0
    int n_hyper;
};

// Triangle terms (tri_kernels.cu). rest_pose [4][N], u / z [6][N], contributions 9 doubles per triangle.
struct TriArrays {
    int n_tris, n_free;
    const int4 *idx;          // [N] vertex ids in x, y, z
    const double *rest_pose;  // [4][N] column-major 2x2
    const double *w;          // [N]
    const double *limit_min, *limit_max;  // [N]
    double rho_dt2;
};

// Collision terms (tri_kernels.cu): one per constrained free vertex, u / z [3][P], contributions 3 doubles per term.
struct PointArrays {
    int n_pts;
    const int *vert;   // [P] free vertex ids
    const double *w;   // [P]
    int n_objs;
    const int *types;  // [n_objs] PASSIVE_*
    const double *prm; // [n_objs][7]
    double rho_dt2;
};

enum { MODE_WARM = 0, MODE_ITER = 1, MODE_REDO = 2 };

// Host launchers (tet_kernels.cu is compiled with -fmad=false: the per-element arithmetic follows
// the reference operation by operation, and fused multiply-adds would change its rounding and,
// for degenerate elements, its branch decisions).
void launch_update_z_hard(int mode, int grid, cudaStream_t s, const TetArrays &A, const double *pos, const double *u,
                          double *z, double *contrib, SolveState *st, double *partials);
void launch_update_u_hard(int mode, int grid, cudaStream_t s, const TetArrays &A, const double *pos_new,
                          const double *pos_last, const double *z, const double *u_in, double *u_out, SolveState *st,
                          double *partials, double *hist_prim, double *hist_comb, int *hist_rej);
// Triangle terms: run before the tet launcher of the same phase (residual shares st->tri_prim2 / tri_comb).
void launch_tri_update_z_hard(int mode, cudaStream_t s, const TriArrays &A, const double *pos, const double *u, double *z,
                              double *contrib, SolveState *st, double *partials);
void launch_tri_update_u_hard(int mode, cudaStream_t s, const TriArrays &A, const double *pos_new, const double *pos_last,
                              const double *z, const double *u_in, double *u_out, SolveState *st, double *partials);
void launch_pt_update_z_hard(int mode, cudaStream_t s, const PointArrays &A, const double *pos, const double *u, double *z,
                             double *contrib, SolveState *st, double *partials);
void launch_pt_update_u_hard(int mode, cudaStream_t s, const PointArrays &A, const double *pos_new, const double *pos_last,
                             const double *z, const double *u_in, double *u_out, SolveState *st, double *partials);
void launch_tri_bconst(cudaStream_t s, const TriArrays &A, int slot0, const int64_t *inc_ptr, const int *inc,
                       const double *pos, double *bconst);
void launch_restore_if_reject(int grid, cudaStream_t s, double *ucur, const double *gdef, int64_t n, SolveState *st);
void launch_rhs_gather(cudaStream_t s, int n_free, const int64_t *inc_ptr, const int *inc, const double *contrib,
                       const double *bconst, const int *iperm, double *W, const SolveState *st, int when = 0,
                       int dof_factor = 0);
void launch_bconst(cudaStream_t s, const TetArrays &A, const int64_t *inc_ptr, const int *inc, const double *pos,
                   const double *mass, const double *xbar, double *bconst);
void launch_copy_if_not_done(cudaStream_t s, double *dst, const double *src, int64_t n, const SolveState *st);
// xzu ordering (admm_anderson_xzu/src/Solver.cpp:78-257); `when` = 1 runs only on a rejected iterate
void launch_grad_u_xzu(int grid, cudaStream_t s, const TetArrays &A, const double *z, double *u, const SolveState *st);
void launch_z_from_x(int grid, cudaStream_t s, const TetArrays &A, const double *pos, double *z);
void launch_contrib(int grid, cudaStream_t s, const TetArrays &A, const double *z, const double *u, double *contrib,
                    const SolveState *st, int when);
void launch_prim_xzu(int mode, int grid, cudaStream_t s, const TetArrays &A, const double *pos, const double *z,
                     SolveState *st, double *partials);
void launch_restore_xzu(int grid, cudaStream_t s, double *u, const double *u_def, double *z, const double *z_def,
                        int64_t nz, double *x, const double *x_def, int64_t nx, const SolveState *st);
void launch_update_z_plain(int grid, cudaStream_t s, const TetArrays &A, const double *pos, const double *u,
                           double *z_out, const SolveState *st);
void launch_update_u_plain(int grid, cudaStream_t s, const TetArrays &A, const double *pos, const double *z, double *u,
                           const SolveState *st, int when);
void launch_comb_xzu(int grid, cudaStream_t s, const TetArrays &A, const double *pos, const double *za,
                     const double *zb, SolveState *st, double *partials, double *hist_prim, double *hist_comb,
                     int *hist_rej);
void launch_copy2_if_not_done(int grid, cudaStream_t s, double *d0, const double *s0, int64_t n0, double *d1,
                              const double *s1, int64_t n1, const SolveState *st);
void launch_prox_hyper_batch(int material, double mu, double lambda, double vol, double *d_z, double *d_g, int64_t n);
void launch_prox_batch(double *d_z, int64_t n);
void launch_fmuvt_batch(const double *d_z, double *d_out, int64_t n);

}  // namespace aaadmm
