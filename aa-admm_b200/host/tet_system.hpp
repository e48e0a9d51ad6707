// Host-side setup of the tet ADMM operators: everything admm::Solver::initialize builds
// (hard/src/Solver.cpp:361-491, xzu/src/Solver.cpp:373-498) restated matrix-free.
//
//  * per-tet constants as TetEnergyTerm's ctor computes them (xzu/src/TetEnergyTerm.cpp:32-65):
//    B^-1 of the rest edges, volume = det/6 (error if < 0), weight = sqrt((lambda+2/3 mu) vol)
//  * the free/pinned split of reset_fix_free_S_matrix (hard/src/Solver.cpp:236-278):
//    free vertices keep their ascending order
//  * A = M + rho dt^2 D^T W^2 D (hard/src/Solver.cpp:462-467). Every reduction row acts on one
//    coordinate with the same coefficient for x, y, z (TetEnergyTerm.cpp:67-88), so
//    A = Ahat (x) I3 and only the scalar n_free x n_free matrix Ahat is built and factored.
//  * per-vertex incidence lists (tet, corner) so that D^T(.) is a deterministic gather.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "sparse_ldlt.hpp"

namespace aaadmm {

struct Lame {
    double mu, lambda;
    Lame(double youngs, double poisson)
        : mu(youngs / (2.0 * (1.0 + poisson))),
          lambda(youngs * poisson / ((1.0 + poisson) * (1.0 - 2.0 * poisson))) {}
    double bulk_modulus() const { return lambda + (2.0 / 3.0) * mu; }
};

enum TetMaterial { TET_LINEAR = 0, TET_NEOHOOKEAN = 1, TET_STVK = 2 };

struct TetSystem {
    int n_verts = 0, n_tets = 0, n_free = 0, n_pin = 0;
    // vertex numbering used on the device: free vertices first (reference free order), then
    // pinned vertices in ascending vertex id (the order of std::map<int,Vec3> pins).
    std::vector<int> vert_to_dev;  // [n_verts]
    std::vector<int> dev_to_vert;  // [n_verts]
    std::vector<int> tet_dev;      // 4 per tet, device vertex ids
    std::vector<double> binv;      // 9 per tet, column-major B^-1 (AoS here; SoA on device)
    std::vector<double> weight;    // per tet
    std::vector<double> volume;    // per tet
    std::vector<double> kvol;      // K*vol per tet (xzu gradient)
    std::vector<int> material;     // per tet
    std::vector<double> mu, lambda;  // per tet (hyper-elastic prox)
    std::vector<double> mass_free;   // per free vertex (scalar; the reference stores it x3)
    // triangle (cloth) terms, hard/src/TriEnergyTerm.cpp:29-72: 6 rows each, F = [x1-x0, x2-x0] * rest_pose
    int n_tris = 0;
    std::vector<int> tri_dev;          // 3 per triangle, device vertex ids
    std::vector<double> tri_binv;      // 4 per triangle, column-major rest_pose (2x2)
    std::vector<double> tri_weight;    // sqrt(bulk modulus * area)
    std::vector<double> tri_area;
    std::vector<double> tri_limit_min, tri_limit_max;  // Lame::limit_min / limit_max of the term
    // collision terms (hard/src/CollisionEnergyTerm.hpp:40-91): 3 rows each, D_i x = w x_idx, free vertices only
    int n_pts = 0;
    std::vector<int> pt_dev;        // device vertex id per term
    std::vector<double> pt_weight;
    // incidence CSR over free vertices: entries are contribution slots, tet*4 + corner for the tets,
    // 4*n_tets + tri*3 + corner for the triangles, 4*n_tets + 3*n_tris + term for the collision terms
    // (a slot holds 3 doubles)
    std::vector<int64_t> inc_ptr;
    std::vector<int> inc;
    SymLower Ahat;  // n_free x n_free
    std::vector<int64_t> contrib_dst;  // per contribution (fixed emission order): its entry of Ahat.x
    std::string error;
};

// TetEnergyTerm ctor constants for one tet (false if the rest tet is inverted).
bool tet_constants(const double *rest12, double youngs, double poisson, double *binv9, double *vol, double *weight);
// TriEnergyTerm constructor (hard/src/TriEnergyTerm.cpp:34-56): rest pose inverse (column-major 2x2) in the
// triangle's own 2-D basis, area, weight = sqrt(bulk modulus * area). False for an inverted rest triangle.
bool tri_constants(const double *rest9, double youngs, double poisson, double *rest_pose4, double *area, double *weight);

// rest12: the 4 rest vertices of every tet (12 doubles per tet, as handed to the TetEnergyTerm
// ctor); tets: 4 ints per tet; masses: 1 per vertex; pinned: vertex ids. rho_dt2 = penalty * dt^2.
// Optional triangle terms of the same scene (rest9: the 3 rest vertices of every triangle).
struct TriInput {
    int n_tris = 0;
    const double *rest9 = nullptr;
    const int *tris = nullptr;
    const double *youngs = nullptr, *poisson = nullptr, *limit_min = nullptr, *limit_max = nullptr;
};
// Optional collision terms: one per listed vertex (must be free), with its weight.
struct PointInput {
    int n = 0;
    const int *verts = nullptr;
    const double *weight = nullptr;
};
bool build_tet_system(TetSystem &S, int n_verts, const double *rest12, int n_tets, const int *tets,
                      const int *material, const double *youngs, const double *poisson,
                      const double *masses, const std::vector<int> &pinned, double rho_dt2,
                      const TriInput *tri = nullptr, const PointInput *pts = nullptr);

// Another material on the SAME mesh, pins and terms (a member of a material sweep): per-element weights / moduli and
// the values of Ahat are recomputed in place - bit-identical to what build_tet_system gives for these inputs - while
// numbering, incidence lists, B^-1 and the pattern of Ahat stay. youngs / poisson: per tet; tri: per triangle.
bool update_tet_system_materials(TetSystem &S, const double *youngs, const double *poisson, double rho_dt2,
                                 const TriInput *tri = nullptr);

}  // namespace aaadmm
