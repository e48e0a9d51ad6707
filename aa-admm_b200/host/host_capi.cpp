// C entry points of libaaadmm_host.so: the host-side mirror classes (admm::Solver, the beam
// scene builder, the setup factorisation) made callable from ctypes for tests, bench.py and
// smoke(). The compute entry points live in libaaadmm_b200.so (include/aaadmm.h).
#include <cmath>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/aaadmm_host.h"
#include "GeometryApps.hpp"
#include "MeshIO.hpp"
#include "Solver.hpp"
#include "beam_scene.hpp"
#include "sparse_ldlt.hpp"
#include "tet_system.hpp"

namespace {
thread_local std::string g_err;
struct SolverHandle {
    admm::Solver solver;
    admm::Solver::Settings settings;
};
struct FactorHandle {
    aaadmm::LdltFactor F;
};
struct BeamHandle {
    aaadmm::BeamMesh mesh;
    aaadmm::BeamPins pins;
};
}  // namespace

#define HOST_TRY try {
#define HOST_CATCH                      \
    }                                   \
    catch (const std::exception &e) {   \
        g_err = e.what();               \
        return -1;                      \
    }

extern "C" {

const char *aaadmm_host_last_error(void) { return g_err.c_str(); }

// ---- beam scenes ---------------------------------------------------------------------------
void *aaadmm_host_beam_new(void) { return new BeamHandle(); }
void aaadmm_host_beam_free(void *h) { delete static_cast<BeamHandle *>(h); }
int aaadmm_host_beam_add(void *h, int cx, int cy, int cz, float y_shift, float density) {
    HOST_TRY
    BeamHandle *b = static_cast<BeamHandle *>(h);
    aaadmm::BeamMesh m = aaadmm::make_beam(cx, cy, cz, y_shift, density);
    const int off = b->mesh.n_verts();
    aaadmm::find_pins(m, off, b->pins);
    aaadmm::append_mesh(b->mesh, m);
    return b->mesh.n_verts();
    HOST_CATCH
}
int aaadmm_host_beam_counts(void *h, int *n_verts, int *n_tets, int *n_pins) {
    BeamHandle *b = static_cast<BeamHandle *>(h);
    *n_verts = b->mesh.n_verts();
    *n_tets = b->mesh.n_tets();
    *n_pins = (int)b->pins.idx.size();
    return 0;
}
int aaadmm_host_beam_copy(void *h, float *verts, int *tets, float *masses, int *pin_idx, double *pin_pts, int *pin_side) {
    BeamHandle *b = static_cast<BeamHandle *>(h);
    if (verts) memcpy(verts, b->mesh.verts.data(), b->mesh.verts.size() * sizeof(float));
    if (tets) memcpy(tets, b->mesh.tets.data(), b->mesh.tets.size() * sizeof(int));
    if (masses) memcpy(masses, b->mesh.masses.data(), b->mesh.masses.size() * sizeof(float));
    if (pin_idx) memcpy(pin_idx, b->pins.idx.data(), b->pins.idx.size() * sizeof(int));
    if (pin_pts) memcpy(pin_pts, b->pins.points.data(), b->pins.points.size() * sizeof(double));
    if (pin_side) memcpy(pin_side, b->pins.side.data(), b->pins.side.size() * sizeof(int));
    return 0;
}
int aaadmm_host_beam_stretch(void *h, double dt) {
    aaadmm::stretch_pins(static_cast<BeamHandle *>(h)->pins, dt);
    return 0;
}

// ---- setup factorisation ---------------------------------------------------------------------
void *aaadmm_host_factor_new(int n, const int64_t *Ap, const int *Ai, const double *Ax, const double *coords,
                             int leaf_size, int n_threads) {
    try {
        aaadmm::SymLower A;
        A.n = n;
        A.p.assign(Ap, Ap + n + 1);
        A.i.assign(Ai, Ai + Ap[n]);
        A.x.assign(Ax, Ax + Ap[n]);
        std::vector<int> perm = aaadmm::nested_dissection(A, coords, leaf_size > 0 ? leaf_size : 96);
        FactorHandle *f = new FactorHandle();
        f->F = aaadmm::ldlt_factorize(A, perm, n_threads);
        if (!f->F.ok) {
            g_err = "ldlt_factorize: zero or non-finite pivot";
            delete f;
            return nullptr;
        }
        return f;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
// Factor cache: save the factor under the key of its matrix; load returns NULL unless the file holds the factor
// of exactly this matrix.
int aaadmm_host_factor_save(void *h, int n, const int64_t *Ap, const int *Ai, const double *Ax, const char *path) {
    aaadmm::SymLower A;
    A.n = n;
    A.p.assign(Ap, Ap + n + 1);
    A.i.assign(Ai, Ai + Ap[n]);
    A.x.assign(Ax, Ax + Ap[n]);
    return aaadmm::ldlt_save(static_cast<FactorHandle *>(h)->F, aaadmm::matrix_key(A), path) ? 0 : -1;
}
void *aaadmm_host_factor_load(int n, const int64_t *Ap, const int *Ai, const double *Ax, const char *path) {
    aaadmm::SymLower A;
    A.n = n;
    A.p.assign(Ap, Ap + n + 1);
    A.i.assign(Ai, Ai + Ap[n]);
    A.x.assign(Ax, Ax + Ap[n]);
    FactorHandle *f = new FactorHandle();
    if (!aaadmm::ldlt_load(path, aaadmm::matrix_key(A), f->F) || f->F.n != n) {
        delete f;
        g_err = "factor cache: no factor of this matrix in the file";
        return nullptr;
    }
    return f;
}
void aaadmm_host_factor_free(void *h) { delete static_cast<FactorHandle *>(h); }
int64_t aaadmm_host_factor_nnz(void *h) {
    FactorHandle *f = static_cast<FactorHandle *>(h);
    return f->F.Lp[f->F.n];
}
int aaadmm_host_factor_copy(void *h, int64_t *Lp, int *Li, double *Lx, double *D, int *perm) {
    FactorHandle *f = static_cast<FactorHandle *>(h);
    const int n = f->F.n;
    memcpy(Lp, f->F.Lp.data(), sizeof(int64_t) * (n + 1));
    memcpy(Li, f->F.Li.data(), sizeof(int) * f->F.Li.size());
    memcpy(Lx, f->F.Lx.data(), sizeof(double) * f->F.Lx.size());
    memcpy(D, f->F.D.data(), sizeof(double) * n);
    memcpy(perm, f->F.perm.data(), sizeof(int) * n);
    return 0;
}
int aaadmm_host_factor_solve(void *h, const double *b, double *x, int nrhs) {
    aaadmm::ldlt_solve_host(static_cast<FactorHandle *>(h)->F, b, x, nrhs);
    return 0;
}
int aaadmm_host_factor_stats(void *h, double *s6) {
    FactorHandle *f = static_cast<FactorHandle *>(h);
    s6[0] = f->F.n_supernodes;
    s6[1] = f->F.flops;
    s6[2] = f->F.seconds_order;
    s6[3] = f->F.seconds_symbolic;
    s6[4] = f->F.seconds_numeric;
    s6[5] = (double)f->F.Lp[f->F.n];
    return 0;
}

// ---- admm::Solver ---------------------------------------------------------------------------
void *aaadmm_host_solver_new(void) { return new SolverHandle(); }
void aaadmm_host_solver_free(void *h) { delete static_cast<SolverHandle *>(h); }

// binding::add_tetmesh (samples/utils/AddMeshes.hpp:97-177): float32 vertices and masses.
int aaadmm_host_solver_add_tetmesh(void *h, const float *verts, int n_verts, const int *tets, int n_tets,
                                   const float *masses, double youngs, double poisson, int material) {
    HOST_TRY
    admm::Solver &s = static_cast<SolverHandle *>(h)->solver;
    const int prev = (int)s.m_x.size() / 3;
    s.m_x.resize((size_t)(prev + n_verts) * 3);
    s.m_v.resize((size_t)(prev + n_verts) * 3, 0.0);
    s.m_masses.resize((size_t)(prev + n_verts) * 3);
    for (int i = 0; i < n_verts; ++i)
        for (int j = 0; j < 3; ++j) {
            s.m_x[(size_t)(prev + i) * 3 + j] = (double)verts[3 * (size_t)i + j];
            s.m_masses[(size_t)(prev + i) * 3 + j] = (double)masses[i];
        }
    admm::Lame lame(youngs, poisson);
    if (material == 0)
        admm::create_tets_from_mesh<float, admm::TetEnergyTerm>(s.energyterms, verts, tets, n_tets, lame, prev);
    else if (material == 1)
        admm::create_tets_from_mesh<float, admm::NeoHookeanTet>(s.energyterms, verts, tets, n_tets, lame, prev);
    else
        admm::create_tets_from_mesh<float, admm::StVKTet>(s.energyterms, verts, tets, n_tets, lame, prev);
    return prev + n_verts;
    HOST_CATCH
}
// ---- mesh files (mcl::meshio + the masses binding::add_tetmesh / add_trimesh give the nodes) ----------------
struct MeshHandle {
    int kind = 0;  // 0 tet mesh (.ele/.node), 1 triangle mesh (.obj)
    mcl::TetMesh tet;
    mcl::TriangleMesh tri;
};
void *aaadmm_host_mesh_load(const char *path, int kind) {
    try {
        std::unique_ptr<MeshHandle> h(new MeshHandle());
        h->kind = kind;
        const bool ok = kind == 0 ? mcl::meshio::load_elenode(&h->tet, path) : mcl::meshio::load_obj(&h->tri, path);
        if (!ok) {
            g_err = std::string("could not load mesh ") + path;
            return nullptr;
        }
        return h.release();
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
void aaadmm_host_mesh_free(void *h) { delete static_cast<MeshHandle *>(h); }
int aaadmm_host_mesh_counts(void *hp, int *n_verts, int *n_elems) {
    MeshHandle *h = static_cast<MeshHandle *>(hp);
    *n_verts = (int)(h->kind == 0 ? h->tet.vertices.size() : h->tri.vertices.size());
    *n_elems = (int)(h->kind == 0 ? h->tet.tets.size() : h->tri.faces.size());
    return 0;
}
int aaadmm_host_mesh_copy(void *hp, float *verts, int *elems, float *masses) {
    HOST_TRY
    MeshHandle *h = static_cast<MeshHandle *>(hp);
    std::vector<float> m;
    if (h->kind == 0) {
        if (!h->tet.vertices.empty()) memcpy(verts, &h->tet.vertices[0][0], sizeof(float) * 3 * h->tet.vertices.size());
        if (!h->tet.tets.empty()) memcpy(elems, &h->tet.tets[0][0], sizeof(int) * 4 * h->tet.tets.size());
        h->tet.weighted_masses(m, 1522.f);  // binding::add_tetmesh
    } else {
        if (!h->tri.vertices.empty()) memcpy(verts, &h->tri.vertices[0][0], sizeof(float) * 3 * h->tri.vertices.size());
        if (!h->tri.faces.empty()) memcpy(elems, &h->tri.faces[0][0], sizeof(int) * 3 * h->tri.faces.size());
        h->tri.weighted_masses(m, 1.0f);  // binding::add_trimesh
    }
    std::copy(m.begin(), m.end(), masses);
    return 0;
    HOST_CATCH
}
int aaadmm_host_mesh_save(void *hp, const char *path) {
    HOST_TRY
    MeshHandle *h = static_cast<MeshHandle *>(hp);
    const bool ok = h->kind == 0 ? mcl::meshio::save_elenode(&h->tet, path) : mcl::meshio::save_obj(&h->tri, path);
    return ok ? 0 : -1;
    HOST_CATCH
}

// ---- operator setup alone (no device): the scalar system matrix of a scene of tets and triangles ------------
struct SystemHandle {
    aaadmm::TetSystem S;
};
void *aaadmm_host_system_new(const float *verts, int n_verts, const int *tets, int n_tets, const int *tris, int n_tris,
                             const float *masses, double youngs, double poisson, const int *pins, int n_pins,
                             double rho_dt2, const int *collision_verts, int n_collisions) {
    try {
        std::vector<double> rest12((size_t)12 * n_tets), rest9((size_t)9 * n_tris), m(n_verts);
        std::vector<double> ey(std::max(n_tets, 1), youngs), ep(std::max(n_tets, 1), poisson);
        std::vector<double> ty(std::max(n_tris, 1), youngs), tp(std::max(n_tris, 1), poisson);
        for (int t = 0; t < n_tets; ++t)
            for (int k = 0; k < 4; ++k)
                for (int j = 0; j < 3; ++j) rest12[12 * (size_t)t + 3 * k + j] = (double)verts[3 * (size_t)tets[4 * (size_t)t + k] + j];
        for (int t = 0; t < n_tris; ++t)
            for (int k = 0; k < 3; ++k)
                for (int j = 0; j < 3; ++j) rest9[9 * (size_t)t + 3 * k + j] = (double)verts[3 * (size_t)tris[3 * (size_t)t + k] + j];
        for (int v = 0; v < n_verts; ++v) m[v] = (double)masses[v];
        aaadmm::TriInput ti;
        ti.n_tris = n_tris;
        ti.rest9 = rest9.data();
        ti.tris = tris;
        ti.youngs = ty.data();
        ti.poisson = tp.data();
        std::unique_ptr<SystemHandle> h(new SystemHandle());
        std::vector<int> pinned(pins, pins + n_pins);
        // weight of the reference's Collision ctor (CollisionEnergyTerm.hpp:63-69)
        std::vector<double> cw(std::max(n_collisions, 1), std::sqrt(admm::Lame::soft_rubber().bulk_modulus() * 2.0));
        aaadmm::PointInput pi;
        pi.n = n_collisions;
        pi.verts = collision_verts;
        pi.weight = cw.data();
        if (!aaadmm::build_tet_system(h->S, n_verts, rest12.data(), n_tets, tets, nullptr, ey.data(), ep.data(), m.data(),
                                      pinned, rho_dt2, &ti, &pi)) {
            g_err = h->S.error;
            return nullptr;
        }
        return h.release();
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
// update_tet_system_materials on the handle: another (uniform) material and rho dt^2 on the same mesh
int aaadmm_host_system_update(void *h, double youngs, double poisson, double rho_dt2) {
    HOST_TRY
    aaadmm::TetSystem &S = static_cast<SystemHandle *>(h)->S;
    std::vector<double> ey(std::max(S.n_tets, 1), youngs), ep(std::max(S.n_tets, 1), poisson);
    std::vector<double> ty(std::max(S.n_tris, 1), youngs), tp(std::max(S.n_tris, 1), poisson);
    aaadmm::TriInput ti;
    ti.n_tris = S.n_tris;
    ti.youngs = ty.data();
    ti.poisson = tp.data();
    if (!aaadmm::update_tet_system_materials(S, ey.data(), ep.data(), rho_dt2, &ti)) {
        g_err = S.error;
        return -1;
    }
    return 0;
    HOST_CATCH
}
void aaadmm_host_system_free(void *h) { delete static_cast<SystemHandle *>(h); }
int aaadmm_host_system_counts(void *h, int *n_free, int64_t *nnz) {
    const aaadmm::TetSystem &S = static_cast<SystemHandle *>(h)->S;
    *n_free = S.n_free;
    *nnz = S.Ahat.p[S.n_free];
    return 0;
}
int aaadmm_host_system_copy(void *h, int64_t *Ap, int *Ai, double *Ax, int *dev_to_vert) {
    const aaadmm::TetSystem &S = static_cast<SystemHandle *>(h)->S;
    std::copy(S.Ahat.p.begin(), S.Ahat.p.end(), Ap);
    std::copy(S.Ahat.i.begin(), S.Ahat.i.end(), Ai);
    std::copy(S.Ahat.x.begin(), S.Ahat.x.end(), Ax);
    std::copy(S.dev_to_vert.begin(), S.dev_to_vert.end(), dev_to_vert);
    return 0;
}

// binding::add_trimesh (samples/utils/AddMeshes.hpp:180-230): float32 vertices and masses, strain limits in Lame.
int aaadmm_host_solver_add_trimesh(void *h, const float *verts, int n_verts, const int *tris, int n_tris,
                                   const float *masses, double youngs, double poisson, double limit_min,
                                   double limit_max) {
    HOST_TRY
    admm::Solver &s = static_cast<SolverHandle *>(h)->solver;
    const int prev = (int)s.m_x.size() / 3;
    s.m_x.resize((size_t)(prev + n_verts) * 3);
    s.m_v.resize((size_t)(prev + n_verts) * 3, 0.0);
    s.m_masses.resize((size_t)(prev + n_verts) * 3);
    for (int i = 0; i < n_verts; ++i)
        for (int j = 0; j < 3; ++j) {
            s.m_x[(size_t)(prev + i) * 3 + j] = (double)verts[3 * (size_t)i + j];
            s.m_masses[(size_t)(prev + i) * 3 + j] = (double)masses[i];
        }
    admm::Lame lame(youngs, poisson);
    lame.limit_min = limit_min;
    lame.limit_max = limit_max;
    admm::create_tris_from_mesh<float, admm::TriEnergyTerm>(s.energyterms, verts, tris, n_tris, lame, prev);
    return prev + n_verts;
    HOST_CATCH
}
// WindForce over the listed triangles, as samples/Asia2019/windyflag.cpp:124-126 adds it to Solver::ext_forces.
int aaadmm_host_solver_add_wind(void *h, const int *tris, int n_tris, const double *dir3) {
    HOST_TRY
    admm::Solver &s = static_cast<SolverHandle *>(h)->solver;
    std::vector<int> faces(tris, tris + 3 * (size_t)n_tris);
    std::shared_ptr<admm::WindForce> wind(new admm::WindForce(faces));
    wind->direction = {dir3[0], dir3[1], dir3[2]};
    s.ext_forces.push_back(wind);
    return 0;
    HOST_CATCH
}
// Solver::set_collisions (in place) + Solver::add_obstacle with an analytic obstacle given by tag and parameters.
int aaadmm_host_solver_set_collisions(void *h, const int *idx, int n) {
    HOST_TRY
    std::vector<int> inds(idx, idx + n);
    static_cast<SolverHandle *>(h)->solver.set_collisions(inds);
    return 0;
    HOST_CATCH
}
int aaadmm_host_solver_add_obstacle(void *h, int type, const double *prm7) {
    HOST_TRY
    admm::Solver &s = static_cast<SolverHandle *>(h)->solver;
    const admm::Vec3 c = {prm7[0], prm7[1], prm7[2]}, n = {prm7[3], prm7[4], prm7[5]};
    std::shared_ptr<admm::PassiveCollision> o;
    switch (type) {
        case AAADMM_PASSIVE_FLOOR: o = std::make_shared<admm::Floor>(prm7[0]); break;
        case AAADMM_PASSIVE_SLIDE_FLOOR: o = std::make_shared<admm::SlideFloor>(c, n); break;
        case AAADMM_PASSIVE_SPHERE: o = std::make_shared<admm::Sphere>(c, prm7[6]); break;
        case AAADMM_PASSIVE_PLANE_HALF_SPHERE: o = std::make_shared<admm::PlaneAndHalfSphere>(c, prm7[6]); break;
        case AAADMM_PASSIVE_CYLINDER: o = std::make_shared<admm::Cylinder>(c, prm7[6]); break;
        default: throw std::runtime_error("add_obstacle: unknown obstacle type");
    }
    s.add_obstacle(o);
    return 0;
    HOST_CATCH
}
int aaadmm_host_wind_project(const int *tris, int n_tris, const double *dir3, double dt, const double *x, double *v,
                             int n_verts) {
    HOST_TRY
    std::vector<int> faces(tris, tris + 3 * (size_t)n_tris);
    admm::WindForce wind(faces);
    wind.direction = {dir3[0], dir3[1], dir3[2]};
    std::vector<double> xx(x, x + 3 * (size_t)n_verts), vv(v, v + 3 * (size_t)n_verts), mm(3 * (size_t)n_verts, 1.0);
    wind.project(dt, xx, vv, mm);
    std::copy(vv.begin(), vv.end(), v);
    return 0;
    HOST_CATCH
}
int aaadmm_host_solver_set_pins(void *h, const int *idx, const double *pts, int n) {
    HOST_TRY
    std::vector<int> inds(idx, idx + n);
    std::vector<admm::Vec3> points(n);
    for (int i = 0; i < n; ++i) points[i] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    static_cast<SolverHandle *>(h)->solver.set_pins(inds, points);
    return 0;
    HOST_CATCH
}
int aaadmm_host_solver_initialize(void *h, double dt, int iters, double gravity, int anderson_m, int accel,
                                  double penalty, int ordering, int nd_leaf) {
    HOST_TRY
    SolverHandle *sh = static_cast<SolverHandle *>(h);
    admm::Solver::Settings &st = sh->settings;
    st.timestep_s = dt;
    st.admm_iters = iters;
    st.gravity = gravity;
    st.Anderson_m = anderson_m;
    st.acceleration_type = accel ? admm::Solver::Settings::ANDERSON : admm::Solver::Settings::NOACC;
    st.penalty = penalty;
    st.ordering = ordering ? admm::Solver::Settings::XZU : admm::Solver::Settings::HARD_ZXU;
    st.verbose = 0;
    st.write_residual_file = false;
    if (nd_leaf > 0) st.nd_leaf_size = nd_leaf;
    return sh->solver.initialize(st) ? 0 : -2;
    HOST_CATCH
}
int aaadmm_host_solver_set_factor(void *h, int n, const int64_t *Lp, const int *Li, const double *Lx, const double *D,
                                  const int *perm) {
    HOST_TRY
    static_cast<SolverHandle *>(h)->solver.set_external_factor(n, Lp, Li, Lx, D, perm);
    return 0;
    HOST_CATCH
}
int aaadmm_host_solver_set_material(void *h, double youngs, double poisson) {
    HOST_TRY
    static_cast<SolverHandle *>(h)->solver.set_material(youngs, poisson);
    return 0;
    HOST_CATCH
}
int aaadmm_host_solver_set_x(void *h, const double *x) {
    admm::Solver &s = static_cast<SolverHandle *>(h)->solver;
    memcpy(s.m_x.data(), x, s.m_x.size() * sizeof(double));
    return 0;
}
int aaadmm_host_solver_was_incremental(void *h) {
    return static_cast<SolverHandle *>(h)->solver.last_initialize_was_incremental() ? 1 : 0;
}
int aaadmm_host_solver_step(void *h) {
    HOST_TRY
    static_cast<SolverHandle *>(h)->solver.step();
    return 0;
    HOST_CATCH
}
int aaadmm_host_solver_set_iters(void *h, int iters, int anderson_m, int accel) {
    admm::Solver &s = static_cast<SolverHandle *>(h)->solver;
    s.m_settings.admm_iters = iters;
    s.m_settings.Anderson_m = anderson_m;
    s.m_settings.acceleration_type = accel ? admm::Solver::Settings::ANDERSON : admm::Solver::Settings::NOACC;
    return 0;
}
int aaadmm_host_solver_n_dof(void *h) { return (int)static_cast<SolverHandle *>(h)->solver.m_x.size(); }
int aaadmm_host_solver_get_x(void *h, double *x) {
    admm::Solver &s = static_cast<SolverHandle *>(h)->solver;
    memcpy(x, s.m_x.data(), s.m_x.size() * sizeof(double));
    return 0;
}
int aaadmm_host_solver_get_v(void *h, double *v) {
    admm::Solver &s = static_cast<SolverHandle *>(h)->solver;
    memcpy(v, s.m_v.data(), s.m_v.size() * sizeof(double));
    return 0;
}
int aaadmm_host_solver_hist_rows(void *h) {
    return (int)static_cast<SolverHandle *>(h)->solver.step_prim_residual.size();
}
int aaadmm_host_solver_hist_copy(void *h, double *prim, double *comb, int *rej) {
    admm::Solver &s = static_cast<SolverHandle *>(h)->solver;
    const size_t n = s.step_prim_residual.size();
    memcpy(prim, s.step_prim_residual.data(), n * sizeof(double));
    memcpy(comb, s.step_comb_residual.data(), n * sizeof(double));
    memcpy(rej, s.is_reject.data(), n * sizeof(int));
    return 0;
}
// out[0..7] = loop_ms, step_ms, kernel_launches, init_ms, iter_num, reject_num, n_free, n_tets
int aaadmm_host_solver_info(void *h, double *out8) {
    admm::Solver &s = static_cast<SolverHandle *>(h)->solver;
    const admm::Solver::RuntimeData &r = s.runtime_data();
    out8[0] = r.loop_ms;
    out8[1] = r.step_ms;
    out8[2] = r.kernel_launches;
    out8[3] = r.initialization_ms;
    out8[4] = s.iter_num;
    out8[5] = s.reject_num;
    out8[6] = s.system().n_free;
    out8[7] = s.system().n_tets;
    return 0;
}
// setup statistics: out[0..5] = nnz(L), supernodes, flops, symbolic s, numeric s, nnz(Ahat lower)
int aaadmm_host_solver_factor_info(void *h, double *out6) {
    admm::Solver &s = static_cast<SolverHandle *>(h)->solver;
    const aaadmm::LdltFactor &F = s.factor();
    out6[0] = (double)F.Lp[F.n];
    out6[1] = F.n_supernodes;
    out6[2] = F.flops;
    out6[3] = F.seconds_symbolic;
    out6[4] = F.seconds_numeric;
    out6[5] = (double)s.system().Ahat.p[s.system().n_free];
    return 0;
}
void *aaadmm_host_solver_device_scene(void *h) { return static_cast<SolverHandle *>(h)->solver.device_scene(); }
void *aaadmm_host_solver_device_factor(void *h) { return static_cast<SolverHandle *>(h)->solver.device_factor(); }

// Host-only pieces of initialize() for CPU tests: tet constants and the scalar system matrix.
int aaadmm_host_tri_constants(const double *rest9, double youngs, double poisson, double *rest_pose4, double *area, double *weight) {
    return aaadmm::tri_constants(rest9, youngs, poisson, rest_pose4, area, weight) ? 0 : -1;
}
int aaadmm_host_tet_constants(const double *rest12, double youngs, double poisson, double *binv9, double *vol,
                              double *weight) {
    return aaadmm::tet_constants(rest12, youngs, poisson, binv9, vol, weight) ? 0 : -1;
}

}  // extern "C"

// ---- ALMGeometrySolver<3> (Geometry/ALMGeometrySolver.h) ---------------------------------------
#include "GeometrySolver.hpp"
namespace {
struct GeoHandle {
    aaadmm::GeometrySolverBase<3> &solver;
    explicit GeoHandle(aaadmm::GeometrySolverBase<3> *s) : solver(*s), owner(s) {}
    std::unique_ptr<aaadmm::GeometrySolverBase<3>> owner;
};
}  // namespace
extern "C" {
void *aaadmm_host_geo_new(void) { return new GeoHandle(new aaadmm::ALMGeometrySolver<3>()); }
void *aaadmm_host_geo_new_variant(int variant) {
    if (variant == AAADMM_GEO_GS) return new GeoHandle(new aaadmm::GeometrySolver<3>());
    return new GeoHandle(new aaadmm::ALMGeometrySolver<3>());
}
void aaadmm_host_geo_free(void *h) { delete static_cast<GeoHandle *>(h); }
int aaadmm_host_geo_add_plane(void *h, const int *idx, int k, double weight) {
    static_cast<GeoHandle *>(h)->solver.add_hard_constraint(new aaadmm::PlaneConstraint(std::vector<int>(idx, idx + k), weight));
    return 0;
}
int aaadmm_host_geo_add_edge(void *h, int i0, int i1, double weight, double len) {
    static_cast<GeoHandle *>(h)->solver.add_hard_constraint(new aaadmm::EdgeLengthConstraint<3>(i0, i1, weight, len));
    return 0;
}
int aaadmm_host_geo_add_angle(void *h, int tip, int s1, int s2, double weight, double amin, double amax) {
    static_cast<GeoHandle *>(h)->solver.add_hard_constraint(new aaadmm::AngleConstraint<3>(tip, s1, s2, weight, amin, amax));
    return 0;
}
int aaadmm_host_geo_add_ref_surface(void *h, int n_points, double weight, const double *V, int nv, const int *F, int nf) {
    aaadmm::Matrix3X Vm(nv);
    memcpy(Vm.data(), V, sizeof(double) * 3 * nv);
    static_cast<GeoHandle *>(h)->solver.add_soft_constraint(
        new aaadmm::ReferenceSurfceConstraint(n_points, weight, Vm, std::vector<int>(F, F + 3 * (size_t)nf)));
    return 0;
}
int aaadmm_host_geo_add_relative_uniform_laplacian(void *h, const int *idx, int n, double weight, const double *ref_pts, int n_pts) {
    aaadmm::Matrix3X R(n_pts);
    memcpy(R.data(), ref_pts, sizeof(double) * 3 * n_pts);
    static_cast<GeoHandle *>(h)->solver.add_relative_uniform_laplacian(std::vector<int>(idx, idx + n), weight, R);
    return 0;
}
int aaadmm_host_geo_add_uniform_laplacian(void *h, const int *idx, int n, double weight) {
    static_cast<GeoHandle *>(h)->solver.add_uniform_laplacian(std::vector<int>(idx, idx + n), weight);
    return 0;
}
int aaadmm_host_geo_add_closeness(void *h, int idx, double weight, const double *target3) {
    static_cast<GeoHandle *>(h)->solver.add_closeness(idx, weight, target3);
    return 0;
}
int aaadmm_host_geo_setup(void *h, int n_points, double rho) {
    HOST_TRY
    return static_cast<GeoHandle *>(h)->solver.setup_ADMM(n_points, rho) ? 0 : -2;
    HOST_CATCH
}
int aaadmm_host_geo_solve(void *h, const double *init_x, int n_points, int max_iter, int anderson_m) {
    HOST_TRY
    aaadmm::Matrix3X x0(n_points);
    memcpy(x0.data(), init_x, sizeof(double) * 3 * n_points);
    GeoHandle *g = static_cast<GeoHandle *>(h);
    g->solver.function_values_.clear();
    g->solver.elapsed_time_.clear();
    g->solver.solve_ADMM(x0, 1e-8, max_iter, anderson_m);
    return (int)g->solver.function_values_.size();
    HOST_CATCH
}
int aaadmm_host_geo_history(void *h, double *values) {
    GeoHandle *g = static_cast<GeoHandle *>(h);
    memcpy(values, g->solver.function_values_.data(), sizeof(double) * g->solver.function_values_.size());
    return 0;
}
// elapsed_time_ of the logged iterations (seconds, cumulative; measured on the device per iteration)
int aaadmm_host_geo_elapsed(void *h, double *secs) {
    GeoHandle *g = static_cast<GeoHandle *>(h);
    memcpy(secs, g->solver.elapsed_time_.data(), sizeof(double) * g->solver.elapsed_time_.size());
    return 0;
}
int aaadmm_host_geo_solution(void *h, double *x, int n_points) {
    memcpy(x, static_cast<GeoHandle *>(h)->solver.get_solution().data(), sizeof(double) * 3 * n_points);
    return 0;
}
// out[0..3] = loop_ms, kernel_launches, rejects (resets), accepted iterations
int aaadmm_host_geo_info(void *h, double *out4) {
    GeoHandle *g = static_cast<GeoHandle *>(h);
    out4[0] = g->solver.last_result.loop_ms;
    out4[1] = g->solver.last_result.kernel_launches;
    out4[2] = g->solver.last_result.rejects;
    out4[3] = g->solver.last_result.iters_logged;
    return 0;
}
// ---- Geometry front-end (host/GeometryApps.hpp): polygon meshes, subdivision, the two applications ----------------
struct PolyHandle {
    aaadmm::geoapp::PolyMesh mesh;
};
void *aaadmm_host_polymesh_load(const char *path) {
    try {
        std::unique_ptr<PolyHandle> h(new PolyHandle());
        if (!aaadmm::geoapp::read_obj(path, h->mesh)) {
            g_err = std::string("unable to read mesh from file ") + path;
            return nullptr;
        }
        return h.release();
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
void *aaadmm_host_polymesh_new(const double *verts, int n_verts, const int *face_ptr, const int *face_idx, int n_faces) {
    try {
        std::unique_ptr<PolyHandle> h(new PolyHandle());
        h->mesh.V.assign(verts, verts + 3 * (size_t)n_verts);
        h->mesh.face_ptr.assign(face_ptr, face_ptr + n_faces + 1);
        h->mesh.face_idx.assign(face_idx, face_idx + face_ptr[n_faces]);
        return h.release();
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
void aaadmm_host_polymesh_free(void *h) { delete static_cast<PolyHandle *>(h); }
int aaadmm_host_polymesh_save(void *h, const char *path) {
    return aaadmm::geoapp::write_obj(static_cast<PolyHandle *>(h)->mesh, path) ? 0 : -1;
}
// counts[0..3] = vertices, faces, face corners, edges; avg_edge_length as MeshTypes.h:147-161
int aaadmm_host_polymesh_counts(void *h, int *counts4, double *avg_edge_length) {
    HOST_TRY
    const aaadmm::geoapp::PolyMesh &m = static_cast<PolyHandle *>(h)->mesh;
    aaadmm::geoapp::Connectivity C(m);
    counts4[0] = m.n_vertices();
    counts4[1] = m.n_faces();
    counts4[2] = (int)m.face_idx.size();
    counts4[3] = C.n_edges;
    if (avg_edge_length) *avg_edge_length = aaadmm::geoapp::average_edge_length(m);
    return 0;
    HOST_CATCH
}
// edges: 2 per edge (halfedge 0: from, to), in OpenMesh's edge order
int aaadmm_host_polymesh_copy(void *h, double *verts, int *face_ptr, int *face_idx, int *edges) {
    HOST_TRY
    const aaadmm::geoapp::PolyMesh &m = static_cast<PolyHandle *>(h)->mesh;
    if (verts) std::copy(m.V.begin(), m.V.end(), verts);
    if (face_ptr) std::copy(m.face_ptr.begin(), m.face_ptr.end(), face_ptr);
    if (face_idx) std::copy(m.face_idx.begin(), m.face_idx.end(), face_idx);
    if (edges) {
        aaadmm::geoapp::Connectivity C(m);
        for (int e = 0; e < C.n_edges; ++e) edges[2 * e] = C.edge_from[e], edges[2 * e + 1] = C.edge_to[e];
    }
    return 0;
    HOST_CATCH
}
void *aaadmm_host_polymesh_subdivide_and_smooth(void *h) {
    try {
        std::unique_ptr<PolyHandle> o(new PolyHandle());
        o->mesh = aaadmm::geoapp::subdivide_and_smooth_mesh(static_cast<PolyHandle *>(h)->mesh);
        return o.release();
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
// app 0: PlanarityOpt optimize_mesh (prm = {penalty, closeness_w, laplacian_w, relative_laplacian_w});
// app 1: WireMeshOpt optimize_mesh (prm = {penalty, min_angle, max_angle, edge_length, closeness_w, laplacian_w}).
// hist: max_iter residuals (n_hist = rows written), solution: 3 per vertex, info3 = {loop_ms, resets, kernel launches}.
int aaadmm_host_geoapp_optimize(int app, void *mesh_h, void *ref_h, int max_iter, int anderson_m, const double *prm,
                                double *hist, int *n_hist, double *solution, double *info3) {
    HOST_TRY
    const aaadmm::geoapp::PolyMesh &m = static_cast<PolyHandle *>(mesh_h)->mesh, &r = static_cast<PolyHandle *>(ref_h)->mesh;
    aaadmm::geoapp::OptimizeResult R =
        app == 0 ? aaadmm::geoapp::planarity_optimize(m, r, max_iter, anderson_m, prm[0], prm[1], prm[2], prm[3], false)
                 : aaadmm::geoapp::wiremesh_optimize(m, r, max_iter, anderson_m, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], false);
    if (!R.ok) {
        g_err = "geoapp: unable to initialize solver";
        return -1;
    }
    *n_hist = (int)R.function_values.size();
    std::copy(R.function_values.begin(), R.function_values.end(), hist);
    std::copy(R.mesh.V.begin(), R.mesh.V.end(), solution);
    if (info3) {
        info3[0] = R.elapsed_time.empty() ? 0.0 : 1e3 * R.elapsed_time.back();
        info3[1] = R.resets;
        info3[2] = 0;
    }
    return 0;
    HOST_CATCH
}
// GeoApp: setup once, solve repeatedly (bench steps). kind 0 planarity, 1 wiremesh; prm as aaadmm_host_geoapp_optimize.
void *aaadmm_host_geoapp_new(int kind, void *mesh_h, void *ref_h, const double *prm) {
    try {
        std::unique_ptr<aaadmm::geoapp::GeoApp> g(new aaadmm::geoapp::GeoApp(
            kind == 0 ? aaadmm::geoapp::GeoApp::PLANARITY : aaadmm::geoapp::GeoApp::WIREMESH, static_cast<PolyHandle *>(mesh_h)->mesh,
            static_cast<PolyHandle *>(ref_h)->mesh, prm));
        if (!g->ok()) {
            g_err = "geoapp: unable to initialize solver";
            return nullptr;
        }
        return g.release();
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
void aaadmm_host_geoapp_free(void *h) { delete static_cast<aaadmm::geoapp::GeoApp *>(h); }
// developer aid (per-task trace of the applies of a Geometry solve: aaadmm_ldlt_dump_trace)
void *aaadmm_host_geoapp_device_factor(void *h) { return static_cast<aaadmm::geoapp::GeoApp *>(h)->device_factor(); }
int aaadmm_host_geoapp_stats(void *h, double *out8) {
    static_cast<aaadmm::geoapp::GeoApp *>(h)->stats(out8);
    return 0;
}
// info4 = {device loop ms, resets, kernel launches, wall ms of solve_ADMM}
int aaadmm_host_geoapp_solve(void *h, int max_iter, int anderson_m, double *hist, int *n_hist, double *solution, double *info4) {
    HOST_TRY
    aaadmm::geoapp::GeoApp *g = static_cast<aaadmm::geoapp::GeoApp *>(h);
    aaadmm::geoapp::OptimizeResult R = g->solve(max_iter, anderson_m, false);
    if (!R.ok) {
        g_err = "geoapp: solve failed";
        return -1;
    }
    *n_hist = (int)R.function_values.size();
    std::copy(R.function_values.begin(), R.function_values.end(), hist);
    if (solution) std::copy(R.mesh.V.begin(), R.mesh.V.end(), solution);
    if (info4) {
        info4[0] = g->last_result().loop_ms;
        info4[1] = g->last_result().rejects;
        info4[2] = g->last_result().kernel_launches;
        info4[3] = R.elapsed_time.empty() ? 0.0 : 1e3 * R.elapsed_time.back();
    }
    return 0;
    HOST_CATCH
}
}  // extern "C"
