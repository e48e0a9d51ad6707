// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product; never linked by it.
//
// C wrapper around the UNMODIFIED reference admm::Solver, compiled from the
// sources where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/libref_hard.so (-DREF_HARD, admm_anderson_hard_zxu) and
// oracle/_ref/libref_xzu.so (-DREF_XZU, admm_anderson_xzu).
//
// It drives the reference exactly as samples/utils/Application.hpp:232-249 does
// (sim callback -> Solver::step()), headless, and exposes what the parity tests
// need: positions, the logged residual trajectory (read back from the file the
// reference's own save() writes, Solver.hpp:126-151), the protected system
// matrix and the Eigen LDLT factor (through a Probe subclass, no source edits).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include <memory>
#include <omp.h>
#include <unistd.h>
#include <sys/stat.h>

#include "Solver.hpp"
#include "TetEnergyTerm.hpp"
#include "TriEnergyTerm.hpp"
#include "PassiveObject.hpp"
#ifdef REF_HARD
#include "SpringEnergyTerm.hpp"
#include "CollisionEnergyTerm.hpp"
#endif
#include "MCL/MeshIO.hpp"
#include "MCL/TetMesh.hpp"
#include "MCL/ShapeFactory.hpp"
#include "MCL/XForm.hpp"

#ifdef REF_HARD
#define RFN(name) ref_hard_##name
#else
#define RFN(name) ref_xzu_##name
#endif

namespace {

struct Probe : public admm::Solver {
    const SparseMat &termA() const { return solver_termA; }
    const SparseMat &D() const { return m_D; }
    const VecX &x_pin() const { return m_x_pin; }
    admm::LDLTSolver *ldlt() { return static_cast<admm::LDLTSolver *>(m_linsolver.get()); }
};

struct Handle {
    Probe solver;
    admm::Solver::Settings settings;
    std::string workdir;
    std::vector<double> hist;  // rows of the last residual file
    int hist_cols = 0;
    double step_wall_ms = 0;
};

// ProbeTet exposes the protected prox / get_gradient of the LINEAR tet.
struct ProbeTri : public admm::TriEnergyTerm {
    using admm::TriEnergyTerm::TriEnergyTerm;
    void call_prox(double *z6) {
        VecX zi = Eigen::Map<VecX>(z6, 6);
        VecX vi = zi;
        Eigen::Matrix<double, 6, 6> W = Eigen::Matrix<double, 6, 6>::Identity();
        prox(W, zi, vi);
        Eigen::Map<VecX>(z6, 6) = zi;
    }
    double w() const { return weight; }
    double ar() const { return area; }
    const Eigen::Matrix2d &rp() const { return rest_pose; }
};
struct ProbeTet : public admm::TetEnergyTerm {
    using admm::TetEnergyTerm::TetEnergyTerm;
    void call_prox(double *z9) {
        VecX zi = Eigen::Map<VecX>(z9, 9);
        VecX vi = zi;
        Eigen::Matrix<double, 9, 9> W = Eigen::Matrix<double, 9, 9>::Identity();
        prox(W, zi, vi);
        Eigen::Map<VecX>(z9, 9) = zi;
    }
    void call_grad(const double *z9, double *g9) {
        VecX zi = Eigen::Map<const VecX>(z9, 9);
        VecX g = VecX::Zero(9);
        get_gradient(zi, g);
        Eigen::Map<VecX>(g9, 9) = g;
    }
    double w() const { return weight; }
    double vol() const { return volume; }
    const Eigen::Matrix3d &binv() const { return edges_inv; }
};

template <typename TET>
struct ProbeHyper : public TET {
    using TET::TET;
    void call_prox(double *z9) {
        typename TET::VecX zi = Eigen::Map<typename TET::VecX>(z9, 9);
        typename TET::VecX vi = zi;
        Eigen::Matrix<double, 9, 9> W = Eigen::Matrix<double, 9, 9>::Identity();
        this->prox(W, zi, vi);
        Eigen::Map<typename TET::VecX>(z9, 9) = zi;
    }
    void call_grad(const double *z9, double *g9) {
        typename TET::VecX zi = Eigen::Map<const typename TET::VecX>(z9, 9);
        typename TET::VecX g = TET::VecX::Zero(9);
        this->get_gradient(zi, g);
        Eigen::Map<typename TET::VecX>(g9, 9) = g;
    }
};

}  // namespace

extern "C" {

void *RFN(new)(const char *workdir) {
    Handle *h = new Handle();
    h->workdir = workdir ? workdir : ".";
    return h;
}

void RFN(free)(void *hp) { delete static_cast<Handle *>(hp); }

// material: 0 LINEAR (TetEnergyTerm), 1 NEOHOOKEAN, 2 STVK. Mirrors
// binding::add_tetmesh (samples/utils/AddMeshes.hpp:97-177) with the arrays
// supplied by the caller (float32 vertices and masses, as the reference has them).
int RFN(add_tetmesh)(void *hp, const float *verts, int n_verts, const int *tets, int n_tets,
                     const float *masses, double youngs, double poisson, int material) {
    Handle *h = static_cast<Handle *>(hp);
    admm::Solver *s = &h->solver;
    int prev = s->m_x.rows() / 3;
    s->m_x.conservativeResize(prev * 3 + n_verts * 3);
    s->m_masses.conservativeResize(prev * 3 + n_verts * 3);
    for (int i = 0; i < n_verts; ++i) {
        int idx = i + prev;
        Eigen::Vector3f v(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]);
        s->m_x.segment<3>(idx * 3) = v.cast<double>();
        s->m_masses.segment<3>(idx * 3) = Eigen::Vector3d(1, 1, 1) * masses[i];
    }
    admm::Lame lame(youngs, poisson);
    try {
        if (material == 0)
            admm::create_tets_from_mesh<float, admm::TetEnergyTerm>(s->energyterms, verts, tets, n_tets, lame, prev);
        else if (material == 1)
            admm::create_tets_from_mesh<float, admm::NeoHookeanTet>(s->energyterms, verts, tets, n_tets, lame, prev);
        else
            admm::create_tets_from_mesh<float, admm::StVKTet>(s->energyterms, verts, tets, n_tets, lame, prev);
    } catch (std::exception &e) {
        fprintf(stderr, "ref add_tetmesh: %s\n", e.what());
        return -1;
    }
    return prev + n_verts;
}

// mcl::meshio readers + weighted_masses with the densities of binding::add_tetmesh / add_trimesh
// (samples/utils/AddMeshes.hpp:105-106,188-189). kind 0: path.ele/.node, 1: .obj. Two calls: sizes, then arrays.
int RFN(mesh_load)(const char *path, int kind, int *n_verts, int *n_elems, float *verts, int *elems, float *masses) {
    try {
        std::vector<float> m;
        if (kind == 0) {
            mcl::TetMesh mesh;
            if (!mcl::meshio::load_elenode(&mesh, path)) return -1;
            *n_verts = (int)mesh.vertices.size();
            *n_elems = (int)mesh.tets.size();
            if (!verts) return 0;
            mesh.weighted_masses(m, 1522.f);
            for (int i = 0; i < *n_verts; ++i)
                for (int j = 0; j < 3; ++j) verts[3 * i + j] = mesh.vertices[i][j];
            for (int i = 0; i < *n_elems; ++i)
                for (int j = 0; j < 4; ++j) elems[4 * i + j] = mesh.tets[i][j];
        } else {
            mcl::TriangleMesh mesh;
            if (!mcl::meshio::load_obj(&mesh, path)) return -1;
            *n_verts = (int)mesh.vertices.size();
            *n_elems = (int)mesh.faces.size();
            if (!verts) return 0;
            mesh.weighted_masses(m, 1.0f);
            for (int i = 0; i < *n_verts; ++i)
                for (int j = 0; j < 3; ++j) verts[3 * i + j] = mesh.vertices[i][j];
            for (int i = 0; i < *n_elems; ++i)
                for (int j = 0; j < 3; ++j) elems[3 * i + j] = mesh.faces[i][j];
        }
        memcpy(masses, m.data(), sizeof(float) * m.size());
        return 0;
    } catch (std::exception &e) {
        fprintf(stderr, "ref mesh_load: %s\n", e.what());
        return -2;
    }
}

// Triangle (cloth) mesh: nodes + create_tris_from_mesh<float, TriEnergyTerm>, as binding::add_trimesh
// (samples/utils/AddMeshes.hpp:180-230) does, with the arrays supplied by the caller; the strain limits of the
// Lame object are set explicitly (the samples use limit_min / limit_max of Lame, EnergyTerm.hpp:47).
int RFN(add_trimesh)(void *hp, const float *verts, int n_verts, const int *tris, int n_tris, const float *masses,
                     double youngs, double poisson, double limit_min, double limit_max) {
    Handle *h = static_cast<Handle *>(hp);
    admm::Solver *s = &h->solver;
    int prev = s->m_x.rows() / 3;
    s->m_x.conservativeResize(prev * 3 + n_verts * 3);
    s->m_masses.conservativeResize(prev * 3 + n_verts * 3);
    for (int i = 0; i < n_verts; ++i) {
        int idx = i + prev;
        Eigen::Vector3f v(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]);
        s->m_x.segment<3>(idx * 3) = v.cast<double>();
        s->m_masses.segment<3>(idx * 3) = Eigen::Vector3d(1, 1, 1) * masses[i];
    }
    admm::Lame lame(youngs, poisson);
    lame.limit_min = limit_min;
    lame.limit_max = limit_max;
    try {
        admm::create_tris_from_mesh<float, admm::TriEnergyTerm>(s->energyterms, verts, tris, n_tris, lame, prev);
    } catch (std::exception &e) {
        fprintf(stderr, "ref add_trimesh: %s\n", e.what());
        return -1;
    }
    return prev + n_verts;
}

// WindForce over the listed triangles (ExplicitForce.hpp:39-46), pushed into Solver::ext_forces as
// samples/Asia2019/windyflag.cpp:124-126 does. The reference's project() updates v from an OpenMP loop whose
// iterations read velocities other iterations write: run it with one thread (RFN(set_threads)) for a
// reproducible result.
int RFN(add_wind)(void *hp, const int *tris, int n_tris, const double *dir3) {
    Handle *h = static_cast<Handle *>(hp);
    std::vector<int> faces(tris, tris + 3 * n_tris);
    std::shared_ptr<admm::WindForce> wind(new admm::WindForce(faces));
    wind->direction = Eigen::Vector3d(dir3[0], dir3[1], dir3[2]);
    h->solver.ext_forces.push_back(wind);
    return 0;
}
void RFN(set_threads)(int n) { omp_set_num_threads(n); }
// WindForce::project on caller arrays (x, v: 3 per vertex; v updated in place), one thread.
void RFN(wind_project)(const int *tris, int n_tris, const double *dir3, double dt, const double *x, double *v, int n_verts) {
    std::vector<int> faces(tris, tris + 3 * n_tris);
    admm::WindForce wind(faces);
    wind.direction = Eigen::Vector3d(dir3[0], dir3[1], dir3[2]);
    Eigen::VectorXd xx = Eigen::Map<const Eigen::VectorXd>(x, 3 * n_verts), vv = Eigen::Map<const Eigen::VectorXd>(v, 3 * n_verts);
    Eigen::VectorXd mm = Eigen::VectorXd::Ones(3 * n_verts);
    omp_set_num_threads(1);
    wind.project(dt, xx, vv, mm);
    memcpy(v, vv.data(), sizeof(double) * 3 * n_verts);
}

int RFN(set_pins)(void *hp, const int *idx, const double *pts, int n) {
    Handle *h = static_cast<Handle *>(hp);
    std::vector<int> inds(idx, idx + n);
    std::vector<Eigen::Vector3d> points(n);
    for (int i = 0; i < n; ++i) points[i] = Eigen::Vector3d(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    try {
        h->solver.set_pins(inds, points);
    } catch (std::exception &e) {
        fprintf(stderr, "ref set_pins: %s\n", e.what());
        return -1;
    }
    return 0;
}

int RFN(initialize)(void *hp, double dt, int iters, double gravity, int anderson_m, int accel, double penalty) {
    Handle *h = static_cast<Handle *>(hp);
    h->settings.timestep_s = dt;
    h->settings.admm_iters = iters;
    h->settings.gravity = gravity;
    h->settings.Anderson_m = anderson_m;
    h->settings.verbose = 0;
    h->settings.acceleration_type =
        accel ? admm::Solver::Settings::ANDERSON : admm::Solver::Settings::NOACC;
#ifdef REF_HARD
    h->settings.penalty = penalty;
#else
    (void)penalty;
#endif
    try {
        FILE *saved = nullptr;
        (void)saved;
        bool ok = h->solver.initialize(h->settings);
        return ok ? 0 : -2;
    } catch (std::exception &e) {
        fprintf(stderr, "ref initialize: %s\n", e.what());
        return -1;
    }
}

// One Solver::step(). The reference writes ./result/residual-*.txt relative to the
// cwd (Solver.hpp:126-151): run inside workdir and read the file back.
int RFN(step)(void *hp) {
    Handle *h = static_cast<Handle *>(hp);
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) return -3;
    std::string res = h->workdir + "/result";
    mkdir(h->workdir.c_str(), 0777);
    mkdir(res.c_str(), 0777);
    if (chdir(h->workdir.c_str()) != 0) return -3;
    int rc = 0;
    try {
        mcl::MicroTimer t;
        h->solver.step();
        h->step_wall_ms = t.elapsed_ms();
    } catch (std::exception &e) {
        fprintf(stderr, "ref step: %s\n", e.what());
        rc = -1;
    }
    if (chdir(cwd) != 0) return -3;
    if (rc) return rc;
    std::string file = res + "/residual-" +
                       (h->settings.acceleration_type ? std::to_string(h->settings.Anderson_m) : std::string("no")) +
                       ".txt";
    std::ifstream in(file);
    h->hist.clear();
    h->hist_cols = 0;
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        double v;
        int c = 0;
        while (ls >> v) {
            h->hist.push_back(v);
            ++c;
        }
        if (c) h->hist_cols = c;
    }
    return 0;
}

double RFN(step_wall_ms)(void *hp) { return static_cast<Handle *>(hp)->step_wall_ms; }
// reference RuntimeData totals of the last step: local, global, acceleration, initialization
void RFN(runtime)(void *hp, double *out4) {
    const admm::Solver::RuntimeData &r = static_cast<Handle *>(hp)->solver.runtime_data();
    out4[0] = r.local_ms;
    out4[1] = r.global_ms;
    out4[2] = r.acceleration_ms;
    out4[3] = r.initialization_ms;
}

int RFN(hist_rows)(void *hp) {
    Handle *h = static_cast<Handle *>(hp);
    return h->hist_cols ? (int)(h->hist.size() / h->hist_cols) : 0;
}
int RFN(hist_cols)(void *hp) { return static_cast<Handle *>(hp)->hist_cols; }
void RFN(hist_copy)(void *hp, double *out) {
    Handle *h = static_cast<Handle *>(hp);
    memcpy(out, h->hist.data(), h->hist.size() * sizeof(double));
}

int RFN(n_dof)(void *hp) { return (int)static_cast<Handle *>(hp)->solver.m_x.rows(); }
void RFN(get_x)(void *hp, double *out) {
    Handle *h = static_cast<Handle *>(hp);
    memcpy(out, h->solver.m_x.data(), h->solver.m_x.rows() * sizeof(double));
}
void RFN(get_v)(void *hp, double *out) {
    Handle *h = static_cast<Handle *>(hp);
    memcpy(out, h->solver.m_v.data(), h->solver.m_v.rows() * sizeof(double));
}

// System matrix A = M + rho dt^2 D^T D (3n_free x 3n_free), CSR.
long RFN(termA_nnz)(void *hp) { return static_cast<Handle *>(hp)->solver.termA().nonZeros(); }
int RFN(termA_rows)(void *hp) { return static_cast<Handle *>(hp)->solver.termA().rows(); }
void RFN(termA_copy)(void *hp, int *rowptr, int *col, double *val) {
    Eigen::SparseMatrix<double, Eigen::RowMajor> A = static_cast<Handle *>(hp)->solver.termA();
    A.makeCompressed();
    memcpy(rowptr, A.outerIndexPtr(), (A.rows() + 1) * sizeof(int));
    memcpy(col, A.innerIndexPtr(), A.nonZeros() * sizeof(int));
    memcpy(val, A.valuePtr(), A.nonZeros() * sizeof(double));
}

// Eigen SimplicialLDLT factor of A: strictly-lower L (CSC), D, and perm with
// perm[new] = old (P A P^T = L D L^T).
long RFN(factor_nnz)(void *hp) {
    Handle *h = static_cast<Handle *>(hp);
    Eigen::SparseMatrix<double> L = h->solver.ldlt()->m_cholesky->matrixL();
    return (long)L.nonZeros();  // includes the unit diagonal if stored
}
int RFN(factor_copy)(void *hp, int *colptr, int *row, double *val, double *D, int *perm, long cap) {
    Handle *h = static_cast<Handle *>(hp);
    auto *chol = h->solver.ldlt()->m_cholesky.get();
    Eigen::SparseMatrix<double> L = chol->matrixL();
    L.makeCompressed();
    int n = L.rows();
    long k = 0;
    for (int j = 0; j < n; ++j) {
        colptr[j] = (int)k;
        for (Eigen::SparseMatrix<double>::InnerIterator it(L, j); it; ++it) {
            if (it.row() == j) continue;  // unit diagonal
            if (k >= cap) return -1;
            row[k] = it.row();
            val[k] = it.value();
            ++k;
        }
    }
    colptr[n] = (int)k;
    Eigen::VectorXd d = chol->vectorD();
    memcpy(D, d.data(), n * sizeof(double));
    // Eigen: permutationP() maps old -> new (indices()[old] = new)
    const auto &P = chol->permutationP();
    for (int i = 0; i < n; ++i) perm[P.indices()[i]] = i;
    return (int)k;
}

// Apply the reference's own solve (LinearSolver.hpp:87-90) to an arbitrary rhs.
void RFN(solve)(void *hp, const double *b, double *x) {
    Handle *h = static_cast<Handle *>(hp);
    int n = h->solver.termA().rows();
    Eigen::VectorXd bb = Eigen::Map<const Eigen::VectorXd>(b, n), xx(n);
    h->solver.ldlt()->solve(xx, bb);
    memcpy(x, xx.data(), n * sizeof(double));
}

// ---- unit-level access to the reference's per-element code -------------------------------
// LINEAR tet built from 4 double vertices; returns weight, volume, B^-1 (col-major).
int RFN(tet_constants)(const double *verts12, double youngs, double poisson, double *w, double *vol,
                       double *binv9) {
    std::vector<Eigen::Vector3d> v(4);
    for (int i = 0; i < 4; ++i) v[i] = Eigen::Vector3d(verts12[3 * i], verts12[3 * i + 1], verts12[3 * i + 2]);
    try {
        ProbeTet t(Eigen::Vector4i(0, 1, 2, 3), v, admm::Lame(youngs, poisson));
        *w = t.w();
        *vol = t.vol();
        memcpy(binv9, t.binv().data(), 9 * sizeof(double));
    } catch (std::exception &e) {
        return -1;
    }
    return 0;
}
// TetEnergyTerm::prox on n column-major 3x3 blocks, in place.
void RFN(tet_prox)(double *z, int n) {
    std::vector<Eigen::Vector3d> v = {Eigen::Vector3d(0, 0, 0), Eigen::Vector3d(1, 0, 0),
                                      Eigen::Vector3d(0, 1, 0), Eigen::Vector3d(0, 0, 1)};
    ProbeTet t(Eigen::Vector4i(0, 1, 2, 3), v, admm::Lame(1e7, 0.399));
    for (int i = 0; i < n; ++i) t.call_prox(z + 9 * i);
}
// TetEnergyTerm::get_gradient / (K vol) on n blocks: out = F - U V^T.
void RFN(tet_F_minus_UVt)(const double *z, double *out, int n) {
    std::vector<Eigen::Vector3d> v = {Eigen::Vector3d(0, 0, 0), Eigen::Vector3d(1, 0, 0),
                                      Eigen::Vector3d(0, 1, 0), Eigen::Vector3d(0, 0, 1)};
    admm::Lame lame(1e7, 0.399);
    ProbeTet t(Eigen::Vector4i(0, 1, 2, 3), v, lame);
    double kv = lame.bulk_modulus() * t.vol();
    for (int i = 0; i < n; ++i) {
        t.call_grad(z + 9 * i, out + 9 * i);
        for (int k = 0; k < 9; ++k) out[9 * i + k] /= kv;
    }
}

// HyperElasticTet::prox (NeoHookeanTet material 1 / StVKTet material 2) on n blocks for the tet with the
// given rest vertices; in place. Also get_gradient (vol * dPsi/dF).
int RFN(tet_prox_hyper)(int material, const double *verts12, double youngs, double poisson, double *z, double *grad_out, int n) {
    std::vector<Eigen::Vector3d> v(4);
    for (int i = 0; i < 4; ++i) v[i] = Eigen::Vector3d(verts12[3 * i], verts12[3 * i + 1], verts12[3 * i + 2]);
    try {
        if (material == 1) {
            ProbeHyper<admm::NeoHookeanTet> t(Eigen::Vector4i(0, 1, 2, 3), v, admm::Lame(youngs, poisson));
            for (int i = 0; i < n; ++i) { if (grad_out) t.call_grad(z + 9 * i, grad_out + 9 * i); t.call_prox(z + 9 * i); }
        } else {
            ProbeHyper<admm::StVKTet> t(Eigen::Vector4i(0, 1, 2, 3), v, admm::Lame(youngs, poisson));
            for (int i = 0; i < n; ++i) { if (grad_out) t.call_grad(z + 9 * i, grad_out + 9 * i); t.call_prox(z + 9 * i); }
        }
    } catch (std::exception &e) {
        fprintf(stderr, "ref tet_prox_hyper: %s\n", e.what());
        return -1;
    }
    return 0;
}

// The reference's own scene generator (ShapeFactory.hpp:436-497 + TetMesh::refine +
// weighted_masses, and the centre/scale of beams.cpp:83-90) for validating the O(n)
// restatement. Returns counts; call with null outputs first.
int RFN(make_beam)(int cx, int cy, int cz, float y_shift, float density, float *verts, int *tets,
                   float *masses, int *n_verts, int *n_tets) {
    std::shared_ptr<mcl::TetMesh> mesh = mcl::factory::make_tet_blocks(cx, cy, cz);
    Eigen::AlignedBox<float, 3> aabb = mesh->bounds();
    mcl::XForm<float> center = mcl::xform::make_trans<float>(-aabb.center());
    float y = aabb.sizes()[1];
    mcl::XForm<float> scale = mcl::xform::make_scale<float>(1.f / y, 1.f / y, 1.f / y);
    mesh->apply_xform(scale * center);
    if (y_shift != 0.f) mesh->apply_xform(mcl::xform::make_trans(0.f, y_shift, 0.f));
    *n_verts = (int)mesh->vertices.size();
    *n_tets = (int)mesh->tets.size();
    if (!verts) return 0;
    std::vector<float> m;
    mesh->weighted_masses(m, density);
    for (int i = 0; i < *n_verts; ++i) {
        for (int k = 0; k < 3; ++k) verts[3 * i + k] = mesh->vertices[i][k];
        masses[i] = m[i];
    }
    for (int i = 0; i < *n_tets; ++i)
        for (int k = 0; k < 4; ++k) tets[4 * i + k] = mesh->tets[i][k];
    return 0;
}


// ---- rows I / J of SURVEY 8: triangle, collision and spring-pin terms --------------------------------
// TriEnergyTerm::prox on n column-major 3x2 blocks (in place) with the given strain limits.
void RFN(tri_prox)(double *z, int n, double limit_min, double limit_max) {
    std::vector<Eigen::Vector3d> v = {Eigen::Vector3d(0, 0, 0), Eigen::Vector3d(1, 0, 0), Eigen::Vector3d(0, 1, 0)};
    admm::Lame lame(1e7, 0.399);
    lame.limit_min = limit_min;
    lame.limit_max = limit_max;
    ProbeTri t(Eigen::Vector3i(0, 1, 2), v, lame);
    for (int i = 0; i < n; ++i) t.call_prox(z + 6 * i);
}
// TriEnergyTerm constructor: rest_pose (column-major 2x2), area, weight.
int RFN(tri_constants)(const double *verts9, double youngs, double poisson, double *rest_pose4, double *area, double *weight) {
    std::vector<Eigen::Vector3d> v(3);
    for (int i = 0; i < 3; ++i) v[i] = Eigen::Vector3d(verts9[3 * i], verts9[3 * i + 1], verts9[3 * i + 2]);
    try {
        ProbeTri t(Eigen::Vector3i(0, 1, 2), v, admm::Lame(youngs, poisson));
        Eigen::Map<Eigen::Matrix2d> rp_out(rest_pose4);
        rp_out = t.rp();
        *area = t.ar();
        *weight = t.w();
    } catch (std::exception &) {
        return -1;
    }
    return 0;
}
#ifdef REF_HARD
// SpringPin::prox on n points (in place): z = pin where active (hard_zxu only: the xzu class is abstract).
void RFN(spring_prox)(double *z, const double *pins, const int *active, int n) {
    for (int i = 0; i < n; ++i) {
        admm::SpringPin sp(0, Eigen::Vector3d(pins[3 * i], pins[3 * i + 1], pins[3 * i + 2]));
        sp.set_active(active[i] != 0);
        Eigen::VectorXd zi = Eigen::Map<Eigen::VectorXd>(z + 3 * i, 3), vi = zi;
        Eigen::MatrixXd W = Eigen::Matrix3d::Identity();
        sp.prox(W, zi, vi);
        Eigen::Map<Eigen::VectorXd>(z + 3 * i, 3) = zi;
    }
}
// Solver::set_collisions (in place) and Solver::add_obstacle with one analytic passive object, as
// samples/Asia2019/plinkohit.cpp:84-96 and plinkopony.cpp:60-117 use them.
int RFN(set_collisions)(void *hp, const int *idx, int n) {
    Handle *h = static_cast<Handle *>(hp);
    std::vector<int> inds(idx, idx + n);
    try {
        h->solver.set_collisions(inds, std::vector<Eigen::Vector3d>());
    } catch (std::exception &e) {
        fprintf(stderr, "ref set_collisions: %s\n", e.what());
        return -1;
    }
    return 0;
}
int RFN(add_obstacle)(void *hp, int type, const double *p) {
    Handle *h = static_cast<Handle *>(hp);
    Eigen::Vector3d c(p[0], p[1], p[2]), nrm(p[3], p[4], p[5]);
    std::shared_ptr<admm::PassiveCollision> o;
    switch (type) {
    case 0: o = std::make_shared<admm::Floor>(p[0]); break;
    case 1: o = std::make_shared<admm::SlideFloor>(c, nrm); break;
    case 2: o = std::make_shared<admm::Sphere>(c, p[6]); break;
    case 3: o = std::make_shared<admm::PlaneAndHalfSphere>(c, p[6]); break;
    case 4: o = std::make_shared<admm::Cylinder>(c, p[6]); break;
    default: return -1;
    }
    h->solver.add_obstacle(o);
    return 0;
}
// Collision::prox against a list of analytic passive objects (types: 0 Floor{y}, 1 SlideFloor{c,n},
// 2 Sphere{c,r}, 3 PlaneAndHalfSphere{c,r}, 4 Cylinder{c,r}; 7 parameters per object), n points in place.
int RFN(collision_prox)(int n_objs, const int *types, const double *prm, double *z, int n) {
    auto cs = std::make_shared<admm::ConstraintSet>();
    for (int j = 0; j < n_objs; ++j) {
        const double *p = prm + 7 * j;
        Eigen::Vector3d c(p[0], p[1], p[2]), nrm(p[3], p[4], p[5]);
        std::shared_ptr<admm::PassiveCollision> o;
        switch (types[j]) {
        case 0: o = std::make_shared<admm::Floor>(p[0]); break;
        case 1: o = std::make_shared<admm::SlideFloor>(c, nrm); break;
        case 2: o = std::make_shared<admm::Sphere>(c, p[6]); break;
        case 3: o = std::make_shared<admm::PlaneAndHalfSphere>(c, p[6]); break;
        case 4: o = std::make_shared<admm::Cylinder>(c, p[6]); break;
        default: return -1;
        }
        cs->collider->passive_objs.push_back(o);
    }
    admm::Collision col(0, cs);
    Eigen::MatrixXd W = Eigen::Matrix3d::Identity();
    for (int i = 0; i < n; ++i) {
        Eigen::VectorXd zi = Eigen::Map<Eigen::VectorXd>(z + 3 * i, 3), vi = zi;
        col.prox(W, zi, vi);
        Eigen::Map<Eigen::VectorXd>(z + 3 * i, 3) = zi;
    }
    return 0;
}
#endif
}  // extern "C"
