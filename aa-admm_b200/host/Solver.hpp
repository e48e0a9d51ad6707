// Host-side mirror of the reference's admm::Solver for the tet ADMM path, over the C ABI.
//
// Same public surface as admm_anderson_hard_zxu/src/Solver.hpp:38-261 and
// admm_anderson_xzu/src/Solver.hpp (add_nodes, set_pins, initialize, step, runtime_data,
// settings, public m_x / m_v / m_masses / energyterms, nested Settings and RuntimeData);
// the two reference projects are one class here, selected by Settings::ordering.
// Eigen types are replaced by std::vector<double> / Vec3 (no Eigen in this image).
// Error behaviour follows the reference: initialize() returns false on bad node data and
// throws std::runtime_error for weight <= 0 / inverted rest tets; set_pins throws on bad input.
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/aaadmm.h"
#include "sparse_ldlt.hpp"
#include "tet_system.hpp"

namespace admm {

typedef std::array<double, 3> Vec3;
typedef std::array<int, 4> Vec4i;
typedef std::array<int, 3> Vec3i;

// admm::Lame (src/EnergyTerm.hpp:33-59)
class Lame {
public:
    static Lame rubber() { return Lame(10000000, 0.499); }
    static Lame soft_rubber() { return Lame(10000000, 0.399); }
    static Lame very_soft_rubber() { return Lame(1000000, 0.299); }
    double mu, lambda;
    double youngs, poisson;
    double bulk_modulus() const { return lambda + (2.0 / 3.0) * mu; }
    double limit_min, limit_max;
    Lame(double k, double v)
        : mu(k / (2.0 * (1.0 + v))), lambda(k * v / ((1.0 + v) * (1.0 - 2.0 * v))), youngs(k), poisson(v),
          limit_min(-100.0), limit_max(100.0) {}
};

// Element description. The reference's virtual prox/update_z/update_u run per element on the
// CPU; here elements only carry their constants and initialize() turns them into SoA batches.
class EnergyTerm {
public:
    virtual ~EnergyTerm() {}
    virtual int get_dim() const = 0;
    virtual double get_weight() const = 0;
    virtual double get_volume() const = 0;
};

class TetEnergyTerm : public EnergyTerm {
public:
    Vec4i tet;
    Lame lame;
    std::array<Vec3, 4> rest;  // rest vertices (doubles cast from the caller's scalars)
    int material;              // aaadmm::TetMaterial
    double volume = 0, weight = 0;
    int get_dim() const { return 9; }
    double get_weight() const { return weight; }
    double get_volume() const { return volume; }
    // throws std::runtime_error("**TetEnergyTerm Error: Inverted initial tet") like the reference ctor
    TetEnergyTerm(const Vec4i &tet_, const std::vector<Vec3> &verts, const Lame &lame_, int material_ = 0);
    // another material on the same element (material sweeps re-initialize one Solver, see Solver::initialize)
    void set_lame(const Lame &l);
};
class NeoHookeanTet : public TetEnergyTerm {
public:
    NeoHookeanTet(const Vec4i &t, const std::vector<Vec3> &v, const Lame &l) : TetEnergyTerm(t, v, l, 1) {}
};
class StVKTet : public TetEnergyTerm {
public:
    StVKTet(const Vec4i &t, const std::vector<Vec3> &v, const Lame &l) : TetEnergyTerm(t, v, l, 2) {}
};

// src/TetEnergyTerm.hpp:35-51
template <typename IN_SCALAR, typename TYPE>
inline void create_tets_from_mesh(std::vector<std::shared_ptr<EnergyTerm>> &energyterms, const IN_SCALAR *verts,
                                  const int *inds, int n_tets, const Lame &lame, const int vertex_offset) {
    for (int i = 0; i < n_tets; ++i) {
        Vec4i tet = {inds[i * 4 + 0], inds[i * 4 + 1], inds[i * 4 + 2], inds[i * 4 + 3]};
        std::vector<Vec3> tetverts(4);
        for (int k = 0; k < 4; ++k)
            tetverts[k] = {(double)verts[tet[k] * 3 + 0], (double)verts[tet[k] * 3 + 1], (double)verts[tet[k] * 3 + 2]};
        for (int k = 0; k < 4; ++k) tet[k] += vertex_offset;
        energyterms.emplace_back(std::make_shared<TYPE>(tet, tetverts, lame));
    }
}

// hard/src/TriEnergyTerm.hpp:53-83: linear-elastic triangle with strain limiting (Lame::limit_min / limit_max).
class TriEnergyTerm : public EnergyTerm {
public:
    Vec3i tri;
    Lame lame;
    std::array<Vec3, 3> rest;
    double area = 0, weight = 0;
    std::array<double, 4> rest_pose;  // column-major 2x2
    int get_dim() const { return 6; }
    double get_weight() const { return weight; }
    double get_volume() const { return area; }
    // throws std::runtime_error like the reference ctor (bad strain limits, inverted initial pose)
    TriEnergyTerm(const Vec3i &tri_, const std::vector<Vec3> &verts, const Lame &lame_);
    void set_lame(const Lame &l);
};

// hard/src/TriEnergyTerm.hpp:32-47
template <typename IN_SCALAR, typename TYPE>
inline void create_tris_from_mesh(std::vector<std::shared_ptr<EnergyTerm>> &energyterms, const IN_SCALAR *verts,
                                  const int *inds, int n_tris, const Lame &lame, const int vertex_offset) {
    for (int i = 0; i < n_tris; ++i) {
        Vec3i tri = {inds[i * 3 + 0], inds[i * 3 + 1], inds[i * 3 + 2]};
        std::vector<Vec3> triverts(3);
        for (int k = 0; k < 3; ++k)
            triverts[k] = {(double)verts[tri[k] * 3 + 0], (double)verts[tri[k] * 3 + 1], (double)verts[tri[k] * 3 + 2]};
        for (int k = 0; k < 3; ++k) tri[k] += vertex_offset;
        energyterms.emplace_back(std::make_shared<TYPE>(tri, triverts, lame));
    }
}

// Analytic passive obstacles (hard/src/PassiveObject.hpp:32-136). The reference evaluates signed_distance()
// through a virtual call per vertex on the CPU; here an obstacle only carries its tag and parameters
// ({cx, cy, cz, nx, ny, nz, radius}; Floor: cx = y) and the projection runs in the device batch.
class PassiveCollision {
public:
    virtual ~PassiveCollision() {}
    int type = AAADMM_PASSIVE_FLOOR;
    std::array<double, 7> prm{{0, 0, 0, 0, 0, 0, 0}};
};
class Floor : public PassiveCollision {
public:
    double m_y;
    Floor(double y) : m_y(y) {
        type = AAADMM_PASSIVE_FLOOR;
        prm[0] = y;
    }
};
class SlideFloor : public PassiveCollision {
public:
    Vec3 center, normal;  // the normal is normalised in the device batch like the reference ctor does
    SlideFloor(const Vec3 &c, const Vec3 &n) : center(c), normal(n) {
        type = AAADMM_PASSIVE_SLIDE_FLOOR;
        for (int j = 0; j < 3; ++j) prm[j] = c[j], prm[3 + j] = n[j];
    }
};
class Sphere : public PassiveCollision {
public:
    Vec3 center;
    double rad;
    Sphere(const Vec3 &c, double r) : center(c), rad(r) {
        type = AAADMM_PASSIVE_SPHERE;
        for (int j = 0; j < 3; ++j) prm[j] = c[j];
        prm[6] = r;
    }
};
class PlaneAndHalfSphere : public PassiveCollision {
public:
    Vec3 center;
    double rad;
    PlaneAndHalfSphere(const Vec3 &c, double r) : center(c), rad(r) {
        type = AAADMM_PASSIVE_PLANE_HALF_SPHERE;
        for (int j = 0; j < 3; ++j) prm[j] = c[j];
        prm[6] = r;
    }
};
class Cylinder : public PassiveCollision {
public:
    Vec3 center;
    double rad;
    Cylinder(const Vec3 &c, double r) : center(c), rad(r) {
        type = AAADMM_PASSIVE_CYLINDER;
        for (int j = 0; j < 3; ++j) prm[j] = c[j];
        prm[6] = r;
    }
};

// src/ExplicitForce.hpp:33-46: explicit velocity updates applied on the host at the start of step().
class ExplicitForce {
public:
    virtual ~ExplicitForce() {}
    virtual void project(double dt, std::vector<double> &x, std::vector<double> &v, std::vector<double> &m) const = 0;
};
// Wejchert / Haumann aerodynamic force on a list of triangles (src/ExplicitForce.cpp:47-105). The reference
// accumulates into v from an OpenMP loop that also reads v; here the triangles are visited in order (what the
// reference computes with one thread).
class WindForce : public ExplicitForce {
public:
    WindForce(std::vector<int> &tris_) : tris(tris_), direction{0.0, 0.0, 0.0} {}
    void project(double dt, std::vector<double> &x, std::vector<double> &v, std::vector<double> &m) const;
    std::vector<int> tris;
    Vec3 direction;
};

class Solver {
public:
    struct Settings {
        enum AccelationType { NOACC = 0, ANDERSON = 1 };
        enum Ordering { HARD_ZXU = AAADMM_ORDER_HARD_ZXU, XZU = AAADMM_ORDER_XZU };
        bool parse_args(int argc, char **argv);  // returns true if help()
        void help();
        double timestep_s;    // -dt
        int verbose;          // -v
        int admm_iters;       // -it
        double gravity;       // -g
        double constraint_w;  // -ck
        int Anderson_m;       // -am
        double penalty;       // -ap (hard_zxu)
        double beta;          // -ab (xzu; parsed, unused, as in the reference)
        AccelationType acceleration_type;  // -a
        Ordering ordering;                 // which reference project's loop to run
        bool write_residual_file;          // save() to ./result/residual-*.txt like the reference
        int nd_leaf_size;                  // nested-dissection leaf size of the setup factorisation
        std::string factor_cache;          // if not empty: file the LDL^T factor is cached in (keyed by the matrix)
        bool host_factorization;           // numeric LDL^T on the host cores instead of the device (default: device)
        Settings()
            : timestep_s(1.0 / 30.0), verbose(1), admm_iters(500), gravity(-9.8), constraint_w(-1), Anderson_m(2),
              penalty(1.0), beta(1.0), acceleration_type(NOACC), ordering(HARD_ZXU), write_residual_file(true),
              nd_leaf_size(96), host_factorization(false) {}
    };
    struct RuntimeData {
        double global_ms, local_ms, acceleration_ms, initialization_ms;
        int inner_iters;
        std::vector<double> step_time;
        // device-side figures of the last step
        double loop_ms, step_ms;
        int kernel_launches;
        RuntimeData()
            : global_ms(0), local_ms(0), acceleration_ms(0), initialization_ms(0), inner_iters(0), loop_ms(0),
              step_ms(0), kernel_launches(0) {}
        void print(const Settings &settings);
    };

    Solver();
    ~Solver();
    Solver(const Solver &) = delete;
    Solver &operator=(const Solver &) = delete;

    // Node data, scaled x3. The reference's VecX (Eigen::VectorXd) is a std::vector<double> with the handful of Eigen
    // members its samples use on these three vectors: segment<3>(i), rows(), setZero().
    class VecX : public std::vector<double> {
    public:
        using std::vector<double>::vector;
        using std::vector<double>::operator=;
        template <int N>
        struct Segment {
            double *p;
            double &operator[](int k) { return p[k]; }
            double operator[](int k) const { return p[k]; }
            template <typename V>
            Segment &operator=(const V &v) {
                for (int k = 0; k < N; ++k) p[k] = v[k];
                return *this;
            }
            operator std::array<double, N>() const {
                std::array<double, N> a;
                for (int k = 0; k < N; ++k) a[k] = p[k];
                return a;
            }
        };
        template <int N>
        Segment<N> segment(int i) { return Segment<N>{data() + i}; }
        int rows() const { return (int)size(); }
        void setZero() { std::fill(begin(), end(), 0.0); }
    };
    VecX m_x, m_v, m_masses;         // per node x3
    std::vector<int> surface_inds;   // indices of surface vertices (filled by binding::add_tetmesh)
    std::vector<std::shared_ptr<EnergyTerm>> energyterms;
    std::vector<std::shared_ptr<ExplicitForce>> ext_forces;

    template <typename T>
    int add_nodes(T *x, T *m, int n_verts) {
        const size_t prev_n = m_x.size();
        const int n3 = n_verts * 3;
        m_x.resize(prev_n + n3);
        m_v.resize(prev_n + n3);
        m_masses.resize(prev_n + n3);
        for (int i = 0; i < n3; ++i) {
            m_x[prev_n + i] = x[i];
            m_v[prev_n + i] = 0.0;
            m_masses[prev_n + i] = m[i];
        }
        return (int)((prev_n + n3) / 3);
    }

    void set_pins(const std::vector<int> &inds, const std::vector<Vec3> &points = std::vector<Vec3>());
    // any other 3-vector type with operator[] (the reference passes std::vector<Eigen::Vector3d>)
    template <typename V3, typename = typename std::enable_if<!std::is_same<V3, Vec3>::value>::type>
    void set_pins(const std::vector<int> &inds, const std::vector<V3> &points) {
        std::vector<Vec3> p(points.size());
        for (size_t i = 0; i < points.size(); ++i) p[i] = {(double)points[i][0], (double)points[i][1], (double)points[i][2]};
        set_pins(inds, p);
    }
    // hard/src/Solver.cpp:318-348: the listed vertices get one Collision energy term each at initialize();
    // obstacles must be added before initialize() (the device scene keeps their parameters).
    void set_collisions(const std::vector<int> &inds, const std::vector<Vec3> &points = std::vector<Vec3>());
    void add_obstacle(std::shared_ptr<PassiveCollision> obj);
    // hard/src/Solver.hpp:113: a dynamic obstacle has its vertices in m_x and is updated every frame. Triangle-mesh
    // obstacles (DynamicCollision / PassiveMesh and the BVH behind them) are outside the device path:
    // throws std::runtime_error, nothing is silently ignored.
    void add_dynamic_collider(std::shared_ptr<PassiveCollision> obj);
    // hard/src/Solver.hpp:125, Solver.cpp:493-498: writes solver_termA = M + rho dt^2 D^T W^2 D of the last initialize()
    // (3 n_free x 3 n_free, = Ahat (x) I3) as a Matrix Market coordinate file (symmetric, lower triangle; the reference
    // streams Eigen's text dump of the same matrix)
    void save_matrix(const std::string &filename);
    // Calling initialize() again on a Solver whose nodes, elements, pins and collision set are unchanged (only the
    // elements' Lame parameters, the time step or the penalty differ: the next member of a parameter sweep) keeps the
    // whole analysis and every device buffer and redoes the numeric part only (system-matrix values, numeric LDL^T on
    // the device, element moduli): last_initialize_was_incremental() then returns true. Velocities are zeroed as in the
    // reference; positions are the caller's.
    bool initialize(const Settings &settings_ = Settings());
    // every energy term gets the material (youngs, poisson); strain limits of triangle terms are kept
    void set_material(double youngs, double poisson);
    // Use a factor computed elsewhere instead of this library's nested-dissection LDL^T at the next initialize():
    // P A P^T = L D L^T with L strictly lower CSC (unit diagonal implied, rows ascending), perm[new] = old.
    // n = number of free vertices (factor of Ahat, A = Ahat (x) I3) or 3 x that (factor of the full system in the
    // reference's degree-of-freedom order, e.g. Eigen::SimplicialLDLT of solver_termA exactly as
    // LinearSolver.hpp:79-84 computes it). The arrays are copied.
    void set_external_factor(int n, const int64_t *Lp, const int *Li, const double *Lx, const double *D, const int *perm);
    void step();
    const RuntimeData &runtime_data() { return m_runtime; }
    const Settings &settings() { return m_settings; }
    void save();

    // Logged trajectory of the last step (the reference keeps these protected and writes them
    // to ./result/residual-*.txt in save()).
    std::vector<double> step_prim_residual, step_comb_residual;
    std::vector<int> is_reject;
    int iter_num = 0;
    int reject_num = 0;

    // setup statistics
    const aaadmm::LdltFactor &factor() const { return m_factor; }
    aaadmm_ldlt *device_factor() { return m_ldlt; }
    aaadmm_tetscene *device_scene() { return m_scene; }
    const aaadmm::TetSystem &system() const { return m_sys; }
    bool last_initialize_was_incremental() const { return reinitialized; }

    Settings m_settings;

protected:
    RuntimeData m_runtime;
    bool initialized;
    std::map<int, Vec3> m_pins;
    std::map<int, Vec3> m_collisions;  // ConstraintSet::collisions
    std::vector<std::shared_ptr<PassiveCollision>> m_obstacles;
    std::vector<double> m_x_pin;  // in the order set_pins received them (reference: m_x_pin)
    std::vector<int> positive_pin;
    bool device_numeric = false;       // the factor's values were computed on the device (re-initialisation can refactor)
    bool reinitialized = false;        // the last initialize() took the fast path (same structure, new moduli)
    uint64_t m_structure_key = 0;      // hash of mesh, terms, pins, collision set of the last full initialize()
    uint64_t m_identity_key = 0;       // hash of the energy-term POINTERS, masses, pin / collision sets of the last initialize()
    std::vector<char> m_term_is_tri;   // per energy term: 0 tet, 1 triangle (as found at the last full pass over the terms)
    bool factor_external = false;      // m_factor was handed in through set_external_factor
    bool factor_from_cache = false;    // the last initialize() took the factor from Settings::factor_cache
    std::vector<int> slot_of_node;  // index among the free nodes (positive_pin) or among the pinned ones
    aaadmm::TetSystem m_sys;
    aaadmm::LdltFactor m_factor;
    aaadmm_ldlt *m_ldlt = nullptr;
    aaadmm_tetscene *m_scene = nullptr;
    std::vector<double> m_xbar, m_xout, m_hist_prim, m_hist_comb;
    std::vector<int> m_hist_rej;
};

}  // namespace admm
