// Host-side mirror of the reference's Geometry solver interface over the C ABI.
//
//   Constraint<3> + shipped subclasses   Geometry/Constraint.h:48-414 (data carriers here: the virtual
//                                        project_impl of the reference runs as device batches)
//   LinearRegularization<3>              Geometry/LinearRegularization.h:37-153
//   ALMGeometrySolver<3>                 Geometry/ALMGeometrySolver.h:52-287: add_hard_constraint,
//                                        add_soft_constraint, add_closeness, add_*laplacian, setup_ADMM,
//                                        solve_ADMM, get_solution, function_values_, elapsed_time_
//   GeometrySolver<3>                    Geometry/GeometrySolver.h:52-267: the older variant with the same
//                                        interface (soft rows inside z/u, swap-back on a growing residual)
// Eigen's Matrix3X is replaced by aaadmm::Matrix3X (3 x n, column-major, data()/cols()/size()).
// User subclasses of Constraint with their own project_impl cannot run on the GPU and are rejected
// by setup_ADMM (returns false), as SURVEY 8b prescribes.
#pragma once
#include <cmath>
#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "../../include/aaadmm.h"
#include "sparse_ldlt.hpp"

namespace aaadmm {

class Matrix3X {
public:
    Matrix3X() {}
    explicit Matrix3X(int n) : v_((size_t)3 * n, 0.0) {}
    void setZero(int rows, int n) { (void)rows; v_.assign((size_t)3 * n, 0.0); }
    void resize(int rows, int n) { (void)rows; v_.resize((size_t)3 * n); }
    int cols() const { return (int)(v_.size() / 3); }
    int rows() const { return 3; }
    size_t size() const { return v_.size(); }
    double *data() { return v_.data(); }
    const double *data() const { return v_.data(); }
    double &operator()(int r, int c) { return v_[(size_t)3 * c + r]; }
    double operator()(int r, int c) const { return v_[(size_t)3 * c + r]; }
private:
    std::vector<double> v_;
};

struct RefSurface {  // reference triangle mesh of the closest-point constraints
    std::vector<double> verts;  // 3 per vertex
    std::vector<int> tris;      // 3 per triangle
};
// Stand-in for the reference's TriMeshAABB (Geometry/TriMeshAABB.h): only carries the mesh; the
// bounding-volume hierarchy is built by the C-ABI library.
typedef RefSurface TriMeshAABB;

template <unsigned int N>
class Constraint {
public:
    enum Kind { UNSUPPORTED = -1, PLANE = AAADMM_GEO_PLANE, EDGE = AAADMM_GEO_EDGE, ANGLE = AAADMM_GEO_ANGLE, CLOSEST = 100 };
    virtual ~Constraint() {}
    int num_indices() const { return (int)idI_.size(); }
    int num_transformed_points() const { return kind_ == PLANE || kind_ == CLOSEST ? num_indices() : num_indices() - 1; }
    Kind kind() const { return kind_; }
    const std::vector<int> &indices() const { return idI_; }
    double weight() const { return weight_ * weight_; }
    const double *params() const { return param_; }
    std::shared_ptr<RefSurface> surface;
protected:
    Constraint(const std::vector<int> &idI, double weight, Kind kind) : idI_(idI), weight_(std::sqrt(weight)), kind_(kind) {
        param_[0] = param_[1] = param_[2] = param_[3] = 0.0;
    }
    std::vector<int> idI_;
    double weight_;  // square root of the weight, as in the reference
    Kind kind_;
    double param_[4];
};

template <unsigned int N>
class EdgeLengthConstraint : public Constraint<N> {
public:
    EdgeLengthConstraint(int idx1, int idx2, double weight, double target_length)
        : Constraint<N>(std::vector<int>({idx1, idx2}), weight, Constraint<N>::EDGE) { this->param_[0] = target_length; }
};

template <unsigned int N>
class AngleConstraint : public Constraint<N> {
public:
    AngleConstraint(int tip_idx, int side_idx1, int side_idx2, double weight, double min_radian, double max_radian)
        : Constraint<N>(std::vector<int>({tip_idx, side_idx1, side_idx2}), weight, Constraint<N>::ANGLE) {
        const double pi = 3.14159265358979323846;
        const double mn = std::max(0.0, min_radian), mx = std::min(pi, max_radian);
        this->param_[0] = mn;
        this->param_[1] = mx;
        this->param_[2] = std::min(std::max(-1.0, std::cos(mn)), 1.0);
        this->param_[3] = std::min(std::max(-1.0, std::cos(mx)), 1.0);
    }
};

class PlaneConstraint : public Constraint<3> {
public:
    PlaneConstraint(const std::vector<int> &idI, double weight) : Constraint<3>(idI, weight, PLANE) {}
};

class PointToRefSurfaceConstraint : public Constraint<3> {
public:
    PointToRefSurfaceConstraint(int pt_idx, double weight, const std::shared_ptr<TriMeshAABB> &aabb)
        : Constraint<3>(std::vector<int>({pt_idx}), weight, CLOSEST) { surface = aabb; }
};

class ReferenceSurfceConstraint : public Constraint<3> {
public:
    // ref_surface_vtx: 3 x nv, ref_surface_faces: 3 ints per face (column-major like Eigen::Matrix3Xi)
    ReferenceSurfceConstraint(int n_points, double weight, const Matrix3X &ref_surface_vtx, const std::vector<int> &ref_surface_faces)
        : Constraint<3>(std::vector<int>(), weight, CLOSEST) {
        idI_.resize(n_points);
        for (int i = 0; i < n_points; ++i) idI_[i] = i;
        surface = std::make_shared<RefSurface>();
        surface->verts.assign(ref_surface_vtx.data(), ref_surface_vtx.data() + ref_surface_vtx.size());
        surface->tris = ref_surface_faces;
    }
};

enum SPDSolverType { LDLT_SOLVER, LLT_SOLVER, CG_SOLVER };
// Geometry/SolverCommon.h:41-44
enum SolverType { SHAPE_UP_SOLVER, AA_SOLVER };

// Both solver classes of the reference share their interface; the variant picks the setup matrices
// and the device loop (AAADMM_GEO_ALM / AAADMM_GEO_GS).
template <unsigned int N>
class GeometrySolverBase {
public:
    typedef Matrix3X MatrixNX;
    virtual ~GeometrySolverBase();
    GeometrySolverBase(const GeometrySolverBase &) = delete;
    GeometrySolverBase &operator=(const GeometrySolverBase &) = delete;

    void add_hard_constraint(Constraint<N> *c) { hard_constraints_.push_back(c); }
    void add_soft_constraint(Constraint<N> *c) { soft_constraints_.push_back(c); }
    void add_closeness(int idx, double weight, const double *target_pt);
    void add_uniform_laplacian(const std::vector<int> &indices, double weight);
    void add_laplacian(const std::vector<int> &indices, const std::vector<double> coefs, double weight);
    void add_relative_uniform_laplacian(const std::vector<int> &indices, double weight, const MatrixNX &ref_points);
    void add_relative_laplacian(const std::vector<int> &indices, const std::vector<double> coefs, double weight, const MatrixNX &ref_points);

    // returns false (with a message on stderr) if the factorisation fails or a constraint has no device batch
    bool setup_ADMM(int n_points, double penalty_param, SPDSolverType spd_solver_type = LDLT_SOLVER);
    void solve_ADMM(const MatrixNX &init_x, double rel_residual_eps, int max_iter, int Anderson_m);
    const MatrixNX &get_solution() { return default_x_; }
    void save(int Anderson_m);

    // Geometry/ALMGeometrySolver.h:322-340: one line per logged iteration (time, residual, "(reject accelerator)")
    void output_iteration_history(SolverType solver_type);

    // History of iterations (Geometry/ALMGeometrySolver.h:391-393). Anderson_reset_[i]: iteration i was logged right after
    // a reset of the accelerator (the reference declares the member but never fills it).
    std::deque<bool> Anderson_reset_;
    std::vector<double> function_values_, elapsed_time_;
    int reset_count = 0;
    // sizes of the last setup_ADMM (measurement reports): points, hard constraints, columns of z / u, soft constraints
    aaadmm_ldlt *device_factor() { return ldlt_; }
    int setup_counts[4] = {0, 0, 0, 0};
    aaadmm_step_result last_result;

protected:
    explicit GeometrySolverBase(int variant);
    int variant_;
    std::vector<Constraint<N> *> soft_constraints_, hard_constraints_;
    double penalty_parameter_;
    int n_points_ = 0;
    // LinearRegularization rows
    std::vector<std::vector<int>> reg_idx_;
    std::vector<std::vector<double>> reg_coef_;
    std::vector<double> reg_target_;  // 3 per row
    void add_laplacian_helper(const std::vector<int> &indices, const std::vector<double> &coefs, double weight, const MatrixNX *ref);
    MatrixNX default_x_;
    bool solver_initialized_;
    LdltFactor factor_;
    aaadmm_ldlt *ldlt_ = nullptr;
    aaadmm_geo *geo_ = nullptr;
};

template <unsigned int N>
class ALMGeometrySolver : public GeometrySolverBase<N> {
public:
    ALMGeometrySolver() : GeometrySolverBase<N>(AAADMM_GEO_ALM) {}
};

template <unsigned int N>
class GeometrySolver : public GeometrySolverBase<N> {
public:
    GeometrySolver() : GeometrySolverBase<N>(AAADMM_GEO_GS) {}
};

}  // namespace aaadmm
