#include "GeometrySolver.hpp"

#include <chrono>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>

namespace aaadmm {

template <unsigned int N>
GeometrySolverBase<N>::GeometrySolverBase(int variant) : variant_(variant), penalty_parameter_(1.0), solver_initialized_(false) {
    last_result = aaadmm_step_result();
}

template <unsigned int N>
GeometrySolverBase<N>::~GeometrySolverBase() {
    // the solver owns the constraints (Geometry/ALMGeometrySolver.h:67-79)
    for (auto *c : hard_constraints_) delete c;
    for (auto *c : soft_constraints_) delete c;
    if (geo_) aaadmm_geo_destroy(geo_);
    if (ldlt_) aaadmm_ldlt_destroy(ldlt_);
}

// LinearRegularization<N> (Geometry/LinearRegularization.h:47-153)
template <unsigned int N>
void GeometrySolverBase<N>::add_laplacian_helper(const std::vector<int> &indices, const std::vector<double> &coefs,
                                                double weight, const MatrixNX *ref) {
    const double sw = std::sqrt(weight);
    reg_idx_.push_back(indices);
    std::vector<double> c(coefs.size());
    for (size_t i = 0; i < coefs.size(); ++i) c[i] = coefs[i] * sw;
    reg_coef_.push_back(c);
    double t[3] = {0, 0, 0};
    if (ref)
        for (size_t i = 0; i < indices.size(); ++i)
            for (int r = 0; r < 3; ++r) t[r] += (*ref)(r, indices[i]) * coefs[i];
    for (int r = 0; r < 3; ++r) reg_target_.push_back(t[r] * sw);
}
template <unsigned int N>
void GeometrySolverBase<N>::add_uniform_laplacian(const std::vector<int> &indices, double weight) {
    const int n = (int)indices.size();
    std::vector<double> coefs(1, 1.0);
    coefs.insert(coefs.end(), n - 1, -1.0 / double(n - 1));
    add_laplacian_helper(indices, coefs, weight, nullptr);
}
template <unsigned int N>
void GeometrySolverBase<N>::add_laplacian(const std::vector<int> &indices, const std::vector<double> coefs, double weight) {
    add_laplacian_helper(indices, coefs, weight, nullptr);
}
template <unsigned int N>
void GeometrySolverBase<N>::add_relative_uniform_laplacian(const std::vector<int> &indices, double weight, const MatrixNX &ref) {
    const int n = (int)indices.size();
    std::vector<double> coefs(1, 1.0);
    coefs.insert(coefs.end(), n - 1, -1.0 / double(n - 1));
    add_laplacian_helper(indices, coefs, weight, &ref);
}
template <unsigned int N>
void GeometrySolverBase<N>::add_relative_laplacian(const std::vector<int> &indices, const std::vector<double> coefs,
                                                  double weight, const MatrixNX &ref) {
    add_laplacian_helper(indices, coefs, weight, &ref);
}
template <unsigned int N>
void GeometrySolverBase<N>::add_closeness(int idx, double weight, const double *target_pt) {
    const double sw = std::sqrt(weight);
    reg_idx_.push_back(std::vector<int>(1, idx));
    reg_coef_.push_back(std::vector<double>(1, sw));
    for (int r = 0; r < 3; ++r) reg_target_.push_back(target_pt[r] * sw);
}

// Geometry/ALMGeometrySolver.h:81-161, Geometry/GeometrySolver.h:85-155
template <unsigned int N>
bool GeometrySolverBase<N>::setup_ADMM(int n_points, double penalty_param, SPDSolverType) {
    penalty_parameter_ = penalty_param;
    n_points_ = n_points;
    const int nh = (int)hard_constraints_.size();
    // ---- hard constraints: D_hard rows (Constraint::add_constraint, Constraint.h:132-159) ----
    std::vector<int> type(nh), idx_ptr(nh + 1, 0), idx;
    std::vector<double> param((size_t)4 * nh);
    std::vector<std::vector<std::pair<int, double>>> rows;  // D_hard row -> (point, coef)
    for (int c = 0; c < nh; ++c) {
        const Constraint<N> *k = hard_constraints_[c];
        if (k->kind() != Constraint<N>::PLANE && k->kind() != Constraint<N>::EDGE && k->kind() != Constraint<N>::ANGLE) {
            std::cerr << "Error: hard constraint " << c << " has no device implementation" << std::endl;
            return false;
        }
        type[c] = (int)k->kind();
        const std::vector<int> &ids = k->indices();
        idx.insert(idx.end(), ids.begin(), ids.end());
        idx_ptr[c + 1] = (int)idx.size();
        for (int q = 0; q < 4; ++q) param[4 * (size_t)c + q] = k->params()[q];
        const int n = (int)ids.size();
        if (k->kind() == Constraint<N>::PLANE) {
            const double c1 = 1.0 - 1.0 / n, c2 = -1.0 / n;
            for (int i = 0; i < n; ++i) {
                rows.emplace_back();
                for (int j = 0; j < n; ++j) rows.back().emplace_back(ids[j], i == j ? c1 : c2);
            }
        } else {
            for (int i = 1; i < n; ++i) {
                rows.emplace_back();
                rows.back().emplace_back(ids[0], -1.0);
                rows.back().emplace_back(ids[i], 1.0);
            }
        }
    }
    const int zc = (int)rows.size();
    // rho * D_hard^T as CSR over points
    std::vector<int64_t> dt_ptr(n_points + 1, 0);
    for (auto &r : rows)
        for (auto &e : r) dt_ptr[e.first + 1]++;
    for (int p = 0; p < n_points; ++p) dt_ptr[p + 1] += dt_ptr[p];
    std::vector<int> dt_col(dt_ptr[n_points]);
    std::vector<double> dt_val(dt_ptr[n_points]);
    {
        std::vector<int64_t> pos(dt_ptr.begin(), dt_ptr.end() - 1);
        for (int r = 0; r < zc; ++r)
            for (auto &e : rows[r]) {
                dt_col[pos[e.first]] = r;
                dt_val[pos[e.first]] = e.second * penalty_param;
                pos[e.first]++;
            }
    }
    // ---- system matrix: rho D^T D + D_soft^T D_soft + L^T L (lower triangle) ----
    std::vector<int> tr, tc;
    std::vector<double> tv;
    auto add_outer = [&](const std::vector<std::pair<int, double>> &r, double scale) {
        for (auto &a : r)
            for (auto &b : r)
                if (a.first >= b.first) {
                    tr.push_back(a.first);
                    tc.push_back(b.first);
                    tv.push_back(scale * a.second * b.second);
                }
    };
    for (auto &r : rows) add_outer(r, penalty_param);
    // soft constraints: closest point to a reference surface, one row per point with coefficient sqrt(w)
    std::vector<int> soft_point;
    double soft_weight = 0.0;
    std::shared_ptr<RefSurface> surf;
    for (auto *c : soft_constraints_) {
        if (c->kind() != Constraint<N>::CLOSEST || !c->surface) {
            std::cerr << "Error: soft constraint has no device implementation" << std::endl;
            return false;
        }
        if (!soft_point.empty() && (c->weight() != soft_weight || c->surface != surf) &&
            !(std::fabs(c->weight() - soft_weight) <= 1e-15 * soft_weight && c->surface.get() == surf.get())) {
            std::cerr << "Error: closest-point soft constraints must share one weight and one reference surface" << std::endl;
            return false;
        }
        soft_weight = c->weight();
        surf = c->surface;
        for (int p : c->indices()) {
            soft_point.push_back(p);
            tr.push_back(p);
            tc.push_back(p);
            // ALM: weighted identity rows outside the penalty term (ALMGeometrySolver.h:103-112);
            // GS: unweighted rows of the one D (GeometrySolver.h:106-108,120-121)
            tv.push_back(variant_ == AAADMM_GEO_GS ? penalty_param : soft_weight);
        }
    }
    if (variant_ == AAADMM_GEO_GS && !soft_point.empty()) {
        // append the soft rows to rho * D^T: point p gets column zc + i with coefficient rho
        std::vector<std::vector<std::pair<int, double>>> extra(n_points);
        for (size_t i = 0; i < soft_point.size(); ++i) extra[soft_point[i]].emplace_back(zc + (int)i, penalty_param);
        std::vector<int64_t> np(n_points + 1, 0);
        std::vector<int> ncol;
        std::vector<double> nval;
        for (int p = 0; p < n_points; ++p) {
            for (int64_t e = dt_ptr[p]; e < dt_ptr[p + 1]; ++e) {
                ncol.push_back(dt_col[e]);
                nval.push_back(dt_val[e]);
            }
            for (auto &e : extra[p]) {
                ncol.push_back(e.first);
                nval.push_back(e.second);
            }
            np[p + 1] = (int64_t)ncol.size();
        }
        dt_ptr.swap(np);
        dt_col.swap(ncol);
        dt_val.swap(nval);
    }
    std::vector<double> rhs_fixed((size_t)3 * n_points, 0.0);
    for (size_t r = 0; r < reg_idx_.size(); ++r) {
        std::vector<std::pair<int, double>> row;
        for (size_t j = 0; j < reg_idx_[r].size(); ++j) row.emplace_back(reg_idx_[r][j], reg_coef_[r][j]);
        add_outer(row, 1.0);
        for (auto &e : row)
            for (int k = 0; k < 3; ++k) rhs_fixed[3 * (size_t)e.first + k] += e.second * reg_target_[3 * r + k];
    }
    SymLower G = sym_from_triplets(n_points, tr, tc, tv, false);
    for (int j = 0; j < n_points; ++j)
        if (G.p[j + 1] == G.p[j] || G.i[G.p[j]] != j) {
            std::cerr << "Error: SPD solver initialization failed (point " << j << " is unconstrained)" << std::endl;
            return false;
        }
    if (const char *dump = getenv("AAADMM_GEO_DUMP_MATRIX")) {  // developer aid: the global matrix (lower CSC) as text
        std::ofstream ofs(dump);
        ofs << std::setprecision(17) << G.n << " " << G.p[G.n] << "\n";
        for (int j = 0; j < G.n; ++j)
            for (int64_t e = G.p[j]; e < G.p[j + 1]; ++e) ofs << G.i[e] << " " << j << " " << G.x[e] << "\n";
    }
    std::vector<int> perm = nested_dissection(G, nullptr, 64);
    factor_ = ldlt_factorize(G, perm);
    if (!factor_.ok) {
        std::cerr << "Error: SPD solver initialization failed" << std::endl;
        return false;
    }
    for (double d : factor_.D)
        if (!(d > 0.0)) {
            std::cerr << "Error: SPD solver initialization failed" << std::endl;
            return false;
        }
    if (geo_) aaadmm_geo_destroy(geo_), geo_ = nullptr;
    if (ldlt_) aaadmm_ldlt_destroy(ldlt_), ldlt_ = nullptr;
    if (aaadmm_ldlt_create(&ldlt_, factor_.n, factor_.Lp.data(), factor_.Li.data(), factor_.Lx.data(), factor_.D.data(),
                           factor_.perm.data(), 3) != 0) {
        std::cerr << "Error: " << aaadmm_last_error() << std::endl;
        return false;
    }
    aaadmm_geo_desc d;
    d.n_points = n_points;
    d.n_hard = nh;
    d.type = type.data();
    d.idx_ptr = idx_ptr.data();
    d.idx = idx.data();
    d.param = param.data();
    d.n_zcols = zc;
    d.dt_ptr = dt_ptr.data();
    d.dt_col = dt_col.data();
    d.dt_val = dt_val.data();
    d.n_soft = (int)soft_point.size();
    d.soft_point = soft_point.data();
    d.soft_weight = soft_weight;
    d.n_ref_verts = surf ? (int)(surf->verts.size() / 3) : 0;
    d.ref_verts = surf ? surf->verts.data() : nullptr;
    d.n_ref_tris = surf ? (int)(surf->tris.size() / 3) : 0;
    d.ref_tris = surf ? surf->tris.data() : nullptr;
    d.rhs_fixed = rhs_fixed.data();
    d.variant = variant_;
    d.rho = penalty_param;
    if (aaadmm_geo_create(&geo_, &d, ldlt_) != 0) {
        std::cerr << "Error: " << aaadmm_last_error() << std::endl;
        return false;
    }
    setup_counts[0] = n_points;
    setup_counts[1] = nh;
    setup_counts[2] = zc;
    setup_counts[3] = (int)soft_constraints_.size();
    solver_initialized_ = true;
    return true;
}

// Geometry/ALMGeometrySolver.h:163-283, Geometry/GeometrySolver.h:156-263
template <unsigned int N>
void GeometrySolverBase<N>::solve_ADMM(const MatrixNX &init_x, double, int max_iter, int Anderson_m) {
    if (!solver_initialized_) {
        std::cerr << "Error: solver not initialized yet" << std::endl;
        return;
    }
    default_x_ = init_x;
    // (the reference defines clear_iteration_history, ALMGeometrySolver.h:398-402, but never calls it: a second solve
    // appends to the history; so does this one)
    std::vector<double> hist(std::max(1, max_iter));
    const auto t0 = std::chrono::steady_clock::now();
    if (aaadmm_geo_solve(geo_, init_x.data(), max_iter, Anderson_m, default_x_.data(), hist.data(), &last_result) != 0) {
        std::cerr << "Error: " << aaadmm_last_error() << std::endl;
        return;
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const int n = last_result.iters_logged;
    std::vector<int> flags(std::max(1, n), 0);
    if (n > 0 && aaadmm_geo_reset_flags(geo_, flags.data(), n) != 0) std::cerr << "Error: " << aaadmm_last_error() << std::endl;
    // measured on the device when each iteration was logged (a rejected turn makes its iteration longer), plus the
    // host time before the loop started (uploads)
    std::vector<double> ms(std::max(1, n), 0.0);
    const bool stamped = n > 0 && aaadmm_geo_iteration_times(geo_, ms.data(), n) == 0;
    const double before = stamped ? std::max(0.0, secs - 1e-3 * last_result.loop_ms) : 0.0;
    for (int i = 0; i < n; ++i) {
        function_values_.push_back(hist[i]);
        elapsed_time_.push_back(stamped ? before + 1e-3 * ms[i] : secs * (i + 1) / n);
        Anderson_reset_.push_back(flags[i] != 0);
    }
    reset_count = last_result.rejects;
}

template <unsigned int N>
void GeometrySolverBase<N>::output_iteration_history(SolverType solver_type) {
    const int n_iter = (int)function_values_.size();
    for (int i = 0; i < n_iter; ++i) {
        std::cout << "Iteration " << i << ": ";
        std::cout << std::setprecision(6) << elapsed_time_[i] << " secs, ";
        std::cout << " target value " << std::setprecision(16) << function_values_[i];
        if (solver_type == AA_SOLVER && i < (int)Anderson_reset_.size() && Anderson_reset_[i]) std::cout << " (reject accelerator)";
        std::cout << std::endl;
    }
    std::cout << std::endl;
}

template <unsigned int N>
void GeometrySolverBase<N>::save(int Anderson_m) {
    std::string file = Anderson_m > 0 ? "./result/residual-" + std::to_string(Anderson_m) + ".txt" : "./result/residual-no.txt";
    std::ofstream ofs(file, std::ios::out | std::ios::ate);
    if (!ofs.is_open()) {
        std::cout << "Cannot open: " << file << std::endl;
        return;
    }
    ofs << std::setprecision(16);
    for (size_t i = 0; i < elapsed_time_.size(); i++) ofs << elapsed_time_[i] << '\t' << function_values_[i] << std::endl;
}

template class GeometrySolverBase<3>;

}  // namespace aaadmm
