#include "GeometryApps.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace aaadmm {
namespace geoapp {

namespace {
inline uint64_t dkey(int a, int b) { return ((uint64_t)(uint32_t)a << 32) | (uint32_t)b; }
inline const double *P(const PolyMesh &m, int v) { return m.V.data() + 3 * (size_t)v; }
}  // namespace

// ---- connectivity -----------------------------------------------------------------------------------------------
Connectivity::Connectivity(const PolyMesh &m) {
    const int nv = m.n_vertices(), nf = m.n_faces();
    std::unordered_map<uint64_t, int> edge_of;  // undirected (min, max) -> edge
    std::unordered_map<uint64_t, int> he;       // directed (a, b) -> face * 64 + corner is too small for big faces: store index into face_idx
    edge_of.reserve((size_t)m.face_idx.size());
    he.reserve((size_t)m.face_idx.size());
    face_edge.assign(m.face_idx.size(), -1);
    for (int f = 0; f < nf; ++f) {
        const int n = m.valence(f);
        const int *fv = m.face(f);
        for (int i = 0; i < n; ++i) {
            const int a = fv[i], b = fv[(i + 1) % n];
            if (a < 0 || a >= nv || b < 0 || b >= nv || a == b) throw std::runtime_error("PolyMesh: bad face vertex index");
            if (!he.emplace(dkey(a, b), m.face_ptr[f] + i).second)
                throw std::runtime_error("PolyMesh: complex edge (two faces use the same directed edge)");
            const uint64_t k = dkey(std::min(a, b), std::max(a, b));
            auto it = edge_of.find(k);
            int e;
            if (it == edge_of.end()) {
                e = n_edges++;
                edge_of.emplace(k, e);
                edge_from.push_back(a);
                edge_to.push_back(b);
                edge_face0.push_back(f);
                edge_face1.push_back(-1);
            } else {
                e = it->second;
                if (edge_face1[e] >= 0) throw std::runtime_error("PolyMesh: complex edge (more than two faces at an edge)");
                edge_face1[e] = f;
            }
            face_edge[m.face_ptr[f] + i] = e;
        }
    }
    vertex_boundary.assign(nv, 0);
    std::vector<char> used(nv, 0);
    for (int e = 0; e < n_edges; ++e) {
        used[edge_from[e]] = used[edge_to[e]] = 1;
        if (edge_face1[e] < 0) vertex_boundary[edge_from[e]] = vertex_boundary[edge_to[e]] = 1;
    }
    for (int v = 0; v < nv; ++v)
        if (!used[v]) vertex_boundary[v] = 1;
    // one-rings. Interior vertex: rotate through its faces - the halfedge that precedes (v -> cur) in its face ends at
    // v, its tail is the next neighbour. Boundary vertex: the neighbours in order of first appearance.
    std::vector<int> first_out(nv, -1);
    ring.assign(nv, {});
    for (int f = 0; f < nf; ++f) {
        const int n = m.valence(f);
        const int *fv = m.face(f);
        for (int i = 0; i < n; ++i)
            if (first_out[fv[i]] < 0) first_out[fv[i]] = fv[(i + 1) % n];
    }
    for (int v = 0; v < nv; ++v) {
        if (!used[v]) continue;
        if (vertex_boundary[v]) continue;
        const int start = first_out[v];
        int cur = start;
        for (int guard = 0; guard < 4096; ++guard) {
            ring[v].push_back(cur);
            const int pos = he.at(dkey(v, cur));
            // face and corner of (v -> cur)
            int f = (int)(std::upper_bound(m.face_ptr.begin(), m.face_ptr.end(), pos) - m.face_ptr.begin()) - 1;
            const int n = m.valence(f), k = pos - m.face_ptr[f];
            cur = m.face(f)[(k - 1 + n) % n];
            if (cur == start) break;
        }
    }
    for (int e = 0; e < n_edges; ++e) {
        const int a = edge_from[e], b = edge_to[e];
        if (vertex_boundary[a]) ring[a].push_back(b);
        if (vertex_boundary[b]) ring[b].push_back(a);
    }
}

// ---- files --------------------------------------------------------------------------------------------------------
bool read_obj(const std::string &path, PolyMesh &mesh) {
    std::ifstream in(path.c_str());
    if (!in.is_open()) return false;
    mesh = PolyMesh();
    std::string line, tok;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        if (!(ls >> tok)) continue;
        if (tok == "v") {
            // OpenMesh's OBJ reader parses the coordinates as float (Core/IO/reader/OBJReader.cc:294,330): the meshes the
            // reference's applications see carry single-precision coordinates
            float x, y, z;
            if (!(ls >> x >> y >> z)) return false;
            mesh.add_vertex((double)x, (double)y, (double)z);
        } else if (tok == "f") {
            std::vector<int> ids;
            while (ls >> tok) {
                const int id = atoi(tok.substr(0, tok.find('/')).c_str());
                if (id == 0) return false;
                ids.push_back(id > 0 ? id - 1 : mesh.n_vertices() + id);
            }
            if (ids.size() < 3) return false;
            mesh.add_face(ids);
        }
    }
    if (mesh.face_ptr.empty()) mesh.face_ptr.push_back(0);
    for (int id : mesh.face_idx)
        if (id < 0 || id >= mesh.n_vertices()) return false;
    return mesh.n_vertices() > 0;
}

bool write_obj(const PolyMesh &mesh, const std::string &path) {
    std::ofstream out(path.c_str());
    if (!out.is_open()) return false;
    out << std::setprecision(16);
    for (int v = 0; v < mesh.n_vertices(); ++v) out << "v " << P(mesh, v)[0] << " " << P(mesh, v)[1] << " " << P(mesh, v)[2] << "\n";
    for (int f = 0; f < mesh.n_faces(); ++f) {
        out << "f";
        for (int i = 0; i < mesh.valence(f); ++i) out << " " << mesh.face(f)[i] + 1;
        out << "\n";
    }
    return (bool)out;
}

double average_edge_length(const PolyMesh &mesh) {
    Connectivity C(mesh);
    if (C.n_edges == 0) return 0.0;
    double length = 0;
    for (int e = 0; e < C.n_edges; ++e) {
        const double *a = P(mesh, C.edge_from[e]), *b = P(mesh, C.edge_to[e]);
        const double d0 = b[0] - a[0], d1 = b[1] - a[1], d2 = b[2] - a[2];
        length += std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    }
    return length / C.n_edges;
}

// ---- Options.txt --------------------------------------------------------------------------------------------------
Parameters::Parameters() : elasticity(std::sqrt(5000000.0)) {}
bool Parameters::load(const char *filename) {
    std::ifstream in(filename);
    if (!in.is_open()) {
        std::cerr << "Error while opening file " << filename << std::endl;
        return false;
    }
    std::string line;
    while (std::getline(in, line)) {
        const std::string::size_type pos = line.find_first_not_of(' ');
        if (pos == std::string::npos || line.at(pos) == '#') continue;
        std::istringstream ls(line.substr(pos));
        std::string key;
        double val;
        if (!(ls >> key >> val)) continue;
        if (key == "Iterations")
            iter = (int)val;
        else if (key == "AndersonM")
            anderson_m = (int)val;
        else if (key == "SquareElasticity")
            elasticity = std::sqrt(val);
        else if (key == "TimeStep")
            time_step = val;
    }
    std::cout << "Successfully loaded options from file " << filename << std::endl;
    return true;
}
bool Parameters::valid_parameters() const {
    if (iter < 1) {
        std::cerr << "Error: Iterations must be at least 1" << std::endl;
        return false;
    }
    if (anderson_m < 0) {
        std::cerr << "Error: AndersonM must not be negative" << std::endl;
        return false;
    }
    return true;
}
void Parameters::output() const {
    std::cout << std::endl << "====== Filter parameters =========" << std::endl;
    std::cout << "Iterations: " << iter << std::endl << "AndersonM: " << anderson_m << std::endl;
    std::cout << "Elasticity: " << elasticity << std::endl << "TimeStep: " << time_step << std::endl;
    std::cout << "==================================" << std::endl;
}

// ---- subdivision + smoothing (MeshTypes.h:214-342) --------------------------------------------------------------
PolyMesh subdivide_and_smooth_mesh(const PolyMesh &in) {
    Connectivity C(in);
    const int nv = in.n_vertices(), ne = C.n_edges, nf = in.n_faces();
    PolyMesh out;
    out.V = in.V;
    for (int e = 0; e < ne; ++e) {
        const double *a = P(in, C.edge_from[e]), *b = P(in, C.edge_to[e]);
        out.add_vertex((a[0] + b[0]) * 0.5, (a[1] + b[1]) * 0.5, (a[2] + b[2]) * 0.5);
    }
    for (int f = 0; f < nf; ++f) {
        const int n = in.valence(f);
        const int *fv = in.face(f);
        double c[3] = {0, 0, 0};
        for (int i = 0; i < n; ++i)
            for (int r = 0; r < 3; ++r) c[r] += P(in, fv[i])[r];
        const int cv = out.n_vertices();
        out.add_vertex(c[0] / n, c[1] / n, c[2] / n);
        for (int i = 0; i < n; ++i) {
            const int e_prev = C.face_edge[in.face_ptr[f] + (i - 1 + n) % n], e_next = C.face_edge[in.face_ptr[f] + i];
            out.add_face({nv + e_prev, fv[i], nv + e_next, cv});
        }
    }
    // uniform Laplacian rows of the new mesh: interior vertices over their one-ring, boundary vertices over their two
    // boundary neighbours (when these belong to two different faces); weight 1
    Connectivity O(out);
    const int n_out = out.n_vertices();
    std::vector<std::vector<int>> rows(n_out);
    for (int v = 0; v < n_out; ++v)
        if (!O.vertex_boundary[v]) {
            rows[v].push_back(v);
            rows[v].insert(rows[v].end(), O.ring[v].begin(), O.ring[v].end());
        }
    // boundary rows need the edges: one pass over the boundary edges
    {
        std::vector<std::vector<int>> bn(n_out), bf(n_out);
        for (int e = 0; e < O.n_edges; ++e)
            if (O.edge_boundary(e)) {
                bn[O.edge_from[e]].push_back(O.edge_to[e]);
                bf[O.edge_from[e]].push_back(O.edge_face0[e]);
                bn[O.edge_to[e]].push_back(O.edge_from[e]);
                bf[O.edge_to[e]].push_back(O.edge_face0[e]);
            }
        for (int v = 0; v < n_out; ++v)
            if (O.vertex_boundary[v]) {
                rows[v].clear();
                if (bf[v].size() == 2 && bf[v][0] != bf[v][1]) rows[v] = {v, bn[v][0], bn[v][1]};
            }
    }
    // variables = the new vertices; least squares  min |L x|^2  with the original vertices fixed:
    //   (A^T A) x_var = -A^T L_fix x_fix,   A = L(:, var)
    std::vector<int> var_of(n_out, -1);
    int n_var = 0;
    for (int v = nv; v < n_out; ++v) var_of[v] = n_var++;
    std::vector<int> tr, tc;
    std::vector<double> tv;
    std::vector<double> rhs((size_t)3 * n_var, 0.0);
    for (const std::vector<int> &row : rows) {
        if (row.empty()) continue;
        const int n = (int)row.size();
        std::vector<double> coef(n, -1.0 / double(n - 1));
        coef[0] = 1.0;
        double fixed[3] = {0, 0, 0};
        for (int i = 0; i < n; ++i)
            if (var_of[row[i]] < 0)
                for (int r = 0; r < 3; ++r) fixed[r] += coef[i] * P(out, row[i])[r];
        for (int i = 0; i < n; ++i) {
            const int a = var_of[row[i]];
            if (a < 0) continue;
            for (int r = 0; r < 3; ++r) rhs[3 * (size_t)a + r] -= coef[i] * fixed[r];
            for (int j = 0; j < n; ++j) {
                const int b = var_of[row[j]];
                if (b < 0 || b > a) continue;
                tr.push_back(a);
                tc.push_back(b);
                tv.push_back(coef[i] * coef[j]);
            }
        }
    }
    if (n_var == 0) return out;
    SymLower M = sym_from_triplets(n_var, tr, tc, tv, false);
    std::vector<double> coords((size_t)3 * n_var);
    for (int v = nv; v < n_out; ++v)
        for (int r = 0; r < 3; ++r) coords[3 * (size_t)var_of[v] + r] = P(out, v)[r];
    LdltFactor F = ldlt_factorize(M, nested_dissection(M, coords.data(), 64));
    if (!F.ok) {
        std::cerr << "Error: unable to construct regularization system" << std::endl;
        return out;
    }
    std::vector<double> sol((size_t)3 * n_var);
    ldlt_solve_host(F, rhs.data(), sol.data(), 3);
    for (int v = nv; v < n_out; ++v)
        for (int r = 0; r < 3; ++r) out.V[3 * (size_t)v + r] = sol[3 * (size_t)var_of[v] + r];
    return out;
}

// ---- the two applications -------------------------------------------------------------------------------------------
namespace {
void to_matrix(const PolyMesh &m, Matrix3X &p) {
    p.resize(3, m.n_vertices());
    std::copy(m.V.begin(), m.V.end(), p.data());
}
std::vector<int> triangles_of(const PolyMesh &ref) {  // fan triangulation, as a TriMesh read does
    std::vector<int> t;
    for (int f = 0; f < ref.n_faces(); ++f)
        for (int i = 1; i + 1 < ref.valence(f); ++i) {
            t.push_back(ref.face(f)[0]);
            t.push_back(ref.face(f)[i]);
            t.push_back(ref.face(f)[i + 1]);
        }
    return t;
}
OptimizeResult finish(ALMGeometrySolver<3> &solver, const PolyMesh &mesh, const Matrix3X &p, double penalty, int max_iter, int m,
                      bool save_history) {
    OptimizeResult R;
    const double eps_ratio = 1e-8;
    const double rel_residual_eps = eps_ratio * average_edge_length(mesh);
    std::cout << "Relative residual eps (normalized by edge length): " << eps_ratio << std::endl;
    if (!solver.setup_ADMM(p.cols(), penalty)) {
        std::cerr << "Error: unable to initialize solver" << std::endl;
        return R;
    }
    solver.solve_ADMM(p, rel_residual_eps, max_iter, m);
    if (save_history) solver.save(m);
    R.ok = true;
    R.function_values = solver.function_values_;
    R.elapsed_time = solver.elapsed_time_;
    R.resets = solver.reset_count;
    R.mesh = mesh;
    const Matrix3X &x = solver.get_solution();
    std::copy(x.data(), x.data() + x.size(), R.mesh.V.begin());
    return R;
}
// boundary vertex: its neighbours across boundary edges and the face next to each of those edges
void boundary_fan(const Connectivity &C, int n_vertices, std::vector<std::vector<int>> &bn, std::vector<std::vector<int>> &bf) {
    bn.assign(n_vertices, {});
    bf.assign(n_vertices, {});
    for (int e = 0; e < C.n_edges; ++e)
        if (C.edge_boundary(e)) {
            bn[C.edge_from[e]].push_back(C.edge_to[e]);
            bf[C.edge_from[e]].push_back(C.edge_face0[e]);
            bn[C.edge_to[e]].push_back(C.edge_from[e]);
            bf[C.edge_to[e]].push_back(C.edge_face0[e]);
        }
}
}  // namespace

namespace {
void planarity_constraints(ALMGeometrySolver<3> &solver, const PolyMesh &mesh, const PolyMesh &ref_mesh, const Matrix3X &p,
                           double closeness_weight, double laplacian_weight, double relative_laplacian_weight) {
    std::shared_ptr<TriMeshAABB> aabb = std::make_shared<TriMeshAABB>();
    aabb->verts = ref_mesh.V;
    aabb->tris = triangles_of(ref_mesh);
    if (closeness_weight > 0)
        for (int i = 0; i < p.cols(); ++i) solver.add_soft_constraint(new PointToRefSurfaceConstraint(i, closeness_weight, aabb));
    Connectivity C(mesh);
    std::vector<std::vector<int>> bn, bf;
    boundary_fan(C, mesh.n_vertices(), bn, bf);
    auto add_lap = [&](const std::vector<int> &ids) {
        if (relative_laplacian_weight > 0) solver.add_relative_uniform_laplacian(ids, relative_laplacian_weight, p);
        if (laplacian_weight > 0) solver.add_uniform_laplacian(ids, laplacian_weight);
    };
    for (int v = 0; v < mesh.n_vertices(); ++v) {
        if (laplacian_weight <= 0 && relative_laplacian_weight <= 0) continue;
        if (!C.vertex_boundary[v]) {
            std::vector<int> vhs(1, v);
            vhs.insert(vhs.end(), C.ring[v].begin(), C.ring[v].end());
            if ((int)vhs.size() == 5) {  // regular vertex: the two straight lines through it
                add_lap({vhs[0], vhs[1], vhs[3]});
                add_lap({vhs[0], vhs[2], vhs[4]});
            } else {
                add_lap(vhs);
            }
        } else if (bf[v].size() == 2 && bf[v][0] != bf[v][1]) {
            add_lap({v, bn[v][0], bn[v][1]});
        }
    }
    for (int f = 0; f < mesh.n_faces(); ++f)
        if (mesh.valence(f) > 3)
            solver.add_hard_constraint(new PlaneConstraint(std::vector<int>(mesh.face(f), mesh.face(f) + mesh.valence(f)), 1.0));
}
}  // namespace

OptimizeResult planarity_optimize(const PolyMesh &mesh, const PolyMesh &ref_mesh, int max_iter, int Anderson_m, double penalty,
                                  double closeness_weight, double laplacian_weight, double relative_laplacian_weight,
                                  bool save_history) {
    Matrix3X p;
    to_matrix(mesh, p);
    ALMGeometrySolver<3> solver;
    planarity_constraints(solver, mesh, ref_mesh, p, closeness_weight, laplacian_weight, relative_laplacian_weight);
    return finish(solver, mesh, p, penalty, max_iter, Anderson_m, save_history);
}

namespace {
void wiremesh_constraints(ALMGeometrySolver<3> &solver, const PolyMesh &mesh, const PolyMesh &ref_mesh, const Matrix3X &p,
                          double min_angle_radian, double max_angle_radian, double edge_length, double closeness_weight,
                          double laplacian_weight) {
    Matrix3X ref_pts;
    to_matrix(ref_mesh, ref_pts);
    if (closeness_weight > 0)
        solver.add_soft_constraint(new ReferenceSurfceConstraint(p.cols(), closeness_weight, ref_pts, triangles_of(ref_mesh)));
    for (int f = 0; f < mesh.n_faces(); ++f) {
        if (mesh.valence(f) != 4) throw std::runtime_error("wiremesh_optimize: the mesh must consist of quads");
        const int *id = mesh.face(f);
        for (int i = 0; i < 4; ++i)
            solver.add_hard_constraint(new AngleConstraint<3>(id[i], id[(i + 1) % 4], id[(i + 3) % 4], 1.0, min_angle_radian, max_angle_radian));
    }
    Connectivity C(mesh);
    for (int e = 0; e < C.n_edges; ++e) solver.add_hard_constraint(new EdgeLengthConstraint<3>(C.edge_from[e], C.edge_to[e], 1.0, edge_length));
    if (laplacian_weight > 0) {  // setup_quad_laplacian_matrix (WireMeshOpt.cpp:184-224)
        std::vector<std::vector<int>> bn, bf;
        boundary_fan(C, mesh.n_vertices(), bn, bf);
        const std::vector<double> coefs = {2.0, -1.0, -1.0};
        for (int v = 0; v < mesh.n_vertices(); ++v) {
            const std::vector<int> &nb = C.ring[v];
            if (nb.size() > 4) {
                std::cout << "Invalid valence" << std::endl;
                break;
            } else if (nb.size() == 4 && !C.vertex_boundary[v]) {
                solver.add_laplacian({v, nb[0], nb[2]}, coefs, laplacian_weight);
                solver.add_laplacian({v, nb[1], nb[3]}, coefs, laplacian_weight);
            } else if (nb.size() == 3) {
                if (!C.vertex_boundary[v]) {
                    std::cout << "Not a regular quad mesh" << std::endl;
                    break;
                }
                std::vector<int> ids(1, v);
                ids.insert(ids.end(), bn[v].begin(), bn[v].end());
                if (ids.size() == 3) solver.add_laplacian(ids, coefs, laplacian_weight);
            }
        }
    }
}
}  // namespace

OptimizeResult wiremesh_optimize(const PolyMesh &mesh, const PolyMesh &ref_mesh, int max_iter, int Anderson_m, double penalty,
                                 double min_angle_radian, double max_angle_radian, double edge_length, double closeness_weight,
                                 double laplacian_weight, bool save_history) {
    Matrix3X p;
    to_matrix(mesh, p);
    ALMGeometrySolver<3> solver;
    wiremesh_constraints(solver, mesh, ref_mesh, p, min_angle_radian, max_angle_radian, edge_length, closeness_weight, laplacian_weight);
    return finish(solver, mesh, p, penalty, max_iter, Anderson_m, save_history);
}

GeoApp::GeoApp(Kind kind, const PolyMesh &mesh, const PolyMesh &ref_mesh, const double *prm) : mesh_(mesh) {
    to_matrix(mesh_, p_);
    if (kind == PLANARITY)
        planarity_constraints(solver_, mesh_, ref_mesh, p_, prm[1], prm[2], prm[3]);
    else
        wiremesh_constraints(solver_, mesh_, ref_mesh, p_, prm[1], prm[2], prm[3], prm[4], prm[5]);
    rel_residual_eps_ = 1e-8 * average_edge_length(mesh_);
    ok_ = solver_.setup_ADMM(p_.cols(), prm[0]);
    if (!ok_) std::cerr << "Error: unable to initialize solver" << std::endl;
}

void GeoApp::stats(double *out8) {
    for (int i = 0; i < 4; ++i) out8[i] = solver_.setup_counts[i];
    double st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (solver_.device_factor()) aaadmm_ldlt_stats(solver_.device_factor(), st);
    out8[4] = st[4];
    out8[5] = st[1];
    out8[6] = st[2];
    out8[7] = st[7];
}

OptimizeResult GeoApp::solve(int max_iter, int Anderson_m, bool save_history) {
    OptimizeResult R;
    if (!ok_) return R;
    const size_t before = solver_.function_values_.size();
    solver_.solve_ADMM(p_, rel_residual_eps_, max_iter, Anderson_m);
    if (save_history) solver_.save(Anderson_m);
    R.ok = true;
    R.function_values.assign(solver_.function_values_.begin() + before, solver_.function_values_.end());
    R.elapsed_time.assign(solver_.elapsed_time_.begin() + before, solver_.elapsed_time_.end());
    R.resets = solver_.reset_count;
    R.mesh = mesh_;
    const Matrix3X &x = solver_.get_solution();
    std::copy(x.data(), x.data() + x.size(), R.mesh.V.begin());
    return R;
}

// ---- reports --------------------------------------------------------------------------------------------------------
void planarity_error(const PolyMesh &mesh, std::vector<double> &per_face, double *max_err, double *mean_err) {
    const double el = average_edge_length(mesh);
    per_face.assign(mesh.n_faces(), 0.0);
    for (int f = 0; f < mesh.n_faces(); ++f) {
        const int n = mesh.valence(f);
        double c[3] = {0, 0, 0};
        for (int i = 0; i < n; ++i)
            for (int r = 0; r < 3; ++r) c[r] += P(mesh, mesh.face(f)[i])[r] / n;
        // smallest eigenvector of the 3 x 3 scatter matrix (cyclic Jacobi) = normal of the best-fitting plane
        double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, Q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int i = 0; i < n; ++i) {
            double d[3];
            for (int r = 0; r < 3; ++r) d[r] = P(mesh, mesh.face(f)[i])[r] - c[r];
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) S[a][b] += d[a] * d[b];
        }
        for (int sweep = 0; sweep < 30; ++sweep)
            for (int a = 0; a < 2; ++a)
                for (int b = a + 1; b < 3; ++b) {
                    if (std::fabs(S[a][b]) < 1e-300) continue;
                    const double th = 0.5 * std::atan2(2.0 * S[a][b], S[b][b] - S[a][a]);
                    const double cs = std::cos(th), sn = std::sin(th);
                    for (int k = 0; k < 3; ++k) {
                        const double x = S[k][a], y = S[k][b];
                        S[k][a] = cs * x - sn * y;
                        S[k][b] = sn * x + cs * y;
                    }
                    for (int k = 0; k < 3; ++k) {
                        const double x = S[a][k], y = S[b][k];
                        S[a][k] = cs * x - sn * y;
                        S[b][k] = sn * x + cs * y;
                    }
                    for (int k = 0; k < 3; ++k) {
                        const double x = Q[k][a], y = Q[k][b];
                        Q[k][a] = cs * x - sn * y;
                        Q[k][b] = sn * x + cs * y;
                    }
                }
        int lo = 0;
        for (int a = 1; a < 3; ++a)
            if (S[a][a] < S[lo][lo]) lo = a;
        double worst = 0;
        for (int i = 0; i < n; ++i) {
            double s = 0;
            for (int r = 0; r < 3; ++r) s += Q[r][lo] * (P(mesh, mesh.face(f)[i])[r] - c[r]);
            worst = std::max(worst, std::fabs(s));
        }
        per_face[f] = el > 0 ? worst / el : worst;
    }
    double mx = 0, sum = 0;
    for (double e : per_face) mx = std::max(mx, e), sum += e;
    if (max_err) *max_err = mx;
    if (mean_err) *mean_err = per_face.empty() ? 0.0 : sum / per_face.size();
}

bool ref_surface_distance(const PolyMesh &mesh, const PolyMesh &ref_mesh, double *max_err, double *mean_err) {
    const std::vector<int> tris = triangles_of(ref_mesh);
    const int nq = mesh.n_vertices();
    std::vector<double> closest((size_t)3 * nq);
    std::vector<int> tri(nq);
    if (aaadmm_geo_closest_points(ref_mesh.V.data(), ref_mesh.n_vertices(), tris.data(), (int)tris.size() / 3, mesh.V.data(), nq,
                                  closest.data(), tri.data()) != 0)
        return false;
    const double el = average_edge_length(mesh);
    double mx = 0, sum = 0;
    for (int i = 0; i < nq; ++i) {
        double d2 = 0;
        for (int r = 0; r < 3; ++r) d2 += (closest[3 * (size_t)i + r] - mesh.V[3 * (size_t)i + r]) * (closest[3 * (size_t)i + r] - mesh.V[3 * (size_t)i + r]);
        const double d = std::sqrt(d2) / el;
        mx = std::max(mx, d);
        sum += d;
    }
    if (max_err) *max_err = mx;
    if (mean_err) *mean_err = nq ? sum / nq : 0.0;
    return true;
}

}  // namespace geoapp
}  // namespace aaadmm
