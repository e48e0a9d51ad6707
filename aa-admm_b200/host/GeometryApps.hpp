// Front-end of the reference's two Geometry applications without OpenMesh / libigl:
//   * polygon meshes from / to Wavefront .obj (what OpenMesh::IO::read_mesh / write_mesh(..., 16) do for the shipped
//     Geometry_model files), with the connectivity numbered as OpenMesh numbers it (vertices and faces in file order,
//     an edge when add_face first meets it, halfedge 0 of an edge in the direction of the face that created it);
//   * average_edge_length, subdivide_and_smooth_mesh           Geometry/MeshTypes.h:147-161, 214-342
//   * Parameters (Options.txt)                                  Geometry/Parameters.h:36-140
//   * the constraint recipes of optimize_mesh                   Geometry/PlanarityOpt.cpp:134-246,
//                                                               Geometry/WireMeshOpt.cpp:226-289
// The solver underneath is the device-backed ALMGeometrySolver<3> mirror (GeometrySolver.hpp).
// samples/planarity.cpp and samples/wiremesh.cpp are the two mains (same command lines as the reference's).
#pragma once
#include <string>
#include <vector>

#include "GeometrySolver.hpp"

namespace aaadmm {
namespace geoapp {

struct PolyMesh {
    std::vector<double> V;      // 3 per vertex
    std::vector<int> face_ptr;  // n_faces + 1
    std::vector<int> face_idx;  // vertex ids, face by face
    int n_vertices() const { return (int)(V.size() / 3); }
    int n_faces() const { return face_ptr.empty() ? 0 : (int)face_ptr.size() - 1; }
    int valence(int f) const { return face_ptr[f + 1] - face_ptr[f]; }
    const int *face(int f) const { return face_idx.data() + face_ptr[f]; }
    void add_vertex(double x, double y, double z) {
        V.push_back(x);
        V.push_back(y);
        V.push_back(z);
    }
    void add_face(const std::vector<int> &ids) {
        if (face_ptr.empty()) face_ptr.push_back(0);
        face_idx.insert(face_idx.end(), ids.begin(), ids.end());
        face_ptr.push_back((int)face_idx.size());
    }
};

// Half-edge connectivity in OpenMesh's numbering. Throws std::runtime_error on a non-manifold edge (OpenMesh's add_face
// refuses such faces: "complex edge").
struct Connectivity {
    int n_edges = 0;
    std::vector<int> edge_from, edge_to;        // halfedge 0 of every edge
    std::vector<int> edge_face0, edge_face1;    // face of halfedge 0 / of the opposite halfedge (-1: boundary)
    std::vector<int> face_edge;                 // per face corner i: the edge (v_i, v_{i+1})
    std::vector<char> vertex_boundary;          // is_boundary(vertex); isolated vertices count as boundary
    std::vector<std::vector<int>> ring;         // interior vertices: one-ring in rotation order; boundary: neighbours
    explicit Connectivity(const PolyMesh &m);
    bool edge_boundary(int e) const { return edge_face1[e] < 0; }
};

bool read_obj(const std::string &path, PolyMesh &mesh);            // v / f records; f tokens "a", "a/b", "a/b/c", "a//c"; 1-based or negative
bool write_obj(const PolyMesh &mesh, const std::string &path);     // 16 significant digits, as write_mesh(..., Default, 16)
double average_edge_length(const PolyMesh &mesh);
PolyMesh subdivide_and_smooth_mesh(const PolyMesh &mesh);

struct Parameters {
    int iter = 1, anderson_m = 5;
    double elasticity, time_step = 0.033;
    Parameters();
    bool load(const char *filename);
    bool valid_parameters() const;
    void output() const;
};

struct OptimizeResult {
    bool ok = false;
    std::vector<double> function_values, elapsed_time;  // logged combined residuals / seconds (ALMGeometrySolver::save)
    int resets = 0;
    PolyMesh mesh;  // input mesh with the optimised vertex positions
};

// PlanarityOpt.cpp: optimize_mesh (closeness to the reference surface, relative / plain uniform Laplacians, one plane
// constraint per face with more than 3 vertices). Anderson_m = 0 means no acceleration. Writes result/residual-*.txt
// like the reference when save_history is set.
OptimizeResult planarity_optimize(const PolyMesh &mesh, const PolyMesh &ref_mesh, int max_iter, int Anderson_m,
                                  double penalty_parameter, double closeness_weight, double laplacian_weight,
                                  double relative_laplacian_weight, bool save_history = true);
// WireMeshOpt.cpp: optimize_mesh on an all-quad mesh (4 angle constraints per face, one edge-length constraint per edge,
// closeness to the reference surface; laplacian_weight > 0 adds setup_quad_laplacian_matrix).
OptimizeResult wiremesh_optimize(const PolyMesh &mesh, const PolyMesh &ref_mesh, int max_iter, int Anderson_m,
                                 double penalty_parameter, double min_angle_radian, double max_angle_radian,
                                 double edge_length, double closeness_weight, double laplacian_weight,
                                 bool save_history = true);

// The same two recipes with the setup kept: build once (constraints, setup_ADMM = system matrix + factorisation + device
// upload), then solve any number of times from the mesh's own positions (what a bench step or a parameter study does).
class GeoApp {
public:
    enum Kind { PLANARITY = 0, WIREMESH = 1 };
    // prm: PLANARITY {penalty, closeness_w, laplacian_w, relative_laplacian_w};
    //      WIREMESH  {penalty, min_angle, max_angle, edge_length, closeness_w, laplacian_w}
    GeoApp(Kind kind, const PolyMesh &mesh, const PolyMesh &ref_mesh, const double *prm);
    bool ok() const { return ok_; }
    OptimizeResult solve(int max_iter, int Anderson_m, bool save_history = false);
    const aaadmm_step_result &last_result() const { return solver_.last_result; }
    // out8 = points, hard constraints, columns of z / u, soft constraints (one per point or one batch), nnz(L), fronts,
    // tree levels, algorithmic bytes of one factor apply
    void stats(double *out8);
    aaadmm_ldlt *device_factor() { return solver_.device_factor(); }  // developer aid: per-task trace of the applies
private:
    ALMGeometrySolver<3> solver_;
    PolyMesh mesh_;
    Matrix3X p_;
    double rel_residual_eps_ = 0.0;
    bool ok_ = false;
};

// Reports of the reference's mains (normalised by the average edge length): per-face planarity error
// (PlanarityOpt.cpp:57-107) and distance to the reference surface (PlanarityOpt.cpp:109-132; needs the GPU library).
void planarity_error(const PolyMesh &mesh, std::vector<double> &per_face, double *max_err, double *mean_err);
bool ref_surface_distance(const PolyMesh &mesh, const PolyMesh &ref_mesh, double *max_err, double *mean_err);

}  // namespace geoapp
}  // namespace aaadmm
