// Mesh containers, file formats and the solver glue of the reference's samples, without Eigen / mclscene:
//   mcl::TetMesh / mcl::TriangleMesh   deps/mclscene/include/MCL/TetMesh.hpp, TriangleMesh.hpp (vertices, tets / faces,
//                                      flags, weighted_masses :297-315 / :281-296)
//   mcl::meshio::load_elenode / save_elenode / load_obj / save_obj
//                                      deps/mclscene/include/MCL/MeshIO.hpp:55-330 (TetGen ASCII .ele/.node with
//                                      0- or 1-based indices detected from the first row, inverted tets re-ordered;
//                                      Wavefront .obj with v / f records, `a/b/c` face tokens, triangles only)
//   binding::add_tetmesh / add_trimesh samples/utils/AddMeshes.hpp:97-239 (float32 nodes, volume- / area-weighted
//                                      masses with the reference's densities, one energy term per element)
// Same names, arguments and error behaviour, so that a reference sample compiles against this header with only its
// include lines changed. Vertices are float32 at the same points as in the reference; masses are accumulated in
// float32 in element order.
#pragma once
#include <array>
#include <memory>
#include <string>
#include <vector>

#include "Solver.hpp"

namespace mcl {

typedef std::array<float, 3> Vec3f;
typedef std::array<int, 3> Vec3i;
typedef std::array<int, 4> Vec4i;

class TetMesh {
public:
    typedef std::shared_ptr<TetMesh> Ptr;
    static Ptr create() { return std::make_shared<TetMesh>(); }
    std::vector<Vec3f> vertices;
    std::vector<Vec4i> tets;
    int flags = 0;
    void weighted_masses(std::vector<float> &m, float density_kgm3 = 1100.0f);
    // vertices of the faces that belong to exactly one tet (MCL/TetMesh.hpp:317-342; ascending here, hash order there)
    void surface_inds(std::vector<int> &surf_inds);
    void clear() {
        vertices.clear();
        tets.clear();
    }
};

class TriangleMesh {
public:
    typedef std::shared_ptr<TriangleMesh> Ptr;
    static Ptr create() { return std::make_shared<TriangleMesh>(); }
    std::vector<Vec3f> vertices;
    std::vector<Vec3i> faces;
    int flags = 0;
    void weighted_masses(std::vector<float> &m, float density_kgm2 = 0.4f);
    void clear() {
        vertices.clear();
        faces.clear();
    }
};

namespace meshio {
// false + message on std::cerr for a missing / inconsistent file; load_elenode throws std::runtime_error for an empty mesh
bool load_obj(TriangleMesh *mesh, std::string file);
bool save_obj(const TriangleMesh *mesh, std::string filename);
bool load_elenode(TetMesh *mesh, std::string file);  // `file` without the .ele / .node extension
bool save_elenode(const TetMesh *mesh, std::string file);
}  // namespace meshio

}  // namespace mcl

namespace binding {

enum MeshFlags {
    NOSELFCOLLISION = 1 << 1,
    LINEAR = 1 << 2,  // default when mesh->flags == 0
    NEOHOOKEAN = 1 << 3,
    STVK = 1 << 4,
};

// Nodes (masses from rubber, 1522 kg/m^3) + one tet energy term per tet of the flagged material. The reference also
// registers a self-collision mesh unless NOSELFCOLLISION is set; that collider is outside the device path.
void add_tetmesh(admm::Solver *solver, std::shared_ptr<mcl::TetMesh> &mesh, const admm::Lame &lame = admm::Lame::rubber(),
                 bool verbose = true);
// Nodes (masses from 1 kg/m^2) + one TriEnergyTerm per face.
void add_trimesh(admm::Solver *solver, std::shared_ptr<mcl::TriangleMesh> &mesh,
                 const admm::Lame &lame = admm::Lame::rubber(), bool verbose = true);

}  // namespace binding
