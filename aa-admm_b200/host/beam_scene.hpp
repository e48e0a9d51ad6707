// Host-side scene authoring for the beam workloads (SURVEY §8d cfg 1/4/5).
//
// O(n) restatement of what the reference builds with
//   mcl::factory::make_tet_blocks  (deps/mclscene/include/MCL/ShapeFactory.hpp:436-497)
//   mcl::TetMesh::refine           (deps/mclscene/include/MCL/TetMesh.hpp:235-295)   [O(tets x verts) there]
//   the centre/scale/offset of     samples/Asia2019/beams.cpp:83-103
//   mcl::TetMesh::weighted_masses  (deps/mclscene/include/MCL/TetMesh.hpp:297-315)
//   find_pins / stretch_beams      samples/Asia2019/beams.cpp:66-92,133-160
// Vertices and masses are float32 at the same points as in the reference.
#pragma once
#include <cstdint>
#include <vector>

namespace aaadmm {

struct BeamMesh {
    std::vector<float> verts;   // 3 per vertex, float32 as in mcl::TetMesh
    std::vector<int> tets;      // 4 per tet
    std::vector<float> masses;  // 1 per vertex (lumped, float32)
    int n_verts() const { return (int)(verts.size() / 3); }
    int n_tets() const { return (int)(tets.size() / 4); }
};

// One beam of cx*cy*cz unit cubes (5 tets each), merged vertices numbered by first
// appearance, centred, scaled to height 1 and shifted by y_shift (0 = no shift).
BeamMesh make_beam(int cx, int cy, int cz, float y_shift, float density_kgm3 = 1522.f);

// Appends `b` to `a` (tet indices offset by a.n_verts()).
void append_mesh(BeamMesh &a, const BeamMesh &b);

struct BeamPins {
    std::vector<int> idx;        // vertex ids (ascending within each mesh)
    std::vector<double> points;  // 3 per pin, current target
    std::vector<int> side;       // 0 = -x end, 1 = +x end
};
// beams.cpp:133-160 for ONE mesh whose vertices start at vertex_offset.
void find_pins(const BeamMesh &mesh_single, int vertex_offset, BeamPins &pins);
// beams.cpp:66-92: move -x pins by -dt*(1,0,0) and +x pins by +dt*(1,0,0).
void stretch_pins(BeamPins &pins, double dt);

}  // namespace aaadmm
