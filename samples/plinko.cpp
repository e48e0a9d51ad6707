// Headless drop-onto-obstacles sample in the style of the reference's samples/Asia2019/plinkohit.cpp:38-96 and
// plinkopony.cpp:60-117: a tet mesh read from TetGen files (mcl::meshio::load_elenode), added with
// binding::add_tetmesh, analytic obstacles (add_obstacle) and a Collision term on every vertex (set_collisions);
// written against the drop-in classes of aa-admm_b200/host, the solver calls are the reference's own.
//
//   g++ -std=c++17 -O2 -Iaa-admm_b200/host samples/plinko.cpp -Laa-admm_b200 -laaadmm_host -laaadmm_b200 \
//       -Wl,-rpath,$PWD/aa-admm_b200 -o plinko
//   ./plinko -mesh <path without .ele/.node> [-it 13] [-a 1 -am 5] [-frames 5] [-scale 1] [-lift 0]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "MeshIO.hpp"  // instead of MCL/MeshIO.hpp + samples/utils/AddMeshes.hpp
#include "Solver.hpp"

int main(int argc, char **argv) {
    admm::Solver::Settings settings;
    settings.admm_iters = 13;  // plinkohit.cpp:57
    settings.verbose = 0;
    settings.write_residual_file = false;
    if (settings.parse_args(argc, argv)) return EXIT_SUCCESS;
    std::string path;
    int frames = 5;
    float scale = 1.f, lift = 0.f;
    for (int i = 1; i + 1 < argc; ++i) {
        if (!strcmp(argv[i], "-mesh")) path = argv[i + 1];
        if (!strcmp(argv[i], "-frames")) frames = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "-scale")) scale = (float)atof(argv[i + 1]);
        if (!strcmp(argv[i], "-lift")) lift = (float)atof(argv[i + 1]);
    }
    if (path.empty()) {
        fprintf(stderr, "usage: plinko -mesh <path without .ele/.node> [-it n] [-a 1 -am m] [-frames n] [-scale s] [-lift y]\n");
        return EXIT_FAILURE;
    }
    mcl::TetMesh::Ptr mesh = mcl::TetMesh::create();
    if (!mcl::meshio::load_elenode(mesh.get(), path)) return EXIT_FAILURE;
    mesh->flags |= binding::LINEAR;
    for (mcl::Vec3f &v : mesh->vertices) {  // the reference applies an mcl::XForm (scale, then translate)
        for (int j = 0; j < 3; ++j) v[j] = scale * v[j];
        v[1] += lift;
    }

    admm::Solver solver;
    binding::add_tetmesh(&solver, mesh, admm::Lame(1e6, 0.399), false);

    // obstacles under the mesh: the lowest vertices already touch the floor in the first frame
    float lo = mesh->vertices[0][1];
    for (const mcl::Vec3f &v : mesh->vertices) lo = std::min(lo, v[1]);
    solver.add_obstacle(std::make_shared<admm::Floor>(lo + 0.05));
    solver.add_obstacle(std::make_shared<admm::PlaneAndHalfSphere>(admm::Vec3{0.0, lo + 0.03, 0.0}, 0.3));
    solver.add_obstacle(std::make_shared<admm::Cylinder>(admm::Vec3{1.0, lo - 0.5, 0.0}, 0.55));
    std::vector<int> all_idx(mesh->vertices.size());
    std::iota(all_idx.begin(), all_idx.end(), 0);
    solver.set_collisions(all_idx);  // in place, as plinkohit.cpp:96 with the vertices' own positions

    if (!solver.initialize(settings)) return EXIT_FAILURE;
    for (int f = 0; f < frames; ++f) {
        solver.step();
        const size_t k = solver.step_comb_residual.size();
        printf("frame %d: %zu iterations, %d rejected, combined residual %.6e -> %.6e, loop %.3f ms\n", f, k, solver.reject_num,
               k ? solver.step_comb_residual[0] : 0.0, k ? solver.step_comb_residual[k - 1] : 0.0, solver.runtime_data().loop_ms);
    }
    double s = 0, low = 1e300;
    for (double v : solver.m_x) s += v;
    for (size_t v = 0; v < solver.m_x.size() / 3; ++v) low = std::min(low, solver.m_x[3 * v + 1]);
    printf("checksum of positions %.12e, lowest y %.6f (floor at %.6f)\n", s, low, (double)lo + 0.05);
    return EXIT_SUCCESS;
}
