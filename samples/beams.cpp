// Headless version of the reference's beam sample (admm_anderson_xzu/samples/Asia2019/beams.cpp:94-160 without the
// GLFW viewer), written against the drop-in classes of aa-admm_b200/host: the solver calls are the reference's own
// (add_nodes, create_tets_from_mesh, set_pins, initialize, step), only the include path changed.
//
//   g++ -std=c++17 -O2 -Iaa-admm_b200/host samples/beams.cpp -Laa-admm_b200 -laaadmm_host -laaadmm_b200 \
//       -Wl,-rpath,$PWD/aa-admm_b200 -o beams
//   ./beams -it 100 -a 1 -am 5 [-frames 3] [-dims 12 3 3] [-xzu] [-save]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "Solver.hpp"      // instead of admm_anderson_*/src/Solver.hpp
#include "beam_scene.hpp"  // instead of MCL/ShapeFactory.hpp + MCL/TetMesh.hpp

int main(int argc, char **argv) {
    admm::Solver::Settings settings;
    settings.admm_iters = 100;
    settings.Anderson_m = 5;
    settings.acceleration_type = admm::Solver::Settings::ANDERSON;
    settings.verbose = 0;
    settings.write_residual_file = false;
    if (settings.parse_args(argc, argv)) return EXIT_SUCCESS;
    int frames = 3, dims[3] = {12, 3, 3};
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "-frames") && i + 1 < argc) frames = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "-dims") && i + 3 < argc)
            for (int k = 0; k < 3; ++k) dims[k] = atoi(argv[i + 1 + k]);
        if (!strcmp(argv[i], "-xzu")) settings.ordering = admm::Solver::Settings::XZU;
        // -save: Solver::save() writes ./result/residual-<m|no>.txt after every step like the reference (the directory
        // must exist, as for the reference)
        if (!strcmp(argv[i], "-save")) settings.write_residual_file = true;
    }

    admm::Solver solver;
    aaadmm::BeamPins pins;
    const float shifts[3] = {1.75f, 0.f, -1.75f};
    for (int b = 0; b < 3; ++b) {
        aaadmm::BeamMesh mesh = aaadmm::make_beam(dims[0], dims[1], dims[2], shifts[b]);
        std::vector<float> masses3(mesh.verts.size());
        for (size_t i = 0; i < masses3.size(); ++i) masses3[i] = mesh.masses[i / 3];
        const int prev = (int)solver.m_x.size() / 3;
        solver.add_nodes(mesh.verts.data(), masses3.data(), mesh.n_verts());
        const admm::Lame lame = admm::Lame::soft_rubber();
        if (b == 0 || settings.ordering == admm::Solver::Settings::HARD_ZXU)
            admm::create_tets_from_mesh<float, admm::TetEnergyTerm>(solver.energyterms, mesh.verts.data(), mesh.tets.data(),
                                                                    mesh.n_tets(), lame, prev);
        else if (b == 1)
            admm::create_tets_from_mesh<float, admm::NeoHookeanTet>(solver.energyterms, mesh.verts.data(), mesh.tets.data(),
                                                                    mesh.n_tets(), lame, prev);
        else
            admm::create_tets_from_mesh<float, admm::StVKTet>(solver.energyterms, mesh.verts.data(), mesh.tets.data(),
                                                              mesh.n_tets(), lame, prev);
        aaadmm::find_pins(mesh, prev, pins);
    }
    auto pin_points = [&]() {
        std::vector<admm::Vec3> pts(pins.idx.size());
        for (size_t i = 0; i < pts.size(); ++i) pts[i] = {pins.points[3 * i], pins.points[3 * i + 1], pins.points[3 * i + 2]};
        return pts;
    };
    aaadmm::stretch_pins(pins, settings.timestep_s);  // beams.cpp:126: stretched once before initialize
    solver.set_pins(pins.idx, pin_points());
    if (!solver.initialize(settings)) return EXIT_FAILURE;
    for (int f = 0; f < frames; ++f) {
        aaadmm::stretch_pins(pins, settings.timestep_s);
        solver.set_pins(pins.idx, pin_points());
        solver.step();
        const size_t n = solver.step_comb_residual.size();
        printf("frame %d: %zu iterations, %d rejected, combined residual %.6e -> %.6e, loop %.3f ms\n", f, n, solver.reject_num,
               n ? solver.step_comb_residual[0] : 0.0, n ? solver.step_comb_residual[n - 1] : 0.0, solver.runtime_data().loop_ms);
    }
    double s = 0;
    for (double v : solver.m_x) s += v;
    printf("checksum of positions %.12e\n", s);
    return EXIT_SUCCESS;
}
