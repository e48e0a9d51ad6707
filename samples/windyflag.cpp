// Headless flag-in-the-wind sample in the style of the reference's samples/Asia2019/windyflag.cpp:63-130 (material,
// strain limits, two pinned corners, WindForce over all faces) plus the obstacle pattern of plinkohit.cpp:84-96
// (add_obstacle + set_collisions), written against the drop-in classes of aa-admm_b200/host. The solver calls are the
// reference's own (add_nodes, create_tris_from_mesh, set_pins, ext_forces, add_obstacle, set_collisions, initialize,
// step); the cloth is generated instead of read from samples/data/cloth.obj (mesh I/O is out of scope).
//
//   g++ -std=c++17 -O2 -Iaa-admm_b200/host samples/windyflag.cpp -Laa-admm_b200 -laaadmm_host -laaadmm_b200 \
//       -Wl,-rpath,$PWD/aa-admm_b200 -o windyflag
//   ./windyflag -it 100 -a 1 -am 5 [-frames 3] [-n 20] [-sphere]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "Solver.hpp"  // instead of admm_anderson_hard_zxu/src/Solver.hpp + TriEnergyTerm.hpp + ExplicitForce.hpp

int main(int argc, char **argv) {
    admm::Solver::Settings settings;
    settings.admm_iters = 100;
    settings.penalty = 1.0;
    settings.Anderson_m = 5;
    settings.acceleration_type = admm::Solver::Settings::ANDERSON;
    settings.verbose = 0;
    settings.write_residual_file = false;
    if (settings.parse_args(argc, argv)) return EXIT_SUCCESS;
    int frames = 3, n = 20;
    bool sphere = false;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "-frames") && i + 1 < argc) frames = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "-n") && i + 1 < argc) n = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "-sphere")) sphere = true;
    }

    // a vertical square flag of n x n cells in the x-y plane, two triangles per cell
    std::vector<float> verts, masses;
    std::vector<int> faces;
    for (int i = 0; i <= n; ++i)
        for (int j = 0; j <= n; ++j) {
            verts.push_back((float)i / n);
            verts.push_back(1.f + (float)j / n);
            verts.push_back(0.01f * (float)((i * 7 + j * 3) % 5));
        }
    auto vid = [&](int i, int j) { return i * (n + 1) + j; };
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const int a = vid(i, j), b = vid(i + 1, j), c = vid(i + 1, j + 1), d = vid(i, j + 1);
            faces.insert(faces.end(), {a, b, c, a, c, d});
        }
    const int n_verts = (n + 1) * (n + 1);
    masses.assign((size_t)3 * n_verts, 1.0f / n_verts);

    admm::Solver solver;
    solver.add_nodes(verts.data(), masses.data(), n_verts);
    // windyflag.cpp:84-86; with the obstacle the sheet is soft rubber like the plinko solids: the Collision terms
    // carry the weight of soft rubber (CollisionEnergyTerm.hpp:63-69) and a 50 Pa cloth next to them makes the
    // reference itself run into NaN in the second frame
    admm::Lame very_soft_rubber = sphere ? admm::Lame::soft_rubber() : admm::Lame(50, 0.1);
    very_soft_rubber.limit_min = 0.95;
    very_soft_rubber.limit_max = 1.05;
    admm::create_tris_from_mesh<float, admm::TriEnergyTerm>(solver.energyterms, verts.data(), faces.data(),
                                                            (int)faces.size() / 3, very_soft_rubber, 0);
    std::vector<int> pins = {vid(0, 0), vid(0, n)};  // the two corners at the pole
    solver.set_pins(pins);

    std::shared_ptr<admm::WindForce> wind(new admm::WindForce(faces));  // windyflag.cpp:124-126
    wind->direction = {10 * 2.5, 0, 2 * 2.5};
    solver.ext_forces.push_back(wind);

    if (sphere) {  // plinkohit.cpp:84-96 pattern: an obstacle + a collision term on every free vertex
        solver.add_obstacle(std::make_shared<admm::Sphere>(admm::Vec3{0.7, 1.4, 0.35}, 0.34));
        std::vector<int> all;
        for (int v = 0; v < n_verts; ++v)
            if (v != pins[0] && v != pins[1]) all.push_back(v);
        solver.set_collisions(all);
    }

    if (!solver.initialize(settings)) return EXIT_FAILURE;
    for (int f = 0; f < frames; ++f) {
        solver.step();
        const size_t k = solver.step_comb_residual.size();
        printf("frame %d: %zu iterations, %d rejected, combined residual %.6e -> %.6e, loop %.3f ms\n", f, k, solver.reject_num,
               k ? solver.step_comb_residual[0] : 0.0, k ? solver.step_comb_residual[k - 1] : 0.0, solver.runtime_data().loop_ms);
    }
    double s = 0, far = 0;
    for (double v : solver.m_x) s += v;
    for (int v = 0; v < n_verts; ++v) far = std::max(far, solver.m_x[3 * (size_t)v + 2]);
    printf("checksum of positions %.12e, largest z %.6f\n", s, far);
    return EXIT_SUCCESS;
}
