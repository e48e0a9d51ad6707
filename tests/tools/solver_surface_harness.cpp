// Test harness (tests/test_host_cpu.py, tests/test_gpu_parity.py): the public members of admm::Solver beyond the hot path
// (hard/src/Solver.hpp:83-128): Eigen-style access to m_x, set_pins with a foreign 3-vector type, surface_inds through
// binding::add_tetmesh, add_dynamic_collider, and - when a GPU is present - save_matrix and the per-iteration time stamps.
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>

#include "MeshIO.hpp"
#include "Solver.hpp"
#include "beam_scene.hpp"

struct MyVec3 {  // stands in for Eigen::Vector3d
    double v[3];
    double operator[](int i) const { return v[i]; }
};

int main(int argc, char **argv) {
    const int cx = argc > 1 ? atoi(argv[1]) : 4;
    aaadmm::BeamMesh bm = aaadmm::make_beam(cx, 2, 2, 0.f);
    std::shared_ptr<mcl::TetMesh> mesh = mcl::TetMesh::create();
    for (int i = 0; i < bm.n_verts(); ++i) mesh->vertices.push_back({bm.verts[3 * i], bm.verts[3 * i + 1], bm.verts[3 * i + 2]});
    for (int t = 0; t < bm.n_tets(); ++t) mesh->tets.push_back({bm.tets[4 * t], bm.tets[4 * t + 1], bm.tets[4 * t + 2], bm.tets[4 * t + 3]});
    mesh->flags = binding::LINEAR | binding::NOSELFCOLLISION;
    admm::Solver solver;
    binding::add_tetmesh(&solver, mesh, admm::Lame::soft_rubber(), false);
    printf("surface_inds %zu of %d vertices\n", solver.surface_inds.size(), bm.n_verts());
    {   // four tets around an interior point: the centre (vertex 4) is the only vertex not on the surface
        std::shared_ptr<mcl::TetMesh> star = mcl::TetMesh::create();
        star->vertices = {{0.f, 0.f, 0.f}, {1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}, {0.25f, 0.25f, 0.25f}};
        star->tets = {{4, 1, 2, 3}, {0, 4, 2, 3}, {0, 1, 4, 3}, {0, 1, 2, 4}};
        std::vector<int> surf;
        star->surface_inds(surf);
        if (surf != std::vector<int>({0, 1, 2, 3})) return 2;
        admm::Solver s2;
        s2.m_x.resize(6, 0.0);  // two nodes already there: the indices are offset (AddMeshes.hpp:135)
        s2.m_v.resize(6, 0.0);
        s2.m_masses.resize(6, 1.0);
        binding::add_tetmesh(&s2, star, admm::Lame::soft_rubber(), false);
        if (s2.surface_inds != std::vector<int>({2, 3, 4, 5})) return 2;
    }
    // Eigen-style access
    admm::Vec3 p0 = solver.m_x.segment<3>(0);
    solver.m_x.segment<3>(0) = MyVec3{{p0[0], p0[1], p0[2]}};
    if (solver.m_x.rows() != 3 * bm.n_verts() || solver.m_x[0] != p0[0]) return 3;
    aaadmm::BeamPins pins;
    aaadmm::find_pins(bm, 0, pins);
    std::vector<MyVec3> pts(pins.idx.size());
    for (size_t i = 0; i < pins.idx.size(); ++i) pts[i] = MyVec3{{pins.points[3 * i], pins.points[3 * i + 1], pins.points[3 * i + 2]}};
    solver.set_pins(pins.idx, pts);
    try {
        solver.add_dynamic_collider(std::make_shared<admm::Floor>(0.0));
        return 4;
    } catch (const std::runtime_error &) {
    }
    if (aaadmm_device_count() <= 0) {
        printf("no device: host part ok\n");
        return 0;
    }
    admm::Solver::Settings settings;
    settings.verbose = 0;
    settings.admm_iters = 40;
    settings.Anderson_m = 5;
    settings.acceleration_type = admm::Solver::Settings::ANDERSON;
    settings.write_residual_file = false;
    if (!solver.initialize(settings)) return 5;
    solver.save_matrix(argc > 2 ? argv[2] : "termA.mtx");
    aaadmm::stretch_pins(pins, settings.timestep_s);
    for (size_t i = 0; i < pins.idx.size(); ++i) pts[i] = MyVec3{{pins.points[3 * i], pins.points[3 * i + 1], pins.points[3 * i + 2]}};
    solver.set_pins(pins.idx, pts);
    solver.step();
    const std::vector<double> &t = solver.runtime_data().step_time;
    if (t.size() != solver.step_comb_residual.size() || t.empty()) return 6;
    for (size_t i = 0; i < t.size(); ++i) {
        if (!(t[i] > 0.0) || (i > 0 && !(t[i] > t[i - 1]))) return 7;  // cumulative device time, strictly increasing
    }
    if (!(t.back() <= solver.runtime_data().loop_ms * 1.05 + 0.05)) return 8;
    printf("iterations %zu, logged time %.4f ms of loop %.4f ms, rejects %d\n", t.size(), t.back(), solver.runtime_data().loop_ms, solver.reject_num);
    return 0;
}
