// The reference's PlanarityOpt application (Geometry/PlanarityOpt.cpp:248-332) against the drop-in classes of
// aa-admm_b200/host: same command line, same defaults, same outputs (result/residual-<m>.txt, the optimised mesh as
// .obj with 16 digits), OpenMesh / libigl replaced by host/GeometryApps.hpp, the solver by the device-backed
// ALMGeometrySolver<3>.
//
//   g++ -std=c++17 -O2 -Iaa-admm_b200/host samples/planarity.cpp -Laa-admm_b200 -laaadmm_host -laaadmm_b200 \
//       -Wl,-rpath,$PWD/aa-admm_b200 -o PlanarityOpt
//   ./PlanarityOpt <INPUT_MESH> <REFERENCE_MESH> <OPTION_FILES> <OUTPUT_MESH>
#include <iostream>

#include "GeometryApps.hpp"  // instead of ALMGeometrySolver.h, MeshTypes.h, TriMeshAABB.h, Constraint.h, Parameters.h

using namespace aaadmm::geoapp;

int main(int argc, char **argv) {
    if (argc != 5) {
        std::cout << "Usage:   <PlanarityOpt> <INPUT_MESH> <REFERENCE_MESH> <OPTION_FILES> <OUTPUT_MESH>" << std::endl;
        return 1;
    }
    PolyMesh mesh;
    if (!read_obj(argv[1], mesh)) {
        std::cerr << "Error: unable to read input mesh from file " << argv[1] << std::endl;
        return 1;
    }
    PolyMesh ref_mesh;
    if (!read_obj(argv[2], ref_mesh)) {
        std::cerr << "Error: unable to read reference mesh from file " << argv[2] << std::endl;
        return 1;
    }
    Parameters param;
    if (!param.load(argv[3])) {
        std::cerr << "Error: unable to load option file " << argv[3] << std::endl;
        return 1;
    }
    if (!param.valid_parameters()) {
        std::cerr << "Invalid filter options. Aborting..." << std::endl;
        return 1;
    }
    param.output();

    const double closeness_weight = 1, laplacian_weight = 0.0, relative_laplacian_weight = 0.1, penalty_parameter = 100000;
    try {
        OptimizeResult R = planarity_optimize(mesh, ref_mesh, param.iter, param.anderson_m, penalty_parameter, closeness_weight,
                                              laplacian_weight, relative_laplacian_weight);
        if (!R.ok) return 1;
        std::vector<double> err;
        double mx, mean, dmx, dmean;
        std::cout << "Before optimization:" << std::endl;
        planarity_error(mesh, err, &mx, &mean);
        std::cout << "Planarity error (normalized by edge length): max " << mx << ", average " << mean << std::endl;
        if (ref_surface_distance(mesh, ref_mesh, &dmx, &dmean))
            std::cout << "Reference surface distance (normalized by edge length): Max " << dmx << ", Average " << dmean << std::endl;
        std::cout << "After optimization:" << std::endl;
        planarity_error(R.mesh, err, &mx, &mean);
        std::cout << "Planarity error (normalized by edge length): max " << mx << ", average " << mean << std::endl;
        if (ref_surface_distance(R.mesh, ref_mesh, &dmx, &dmean))
            std::cout << "Reference surface distance (normalized by edge length): Max " << dmx << ", Average " << dmean << std::endl;
        std::cout << "iterations " << R.function_values.size() << ", resets " << R.resets << ", combined residual "
                  << R.function_values.front() << " -> " << R.function_values.back() << std::endl;
        if (!write_obj(R.mesh, argv[4])) std::cerr << "Error: unable to save result mesh to file " << argv[4] << std::endl;
    } catch (const std::exception &e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
