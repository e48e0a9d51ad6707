// The reference's WireMeshOpt application (Geometry/WireMeshOpt.cpp:340-407) against the drop-in classes of
// aa-admm_b200/host: same command line, same defaults (one subdivision + smoothing step, target edge length = half the
// average edge length, angles in [pi/4, 3 pi/4], penalty 1000), same outputs (result/residual-<m>.txt, the optimised
// mesh as .obj with 16 digits).
//
//   g++ -std=c++17 -O2 -Iaa-admm_b200/host samples/wiremesh.cpp -Laa-admm_b200 -laaadmm_host -laaadmm_b200 \
//       -Wl,-rpath,$PWD/aa-admm_b200 -o WireMeshOpt
//   ./WireMeshOpt <INPUT_POLY_MESH> <REF_TRI_MESH> <OPTIONS_FILE> <OUTPUT_MESH>
#include <cmath>
#include <iostream>

#include "GeometryApps.hpp"  // instead of ALMGeometrySolver.h, MeshTypes.h, Constraint.h, Parameters.h

using namespace aaadmm::geoapp;

int main(int argc, char **argv) {
    if (argc != 5) {
        std::cout << "Usage:   <WireMeshOpt>  <INPUT_POLY_MESH>  <REF_TRI_MESH>  <OPTIONS_FILE>  <OUTPUT_MESH>" << std::endl;
        return 1;
    }
    PolyMesh mesh;
    if (!read_obj(argv[1], mesh)) {
        std::cerr << "Error: unable to read input mesh from the file " << argv[1] << std::endl;
        return 1;
    }
    PolyMesh ref_mesh;
    if (!read_obj(argv[2], ref_mesh)) {
        std::cerr << "Error: unable to read referece mesh from file " << argv[2] << std::endl;
        return 1;
    }
    try {
        double edge_length = average_edge_length(mesh);
        const double pi = 3.14159265358979323846;
        const double min_angle_radian = pi * 0.25, max_angle_radian = pi * 0.75;
        PolyMesh sub_mesh = subdivide_and_smooth_mesh(mesh);
        edge_length *= 0.5;
        std::cout << "target length = " << edge_length << std::endl;

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

        const double closeness_weight = 1, laplacian_weight = -1, penalty_parameter = 1000;
        OptimizeResult R = wiremesh_optimize(sub_mesh, ref_mesh, param.iter, param.anderson_m, penalty_parameter, min_angle_radian,
                                             max_angle_radian, edge_length, closeness_weight, laplacian_weight);
        if (!R.ok) return 1;
        double dmx, dmean;
        if (ref_surface_distance(R.mesh, ref_mesh, &dmx, &dmean))
            std::cout << "Reference surface distance (normalized by edge length): Max " << dmx << ", Average " << dmean << std::endl;
        std::cout << "iterations " << R.function_values.size() << ", resets " << R.resets << ", combined residual "
                  << R.function_values.front() << " -> " << R.function_values.back() << std::endl;
        if (!write_obj(R.mesh, argv[4])) std::cerr << "Error: unable to save perturbed mesh to file " << argv[4] << std::endl;
    } catch (const std::exception &e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
