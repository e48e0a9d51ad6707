// Remaining element types of the hard_zxu / xzu solvers, as device batches:
//   TriEnergyTerm::prox        xzu/src/TriEnergyTerm.cpp:77-107 (average with the polar factor, then clamp the
//                              column norms) and hard/src/TriEnergyTerm.cpp:74-105 (clamp (1 + sigma) / 2)
//   Collision::prox            hard/src/CollisionEnergyTerm.hpp:79-91 over the analytic passive objects of
//                              hard/src/PassiveObject.hpp:32-136 (Floor, SlideFloor, Sphere, PlaneAndHalfSphere, Cylinder)
//   SpringPin::prox            hard/src/SpringEnergyTerm.hpp:66-70
// One thread per element. Compiled without FMA contraction like the other per-element kernels.
#pragma once
#include "common.cuh"

namespace aaadmm {

enum { PASSIVE_FLOOR = 0, PASSIVE_SLIDE_FLOOR = 1, PASSIVE_SPHERE = 2, PASSIVE_PLANE_HALF_SPHERE = 3, PASSIVE_CYLINDER = 4 };

// z: n column-major 3x2 blocks (6 doubles each), in place. variant 0 = xzu, 1 = hard_zxu.
void launch_tri_prox(int variant, double *z, int64_t n, double limit_min, double limit_max, cudaStream_t s);
// types [n_objs], prm [n_objs][7] = {cx, cy, cz, nx, ny, nz, radius} (Floor: cx = y); z: n points in place.
void launch_collision_prox(int n_objs, const int *types, const double *prm, double *z, int64_t n, cudaStream_t s);
void launch_spring_prox(double *z, const double *pins, const int *active, int64_t n, cudaStream_t s);

}  // namespace aaadmm
