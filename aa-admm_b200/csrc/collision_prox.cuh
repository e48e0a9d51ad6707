// Collision::prox (hard/src/CollisionEnergyTerm.hpp:79-91) over the analytic passive objects of
// hard/src/PassiveObject.hpp:32-136, for one point; shared by the unit batch kernel (extra_terms.cu) and the
// collision terms of the hard_zxu loop (tri_kernels.cu). Include only from units compiled with -fmad=false.
#pragma once
#include <cfloat>

#include "extra_terms.cuh"

namespace aaadmm {

struct Payload {
    double dx, p[3];
};

__device__ __forceinline__ double norm3(const double *v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

// PassiveCollision::signed_distance of one object: keeps the smallest signed distance seen so far and
// the surface point that belongs to it (`if (dx > p.dx) return;`).
__device__ __forceinline__ void signed_distance(int type, const double *q, const double *x, Payload &pl) {
    const double *c = q;
    const double rad = q[6];
    if (type == PASSIVE_FLOOR) {
        const double dx = x[1] - q[0];
        if (dx > pl.dx) return;
        pl.dx = dx;
        pl.p[0] = x[0];
        pl.p[1] = q[0];
        pl.p[2] = x[2];
    } else if (type == PASSIVE_SLIDE_FLOOR) {
        double nrm[3] = {q[3], q[4], q[5]};
        const double nn = norm3(nrm);  // the constructor normalises
        if (nn > 0.0) {
            nrm[0] /= nn;
            nrm[1] /= nn;
            nrm[2] /= nn;
        }
        const double dx = (x[0] - c[0]) * nrm[0] + (x[1] - c[1]) * nrm[1] + (x[2] - c[2]) * nrm[2];
        if (dx > pl.dx) return;
        pl.dx = dx;
#pragma unroll
        for (int k = 0; k < 3; ++k) pl.p[k] = x[k] - dx * nrm[k];
    } else if (type == PASSIVE_SPHERE) {
        double dir[3] = {x[0] - c[0], x[1] - c[1], x[2] - c[2]};
        const double len = norm3(dir);
        const double dx = len - rad;
        if (dx > pl.dx) return;
        pl.dx = dx;
#pragma unroll
        for (int k = 0; k < 3; ++k) pl.p[k] = c[k] + (len > 0.0 ? dir[k] / len : dir[k]) * rad;
    } else if (type == PASSIVE_PLANE_HALF_SPHERE) {
        const double px = x[0] - c[0], pz = x[2] - c[2];
        const double dc = sqrt(px * px + 0.0 * 0.0 + pz * pz) - rad;
        if (dc > 0.0) {
            const double dx = x[1] - c[1];
            if (dx > pl.dx) return;
            pl.dx = dx;
            pl.p[0] = x[0];
            pl.p[1] = c[1];
            pl.p[2] = x[2];
        } else {
            double dir[3] = {x[0] - c[0], x[1] - c[1], x[2] - c[2]};
            const double len = norm3(dir);
            const double dx = (x[1] - c[1] > 0.0) ? len + rad : rad - len;
            if (dx > pl.dx) return;
            pl.dx = dx;
#pragma unroll
            for (int k = 0; k < 3; ++k) pl.p[k] = c[k] + (len > 0.0 ? dir[k] / len : dir[k]) * rad;
        }
    } else {  // PASSIVE_CYLINDER: axis along z through `center`
        double dir[3] = {x[0] - c[0], x[1] - c[1], 0.0 - c[2]};
        const double len = norm3(dir);
        const double dx = len - rad;
        if (dx > pl.dx) return;
        pl.dx = dx;
#pragma unroll
        for (int k = 0; k < 3; ++k) pl.p[k] = c[k] + (len > 0.0 ? dir[k] / len : dir[k]) * rad + (k == 2 ? x[2] : 0.0);
    }
}

// z <- surface point of the object with the smallest signed distance if that distance is negative
__device__ __forceinline__ void collision_prox_point(int n_objs, const int *__restrict__ types, const double *__restrict__ prm,
                                                     double (&z)[3]) {
    const double x[3] = {z[0], z[1], z[2]};
    Payload pl;
    pl.dx = DBL_MAX;
    pl.p[0] = pl.p[1] = pl.p[2] = 0.0;
    for (int j = 0; j < n_objs; ++j) signed_distance(types[j], prm + 7 * j, x, pl);
    if (pl.dx < 0.0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) z[k] = pl.p[k];
    }
}

}  // namespace aaadmm
