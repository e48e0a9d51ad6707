#include "extra_terms.cuh"

#include <cfloat>

namespace aaadmm {

namespace {

// Thin SVD of a 3x2 block by one one-sided (Hestenes) Jacobi rotation of its two columns: F J = [g1 g2]
// with g1 . g2 = 0, sigma_i = |g_i|, u_i = g_i / sigma_i, V = J. The reference takes U, sigma, V from
// Eigen's JacobiSVD with a full-pivoting QR preconditioner; U f(Sigma) V^T does not depend on how the
// factors were obtained as long as the singular values are distinct from zero.
struct Svd32 {
    double u1[3], u2[3], s1, s2, v11, v12, v21, v22;  // V = [v11 v12; v21 v22], columns v1, v2
};

__device__ Svd32 svd32(const double *F) {
    const double *f1 = F, *f2 = F + 3;
    const double a = f1[0] * f1[0] + f1[1] * f1[1] + f1[2] * f1[2];
    const double c = f2[0] * f2[0] + f2[1] * f2[1] + f2[2] * f2[2];
    const double b = f1[0] * f2[0] + f1[1] * f2[1] + f1[2] * f2[2];
    double cs = 1.0, sn = 0.0;
    if (fabs(b) > DBL_MIN) {
        const double zeta = (c - a) / (2.0 * b);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        cs = 1.0 / sqrt(1.0 + t * t);
        sn = cs * t;
    }
    Svd32 r;
    double g1[3], g2[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        g1[k] = cs * f1[k] - sn * f2[k];
        g2[k] = sn * f1[k] + cs * f2[k];
    }
    r.v11 = cs;
    r.v21 = -sn;
    r.v12 = sn;
    r.v22 = cs;
    r.s1 = sqrt(g1[0] * g1[0] + g1[1] * g1[1] + g1[2] * g1[2]);
    r.s2 = sqrt(g2[0] * g2[0] + g2[1] * g2[1] + g2[2] * g2[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        r.u1[k] = r.s1 > 0.0 ? g1[k] / r.s1 : (k == 0 ? 1.0 : 0.0);
        r.u2[k] = r.s2 > 0.0 ? g2[k] / r.s2 : 0.0;
    }
    if (!(r.s2 > 0.0)) {  // rank deficient: any unit vector orthogonal to u1
        const int m = fabs(r.u1[0]) <= fabs(r.u1[1]) ? (fabs(r.u1[0]) <= fabs(r.u1[2]) ? 0 : 2) : (fabs(r.u1[1]) <= fabs(r.u1[2]) ? 1 : 2);
        double e[3] = {0.0, 0.0, 0.0};
        e[m] = 1.0;
        const double d = r.u1[m];
        double n2 = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            r.u2[k] = e[k] - d * r.u1[k];
            n2 += r.u2[k] * r.u2[k];
        }
        n2 = sqrt(n2);
#pragma unroll
        for (int k = 0; k < 3; ++k) r.u2[k] /= n2;
    }
    return r;
}

__global__ void __launch_bounds__(128)
k_tri_prox(int variant, double *__restrict__ z, int64_t n, double lmin, double lmax) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    double F[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) F[k] = z[6 * i + k];
    const Svd32 s = svd32(F);
    const bool check = lmin > 0.0 || lmax < 99.0;
    double out[6];
    if (variant == 0) {
        // p = U [I;0] V^T, z = (p + z) / 2, then the column norms are clamped into [lmin, lmax]
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            out[k] = 0.5 * ((s.u1[k] * s.v11 + s.u2[k] * s.v12) + F[k]);
            out[3 + k] = 0.5 * ((s.u1[k] * s.v21 + s.u2[k] * s.v22) + F[3 + k]);
        }
        if (check) {
            const double l0 = sqrt(out[0] * out[0] + out[1] * out[1] + out[2] * out[2]);
            const double l1 = sqrt(out[3] * out[3] + out[4] * out[4] + out[5] * out[5]);
            // the reference applies the four tests one after the other with the ORIGINAL norms
            double f0 = 1.0, f1 = 1.0;
            if (l0 < lmin) f0 *= lmin / l0;
            if (l1 < lmin) f1 *= lmin / l1;
            if (l0 > lmax) f0 *= lmax / l0;
            if (l1 > lmax) f1 *= lmax / l1;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                out[k] *= f0;
                out[3 + k] *= f1;
            }
        }
    } else {
        // Sigma' = ((1, 1) + sigma) / 2, clamped into [lmin, lmax]; z = U Sigma' V^T
        double a = (1.0 + s.s1) / 2.0, b = (1.0 + s.s2) / 2.0;
        if (check) {
            const double l0 = a, l1 = b;
            if (l0 < lmin) a = lmin;
            if (l1 < lmin) b = lmin;
            if (l0 > lmax) a = lmax;
            if (l1 > lmax) b = lmax;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            out[k] = a * s.u1[k] * s.v11 + b * s.u2[k] * s.v12;
            out[3 + k] = a * s.u1[k] * s.v21 + b * s.u2[k] * s.v22;
        }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) z[6 * i + k] = out[k];
}

struct Payload {
    double dx, p[3];
};

__device__ __forceinline__ double norm3(const double *v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

// PassiveCollision::signed_distance of one object: keeps the smallest signed distance seen so far and
// the surface point that belongs to it (`if (dx > p.dx) return;`).
__device__ void signed_distance(int type, const double *q, const double *x, Payload &pl) {
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

__global__ void __launch_bounds__(128)
k_collision_prox(int n_objs, const int *__restrict__ types, const double *__restrict__ prm, double *__restrict__ z, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    const double x[3] = {z[3 * i], z[3 * i + 1], z[3 * i + 2]};
    Payload pl;
    pl.dx = DBL_MAX;
    pl.p[0] = pl.p[1] = pl.p[2] = 0.0;
    for (int j = 0; j < n_objs; ++j) signed_distance(types[j], prm + 7 * j, x, pl);
    if (pl.dx < 0.0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) z[3 * i + k] = pl.p[k];
    }
}

__global__ void k_spring_prox(double *__restrict__ z, const double *__restrict__ pins, const int *__restrict__ active, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !active[i]) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) z[3 * i + k] = pins[3 * i + k];
}

}  // namespace

void launch_tri_prox(int variant, double *z, int64_t n, double limit_min, double limit_max, cudaStream_t s) {
    if (n > 0) k_tri_prox<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(variant, z, n, limit_min, limit_max);
}
void launch_collision_prox(int n_objs, const int *types, const double *prm, double *z, int64_t n, cudaStream_t s) {
    if (n > 0) k_collision_prox<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(n_objs, types, prm, z, n);
}
void launch_spring_prox(double *z, const double *pins, const int *active, int64_t n, cudaStream_t s) {
    if (n > 0) k_spring_prox<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(z, pins, active, n);
}

}  // namespace aaadmm
