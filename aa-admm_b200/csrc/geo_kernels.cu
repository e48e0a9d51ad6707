#include "geo_kernels.cuh"

#include <cfloat>

namespace aaadmm {

namespace {

// ---- Constraint<3>::apply_transform (Constraint.h:73-94) ------------------------------------------
// Returns the number of transformed columns; out[j*3 + r].
__device__ __forceinline__ int geo_transform(int type, const int *__restrict__ ids, int k, const double *__restrict__ x,
                                             double *out) {
    if (type == GEO_PLANE) {  // MEAN_CENTERING
        double mean[3] = {0.0, 0.0, 0.0};
        for (int j = 0; j < k; ++j)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double v = x[3 * (size_t)ids[j] + r];
                out[j * 3 + r] = v;
                mean[r] += v;
            }
#pragma unroll
        for (int r = 0; r < 3; ++r) mean[r] /= (double)k;
        for (int j = 0; j < k; ++j)
#pragma unroll
            for (int r = 0; r < 3; ++r) out[j * 3 + r] -= mean[r];
        return k;
    }
    // SUBTRACT_FIRST
    double first[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) first[r] = x[3 * (size_t)ids[0] + r];
    for (int j = 1; j < k; ++j)
#pragma unroll
        for (int r = 0; r < 3; ++r) out[(j - 1) * 3 + r] = x[3 * (size_t)ids[j] + r] - first[r];
    return k - 1;
}

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return fmin(fmax(lo, v), hi); }

// PlaneConstraint::project_impl (Constraint.h:406-413): remove the component along the left singular
// vector of the smallest singular value of the 3 x k block. The reference takes it from Eigen's
// JacobiSVD; here a one-sided (Hestenes) Jacobi on the three rows gives the same vector (up to sign,
// which the projection does not see) without forming A A^T.
__device__ void project_plane(const double *v, int k, double *out) {
    double a[3][GEO_MAX_K], U[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int j = 0; j < k; ++j)
#pragma unroll
        for (int r = 0; r < 3; ++r) a[r][j] = v[j * 3 + r];
    for (int sweep = 0; sweep < 40; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
            double alpha = 0.0, beta = 0.0, gamma = 0.0;
            for (int j = 0; j < k; ++j) {
                alpha += a[p][j] * a[p][j];
                beta += a[q][j] * a[q][j];
                gamma += a[p][j] * a[q][j];
            }
            if (fabs(gamma) > 1e-16 * sqrt(alpha * beta) && fabs(gamma) > DBL_MIN) {
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int j = 0; j < k; ++j) {
                    const double tp = a[p][j], tq = a[q][j];
                    a[p][j] = c * tp - s * tq;
                    a[q][j] = s * tp + c * tq;
                }
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const double up = U[p * 3 + r], uq = U[q * 3 + r];
                    U[p * 3 + r] = c * up - s * uq;
                    U[q * 3 + r] = s * up + c * uq;
                }
            }
        }
        if (!rotated) break;
    }
    double nrm[3] = {0.0, 0.0, 0.0};
    for (int j = 0; j < k; ++j)
#pragma unroll
        for (int r = 0; r < 3; ++r) nrm[r] += a[r][j] * a[r][j];
    int m = 0;
    if (nrm[1] < nrm[m]) m = 1;
    if (nrm[2] < nrm[m]) m = 2;
    double n[3] = {U[m * 3 + 0], U[m * 3 + 1], U[m * 3 + 2]};
    const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    n[0] /= nn;
    n[1] /= nn;
    n[2] /= nn;
    for (int j = 0; j < k; ++j) {
        const double d = n[0] * v[j * 3 + 0] + n[1] * v[j * 3 + 1] + n[2] * v[j * 3 + 2];
#pragma unroll
        for (int r = 0; r < 3; ++r) out[j * 3 + r] = v[j * 3 + r] - n[r] * d;
    }
}

// EdgeLengthConstraint::project_impl (Constraint.h:211-214)
__device__ __forceinline__ void project_edge(const double *v, double len, double *out) {
    const double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
#pragma unroll
    for (int r = 0; r < 3; ++r) out[r] = (n > 0.0 ? v[r] / n : v[r]) * len;
}

__device__ __forceinline__ void normalized3(const double *v, double *o) {
    const double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
#pragma unroll
    for (int r = 0; r < 3; ++r) o[r] = n > 0.0 ? v[r] / n : v[r];
}

// AngleConstraint::project_impl (Constraint.h:243-291)
__device__ void project_angle(const double *v, const double *prm, double *out) {
    const double min_angle = prm[0], max_angle = prm[1], min_cos = prm[2], max_cos = prm[3];
#pragma unroll
    for (int k = 0; k < 6; ++k) out[k] = v[k];
    const double *v1 = v, *v2 = v + 3;
    const double eps = 1e-14;
    const double v1s = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2], v2s = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
    const double v1n = sqrt(v1s), v2n = sqrt(v2s);
    double u1[3], u2[3];
    normalized3(v1, u1);
    normalized3(v2, u2);
    const double cg = clampd(u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2], -1.0, 1.0);
    if ((1.0 - fabs(cg) > eps) && (cg > min_cos || cg < max_cos)) {
        const double gamma = acos(cg);
        double eta = cg > min_cos ? (min_angle - gamma) : (gamma - max_angle);
        eta = fmax(eta, 0.0);
        double theta = 0.5 * atan2(v2s * sin(2 * eta), v1s + v2s * cos(2 * eta));
        theta = fmax(0.0, fmin(eta, theta));
        const double phi = eta - theta;
        double t3[3], t4[3], u3[3], u4[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            t3[r] = u2[r] - u1[r] * cg;
            t4[r] = u1[r] - u2[r] * cg;
        }
        normalized3(t3, u3);
        normalized3(t4, u4);
        if (cg > min_cos) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                u3[r] *= -1.0;
                u4[r] *= -1.0;
            }
        }
        const double ct = cos(theta), st = sin(theta), cp = cos(phi), sp = sin(phi);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            out[r] = (u1[r] * ct + u3[r] * st) * (v1n * ct);
            out[3 + r] = (u2[r] * cp + u4[r] * sp) * (v2n * cp);
        }
    }
}

__device__ __forceinline__ void geo_project(int type, const double *v, int kc, const double *prm, double *out) {
    if (type == GEO_PLANE)
        project_plane(v, kc, out);
    else if (type == GEO_EDGE)
        project_edge(v, prm[0], out);
    else
        project_angle(v, prm, out);
}

__global__ void __launch_bounds__(GEO_BLOCK)
k_geo_local(GeoConstraints C, const double *__restrict__ x, const double *__restrict__ u, double *__restrict__ prev_dx,
            double *__restrict__ z, double *__restrict__ zmu, const SolveState *st) {
    if (st->done) return;
    const int c = blockIdx.x * GEO_BLOCK + threadIdx.x;
    if (c >= C.n) return;
    const int type = C.type[c], p0 = C.idx_ptr[c], k = C.idx_ptr[c + 1] - p0;
    double dx[GEO_MAX_K * 3], v[GEO_MAX_K * 3], zz[GEO_MAX_K * 3];
    const int kc = geo_transform(type, C.idx + p0, k, x, dx);
    const size_t o = 3 * (size_t)C.col0[c];
    for (int j = 0; j < kc * 3; ++j) {
        prev_dx[o + j] = dx[j];
        v[j] = dx[j] + u[o + j];
    }
    geo_project(type, v, kc, C.param + 4 * (size_t)c, zz);
    for (int j = 0; j < kc * 3; ++j) z[o + j] = zz[j];
    // z - u for the right-hand side gather of the same turn (k_geo_rhs then reads one array instead of two; the same
    // subtraction, the same bits)
    if (zmu)
        for (int j = 0; j < kc * 3; ++j) zmu[o + j] = zz[j] - u[o + j];
}

__global__ void __launch_bounds__(GEO_BLOCK)
k_geo_project_only(GeoConstraints C, const double *__restrict__ v, double *__restrict__ z) {
    const int c = blockIdx.x * GEO_BLOCK + threadIdx.x;
    if (c >= C.n) return;
    const int type = C.type[c], k = C.idx_ptr[c + 1] - C.idx_ptr[c];
    const int kc = type == GEO_PLANE ? k : k - 1;
    const size_t o = 3 * (size_t)C.col0[c];
    double vv[GEO_MAX_K * 3], zz[GEO_MAX_K * 3];
    for (int j = 0; j < kc * 3; ++j) vv[j] = v[o + j];
    geo_project(type, vv, kc, C.param + 4 * (size_t)c, zz);
    for (int j = 0; j < kc * 3; ++j) z[o + j] = zz[j];
}

// ---- closest point on a triangle (point_simplex_squared_distance.cpp:44-110, Ericson ch. 5) -------
__device__ __forceinline__ double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

__device__ double closest_on_triangle(const double *p, const double *t, double *c) {
    const double *a = t, *b = t + 3, *cc = t + 6;
    double ab[3], ac[3], ap[3], bp[3], cp[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        ab[r] = b[r] - a[r];
        ac[r] = cc[r] - a[r];
        ap[r] = p[r] - a[r];
    }
    const double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    bool done = false;
    if (d1 <= 0.0 && d2 <= 0.0) {
#pragma unroll
        for (int r = 0; r < 3; ++r) c[r] = a[r];
        done = true;
    }
    double d3 = 0, d4 = 0, d5 = 0, d6 = 0, vc = 0, vb = 0;
    if (!done) {
#pragma unroll
        for (int r = 0; r < 3; ++r) bp[r] = p[r] - b[r];
        d3 = dot3(ab, bp);
        d4 = dot3(ac, bp);
        if (d3 >= 0.0 && d4 <= d3) {
#pragma unroll
            for (int r = 0; r < 3; ++r) c[r] = b[r];
            done = true;
        }
    }
    if (!done) {
        vc = d1 * d4 - d3 * d2;
        const bool a_ne_b = !(a[0] == b[0] && a[1] == b[1] && a[2] == b[2]);
        if (a_ne_b && vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
            const double v = d1 / (d1 - d3);
#pragma unroll
            for (int r = 0; r < 3; ++r) c[r] = a[r] + v * ab[r];
            done = true;
        }
    }
    if (!done) {
#pragma unroll
        for (int r = 0; r < 3; ++r) cp[r] = p[r] - cc[r];
        d5 = dot3(ab, cp);
        d6 = dot3(ac, cp);
        if (d6 >= 0.0 && d5 <= d6) {
#pragma unroll
            for (int r = 0; r < 3; ++r) c[r] = cc[r];
            done = true;
        }
    }
    if (!done) {
        vb = d5 * d2 - d1 * d6;
        if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
            const double w = d2 / (d2 - d6);
#pragma unroll
            for (int r = 0; r < 3; ++r) c[r] = a[r] + w * ac[r];
            done = true;
        }
    }
    if (!done) {
        const double va = d3 * d6 - d5 * d4;
        if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
            const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
#pragma unroll
            for (int r = 0; r < 3; ++r) c[r] = b[r] + w * (cc[r] - b[r]);
        } else {
            const double denom = 1.0 / (va + vb + vc);
            const double v = vb * denom, w = vc * denom;
#pragma unroll
            for (int r = 0; r < 3; ++r) c[r] = a[r] + ab[r] * v + ac[r] * w;
        }
    }
    const double dx = p[0] - c[0], dy = p[1] - c[1], dz = p[2] - c[2];
    return dx * dx + dy * dy + dz * dz;
}

__device__ __forceinline__ double box_dist2(const BvhNode &n, const double *p) {
    double d = 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double e = fmax(fmax(n.lo[r] - p[r], p[r] - n.hi[r]), 0.0);
        d += e * e;
    }
    return d;
}

// exact nearest point by BVH descent with strict-< updates; bounded by the previous call's triangle
__device__ void bvh_closest(const GeoSoft &S, const double *p, int hint, double *best_c, int *best_tri) {
    double best = DBL_MAX;
    int bt = -1;
    if (hint >= 0) {
        best = closest_on_triangle(p, S.tri + 9 * (size_t)hint, best_c);
        bt = hint;
    }
    int stack[64], sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const int ni = stack[--sp];
        const BvhNode &nd = S.nodes[ni];
        if (box_dist2(nd, p) >= best) continue;
        if (nd.left < 0) {
            const int first = -(nd.left + 1);
            for (int k = 0; k < nd.right; ++k) {
                const int ti = S.tri_order[first + k];
                if (ti == hint) continue;
                double c[3];
                const double d = closest_on_triangle(p, S.tri + 9 * (size_t)ti, c);
                if (d < best) {
                    best = d;
                    bt = ti;
                    best_c[0] = c[0];
                    best_c[1] = c[1];
                    best_c[2] = c[2];
                }
            }
        } else {
            const double dl = box_dist2(S.nodes[nd.left], p), dr = box_dist2(S.nodes[nd.right], p);
            // push the farther child first so that the nearer one is visited first
            if (dl <= dr) {
                if (dr < best && sp < 63) stack[sp++] = nd.right;
                if (dl < best && sp < 63) stack[sp++] = nd.left;
            } else {
                if (dl < best && sp < 63) stack[sp++] = nd.left;
                if (dr < best && sp < 63) stack[sp++] = nd.right;
            }
        }
    }
    *best_tri = bt;
}

__global__ void __launch_bounds__(GEO_BLOCK)
k_geo_soft(GeoSoft S, const double *__restrict__ x, double *__restrict__ cp, const SolveState *st) {
    if (st && st->done) return;
    const int j = blockIdx.x * GEO_BLOCK + threadIdx.x;
    if (j >= S.n) return;
    // neighbouring lanes take neighbouring points (Morton order of the first positions): their descents visit the same
    // nodes (ncu before: 13 of 32 lanes active per instruction with the points in mesh order)
    const int i = S.order ? S.order[j] : j;
    const int pt = S.point ? S.point[i] : i;
    double p[3] = {x[3 * (size_t)pt], x[3 * (size_t)pt + 1], x[3 * (size_t)pt + 2]}, c[3];
    int tri;
    bvh_closest(S, p, S.last_tri ? S.last_tri[i] : -1, c, &tri);
    if (S.last_tri) S.last_tri[i] = tri;
    cp[3 * (size_t)i] = c[0];
    cp[3 * (size_t)i + 1] = c[1];
    cp[3 * (size_t)i + 2] = c[2];
}

__global__ void __launch_bounds__(GEO_BLOCK)
k_geo_closest_only(GeoSoft S, const double *__restrict__ q, double *__restrict__ cp, int *__restrict__ tri_out) {
    const int i = blockIdx.x * GEO_BLOCK + threadIdx.x;
    if (i >= S.n) return;
    double p[3] = {q[3 * (size_t)i], q[3 * (size_t)i + 1], q[3 * (size_t)i + 2]}, c[3];
    int tri;
    bvh_closest(S, p, -1, c, &tri);
    tri_out[i] = tri;
    cp[3 * (size_t)i] = c[0];
    cp[3 * (size_t)i + 1] = c[1];
    cp[3 * (size_t)i + 2] = c[2];
}

__global__ void __launch_bounds__(128)
k_geo_rhs(int n_points, const int64_t *__restrict__ dt_ptr, const int *__restrict__ dt_col,
          const double *__restrict__ dt_val, const double *__restrict__ z, const double *__restrict__ u,
          const double *__restrict__ rhs_fixed, const int *__restrict__ soft_of_point, double soft_weight,
          const double *__restrict__ cp, const int *__restrict__ iperm, double *__restrict__ W, const SolveState *st) {
    if (st->done) return;
    const int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= n_points) return;
    double s[3] = {rhs_fixed[3 * (size_t)p], rhs_fixed[3 * (size_t)p + 1], rhs_fixed[3 * (size_t)p + 2]};
    double a[3] = {0.0, 0.0, 0.0};
    const int64_t e1 = dt_ptr[p + 1];
    // the entries of a row are summed in their order; the loads of four entries (column, value, then z and u of that
    // column) are issued before the first use: the kernel is a latency-bound gather (ncu: 255 warps stalled on
    // long_scoreboard per issue, 3 % issue utilisation)
    for (int64_t e = dt_ptr[p]; e < e1; e += 4) {
        size_t o[4];
        double v[4], zz[4][3], uu[4][3];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool in = e + i < e1;
            o[i] = in ? 3 * (size_t)dt_col[e + i] : 0;
            v[i] = in ? dt_val[e + i] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                zz[i][r] = z[o[i] + r];
                uu[i][r] = u ? u[o[i] + r] : 0.0;  // u == null: `z` already holds z - u (k_geo_local)
            }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (e + i < e1) {
#pragma unroll
                for (int r = 0; r < 3; ++r) a[r] += v[i] * (u ? zz[i][r] - uu[i][r] : zz[i][r]);
            }
    }
    const int si = soft_of_point ? soft_of_point[p] : -1;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        s[r] += a[r];
        if (si >= 0) s[r] += soft_weight * cp[3 * (size_t)si + r];
    }
    const size_t o = 3 * (size_t)iperm[p];
    W[o] = s[0];
    W[o + 1] = s[1];
    W[o + 2] = s[2];
}

__global__ void __launch_bounds__(GEO_BLOCK)
k_geo_u_resid(GeoConstraints C, const double *__restrict__ x_new, const double *__restrict__ u,
              const double *__restrict__ z, const double *__restrict__ prev_dx, double *__restrict__ u_new,
              SolveState *st, double *partials, double *hist, int accel) {
    if (st->done) return;
    double acc[2] = {0.0, 0.0};
    for (int c = blockIdx.x * GEO_BLOCK + threadIdx.x; c < C.n; c += gridDim.x * GEO_BLOCK) {
        const int type = C.type[c], p0 = C.idx_ptr[c], k = C.idx_ptr[c + 1] - p0;
        double dx[GEO_MAX_K * 3];
        const int kc = geo_transform(type, C.idx + p0, k, x_new, dx);
        const size_t o = 3 * (size_t)C.col0[c];
        for (int j = 0; j < kc * 3; ++j) {
            const double r = dx[j] - z[o + j];
            const double d = dx[j] - prev_dx[o + j];
            u_new[o + j] = u[o + j] + r;
            acc[0] += r * r;
            acc[1] += d * d;
        }
    }
    double out[2];
    if (grid_reduce<2, GEO_BLOCK>(acc, partials, &st->ticket, out)) {
        if (threadIdx.x == 0) {
            const double r = out[0] + out[1];
            st->comb = r;
            const bool accept = (!accel) || st->reject || r < st->prev_prim;
            if (accept) {
                const int it = st->iter;
                hist[it] = r;
                hist[st->max_iters + it] = st->reject ? 1.0 : 0.0;  // this iteration follows a reset of the accelerator
                hist[2 * st->max_iters + it] = 1e-6 * (double)(global_timer_ns() - st->t0);  // ms since the loop began
                st->iter = it + 1;
                st->prev_prim = r;
                st->reject = 0;   // reset = false
                st->aa_skip = 0;
            } else {
                st->reject = 1;   // reset = true; aa->reset(current_u, current_x)
                st->aa_skip = 1;
                st->n_rejects += 1;
                st->aa_iter = 0;
                st->aa_col = 0;
            }
        }
    }
}

__global__ void k_geo_select(double *__restrict__ cur, double *__restrict__ def, const double *__restrict__ nw,
                             int64_t n, const SolveState *st, int accel) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (st->aa_skip) {  // rejected: current = default
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) cur[i] = def[i];
    } else if (!accel) {  // accepted, no acceleration: default = current = new
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const double v = nw[i];
            cur[i] = v;
            def[i] = v;
        }
    }
}

// ---- GeometrySolver<3> (older variant) -----------------------------------------------------------
// when: 0 always, 1 only on a rejected iterate
__global__ void __launch_bounds__(GEO_BLOCK)
k_gs_dx(GeoConstraints C, GeoSoft S, int zc_hard, const double *__restrict__ x, double *__restrict__ dx,
        const SolveState *st, int when) {
    if (st->done || (when == 1 && !st->reject)) return;
    const int c = blockIdx.x * GEO_BLOCK + threadIdx.x;
    if (c < C.n) {
        const int type = C.type[c], p0 = C.idx_ptr[c], k = C.idx_ptr[c + 1] - p0;
        double d[GEO_MAX_K * 3];
        const int kc = geo_transform(type, C.idx + p0, k, x, d);
        const size_t o = 3 * (size_t)C.col0[c];
        for (int j = 0; j < kc * 3; ++j) dx[o + j] = d[j];
    } else if (c < C.n + S.n) {
        const int i = c - C.n;
        const int pt = S.point ? S.point[i] : i;
        const size_t o = 3 * ((size_t)zc_hard + i);
#pragma unroll
        for (int r = 0; r < 3; ++r) dx[o + r] = x[3 * (size_t)pt + r];
    }
}

template <int MODE>
__global__ void __launch_bounds__(GEO_BLOCK)
k_gs_z(GeoConstraints C, GeoSoft S, int zc_hard, double rho, const double *__restrict__ dx,
       const double *__restrict__ u, double *__restrict__ z, SolveState *st, double *partials, double *hist) {
    if (st->done) return;
    if (MODE == 2 && !st->reject) return;
    double acc[1] = {0.0};
    const int total = C.n + S.n;
    for (int c = blockIdx.x * GEO_BLOCK + threadIdx.x; c < total; c += gridDim.x * GEO_BLOCK) {
        if (c < C.n) {
            const int type = C.type[c], k = C.idx_ptr[c + 1] - C.idx_ptr[c];
            const int kc = type == GEO_PLANE ? k : k - 1;
            const size_t o = 3 * (size_t)C.col0[c];
            double v[GEO_MAX_K * 3], zz[GEO_MAX_K * 3];
            for (int j = 0; j < kc * 3; ++j) v[j] = dx[o + j] + u[o + j];
            geo_project(type, v, kc, C.param + 4 * (size_t)c, zz);
            for (int j = 0; j < kc * 3; ++j) {
                z[o + j] = zz[j];
                const double r = dx[o + j] - zz[j];
                acc[0] += r * r;
            }
        } else {
            const int i = c - C.n;
            const size_t o = 3 * ((size_t)zc_hard + i);
            double v[3], cp[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) v[r] = dx[o + r] + u[o + r];
            int tri;
            bvh_closest(S, v, S.last_tri ? S.last_tri[i] : -1, cp, &tri);
            if (S.last_tri) S.last_tri[i] = tri;
            // Constraint::project_and_combine (Constraint.h:118-130): a = rho / (w + rho)
            const double a = rho / (S.weight + rho);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double zz = v[r] * a + cp[r] * (1 - a);
                z[o + r] = zz;
                const double d = dx[o + r] - zz;
                acc[0] += d * d;
            }
        }
    }
    if (MODE == 0) return;  // warm-up turn of ADMM_init_variables: no residual
    double out[1];
    if (grid_reduce<1, GEO_BLOCK>(acc, partials, &st->ticket, out)) {
        if (threadIdx.x == 0) {
            const double res = sqrt(out[0]);  // get_ADMM_residual: (Dx - z).norm()
            st->comb = res;
            bool log = true;
            if (MODE == 1) {
                st->reject = (st->accel && res > st->prev_prim) ? 1 : 0;  // need_reset
                if (st->reject) {
                    st->n_rejects += 1;
                    log = false;  // logged after the redo
                }
            }
            if (log) {
                const int it = st->iter;
                hist[it] = res;
                hist[st->max_iters + it] = MODE == 2 ? 1.0 : 0.0;  // logged after the redo of a rejected iterate
                hist[2 * st->max_iters + it] = 1e-6 * (double)(global_timer_ns() - st->t0);  // ms since the loop began
                st->iter = it + 1;
                st->prev_prim = res;
                if (st->iter >= st->max_iters) st->done = 1;  // end_iteration: the rest of this turn is skipped
            }
        }
    }
}

// when = 1: the reset of GeometrySolver.h:192-197. The reference swaps the two buffers; the old current
// (the rejected accelerated iterate) is dead afterwards - x_update / u_update overwrite all of default -
// so a one-way copy is equivalent. when = 0: the un-accelerated swap of :243-246, same argument.
__global__ void k_gs_take_default(double *__restrict__ cur, const double *__restrict__ def, int64_t n,
                                  const SolveState *st, int when) {
    if (st->done || (when == 1 && !st->reject)) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) cur[i] = def[i];
}

__global__ void k_gs_u(const double *__restrict__ u_cur, const double *__restrict__ dx, const double *__restrict__ z,
                       double *__restrict__ u_def, int64_t n, const SolveState *st) {
    if (st->done) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        u_def[i] = u_cur[i] + dx[i] - z[i];
}

}  // namespace

void launch_gs_dx(cudaStream_t s, const GeoConstraints &C, const GeoSoft &S, int zc_hard, const double *x, double *dx,
                  const SolveState *st, int when) {
    const int total = C.n + S.n;
    if (total > 0) k_gs_dx<<<(total + GEO_BLOCK - 1) / GEO_BLOCK, GEO_BLOCK, 0, s>>>(C, S, zc_hard, x, dx, st, when);
}
void launch_gs_z(cudaStream_t s, int mode, const GeoConstraints &C, const GeoSoft &S, int zc_hard, double rho,
                 const double *dx, const double *u, double *z, double *cp_scratch, SolveState *st, double *partials,
                 double *hist) {
    (void)cp_scratch;
    const int total = C.n + S.n;
    const int grid = max(1, min((total + GEO_BLOCK - 1) / GEO_BLOCK, stream_grid(8)));
    if (mode == 0)
        k_gs_z<0><<<grid, GEO_BLOCK, 0, s>>>(C, S, zc_hard, rho, dx, u, z, st, partials, hist);
    else if (mode == 1)
        k_gs_z<1><<<grid, GEO_BLOCK, 0, s>>>(C, S, zc_hard, rho, dx, u, z, st, partials, hist);
    else
        k_gs_z<2><<<grid, GEO_BLOCK, 0, s>>>(C, S, zc_hard, rho, dx, u, z, st, partials, hist);
}
void launch_gs_take_default(cudaStream_t s, double *cur, const double *def, int64_t n, const SolveState *st, int when) {
    k_gs_take_default<<<stream_grid(4), 256, 0, s>>>(cur, def, n, st, when);
}
void launch_gs_u(cudaStream_t s, const double *u_cur, const double *dx, const double *z, double *u_def, int64_t n,
                 const SolveState *st) {
    k_gs_u<<<stream_grid(4), 256, 0, s>>>(u_cur, dx, z, u_def, n, st);
}

void launch_geo_local(cudaStream_t s, const GeoConstraints &C, const double *x, const double *u, double *prev_dx,
                      double *z, double *zmu, const SolveState *st) {
    if (C.n > 0) k_geo_local<<<(C.n + GEO_BLOCK - 1) / GEO_BLOCK, GEO_BLOCK, 0, s>>>(C, x, u, prev_dx, z, zmu, st);
}
void launch_geo_soft(cudaStream_t s, const GeoSoft &S, const double *x, double *cp, const SolveState *st) {
    if (S.n > 0) k_geo_soft<<<(S.n + GEO_BLOCK - 1) / GEO_BLOCK, GEO_BLOCK, 0, s>>>(S, x, cp, st);
}
void launch_geo_rhs(cudaStream_t s, int n_points, const int64_t *dt_ptr, const int *dt_col, const double *dt_val,
                    const double *z, const double *u, const double *rhs_fixed, const int *soft_of_point,
                    double soft_weight, const double *cp, const int *iperm, double *W, const SolveState *st) {
    k_geo_rhs<<<(n_points + 127) / 128, 128, 0, s>>>(n_points, dt_ptr, dt_col, dt_val, z, u, rhs_fixed, soft_of_point,
                                                      soft_weight, cp, iperm, W, st);
}
void launch_geo_u_resid(cudaStream_t s, const GeoConstraints &C, const double *x_new, const double *u, const double *z,
                        const double *prev_dx, double *u_new, SolveState *st, double *partials, double *hist,
                        int accel) {
    const int grid = max(1, min((C.n + GEO_BLOCK - 1) / GEO_BLOCK, stream_grid(8)));
    k_geo_u_resid<<<grid, GEO_BLOCK, 0, s>>>(C, x_new, u, z, prev_dx, u_new, st, partials, hist, accel);
}
void launch_geo_select(cudaStream_t s, double *cur, double *def, const double *nw, int64_t n, const SolveState *st,
                       int accel) {
    k_geo_select<<<stream_grid(4), 256, 0, s>>>(cur, def, nw, n, st, accel);
}
void launch_geo_project_only(const GeoConstraints &C, const double *v, double *z) {
    if (C.n > 0) k_geo_project_only<<<(C.n + GEO_BLOCK - 1) / GEO_BLOCK, GEO_BLOCK>>>(C, v, z);
}
void launch_geo_closest_only(const GeoSoft &S, const double *q, double *cp, int *tri_out) {
    if (S.n > 0) k_geo_closest_only<<<(S.n + GEO_BLOCK - 1) / GEO_BLOCK, GEO_BLOCK>>>(S, q, cp, tri_out);
}

}  // namespace aaadmm
