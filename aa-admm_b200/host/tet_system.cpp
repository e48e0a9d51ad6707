#include "tet_system.hpp"

#include <algorithm>
#include <cmath>

namespace aaadmm {

namespace {
// 3x3 cofactor / inverse / determinant in the operation order of the fixed-size Eigen code
// the reference's TetEnergyTerm ctor runs (Eigen/src/LU/InverseImpl.h:120-170,
// Determinant.h bruteforce_det3_helper), m is column-major m[c*3+r].
inline double M(const double *m, int r, int c) { return m[c * 3 + r]; }
inline double cofactor(const double *m, int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return M(m, i1, j1) * M(m, i2, j2) - M(m, i1, j2) * M(m, i2, j1);
}
inline void inverse3(const double *m, double *inv) {
    const double c0 = cofactor(m, 0, 0), c1 = cofactor(m, 1, 0), c2 = cofactor(m, 2, 0);
    const double det = (c0 * M(m, 0, 0) + c1 * M(m, 1, 0)) + c2 * M(m, 2, 0);  // packet (a0,a1) reduced first, then the tail
    const double invdet = 1.0 / det;
    inv[0 * 3 + 0] = c0 * invdet;  // row 0
    inv[1 * 3 + 0] = c1 * invdet;
    inv[2 * 3 + 0] = c2 * invdet;
    inv[0 * 3 + 1] = cofactor(m, 0, 1) * invdet;  // (1,0)
    inv[1 * 3 + 1] = cofactor(m, 1, 1) * invdet;  // (1,1)
    inv[2 * 3 + 1] = cofactor(m, 2, 1) * invdet;  // (1,2)
    inv[0 * 3 + 2] = cofactor(m, 0, 2) * invdet;  // (2,0)
    inv[1 * 3 + 2] = cofactor(m, 1, 2) * invdet;  // (2,1)
    inv[2 * 3 + 2] = cofactor(m, 2, 2) * invdet;  // (2,2)
}
inline double det3_helper(const double *m, int a, int b, int c) {
    return M(m, 0, a) * (M(m, 1, b) * M(m, 2, c) - M(m, 1, c) * M(m, 2, b));
}
inline double det3(const double *m) {
    return det3_helper(m, 0, 1, 2) - det3_helper(m, 1, 0, 2) + det3_helper(m, 2, 0, 1);
}
}  // namespace

bool tri_constants(const double *rest9, double youngs, double poisson, double *rest_pose4, double *area, double *weight) {
    double e12[3], e13[3], n1[3], n2[3];
    for (int r = 0; r < 3; ++r) {
        e12[r] = rest9[3 + r] - rest9[r];
        e13[r] = rest9[6 + r] - rest9[r];
    }
    const double l1 = std::sqrt(e12[0] * e12[0] + e12[1] * e12[1] + e12[2] * e12[2]);
    for (int r = 0; r < 3; ++r) n1[r] = l1 > 0.0 ? e12[r] / l1 : e12[r];
    const double d = e13[0] * n1[0] + e13[1] * n1[1] + e13[2] * n1[2];
    for (int r = 0; r < 3; ++r) n2[r] = e13[r] - d * n1[r];
    const double l2 = std::sqrt(n2[0] * n2[0] + n2[1] * n2[1] + n2[2] * n2[2]);
    for (int r = 0; r < 3; ++r) n2[r] = l2 > 0.0 ? n2[r] / l2 : n2[r];
    // B = basis^T * edges (2x2), rest_pose = B^-1
    const double b00 = n1[0] * e12[0] + n1[1] * e12[1] + n1[2] * e12[2], b01 = n1[0] * e13[0] + n1[1] * e13[1] + n1[2] * e13[2];
    const double b10 = n2[0] * e12[0] + n2[1] * e12[1] + n2[2] * e12[2], b11 = n2[0] * e13[0] + n2[1] * e13[1] + n2[2] * e13[2];
    const double det = b00 * b11 - b01 * b10;
    const double inv = 1.0 / det;
    rest_pose4[0] = b11 * inv;   // (0,0)
    rest_pose4[1] = -b10 * inv;  // (1,0)
    rest_pose4[2] = -b01 * inv;  // (0,1)
    rest_pose4[3] = b00 * inv;   // (1,1)
    *area = 0.5 * det;
    if (*area < 0) return false;
    Lame lame(youngs, poisson);
    *weight = std::sqrt(lame.bulk_modulus() * (*area));
    return true;
}

bool tet_constants(const double *rest12, double youngs, double poisson, double *binv9, double *vol, double *weight) {
    double e[9];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) e[c * 3 + r] = rest12[3 * (c + 1) + r] - rest12[r];
    inverse3(e, binv9);
    *vol = det3(e) / 6.0;
    if (*vol < 0) return false;
    Lame lame(youngs, poisson);
    *weight = std::sqrt(lame.bulk_modulus() * (*vol));
    return true;
}

// Every contribution to the lower triangle of Ahat = M + rho dt^2 sum_t w_t^2 G_t^T G_t (+ triangle and collision
// terms) as (row, column, value), in ONE fixed order: masses, tets, triangles, collision terms.
// G_t(r,c) = sum_k Sel(c,k) Binv(k,r).
template <typename F>
static void for_each_contribution(const TetSystem &S, double rho_dt2, F emit) {
    const int nf = S.n_free, n_tets = S.n_tets, n_tris = S.n_tris, n_pts = S.n_pts;
    for (int v = 0; v < nf; ++v) emit(v, v, S.mass_free[v]);
    for (int t = 0; t < n_tets; ++t) {
        const double *bi = &S.binv[9 * (size_t)t];
        double G[3][4];
        for (int r = 0; r < 3; ++r) {
            // Binv column-major: Binv(k,r) = bi[r*3+k]
            G[r][1] = bi[r * 3 + 0];
            G[r][2] = bi[r * 3 + 1];
            G[r][3] = bi[r * 3 + 2];
            G[r][0] = -G[r][1] - G[r][2] - G[r][3];
        }
        const double w = S.weight[t];
        for (int a = 0; a < 4; ++a) {
            const int va = S.tet_dev[4 * (size_t)t + a];
            if (va >= nf) continue;
            for (int b = 0; b <= a; ++b) {
                const int vb = S.tet_dev[4 * (size_t)t + b];
                if (vb >= nf) continue;
                double s = 0;
                for (int r = 0; r < 3; ++r) s += (rho_dt2 * (w * G[r][a])) * (w * G[r][b]);
                emit(std::max(va, vb), std::min(va, vb), s);
            }
        }
    }
    // triangles: F(:,c) = sum_a Dc(a,c) x_a, Dc = S * rest_pose (hard/src/TriEnergyTerm.cpp:59-72)
    for (int t = 0; t < n_tris; ++t) {
        const double *rp = &S.tri_binv[4 * (size_t)t];  // R(k,c) = rp[c*2+k]
        double G[2][3];
        for (int c = 0; c < 2; ++c) {
            G[c][1] = rp[c * 2 + 0];
            G[c][2] = rp[c * 2 + 1];
            G[c][0] = -G[c][1] - G[c][2];
        }
        const double w = S.tri_weight[t];
        for (int a = 0; a < 3; ++a) {
            const int va = S.tri_dev[3 * (size_t)t + a];
            if (va >= nf) continue;
            for (int b = 0; b <= a; ++b) {
                const int vb = S.tri_dev[3 * (size_t)t + b];
                if (vb >= nf) continue;
                double s = 0;
                for (int c = 0; c < 2; ++c) s += (rho_dt2 * (w * G[c][a])) * (w * G[c][b]);
                emit(std::max(va, vb), std::min(va, vb), s);
            }
        }
    }
    for (int i = 0; i < n_pts; ++i) emit(S.pt_dev[i], S.pt_dev[i], rho_dt2 * (S.pt_weight[i] * S.pt_weight[i]));
}

bool build_tet_system(TetSystem &S, int n_verts, const double *rest12, int n_tets, const int *tets,
                      const int *material, const double *youngs, const double *poisson,
                      const double *masses, const std::vector<int> &pinned, double rho_dt2, const TriInput *tri,
                      const PointInput *pts) {
    S = TetSystem();
    S.n_verts = n_verts;
    S.n_tets = n_tets;
    std::vector<char> is_pin(n_verts, 0);
    for (int p : pinned) {
        if (p < 0 || p >= n_verts) {
            S.error = "pin index out of range";
            return false;
        }
        is_pin[p] = 1;
    }
    S.vert_to_dev.assign(n_verts, -1);
    S.dev_to_vert.assign(n_verts, -1);
    int nf = 0;
    for (int v = 0; v < n_verts; ++v)
        if (!is_pin[v]) {
            S.vert_to_dev[v] = nf;
            S.dev_to_vert[nf] = v;
            ++nf;
        }
    S.n_free = nf;
    int np = 0;
    for (int v = 0; v < n_verts; ++v)
        if (is_pin[v]) {
            S.vert_to_dev[v] = nf + np;
            S.dev_to_vert[nf + np] = v;
            ++np;
        }
    S.n_pin = np;
    S.mass_free.resize(nf);
    for (int k = 0; k < nf; ++k) S.mass_free[k] = masses[S.dev_to_vert[k]];

    S.tet_dev.resize((size_t)4 * n_tets);
    S.binv.resize((size_t)9 * n_tets);
    S.weight.resize(n_tets);
    S.volume.resize(n_tets);
    S.kvol.resize(n_tets);
    S.material.resize(n_tets);
    S.mu.resize(n_tets);
    S.lambda.resize(n_tets);
    for (int t = 0; t < n_tets; ++t) {
        const int *tv = tets + 4 * (size_t)t;
        double vol, w;
        if (!tet_constants(rest12 + 12 * (size_t)t, youngs[t], poisson[t], &S.binv[9 * (size_t)t], &vol, &w)) {
            S.error = "**TetEnergyTerm Error: Inverted initial tet";
            return false;
        }
        Lame lame(youngs[t], poisson[t]);
        const double k = lame.bulk_modulus();
        if (!(w > 0.0)) {
            S.error = "**EnergyTerm::get_reduction Error: Some weight leq 0";
            return false;
        }
        S.volume[t] = vol;
        S.weight[t] = w;
        S.kvol[t] = k * vol;
        S.material[t] = material ? material[t] : 0;
        S.mu[t] = lame.mu;
        S.lambda[t] = lame.lambda;
        for (int c = 0; c < 4; ++c) S.tet_dev[4 * (size_t)t + c] = S.vert_to_dev[tv[c]];
    }

    const int n_tris = tri ? tri->n_tris : 0;
    S.n_tris = n_tris;
    S.tri_dev.resize((size_t)3 * n_tris);
    S.tri_binv.resize((size_t)4 * n_tris);
    S.tri_weight.resize(n_tris);
    S.tri_area.resize(n_tris);
    S.tri_limit_min.resize(n_tris);
    S.tri_limit_max.resize(n_tris);
    for (int t = 0; t < n_tris; ++t) {
        const double lmin = tri->limit_min ? tri->limit_min[t] : -100.0, lmax = tri->limit_max ? tri->limit_max[t] : 100.0;
        if (lmin > 1.0) {
            S.error = "**TriEnergyTerm Error: Strain limit min should be -inf to 1";
            return false;
        }
        if (lmax < 1.0) {
            S.error = "**TriEnergyTerm Error: Strain limit max should be 1 to inf";
            return false;
        }
        if (!tri_constants(tri->rest9 + 9 * (size_t)t, tri->youngs[t], tri->poisson[t], &S.tri_binv[4 * (size_t)t],
                           &S.tri_area[t], &S.tri_weight[t])) {
            S.error = "**TriEnergyTerm Error: Inverted initial pose";
            return false;
        }
        if (!(S.tri_weight[t] > 0.0)) {
            S.error = "**EnergyTerm::get_reduction Error: Some weight leq 0";
            return false;
        }
        S.tri_limit_min[t] = lmin;
        S.tri_limit_max[t] = lmax;
        for (int c = 0; c < 3; ++c) {
            const int v = tri->tris[3 * (size_t)t + c];
            if (v < 0 || v >= n_verts) {
                S.error = "triangle vertex index out of range";
                return false;
            }
            S.tri_dev[3 * (size_t)t + c] = S.vert_to_dev[v];
        }
    }

    const int n_pts = pts ? pts->n : 0;
    S.n_pts = n_pts;
    S.pt_dev.resize(n_pts);
    S.pt_weight.resize(n_pts);
    for (int i = 0; i < n_pts; ++i) {
        const int v = pts->verts[i];
        if (v < 0 || v >= n_verts) {
            S.error = "collision vertex index out of range";
            return false;
        }
        if (is_pin[v]) {
            S.error = "collision term on a pinned vertex is not supported";
            return false;
        }
        if (!(pts->weight[i] > 0.0)) {
            S.error = "**EnergyTerm::get_reduction Error: Some weight leq 0";
            return false;
        }
        S.pt_dev[i] = S.vert_to_dev[v];
        S.pt_weight[i] = pts->weight[i];
    }

    // incidence lists over free vertices (tet slots first, then the triangle slots, then the collision terms)
    S.inc_ptr.assign(nf + 1, 0);
    for (size_t k = 0; k < S.tet_dev.size(); ++k)
        if (S.tet_dev[k] < nf) S.inc_ptr[S.tet_dev[k] + 1]++;
    for (size_t k = 0; k < S.tri_dev.size(); ++k)
        if (S.tri_dev[k] < nf) S.inc_ptr[S.tri_dev[k] + 1]++;
    for (int i = 0; i < n_pts; ++i) S.inc_ptr[S.pt_dev[i] + 1]++;
    for (int v = 0; v < nf; ++v) S.inc_ptr[v + 1] += S.inc_ptr[v];
    S.inc.resize(S.inc_ptr[nf]);
    {
        std::vector<int64_t> pos(S.inc_ptr.begin(), S.inc_ptr.end() - 1);
        for (size_t k = 0; k < S.tet_dev.size(); ++k)
            if (S.tet_dev[k] < nf) S.inc[pos[S.tet_dev[k]]++] = (int)k;
        const size_t slot0 = (size_t)4 * n_tets;
        for (size_t k = 0; k < S.tri_dev.size(); ++k)
            if (S.tri_dev[k] < nf) S.inc[pos[S.tri_dev[k]]++] = (int)(slot0 + k);
        const size_t slot1 = slot0 + (size_t)3 * n_tris;
        for (int i = 0; i < n_pts; ++i) S.inc[pos[S.pt_dev[i]]++] = (int)(slot1 + i);
    }

    // Ahat = M + rho dt^2 sum_t w_t^2 G_t^T G_t, G_t(r,c) = sum_k Sel(c,k) Binv(k,r)
    std::vector<int> tr, tc;
    std::vector<double> tv;
    tr.reserve((size_t)10 * n_tets + nf);
    tc.reserve((size_t)10 * n_tets + nf);
    tv.reserve((size_t)10 * n_tets + nf);
    for_each_contribution(S, rho_dt2, [&](int rr, int cc, double v) {
        tr.push_back(rr);
        tc.push_back(cc);
        tv.push_back(v);
    });
    S.Ahat = sym_from_triplets(nf, tr, tc, tv, false);
    // where every contribution lands in Ahat.x (same emission order): update_tet_system_materials refills the values of
    // another material on the same mesh without sorting anything
    S.contrib_dst.resize(tr.size());
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)tr.size(); ++k) {
        const int *b0 = S.Ahat.i.data() + S.Ahat.p[tc[k]], *b1 = S.Ahat.i.data() + S.Ahat.p[tc[k] + 1];
        S.contrib_dst[k] = (int64_t)(std::lower_bound(b0, b1, tr[k]) - S.Ahat.i.data());
    }
    return true;
}

bool update_tet_system_materials(TetSystem &S, const double *youngs, const double *poisson, double rho_dt2,
                                 const TriInput *tri) {
    for (int t = 0; t < S.n_tets; ++t) {
        Lame lame(youngs[t], poisson[t]);
        const double k = lame.bulk_modulus();
        const double w = std::sqrt(k * S.volume[t]);  // tet_constants
        if (!(w > 0.0)) {
            S.error = "**EnergyTerm::get_reduction Error: Some weight leq 0";
            return false;
        }
        S.weight[t] = w;
        S.kvol[t] = k * S.volume[t];
        S.mu[t] = lame.mu;
        S.lambda[t] = lame.lambda;
    }
    for (int t = 0; t < S.n_tris; ++t) {
        Lame lame(tri->youngs[t], tri->poisson[t]);
        const double w = std::sqrt(lame.bulk_modulus() * S.tri_area[t]);  // tri_constants
        if (!(w > 0.0)) {
            S.error = "**EnergyTerm::get_reduction Error: Some weight leq 0";
            return false;
        }
        S.tri_weight[t] = w;
        if (tri->limit_min) S.tri_limit_min[t] = tri->limit_min[t];
        if (tri->limit_max) S.tri_limit_max[t] = tri->limit_max[t];
    }
    // the same sums in the same order as sym_from_triplets forms them (duplicates are added in emission order)
    std::fill(S.Ahat.x.begin(), S.Ahat.x.end(), 0.0);
    int64_t k = 0;
    double *x = S.Ahat.x.data();
    const int64_t *dst = S.contrib_dst.data();
    for_each_contribution(S, rho_dt2, [&](int, int, double v) { x[dst[k++]] += v; });
    return k == (int64_t)S.contrib_dst.size();
}

}  // namespace aaadmm
