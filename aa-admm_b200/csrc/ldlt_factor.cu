// Numeric phase of the sparse LDL^T factorisation on the device (multifrontal, FP64, no pivoting).
//
// Role in the reference: the numeric part of Eigen::SimplicialLDLT::compute behind LDLTSolver::update_system
// (admm_anderson_xzu/src/LinearSolver.hpp:79-84) - computed once per system matrix at setup. Here the symbolic
// analysis (ordering, pattern of L, fronts) stays on the host and is reusable for every matrix with the same
// pattern (a material sweep over one mesh); the values are factored on the GPU straight into the
// per-front dense layout [T ; P] the triangular-solve setup consumes (ldlt_apply.cu), so no factor value ever
// exists on the host.
//
// Fronts are processed by their height in the front tree (children before parents), all fronts of a height in the
// same launches:
//   1. extend-add: the children's Schur complements are added into the parent's frontal matrix, one launch per
//      child rank (the r-th child of every parent), so that the order of the additions into an entry is fixed;
//   2. right-looking blocked partial factorisation, panels of PB = 32 columns: diagonal block (one warp per front),
//      the rows below it (one thread per row), trailing update F(i,j) -= sum_c L(i,c) d(c) L(j,c) in 64 x 64 tiles
//      (4 x 4 outputs per thread), which lands in the front's own columns or in its Schur complement (k x k).
// Everything is gather/ordered: no float atomics, results are bit-reproducible.
#include "ldlt_factor.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>

namespace aaadmm {

namespace {

constexpr int PB = 32;    // panel width
constexpr int UT = 64;    // trailing-update tile
constexpr int P2R = 128;  // rows per CTA in the panel solve
constexpr int EAC = 16;   // columns of a child's Schur complement per extend-add CTA

struct FrontNum {   // what the numeric kernels need to know about a front
    int64_t m_off;  // [T ; P] in dA (column-major, ld)
    int64_t s_off;  // Schur complement (k x k, column-major, ld = k) in S
    int64_t c_off;  // the front's k rows below as local row indices of the PARENT front (cmap)
    int first, ns, k, ld;
    int parent, pad;
};

// blockIdx.x -> (entry e of the launch's front list, local task t) through the prefix sums of the task counts
__device__ __forceinline__ void find_task(const int *__restrict__ prefix, int count, int &e, int &t) {
    int lo = 0, hi = count;  // prefix[lo] <= blockIdx.x < prefix[hi]
    const int b = blockIdx.x;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= b) lo = mid; else hi = mid;
    }
    e = lo;
    t = b - __ldg(prefix + lo);
}

// A (lower CSC, original numbering) -> its position inside the front matrices
__global__ void k_scatter_matrix(const double *__restrict__ Ax, const int64_t *__restrict__ dst, int64_t nnz,
                                 double *__restrict__ dA) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nnz) dA[dst[p]] = Ax[p];
}

// Extend-add of one child per parent: S_child(a, b), a >= b, is added to the parent's entry (cmap[a], cmap[b]).
// list[e] = child front; one CTA = EAC columns of its Schur complement.
__global__ void __launch_bounds__(256)
k_extend_add(const FrontNum *__restrict__ fronts, const int *__restrict__ list, const int *__restrict__ prefix, int count,
             const int *__restrict__ cmap, double *__restrict__ dA, double *__restrict__ S) {
    int e, t;
    find_task(prefix, count, e, t);
    const FrontNum C = fronts[list[e]];
    const FrontNum P = fronts[C.parent];
    const int *map = cmap + C.c_off;
    const double *Sc = S + C.s_off;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int bb = warp; bb < EAC; bb += 8) {
        const int b = t * EAC + bb;
        if (b >= C.k) break;
        const int q = __ldg(map + b);  // parent-local column
        double *dcol = q < P.ns ? dA + P.m_off + (size_t)q * P.ld : S + P.s_off + (size_t)(q - P.ns) * P.k - P.ns;
        const double *scol = Sc + (size_t)b * C.k;
        for (int a = b + lane; a < C.k; a += 32) dcol[__ldg(map + a)] += scol[a];
    }
}

// Diagonal block of panel j0 of every listed front: unblocked LDL^T of the nb x nb block in shared memory, one warp
// per front (lane = row). Writes the unit-lower block (explicit ones on the diagonal, as the solve setup expects)
// and the pivots; a zero / non-finite pivot raises *bad.
__global__ void __launch_bounds__(32)
k_panel_diag(const FrontNum *__restrict__ fronts, const int *__restrict__ list, int j0, double *__restrict__ dA,
             double *__restrict__ D, int *bad) {
    __shared__ double a[PB][PB + 1];
    __shared__ double wv[PB];
    const FrontNum F = fronts[list[blockIdx.x]];
    const int nb = min(PB, F.ns - j0);
    const int r = threadIdx.x;
    double *blk = dA + F.m_off + (size_t)j0 * F.ld + j0;
    for (int c = 0; c < nb; ++c) a[r][c] = (r < nb && r >= c) ? blk[(size_t)c * F.ld + r] : 0.0;
    __syncwarp();
    for (int c = 0; c < nb; ++c) {
        const double d = a[c][c];
        if (r == 0 && (d == 0.0 || !isfinite(d))) *bad = 1;
        double l = 0.0;
        if (r > c && r < nb) {
            const double w = a[r][c];
            wv[r] = w;
            l = w / d;
            a[r][c] = l;
        }
        __syncwarp();
        if (r > c && r < nb)
            for (int c2 = c + 1; c2 <= r; ++c2) a[r][c2] -= l * wv[c2];
        __syncwarp();
    }
    if (r < nb) {
        D[F.first + j0 + r] = a[r][r];
        for (int c = 0; c < nb; ++c)
            if (r >= c) blk[(size_t)c * F.ld + r] = r == c ? 1.0 : a[r][c];
    }
}

// Rows below the diagonal block of panel j0: L(i, :) = (F(i, :) L11^-T) D^-1, one thread per row.
__global__ void __launch_bounds__(P2R)
k_panel_rows(const FrontNum *__restrict__ fronts, const int *__restrict__ list, const int *__restrict__ prefix, int count,
             int j0, double *__restrict__ dA, const double *__restrict__ D) {
    __shared__ double L11[PB][PB + 1];
    __shared__ double dinv[PB];
    int e, t;
    find_task(prefix, count, e, t);
    const FrontNum F = fronts[list[e]];
    const int nb = min(PB, F.ns - j0), m = F.ns + F.k;
    const double *blk = dA + F.m_off + (size_t)j0 * F.ld + j0;
    for (int idx = threadIdx.x; idx < PB * PB; idx += P2R) {
        const int r = idx % PB, c = idx / PB;
        L11[r][c] = (r < nb && c < nb && r > c) ? blk[(size_t)c * F.ld + r] : 0.0;
    }
    if (threadIdx.x < PB) dinv[threadIdx.x] = threadIdx.x < nb ? 1.0 / D[F.first + j0 + threadIdx.x] : 1.0;
    __syncthreads();
    const int i = j0 + nb + t * P2R + (int)threadIdx.x;
    if (i >= m) return;
    double *row = dA + F.m_off + (size_t)j0 * F.ld + i;
    double y[PB];
#pragma unroll
    for (int c = 0; c < PB; ++c) y[c] = c < nb ? row[(size_t)c * F.ld] : 0.0;
#pragma unroll
    for (int c = 0; c < PB; ++c) {
        double s = y[c];
#pragma unroll
        for (int q = 0; q < c; ++q) s -= y[q] * L11[c][q];
        y[c] = s;
    }
#pragma unroll
    for (int c = 0; c < PB; ++c)
        if (c < nb) row[(size_t)c * F.ld] = y[c] * dinv[c];
}

// Trailing update after panel j0: F(i, j) -= sum_c L(i, c) d(c) L(j, c) for j0 + nb <= j <= i < ns + k. One CTA = one
// UT x UT tile of the block-lower triangle; columns j < ns live in the front matrix, the others in the Schur complement.
__global__ void __launch_bounds__(256)
k_trailing_update(const FrontNum *__restrict__ fronts, const int *__restrict__ list, const int *__restrict__ prefix, int count,
                  int j0, double *__restrict__ dA, double *__restrict__ S, const double *__restrict__ D) {
    __shared__ double Li[PB][UT + 1], Lj[PB][UT + 1];
    int e, t;
    find_task(prefix, count, e, t);
    const FrontNum F = fronts[list[e]];
    const int nb = min(PB, F.ns - j0), m = F.ns + F.k;
    const int base = j0 + nb;
    // t -> (I, J), J <= I, row-major enumeration of the block-lower triangle: t = I (I + 1) / 2 + J
    int I = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((I + 1) * (I + 2) / 2 <= t) ++I;
    while (I * (I + 1) / 2 > t) --I;
    const int J = t - I * (I + 1) / 2;
    const int i0 = base + I * UT, jj0 = base + J * UT;
    const double *pan = dA + F.m_off + (size_t)j0 * F.ld;
    for (int idx = threadIdx.x; idx < PB * UT; idx += 256) {
        const int c = idx / UT, r = idx % UT;
        const bool in = c < nb;
        Li[c][r] = (in && i0 + r < m) ? pan[(size_t)c * F.ld + i0 + r] : 0.0;
        Lj[c][r] = (in && jj0 + r < m) ? pan[(size_t)c * F.ld + jj0 + r] * D[F.first + j0 + c] : 0.0;
    }
    __syncthreads();
    const int tr = threadIdx.x & 15, tj = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 8
    for (int c = 0; c < PB; ++c) {
        double p[4], q[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) p[a] = Li[c][tr + 16 * a];
#pragma unroll
        for (int b = 0; b < 4; ++b) q[b] = Lj[c][tj + 16 * b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] += p[a] * q[b];
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int j = jj0 + tj + 16 * b;
        if (j >= m) continue;
        double *col = j < F.ns ? dA + F.m_off + (size_t)j * F.ld : S + F.s_off + (size_t)(j - F.ns) * F.k - F.ns;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = i0 + tr + 16 * a;
            if (i < m && i >= j) col[i] -= acc[a][b];
        }
    }
}

__global__ void k_invert_pivots(const double *__restrict__ D, double *__restrict__ dinv, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dinv[i] = 1.0 / D[i];
}

template <typename T>
int up(T **dst, const std::vector<T> &src) {
    const size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    AAADMM_CUDA_OK(cudaMalloc((void **)dst, bytes));
    if (!src.empty()) AAADMM_CUDA_OK(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

}  // namespace

struct PanelStep {
    int cnt;             // fronts of the height with ns > j0 (a prefix of the height's list: sorted by ns, descending)
    int rows_off, rows_total;  // prefix sums (cnt + 1 ints at rows_off) / CTAs of the panel-rows launch
    int upd_off, upd_total;    // the same for the trailing update
};
struct HeightPlan {
    int list_off, n_fronts;
    std::vector<PanelStep> steps;
    struct Rank { int list_off, prefix_off, cnt, total; };
    std::vector<Rank> ranks;  // extend-add launches: the r-th child of every front of this height
};

struct FactorPlan {
    int n = 0, nb = 0;
    int64_t nnzA = 0, s_tot = 0, m_tot = 0;
    FrontNum *fronts = nullptr;
    int64_t *a_dst = nullptr;
    int *cmap = nullptr, *lists = nullptr, *prefix = nullptr, *bad = nullptr;
    double *S = nullptr, *Ax = nullptr;
    int *bad_h = nullptr;  // pinned
    std::vector<HeightPlan> heights;
    int n_launches = 0;
};

void factor_plan_destroy(FactorPlan *p) {
    if (!p) return;
    cudaFree(p->fronts);
    cudaFree(p->a_dst);
    cudaFree(p->cmap);
    cudaFree(p->lists);
    cudaFree(p->prefix);
    cudaFree(p->bad);
    cudaFree(p->S);
    cudaFree(p->Ax);
    if (p->bad_h) cudaFreeHost(p->bad_h);
    delete p;
}

int factor_plan_build(FactorPlan **out, int n, const std::vector<FrontDesc> &fr, int nb, const std::vector<int> &rows,
                      const std::vector<int> &parent, const std::vector<int> &level, const std::vector<int> &blk_of,
                      const int *perm, const int64_t *Ap, const int *Ai, int64_t m_tot) {
    FactorPlan *p = new FactorPlan();
    p->n = n;
    p->nb = nb;
    p->m_tot = m_tot;
    std::vector<int> iperm(std::max(n, 1));
    for (int k = 0; k < n; ++k) iperm[perm[k]] = k;
    auto local_row = [&](const FrontDesc &F, int i) -> int {  // row of global (elimination-order) index i in front F, -1 if absent
        if (i < F.first + F.ns) return i - F.first;
        const int *R = rows.data() + F.r_off;
        const int *it = std::lower_bound(R, R + F.k, i);
        return (it != R + F.k && *it == i) ? F.ns + (int)(it - R) : -1;
    };
    // ---- matrix entries -> front positions ----
    const int64_t nnz = Ap[n];
    p->nnzA = nnz;
    std::vector<int64_t> dst((size_t)std::max<int64_t>(nnz, 1));
    int bad_pattern = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad_pattern)
    for (int c = 0; c < n; ++c)
        for (int64_t q = Ap[c]; q < Ap[c + 1]; ++q) {
            int i = iperm[Ai[q]], j = iperm[c];
            if (i < j) std::swap(i, j);
            const FrontDesc &F = fr[blk_of[j]];
            const int r = local_row(F, i);
            if (r < 0) {
                ++bad_pattern;
                dst[q] = 0;
            } else {
                dst[q] = F.m_off + (int64_t)(j - F.first) * F.ld + r;
            }
        }
    if (bad_pattern) {
        set_last_error("ldlt factor: the matrix has entries outside the pattern of L");
        delete p;
        return -1;
    }
    // ---- per-front numeric descriptors, Schur complement storage, child -> parent row maps ----
    std::vector<FrontNum> fn(std::max(nb, 1));
    std::vector<int> cmap((size_t)std::max<size_t>(rows.size(), 1), 0);
    int64_t s_tot = 0;
    for (int b = 0; b < nb; ++b) {
        const FrontDesc &F = fr[b];
        FrontNum &N = fn[b];
        N.m_off = F.m_off;
        N.s_off = s_tot;
        N.c_off = F.r_off;
        N.first = F.first;
        N.ns = F.ns;
        N.k = F.k;
        N.ld = F.ld;
        N.parent = parent[b];
        N.pad = 0;
        if (parent[b] >= 0) s_tot += (int64_t)F.k * F.k;  // a root's Schur complement is never needed (k = 0 anyway)
    }
    p->s_tot = s_tot;
    int bad_nest = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : bad_nest)
    for (int b = 0; b < nb; ++b) {
        if (parent[b] < 0) continue;
        const FrontDesc &C = fr[b], &P = fr[parent[b]];
        for (int q = 0; q < C.k; ++q) {
            const int r = local_row(P, rows[C.r_off + q]);
            if (r < 0) ++bad_nest;
            cmap[C.r_off + q] = std::max(r, 0);
        }
    }
    if (bad_nest) {
        set_last_error("ldlt factor: front patterns do not nest along the front tree");
        delete p;
        return -1;
    }
    // ---- launch lists ----
    int nlev = 0;
    for (int b = 0; b < nb; ++b) nlev = std::max(nlev, level[b] + 1);
    std::vector<std::vector<int>> by_height(nlev), children(nb);
    for (int b = 0; b < nb; ++b) {
        by_height[level[b]].push_back(b);
        if (parent[b] >= 0) children[parent[b]].push_back(b);  // ascending child order
    }
    std::vector<int> lists, prefix;
    p->heights.resize(nlev);
    for (int l = 0; l < nlev; ++l) {
        std::vector<int> &v = by_height[l];
        std::stable_sort(v.begin(), v.end(), [&](int a, int b) { return fr[a].ns > fr[b].ns; });
        HeightPlan &H = p->heights[l];
        H.list_off = (int)lists.size();
        H.n_fronts = (int)v.size();
        lists.insert(lists.end(), v.begin(), v.end());
        // extend-add: rank r = the r-th child of every front of this height that has one
        size_t maxch = 0;
        for (int b : v) maxch = std::max(maxch, children[b].size());
        for (size_t r = 0; r < maxch; ++r) {
            HeightPlan::Rank R;
            R.list_off = (int)lists.size();
            R.prefix_off = (int)prefix.size();
            R.cnt = 0;
            int tot = 0;
            for (int b : v)
                if (children[b].size() > r) {
                    const int c = children[b][r];
                    lists.push_back(c);
                    prefix.push_back(tot);
                    tot += (fr[c].k + EAC - 1) / EAC;
                    ++R.cnt;
                }
            prefix.push_back(tot);
            R.total = tot;
            if (tot > 0) H.ranks.push_back(R);
        }
        const int max_ns = v.empty() ? 0 : fr[v[0]].ns;
        for (int j0 = 0; j0 < max_ns; j0 += PB) {
            PanelStep st;
            st.cnt = 0;
            while (st.cnt < (int)v.size() && fr[v[st.cnt]].ns > j0) ++st.cnt;
            st.rows_off = (int)prefix.size();
            int tot = 0;
            for (int e = 0; e < st.cnt; ++e) {
                const FrontDesc &F = fr[v[e]];
                const int nbp = std::min(PB, F.ns - j0), below = F.ns + F.k - (j0 + nbp);
                prefix.push_back(tot);
                tot += (below + P2R - 1) / P2R;
            }
            prefix.push_back(tot);
            st.rows_total = tot;
            st.upd_off = (int)prefix.size();
            tot = 0;
            for (int e = 0; e < st.cnt; ++e) {
                const FrontDesc &F = fr[v[e]];
                const int nbp = std::min(PB, F.ns - j0), below = F.ns + F.k - (j0 + nbp);
                const int nt = (below + UT - 1) / UT;
                prefix.push_back(tot);
                tot += nt * (nt + 1) / 2;
            }
            prefix.push_back(tot);
            st.upd_total = tot;
            H.steps.push_back(st);
        }
    }
    int rc = 0;
    rc |= up(&p->fronts, fn);
    rc |= up(&p->a_dst, dst);
    rc |= up(&p->cmap, cmap);
    rc |= up(&p->lists, lists);
    rc |= up(&p->prefix, prefix);
    if (rc || cudaMalloc((void **)&p->S, sizeof(double) * (size_t)std::max<int64_t>(s_tot, 1)) != cudaSuccess ||
        cudaMalloc((void **)&p->Ax, sizeof(double) * (size_t)std::max<int64_t>(nnz, 1)) != cudaSuccess ||
        cudaMalloc((void **)&p->bad, sizeof(int)) != cudaSuccess || cudaMallocHost((void **)&p->bad_h, sizeof(int)) != cudaSuccess) {
        set_last_error("ldlt factor: cudaMalloc failed");
        factor_plan_destroy(p);
        return -1;
    }
    *out = p;
    return 0;
}

int64_t factor_plan_nnz(const FactorPlan *p) { return p->nnzA; }
double *factor_plan_values(FactorPlan *p) { return p->Ax; }
int factor_plan_launches(const FactorPlan *p) { return p->n_launches; }

// d_Ax = factor_plan_values(p) must hold the matrix values (lower CSC, original numbering). dA: front matrices
// (m_tot doubles), D / dinv: n pivots and their reciprocals. Everything is enqueued on `s`; factor_plan_check waits.
int factor_plan_run(FactorPlan *p, double *dA, double *D, double *dinv, cudaStream_t s) {
    AAADMM_CUDA_OK(cudaMemsetAsync(dA, 0, sizeof(double) * (size_t)std::max<int64_t>(p->m_tot, 1), s));
    if (p->s_tot > 0) AAADMM_CUDA_OK(cudaMemsetAsync(p->S, 0, sizeof(double) * (size_t)p->s_tot, s));
    AAADMM_CUDA_OK(cudaMemsetAsync(p->bad, 0, sizeof(int), s));
    int L = 0;
    if (p->nnzA > 0) {
        k_scatter_matrix<<<(unsigned)((p->nnzA + 255) / 256), 256, 0, s>>>(p->Ax, p->a_dst, p->nnzA, dA);
        ++L;
    }
    for (const HeightPlan &H : p->heights) {
        for (const HeightPlan::Rank &R : H.ranks) {
            k_extend_add<<<R.total, 256, 0, s>>>(p->fronts, p->lists + R.list_off, p->prefix + R.prefix_off, R.cnt, p->cmap, dA, p->S);
            ++L;
        }
        const int *list = p->lists + H.list_off;
        int j0 = 0;
        for (const PanelStep &st : H.steps) {
            k_panel_diag<<<st.cnt, 32, 0, s>>>(p->fronts, list, j0, dA, D, p->bad);
            ++L;
            if (st.rows_total > 0) {
                k_panel_rows<<<st.rows_total, P2R, 0, s>>>(p->fronts, list, p->prefix + st.rows_off, st.cnt, j0, dA, D);
                ++L;
            }
            if (st.upd_total > 0) {
                k_trailing_update<<<st.upd_total, 256, 0, s>>>(p->fronts, list, p->prefix + st.upd_off, st.cnt, j0, dA, p->S, D);
                ++L;
            }
            j0 += PB;
        }
    }
    if (p->n > 0) {
        k_invert_pivots<<<(p->n + 255) / 256, 256, 0, s>>>(D, dinv, p->n);
        ++L;
    }
    p->n_launches = L;
    AAADMM_CUDA_OK(cudaMemcpyAsync(p->bad_h, p->bad, sizeof(int), cudaMemcpyDeviceToHost, s));
    AAADMM_CUDA_OK(cudaGetLastError());
    return 0;
}

int factor_plan_check(FactorPlan *p, cudaStream_t s) {
    AAADMM_CUDA_OK(cudaStreamSynchronize(s));
    if (*p->bad_h) {
        set_last_error("ldlt factor: zero or non-finite pivot (the matrix is singular or not positive definite enough for LDL^T without pivoting)");
        return -1;
    }
    return 0;
}

}  // namespace aaadmm
