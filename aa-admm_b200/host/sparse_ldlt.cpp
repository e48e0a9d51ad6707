#include "sparse_ldlt.hpp"

#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <numeric>
#include <stdexcept>

namespace aaadmm {

namespace {
double now_s() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}
}  // namespace

SymLower sym_from_triplets(int n, const std::vector<int> &r, const std::vector<int> &c,
                           const std::vector<double> &v, bool input_is_full) {
    SymLower A;
    A.n = n;
    std::vector<int64_t> cnt(n + 1, 0);
    const size_t nt = r.size();
    for (size_t k = 0; k < nt; ++k) {
        int rr = r[k], cc = c[k];
        if (rr < cc) {
            if (input_is_full) continue;
            std::swap(rr, cc);
        }
        cnt[cc + 1]++;
    }
    for (int j = 0; j < n; ++j) cnt[j + 1] += cnt[j];
    std::vector<int> ti(cnt[n]);
    std::vector<double> tx(cnt[n]);
    std::vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
    for (size_t k = 0; k < nt; ++k) {
        int rr = r[k], cc = c[k];
        if (rr < cc) {
            if (input_is_full) continue;
            std::swap(rr, cc);
        }
        int64_t p = pos[cc]++;
        ti[p] = rr;
        tx[p] = v[k];
    }
    A.p.assign(n + 1, 0);
    std::vector<std::pair<int, double>> col;
    for (int j = 0; j < n; ++j) {
        col.clear();
        for (int64_t p = cnt[j]; p < cnt[j + 1]; ++p) col.emplace_back(ti[p], tx[p]);
        std::stable_sort(col.begin(), col.end(),
                         [](const std::pair<int, double> &a, const std::pair<int, double> &b) { return a.first < b.first; });
        size_t k = 0;
        while (k < col.size()) {
            int row = col[k].first;
            double s = 0;
            while (k < col.size() && col[k].first == row) s += col[k++].second;
            A.i.push_back(row);
            A.x.push_back(s);
        }
        A.p[j + 1] = (int64_t)A.i.size();
    }
    return A;
}

// ------------------------------------------------------------------------------------------
// Nested dissection
// ------------------------------------------------------------------------------------------
namespace {
struct NdCtx {
    int n;
    std::vector<int64_t> adjp;
    std::vector<int> adji;
    const double *coords;
    int leaf;
    std::vector<char> side;
    std::vector<int> order;  // order[new] = old
    std::vector<int> dist;   // BFS scratch
};

void sort_along(const NdCtx &c, std::vector<int> &ids) {
    if (!c.coords || ids.size() < 2) {
        std::sort(ids.begin(), ids.end());
        return;
    }
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int v : ids)
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], c.coords[3 * (size_t)v + k]);
            hi[k] = std::max(hi[k], c.coords[3 * (size_t)v + k]);
        }
    int ax[3] = {0, 1, 2};
    std::sort(ax, ax + 3, [&](int a, int b) { return (hi[a] - lo[a]) > (hi[b] - lo[b]); });
    const double *X = c.coords;
    std::sort(ids.begin(), ids.end(), [&](int a, int b) {
        for (int k = 0; k < 3; ++k) {
            double xa = X[3 * (size_t)a + ax[k]], xb = X[3 * (size_t)b + ax[k]];
            if (xa != xb) return xa < xb;
        }
        return a < b;
    });
}

// Splits ids into A (side 1) and B (side 2); returns false if no useful split exists.
bool split_geometric(NdCtx &c, const std::vector<int> &ids, std::vector<int> &A, std::vector<int> &B) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int v : ids)
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], c.coords[3 * (size_t)v + k]);
            hi[k] = std::max(hi[k], c.coords[3 * (size_t)v + k]);
        }
    int ax[3] = {0, 1, 2};
    std::sort(ax, ax + 3, [&](int a, int b) { return (hi[a] - lo[a]) > (hi[b] - lo[b]); });
    std::vector<double> key(ids.size());
    for (int t = 0; t < 3; ++t) {
        const int a = ax[t];
        if (!(hi[a] > lo[a])) break;
        for (size_t k = 0; k < ids.size(); ++k) key[k] = c.coords[3 * (size_t)ids[k] + a];
        std::vector<double> tmp(key);
        std::nth_element(tmp.begin(), tmp.begin() + tmp.size() / 2, tmp.end());
        double med = tmp[tmp.size() / 2];
        // A = key <= med would put the median layer in A; keep B non-empty
        size_t nle = 0;
        for (double kk : key) nle += (kk <= med);
        if (nle == ids.size()) {
            // median equals the max: use strictly-less split
            size_t nlt = 0;
            for (double kk : key) nlt += (kk < med);
            if (nlt == 0) continue;
            A.clear();
            B.clear();
            for (size_t k = 0; k < ids.size(); ++k) (key[k] < med ? A : B).push_back(ids[k]);
            return true;
        }
        A.clear();
        B.clear();
        for (size_t k = 0; k < ids.size(); ++k) (key[k] <= med ? A : B).push_back(ids[k]);
        return true;
    }
    return false;
}

// BFS level-structure bisection for graphs without coordinates.
bool split_bfs(NdCtx &c, const std::vector<int> &ids, std::vector<int> &A, std::vector<int> &B) {
    // side==3 marks membership during this call
    for (int v : ids) c.side[v] = 3;
    int start = ids[0];
    std::vector<int> q;
    for (int pass = 0; pass < 3; ++pass) {
        q.clear();
        q.push_back(start);
        c.dist[start] = 0;
        for (int v : ids) c.dist[v] = -1;
        c.dist[start] = 0;
        for (size_t h = 0; h < q.size(); ++h) {
            int v = q[h];
            for (int64_t p = c.adjp[v]; p < c.adjp[v + 1]; ++p) {
                int u = c.adji[p];
                if (c.side[u] == 3 && c.dist[u] < 0) {
                    c.dist[u] = c.dist[v] + 1;
                    q.push_back(u);
                }
            }
        }
        start = q.back();
    }
    bool ok = false;
    if (q.size() == ids.size()) {
        size_t half = q.size() / 2;
        int dcut = c.dist[q[half]];
        A.clear();
        B.clear();
        for (int v : q) (c.dist[v] < dcut || (c.dist[v] == dcut && dcut == 0) ? A : B).push_back(v);
        if (A.empty()) {
            A.clear();
            B.clear();
            for (int v : q) (c.dist[v] <= dcut ? A : B).push_back(v);
        }
        ok = !A.empty() && !B.empty();
    } else {
        // disconnected: component vs rest
        A = q;
        B.clear();
        for (int v : ids)
            if (c.dist[v] < 0) B.push_back(v);
        ok = !B.empty();
    }
    for (int v : ids) c.side[v] = 0;
    return ok;
}

void nd_rec(NdCtx &c, std::vector<int> ids) {
    if ((int)ids.size() <= c.leaf) {
        sort_along(c, ids);
        c.order.insert(c.order.end(), ids.begin(), ids.end());
        return;
    }
    std::vector<int> A, B;
    bool ok = c.coords ? split_geometric(c, ids, A, B) : false;
    if (!ok) ok = split_bfs(c, ids, A, B);
    if (!ok) {
        sort_along(c, ids);
        c.order.insert(c.order.end(), ids.begin(), ids.end());
        return;
    }
    for (int v : A) c.side[v] = 1;
    for (int v : B) c.side[v] = 2;
    std::vector<int> S, A2;
    for (int v : A) {
        bool touches = false;
        for (int64_t p = c.adjp[v]; p < c.adjp[v + 1] && !touches; ++p) touches = (c.side[c.adji[p]] == 2);
        (touches ? S : A2).push_back(v);
    }
    for (int v : ids) c.side[v] = 0;
    std::vector<int>().swap(ids);
    std::vector<int>().swap(A);
    nd_rec(c, std::move(A2));
    nd_rec(c, std::move(B));
    sort_along(c, S);
    c.order.insert(c.order.end(), S.begin(), S.end());
}
}  // namespace

std::vector<int> nested_dissection(const SymLower &A, const double *coords, int leaf_size) {
    NdCtx c;
    c.n = A.n;
    c.coords = coords;
    c.leaf = std::max(1, leaf_size);
    c.adjp.assign(A.n + 1, 0);
    for (int j = 0; j < A.n; ++j)
        for (int64_t p = A.p[j]; p < A.p[j + 1]; ++p) {
            int i = A.i[p];
            if (i == j) continue;
            c.adjp[i + 1]++;
            c.adjp[j + 1]++;
        }
    for (int j = 0; j < A.n; ++j) c.adjp[j + 1] += c.adjp[j];
    c.adji.resize(c.adjp[A.n]);
    std::vector<int64_t> pos(c.adjp.begin(), c.adjp.end() - 1);
    for (int j = 0; j < A.n; ++j)
        for (int64_t p = A.p[j]; p < A.p[j + 1]; ++p) {
            int i = A.i[p];
            if (i == j) continue;
            c.adji[pos[i]++] = j;
            c.adji[pos[j]++] = i;
        }
    c.side.assign(A.n, 0);
    c.dist.assign(A.n, -1);
    c.order.reserve(A.n);
    std::vector<int> all(A.n);
    std::iota(all.begin(), all.end(), 0);
    nd_rec(c, std::move(all));
    return c.order;
}

// ------------------------------------------------------------------------------------------
// Supernodal multifrontal LDL^T
// ------------------------------------------------------------------------------------------
namespace {

constexpr int NB = 48;  // panel width of the dense partial factorisation

// F: nf x nf column-major (lower part used), factor the first ns columns in place,
// leave the Schur complement in the trailing block. W: scratch nf*NB doubles.
bool partial_ldlt(double *F, int nf, int ns, double *Dout, double *W, bool parallel_update) {
    const size_t ld = (size_t)nf;
    for (int k0 = 0; k0 < ns; k0 += NB) {
        const int kb = std::min(NB, ns - k0);
        for (int k = k0; k < k0 + kb; ++k) {
            double *ck = F + k * ld;
            const double d = ck[k];
            if (d == 0.0 || !std::isfinite(d)) return false;
            Dout[k] = d;
            const double inv = 1.0 / d;
            double *wk = W + (size_t)(k - k0) * ld;
            for (int i = k + 1; i < nf; ++i) {
                wk[i] = ck[i];
                ck[i] *= inv;
            }
            for (int j = k + 1; j < k0 + kb; ++j) {
                const double c = wk[j];
                double *cj = F + j * ld;
                for (int i = j; i < nf; ++i) cj[i] -= ck[i] * c;
            }
        }
        const int j0 = k0 + kb;
        if (j0 >= nf) continue;
        const double *Lp = F + (size_t)k0 * ld;  // panel columns (scaled L)
        const int nblk = (nf - j0 + 3) / 4;
        auto body = [&](int b) {
            const int j = j0 + 4 * b;
            const int jw = std::min(4, nf - j);
            // small triangle rows j..j+jw-1 handled per column, then the common rows
            for (int jj = 0; jj < jw; ++jj) {
                double *cj = F + (size_t)(j + jj) * ld;
                const int iend = std::min(nf, j + jw);
                for (int i = j + jj; i < iend; ++i) {
                    double s = 0;
                    for (int kk = 0; kk < kb; ++kk) s += Lp[(size_t)kk * ld + i] * W[(size_t)kk * ld + j + jj];
                    cj[i] -= s;
                }
            }
            const int i0 = j + jw;
            if (i0 >= nf) return;
            if (jw == 4) {
                double *a0 = F + (size_t)(j + 0) * ld, *a1 = F + (size_t)(j + 1) * ld;
                double *a2 = F + (size_t)(j + 2) * ld, *a3 = F + (size_t)(j + 3) * ld;
                constexpr int TB = 256;
                for (int ib = i0; ib < nf; ib += TB) {
                    const int ie = std::min(nf, ib + TB);
                    int kk = 0;
                    for (; kk + 4 <= kb; kk += 4) {
                        const double *l0 = Lp + (size_t)(kk + 0) * ld, *l1 = Lp + (size_t)(kk + 1) * ld;
                        const double *l2 = Lp + (size_t)(kk + 2) * ld, *l3 = Lp + (size_t)(kk + 3) * ld;
                        const double *w0 = W + (size_t)(kk + 0) * ld + j, *w1 = W + (size_t)(kk + 1) * ld + j;
                        const double *w2 = W + (size_t)(kk + 2) * ld + j, *w3 = W + (size_t)(kk + 3) * ld + j;
                        const double c00 = w0[0], c01 = w0[1], c02 = w0[2], c03 = w0[3];
                        const double c10 = w1[0], c11 = w1[1], c12 = w1[2], c13 = w1[3];
                        const double c20 = w2[0], c21 = w2[1], c22 = w2[2], c23 = w2[3];
                        const double c30 = w3[0], c31 = w3[1], c32 = w3[2], c33 = w3[3];
#pragma omp simd
                        for (int i = ib; i < ie; ++i) {
                            const double x0 = l0[i], x1 = l1[i], x2 = l2[i], x3 = l3[i];
                            a0[i] -= x0 * c00 + x1 * c10 + x2 * c20 + x3 * c30;
                            a1[i] -= x0 * c01 + x1 * c11 + x2 * c21 + x3 * c31;
                            a2[i] -= x0 * c02 + x1 * c12 + x2 * c22 + x3 * c32;
                            a3[i] -= x0 * c03 + x1 * c13 + x2 * c23 + x3 * c33;
                        }
                    }
                    for (; kk < kb; ++kk) {
                        const double *l0 = Lp + (size_t)kk * ld;
                        const double *w0 = W + (size_t)kk * ld + j;
                        const double c0 = w0[0], c1 = w0[1], c2 = w0[2], c3 = w0[3];
#pragma omp simd
                        for (int i = ib; i < ie; ++i) {
                            const double x0 = l0[i];
                            a0[i] -= x0 * c0;
                            a1[i] -= x0 * c1;
                            a2[i] -= x0 * c2;
                            a3[i] -= x0 * c3;
                        }
                    }
                }
            } else {
                for (int jj = 0; jj < jw; ++jj) {
                    double *cj = F + (size_t)(j + jj) * ld;
                    for (int kk = 0; kk < kb; ++kk) {
                        const double c = W[(size_t)kk * ld + j + jj];
                        const double *l0 = Lp + (size_t)kk * ld;
                        for (int i = i0; i < nf; ++i) cj[i] -= l0[i] * c;
                    }
                }
            }
        };
        if (parallel_update && nblk >= 16) {
#pragma omp taskloop grainsize(4) default(shared)
            for (int b = 0; b < nblk; ++b) body(b);
        } else {
            for (int b = 0; b < nblk; ++b) body(b);
        }
    }
    return true;
}

struct Symbolic {
    int n = 0, nsn = 0;
    std::vector<int> sn_first, sn_last, sn_of;
    std::vector<std::vector<int>> sn_rows;  // pattern of the first column (rows > first), sorted
    std::vector<int> sn_parent;
    std::vector<std::vector<int>> sn_children;
};

}  // namespace

// Postorder of the elimination tree of P A P^T, composed with P. It changes neither the fill nor the tree, only the
// numbering: every subtree becomes a run of consecutive columns. The dissection above numbers the nodes of a leaf domain
// along a coordinate, which is a topological order of the tree but not a postorder; the device-side setup merges a small
// subtree into ONE front only when its columns are consecutive (csrc/ldlt_apply.cu, rule 1), and on a surface mesh
// (cfg 3) most leaf domains otherwise fall apart into chains of fronts of 1-10 columns, seven tree levels deep.
static std::vector<int> etree_postorder(const SymLower &A, const std::vector<int> &perm) {
    const int n = A.n;
    std::vector<int> iperm(n);
    for (int k = 0; k < n; ++k) iperm[perm[k]] = k;
    // rows of the strictly lower triangle of B = P A P^T: row r lists the columns c < r
    std::vector<int64_t> rp(n + 1, 0);
    for (int j = 0; j < n; ++j)
        for (int64_t p = A.p[j]; p < A.p[j + 1]; ++p) {
            const int r = iperm[A.i[p]], c = iperm[j];
            if (r != c) rp[std::max(r, c) + 1]++;
        }
    for (int j = 0; j < n; ++j) rp[j + 1] += rp[j];
    std::vector<int> ri(rp[n]);
    {
        std::vector<int64_t> pos(rp.begin(), rp.end() - 1);
        for (int j = 0; j < n; ++j)
            for (int64_t p = A.p[j]; p < A.p[j + 1]; ++p) {
                const int r = iperm[A.i[p]], c = iperm[j];
                if (r != c) ri[pos[std::max(r, c)]++] = std::min(r, c);
            }
    }
    // Liu's algorithm with path compression
    std::vector<int> parent(n, -1), anc(n, -1);
    for (int j = 0; j < n; ++j)
        for (int64_t p = rp[j]; p < rp[j + 1]; ++p) {
            int k = ri[p];
            while (anc[k] != -1 && anc[k] != j) {
                const int t = anc[k];
                anc[k] = j;
                k = t;
            }
            if (anc[k] == -1) {
                anc[k] = j;
                parent[k] = j;
            }
        }
    // depth-first postorder, children in ascending order (chains keep their order)
    std::vector<int> head(n, -1), next(n, -1);
    for (int j = n - 1; j >= 0; --j)
        if (parent[j] >= 0) {
            next[j] = head[parent[j]];
            head[parent[j]] = j;
        }
    std::vector<int> post;
    post.reserve(n);
    std::vector<int> stack;
    for (int root = 0; root < n; ++root) {
        if (parent[root] >= 0) continue;
        stack.push_back(root);
        while (!stack.empty()) {
            const int v = stack.back();
            const int c = head[v];
            if (c >= 0) {
                head[v] = next[c];
                stack.push_back(c);
            } else {
                post.push_back(v);
                stack.pop_back();
            }
        }
    }
    std::vector<int> out(n);
    for (int k = 0; k < n; ++k) out[k] = perm[post[k]];
    return out;
}

static LdltFactor ldlt_factorize_impl(const SymLower &A, const std::vector<int> &perm_in, int n_threads, bool symbolic_only) {
    LdltFactor out;
    const int n = A.n;
    out.n = n;
    static const bool no_postorder = getenv("AAADMM_NO_POSTORDER") != nullptr;  // experiments
    const std::vector<int> perm = no_postorder ? perm_in : etree_postorder(A, perm_in);
    out.perm = perm;
    if (n_threads <= 0) n_threads = omp_get_max_threads();
    double t0 = now_s();

    // ---- permuted lower matrix B = P A P^T (CSC, sorted rows) ----
    std::vector<int> iperm(n);
    for (int k = 0; k < n; ++k) iperm[perm[k]] = k;
    std::vector<int64_t> Bp(n + 1, 0);
    for (int j = 0; j < n; ++j)
        for (int64_t p = A.p[j]; p < A.p[j + 1]; ++p) {
            int r = iperm[A.i[p]], c = iperm[j];
            Bp[std::min(r, c) + 1]++;
        }
    for (int j = 0; j < n; ++j) Bp[j + 1] += Bp[j];
    std::vector<int> Bi(Bp[n]);
    std::vector<double> Bx(Bp[n]);
    {
        std::vector<int64_t> pos(Bp.begin(), Bp.end() - 1);
        for (int j = 0; j < n; ++j)
            for (int64_t p = A.p[j]; p < A.p[j + 1]; ++p) {
                int r = iperm[A.i[p]], c = iperm[j];
                if (r < c) std::swap(r, c);
                int64_t q = pos[c]++;
                Bi[q] = r;
                Bx[q] = A.x[p];
            }
#pragma omp parallel for schedule(dynamic, 256) num_threads(n_threads)
        for (int j = 0; j < n; ++j) {
            int64_t b = Bp[j], e = Bp[j + 1];
            std::vector<std::pair<int, double>> col(e - b);
            for (int64_t p = b; p < e; ++p) col[p - b] = {Bi[p], Bx[p]};
            std::sort(col.begin(), col.end(),
                      [](const std::pair<int, double> &a, const std::pair<int, double> &bb) { return a.first < bb.first; });
            for (int64_t p = b; p < e; ++p) {
                Bi[p] = col[p - b].first;
                Bx[p] = col[p - b].second;
            }
        }
    }

    // ---- symbolic: supernodes with nested column patterns ----
    Symbolic S;
    S.n = n;
    S.sn_of.assign(n, -1);
    std::vector<int> parent(n, -1), child_head(n, -1), child_next(n, -1), stamp(n, -1);
    auto below_begin = [&](int c) -> const int * {  // pattern(c) = rows > c, sorted
        int s = S.sn_of[c];
        return S.sn_rows[s].data() + (c - S.sn_first[s]);
    };
    auto below_end = [&](int c) -> const int * {
        int s = S.sn_of[c];
        return S.sn_rows[s].data() + S.sn_rows[s].size();
    };
    int cur = -1;
    std::vector<int> collect;
    for (int j = 0; j < n; ++j) {
        bool join = false;
        if (cur >= 0 && S.sn_last[cur] == j - 1 && parent[j - 1] == j) {
            join = true;
            for (int64_t p = Bp[j]; p < Bp[j + 1] && join; ++p)
                if (Bi[p] > j && stamp[Bi[p]] != cur) join = false;
            for (int c = child_head[j]; c >= 0 && join; c = child_next[c]) {
                if (c == j - 1) continue;
                for (const int *r = below_begin(c), *e = below_end(c); r < e; ++r)
                    if (*r != j && stamp[*r] != cur) {
                        join = false;
                        break;
                    }
            }
        }
        if (join) {
            S.sn_last[cur] = j;
            S.sn_of[j] = cur;
        } else {
            cur = S.nsn++;
            S.sn_first.push_back(j);
            S.sn_last.push_back(j);
            S.sn_of[j] = cur;
            collect.clear();
            for (int64_t p = Bp[j]; p < Bp[j + 1]; ++p) {
                int r = Bi[p];
                if (r > j && stamp[r] != cur) {
                    stamp[r] = cur;
                    collect.push_back(r);
                }
            }
            for (int c = child_head[j]; c >= 0; c = child_next[c])
                for (const int *r = below_begin(c), *e = below_end(c); r < e; ++r)
                    if (*r != j && stamp[*r] != cur) {
                        stamp[*r] = cur;
                        collect.push_back(*r);
                    }
            std::sort(collect.begin(), collect.end());
            S.sn_rows.emplace_back(collect);
        }
        // parent(j) = smallest row of pattern(j)
        const int *b = below_begin(j), *e = below_end(j);
        if (b < e) {
            parent[j] = *b;
            child_next[j] = child_head[*b];
            child_head[*b] = j;
        }
    }
    const int nsn = S.nsn;
    S.sn_parent.assign(nsn, -1);
    S.sn_children.assign(nsn, {});
    for (int s = 0; s < nsn; ++s) {
        int last = S.sn_last[s];
        if (parent[last] >= 0) {
            S.sn_parent[s] = S.sn_of[parent[last]];
            S.sn_children[S.sn_parent[s]].push_back(s);
        }
    }
    // L column pointers
    out.Lp.assign(n + 1, 0);
    for (int j = 0; j < n; ++j) out.Lp[j + 1] = out.Lp[j] + (below_end(j) - below_begin(j));
    const int64_t nnzL = out.Lp[n];
    out.Li.resize(nnzL);
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads)
    for (int j = 0; j < n; ++j) std::copy(below_begin(j), below_end(j), out.Li.begin() + out.Lp[j]);
    out.n_supernodes = nsn;
    double t1 = now_s();
    out.seconds_symbolic = t1 - t0;
    if (symbolic_only) {  // pattern of L only: the numeric phase runs elsewhere (csrc/ldlt_factor.cu)
        out.ok = true;
        return out;
    }
    out.Lx.resize(nnzL);
    out.D.assign(n, 0.0);

    // ---- numeric multifrontal ----
    std::vector<double> work(nsn, 0.0);
    double flops = 0;
    for (int s = 0; s < nsn; ++s) {
        double ns = S.sn_last[s] - S.sn_first[s] + 1;
        double nf = ns + (double)S.sn_rows[s].size() - (ns - 1);
        double f = 0;
        // sum_{k=0}^{ns-1} (nf-k)^2
        for (int k = 0; k < (int)ns; ++k) f += (nf - k) * (nf - k);
        work[s] += f;
        flops += f;
        if (S.sn_parent[s] >= 0) work[S.sn_parent[s]] += work[s];
    }
    out.flops = flops;

    std::vector<std::vector<double>> update(nsn);  // Schur complements awaiting the parent
    std::vector<std::vector<int>> loc_tl(n_threads);
    bool failed = false;

    std::function<void(int)> process = [&](int s) {
        for (int c : S.sn_children[s]) {
#pragma omp task default(shared) firstprivate(c) if (work[c] > 2e6)
            process(c);
        }
#pragma omp taskwait
        if (failed) return;
        const int first = S.sn_first[s], last = S.sn_last[s];
        const int ns = last - first + 1;
        const std::vector<int> &rows = S.sn_rows[s];
        const int nb = (int)rows.size() - (ns - 1);
        const int nf = ns + nb;
        const int *below = rows.data() + (ns - 1);
        std::vector<int> &loc = loc_tl[omp_get_thread_num()];
        if (loc.empty()) loc.assign(n, -1);
        for (int k = 0; k < ns; ++k) loc[first + k] = k;
        for (int k = 0; k < nb; ++k) loc[below[k]] = ns + k;
        std::vector<double> F((size_t)nf * nf, 0.0);
        for (int j = first; j <= last; ++j) {
            double *col = F.data() + (size_t)(j - first) * nf;
            for (int64_t p = Bp[j]; p < Bp[j + 1]; ++p) col[loc[Bi[p]]] += Bx[p];
        }
        for (int c : S.sn_children[s]) {
            const int cns = S.sn_last[c] - S.sn_first[c] + 1;
            const std::vector<int> &crows = S.sn_rows[c];
            const int cnb = (int)crows.size() - (cns - 1);
            const int *cb = crows.data() + (cns - 1);
            std::vector<double> &U = update[c];
            for (int b = 0; b < cnb; ++b) {
                double *col = F.data() + (size_t)loc[cb[b]] * nf;
                const double *ucol = U.data() + (size_t)b * cnb;
                for (int a = b; a < cnb; ++a) col[loc[cb[a]]] += ucol[a];
            }
            std::vector<double>().swap(U);
        }
        std::vector<double> W((size_t)nf * std::min(NB, ns));
        if (!partial_ldlt(F.data(), nf, ns, out.D.data() + first, W.data(), nf >= 768)) {
            failed = true;
            return;
        }
        for (int k = 0; k < ns; ++k) {
            const double *col = F.data() + (size_t)k * nf;
            double *dst = out.Lx.data() + out.Lp[first + k];
            for (int i = k + 1; i < nf; ++i) dst[i - k - 1] = col[i];
        }
        if (nb > 0 && S.sn_parent[s] >= 0) {
            std::vector<double> &U = update[s];
            U.resize((size_t)nb * nb);
            for (int b = 0; b < nb; ++b) {
                const double *col = F.data() + (size_t)(ns + b) * nf + ns;
                std::copy(col + b, col + nb, U.data() + (size_t)b * nb + b);
            }
        }
    };

#pragma omp parallel num_threads(n_threads)
    {
#pragma omp single
        {
            for (int s = 0; s < nsn; ++s)
                if (S.sn_parent[s] < 0) {
#pragma omp task default(shared) firstprivate(s)
                    process(s);
                }
        }
    }
    out.ok = !failed;
    out.seconds_numeric = now_s() - t1;
    return out;
}

LdltFactor ldlt_factorize(const SymLower &A, const std::vector<int> &perm, int n_threads) {
    return ldlt_factorize_impl(A, perm, n_threads, false);
}

LdltFactor ldlt_symbolic(const SymLower &A, const std::vector<int> &perm, int n_threads) {
    return ldlt_factorize_impl(A, perm, n_threads, true);
}

void ldlt_solve_host(const LdltFactor &F, const double *b, double *x, int nrhs) {
    const int n = F.n;
    std::vector<double> y((size_t)n * nrhs);
    for (int k = 0; k < n; ++k)
        for (int r = 0; r < nrhs; ++r) y[(size_t)k * nrhs + r] = b[(size_t)F.perm[k] * nrhs + r];
    for (int j = 0; j < n; ++j)
        for (int64_t p = F.Lp[j]; p < F.Lp[j + 1]; ++p)
            for (int r = 0; r < nrhs; ++r) y[(size_t)F.Li[p] * nrhs + r] -= F.Lx[p] * y[(size_t)j * nrhs + r];
    for (int j = 0; j < n; ++j)
        for (int r = 0; r < nrhs; ++r) y[(size_t)j * nrhs + r] /= F.D[j];
    for (int j = n - 1; j >= 0; --j)
        for (int64_t p = F.Lp[j]; p < F.Lp[j + 1]; ++p)
            for (int r = 0; r < nrhs; ++r) y[(size_t)j * nrhs + r] -= F.Lx[p] * y[(size_t)F.Li[p] * nrhs + r];
    for (int k = 0; k < n; ++k)
        for (int r = 0; r < nrhs; ++r) x[(size_t)F.perm[k] * nrhs + r] = y[(size_t)k * nrhs + r];
}


// ---- on-disk factor cache ----------------------------------------------------------------------------
namespace {
inline void fnv(uint64_t &h, const void *p, size_t bytes) {
    const unsigned char *c = static_cast<const unsigned char *>(p);
    for (size_t i = 0; i < bytes; ++i) {
        h ^= c[i];
        h *= 1099511628211ull;
    }
}
}  // namespace

uint64_t matrix_key(const SymLower &A) {
    uint64_t h = 1469598103934665603ull;
    fnv(h, &A.n, sizeof(A.n));
    fnv(h, A.p.data(), A.p.size() * sizeof(A.p[0]));
    fnv(h, A.i.data(), A.i.size() * sizeof(A.i[0]));
    fnv(h, A.x.data(), A.x.size() * sizeof(A.x[0]));
    return h;
}

bool ldlt_save(const LdltFactor &F, uint64_t key, const std::string &path) {
    const std::string tmp = path + ".tmp";
    FILE *fp = fopen(tmp.c_str(), "wb");
    if (!fp) return false;
    const int64_t nnz = F.Lp.empty() ? 0 : F.Lp[F.n];
    bool ok = fwrite("AAFACT01", 1, 8, fp) == 8 && fwrite(&key, 8, 1, fp) == 1 && fwrite(&F.n, 4, 1, fp) == 1 &&
              fwrite(&nnz, 8, 1, fp) == 1;
    ok = ok && fwrite(F.perm.data(), 4, (size_t)F.n, fp) == (size_t)F.n;
    ok = ok && fwrite(F.Lp.data(), 8, (size_t)F.n + 1, fp) == (size_t)F.n + 1;
    ok = ok && fwrite(F.Li.data(), 4, (size_t)nnz, fp) == (size_t)nnz;
    ok = ok && fwrite(F.Lx.data(), 8, (size_t)nnz, fp) == (size_t)nnz;
    ok = ok && fwrite(F.D.data(), 8, (size_t)F.n, fp) == (size_t)F.n;
    ok = (fclose(fp) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) {
        remove(tmp.c_str());
        return false;
    }
    return true;
}

bool ldlt_load(const std::string &path, uint64_t key, LdltFactor &F) {
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp) return false;
    char magic[8];
    uint64_t k = 0;
    int n = 0;
    int64_t nnz = 0;
    bool ok = fread(magic, 1, 8, fp) == 8 && memcmp(magic, "AAFACT01", 8) == 0 && fread(&k, 8, 1, fp) == 1 && k == key &&
              fread(&n, 4, 1, fp) == 1 && fread(&nnz, 8, 1, fp) == 1 && n >= 0 && nnz >= 0;
    if (ok) {
        // the header must account for the file's size exactly before anything is allocated from it
        const int64_t expect = 8 + 8 + 4 + 8 + 4 * (int64_t)n + 8 * ((int64_t)n + 1) + 4 * nnz + 8 * nnz + 8 * (int64_t)n;
        const long here = ftell(fp);
        ok = fseek(fp, 0, SEEK_END) == 0 && (int64_t)ftell(fp) == expect && fseek(fp, here, SEEK_SET) == 0;
    }
    if (ok) {
        LdltFactor G;
        G.n = n;
        G.perm.resize(n);
        G.Lp.resize((size_t)n + 1);
        G.Li.resize((size_t)nnz);
        G.Lx.resize((size_t)nnz);
        G.D.resize(n);
        ok = fread(G.perm.data(), 4, (size_t)n, fp) == (size_t)n && fread(G.Lp.data(), 8, (size_t)n + 1, fp) == (size_t)n + 1 &&
             fread(G.Li.data(), 4, (size_t)nnz, fp) == (size_t)nnz && fread(G.Lx.data(), 8, (size_t)nnz, fp) == (size_t)nnz &&
             fread(G.D.data(), 8, (size_t)n, fp) == (size_t)n && G.Lp[n] == nnz && G.Lp[0] == 0;
        // structure checks: a damaged or foreign file is a cache miss, never an out-of-bounds index later on
        if (ok) {
            std::vector<char> seen((size_t)n, 0);
            for (int i = 0; ok && i < n; ++i) {
                const int p = G.perm[i];
                ok = p >= 0 && p < n && !seen[p];
                if (ok) seen[p] = 1;
            }
        }
        for (int j = 0; ok && j < n; ++j) {
            ok = G.Lp[j] <= G.Lp[j + 1] && std::isfinite(G.D[j]) && G.D[j] != 0.0;
            int prev = j;  // strictly lower triangle, rows ascending
            for (int64_t q = G.Lp[j]; ok && q < G.Lp[j + 1]; ++q) {
                ok = G.Li[q] > prev && G.Li[q] < n;
                prev = G.Li[q];
            }
        }
        if (ok) {
            G.ok = true;
            F = std::move(G);
        }
    }
    fclose(fp);
    return ok;
}

}  // namespace aaadmm
