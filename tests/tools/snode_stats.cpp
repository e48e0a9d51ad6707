// Developer probe (not a test): supernode / panel statistics of the beam factor, for the layout of ldlt_apply.cu.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <map>
#include "beam_scene.hpp"
#include "tet_system.hpp"
using namespace aaadmm;
int main(int argc, char **argv) {
    int cx = atoi(argv[1]), cy = atoi(argv[2]), cz = atoi(argv[3]);
    int leaf = argc > 4 ? atoi(argv[4]) : 96;
    int kSmall = argc > 5 ? atoi(argv[5]) : 96;
    BeamMesh m = make_beam(cx, cy, cz, 0.f);
    BeamPins pins; find_pins(m, 0, pins);
    std::vector<double> rest(m.verts.begin(), m.verts.end()), masses(m.masses.begin(), m.masses.end());
    std::vector<double> E(m.n_tets(), 1e7), nu(m.n_tets(), 0.399);
    std::vector<double> rest12(12 * (size_t)m.n_tets());
    for (int t = 0; t < m.n_tets(); ++t) for (int c = 0; c < 4; ++c) for (int j = 0; j < 3; ++j) rest12[12 * (size_t)t + 3 * c + j] = rest[3 * (size_t)m.tets[4 * t + c] + j];
    TetSystem S;
    double dt = 1.0 / 30.0;
    if (!build_tet_system(S, m.n_verts(), rest12.data(), m.n_tets(), m.tets.data(), nullptr, E.data(), nu.data(), masses.data(), pins.idx, dt * dt)) return 1;
    std::vector<double> coords(3 * (size_t)S.n_free);
    for (int k = 0; k < S.n_free; ++k) for (int j = 0; j < 3; ++j) coords[3 * (size_t)k + j] = rest[3 * (size_t)S.dev_to_vert[k] + j];
    std::vector<int> perm = nested_dissection(S.Ahat, coords.data(), leaf);
    LdltFactor F = ldlt_factorize(S.Ahat, perm);
    const int n = F.n; const auto &Lp = F.Lp; const auto &Li = F.Li;
    printf("n %d nnz %ld\n", n, (long)Lp[n]);
    const int kCap = 6144;
    std::vector<int> blk_of(n), blk_first;
    for (int j = 0; j < n; ++j) {
        bool join = false;
        if (j > 0) {
            const int64_t c0 = Lp[j] - Lp[j - 1], c1 = Lp[j + 1] - Lp[j];
            const bool chain = c0 > 0 && Li[Lp[j - 1]] == j;
            const int cur = j - blk_first.back();
            if (chain && cur < kCap && (c0 == c1 + 1 || cur < kSmall)) join = true;
        }
        if (!join) blk_first.push_back(j);
        blk_of[j] = (int)blk_first.size() - 1;
    }
    const int nb = (int)blk_first.size(); blk_first.push_back(n);
    std::vector<int> level(nb, 0);
    std::vector<int64_t> k(nb, 0), offnnz(nb, 0);
    for (int b = 0; b < nb; ++b) {
        const int jl = blk_first[b + 1] - 1;
        int64_t kk = 0;
        for (int64_t p = Lp[jl]; p < Lp[jl + 1]; ++p) kk += 1;  // all rows of the last column are off-block
        k[b] = kk;
        for (int j = blk_first[b]; j < blk_first[b + 1]; ++j)
            for (int64_t p = Lp[j]; p < Lp[j + 1]; ++p) {
                const int bi = blk_of[Li[p]];
                if (bi != b) { offnnz[b]++; if (level[bi] < level[b] + 1) level[bi] = level[b] + 1; }
            }
    }
    int nlev = 0; for (int b = 0; b < nb; ++b) nlev = std::max(nlev, level[b] + 1);
    double padded = 0, off = 0, dense = 0, tri = 0;
    std::vector<double> lev_pad(nlev, 0), lev_dense(nlev, 0); std::vector<int> lev_nb(nlev, 0), lev_maxns(nlev, 0), lev_maxk(nlev, 0);
    for (int b = 0; b < nb; ++b) {
        const double ns = blk_first[b + 1] - blk_first[b];
        padded += ns * k[b]; off += offnnz[b]; dense += ns * ns; tri += ns * (ns - 1) / 2;
        lev_pad[level[b]] += ns * k[b]; lev_dense[level[b]] += ns * ns; lev_nb[level[b]]++;
        lev_maxns[level[b]] = std::max(lev_maxns[level[b]], (int)ns); lev_maxk[level[b]] = std::max<int>(lev_maxk[level[b]], (int)k[b]);
    }
    printf("blocks %d levels %d  off nnz %.4g  padded panel %.4g (x%.3f)  dense sq %.4g  tri %.4g\n", nb, nlev, off, padded, padded / off, dense, tri);
    for (int l = 0; l < nlev; ++l)
        printf("  lev %2d blocks %6d  panel %.4g (%.1f%%)  dense %.4g  max ns %d max k %d\n", l, lev_nb[l], lev_pad[l], 100 * lev_pad[l] / padded, lev_dense[l], lev_maxns[l], lev_maxk[l]);
    // children contributions per front row
    {
        std::vector<int> parent(nb, -1);
        std::vector<std::vector<int>> R(nb);
        for (int b = 0; b < nb; ++b) { const int jl = blk_first[b + 1] - 1; R[b].assign(Li.begin() + Lp[jl], Li.begin() + Lp[jl + 1]); if (!R[b].empty()) parent[b] = blk_of[R[b][0]]; }
        std::map<long, int> cnt; std::map<int,long> hist; std::map<int,int> nchild;
        for (int b = 0; b < nb; ++b) { if (parent[b] < 0) continue; nchild[parent[b]]++; for (int i : R[b]) cnt[((long)parent[b] << 32) | (unsigned)i]++; }
        for (auto &e : cnt) hist[e.second]++;
        for (auto &e : hist) printf("rows with %d contributions: %ld\n", e.first, e.second);
        std::map<int,int> ch; for (auto &e : nchild) ch[e.second]++;
        for (auto &e : ch) printf("fronts with %d children: %d\n", e.first, e.second);
    }
    return 0;
}
