// Developer probe (not a test): times ND ordering + multifrontal LDL^T on a beam and checks A x = b.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <chrono>
#include "beam_scene.hpp"
#include "tet_system.hpp"
using namespace aaadmm;
int main(int argc, char **argv) {
    int cx = atoi(argv[1]), cy = atoi(argv[2]), cz = atoi(argv[3]);
    int leaf = argc > 4 ? atoi(argv[4]) : 96;
    auto t0 = std::chrono::steady_clock::now();
    auto el = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    BeamMesh m = make_beam(cx, cy, cz, 0.f);
    BeamPins pins; find_pins(m, 0, pins);
    printf("verts %d tets %d pins %zu  (%.2fs)\n", m.n_verts(), m.n_tets(), pins.idx.size(), el());
    std::vector<double> rest(m.verts.begin(), m.verts.end()), masses(m.masses.begin(), m.masses.end());
    std::vector<double> E(m.n_tets(), 1e7), nu(m.n_tets(), 0.399);
    std::vector<double> rest12(12 * (size_t)m.n_tets());
    for (int t = 0; t < m.n_tets(); ++t) for (int c = 0; c < 4; ++c) for (int j = 0; j < 3; ++j) rest12[12 * (size_t)t + 3 * c + j] = rest[3 * (size_t)m.tets[4 * t + c] + j];
    TetSystem S;
    double dt = 1.0 / 30.0;
    if (!build_tet_system(S, m.n_verts(), rest12.data(), m.n_tets(), m.tets.data(), nullptr, E.data(), nu.data(), masses.data(), pins.idx, dt * dt)) { printf("err %s\n", S.error.c_str()); return 1; }
    printf("n_free %d nnz(Ahat lower) %ld (%.2fs)\n", S.n_free, (long)S.Ahat.p[S.n_free], el());
    std::vector<double> coords(3 * (size_t)S.n_free);
    for (int k = 0; k < S.n_free; ++k) for (int j = 0; j < 3; ++j) coords[3 * (size_t)k + j] = rest[3 * (size_t)S.dev_to_vert[k] + j];
    double t1 = el();
    std::vector<int> perm = nested_dissection(S.Ahat, coords.data(), leaf);
    printf("ND order %.2fs\n", el() - t1);
    LdltFactor F = ldlt_factorize(S.Ahat, perm);
    printf("ok %d nnz(L) %ld supernodes %d flops %.3g symbolic %.2fs numeric %.2fs (%.2f GF/s)\n", (int)F.ok, (long)F.Lp[F.n], F.n_supernodes, F.flops, F.seconds_symbolic, F.seconds_numeric, F.flops * 1e-9 / F.seconds_numeric);
    // check
    int n = S.n_free; std::vector<double> x(3 * (size_t)n), b(3 * (size_t)n, 0.0), xs(3 * (size_t)n);
    for (size_t i = 0; i < x.size(); ++i) x[i] = sin(0.001 * i) + 0.5;
    for (int j = 0; j < n; ++j) for (int64_t p = S.Ahat.p[j]; p < S.Ahat.p[j + 1]; ++p) { int i = S.Ahat.i[p]; double v = S.Ahat.x[p];
        for (int r = 0; r < 3; ++r) { b[3 * (size_t)i + r] += v * x[3 * (size_t)j + r]; if (i != j) b[3 * (size_t)j + r] += v * x[3 * (size_t)i + r]; } }
    double t2 = el(); ldlt_solve_host(F, b.data(), xs.data(), 3); double ts = el() - t2;
    double err = 0, nrm = 0; for (size_t i = 0; i < x.size(); ++i) { err = fmax(err, fabs(x[i] - xs[i])); nrm = fmax(nrm, fabs(x[i])); }
    printf("solve %.3fs relerr %.3g\n", ts, err / nrm);
    return 0;
}
