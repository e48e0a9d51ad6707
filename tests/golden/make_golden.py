"""Regenerates tests/golden/*.npz from the UNMODIFIED reference compiled into oracle/_ref
(run in the build container, where /root/reference exists):  python tests/golden/make_golden.py
Flags of the reference build are pinned in oracle/Makefile (-O2 -fopenmp -DNDEBUG -ffp-contract=off)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import aa_admm_b200 as A  # noqa: E402  (host-side scene builder only)
from oracle import refbind as R  # noqa: E402
from scenes import beam_arrays  # noqa: E402


def solver_case(variant, dims, m, accel, frames=2, n_beams=1):
    scene = beam_arrays(A, *dims, n_beams=n_beams)
    verts, tets, masses, pidx, ppts, pside = scene.arrays()
    s = R.RefSolver(variant)
    s.add_tetmesh(verts, tets, masses, 1e7, 0.399, 0)
    dt = 1.0 / 30.0
    s.set_pins(pidx, scene.stretch(dt))
    s.initialize(dt, 100, -9.8, max(m, 1), accel, 1.0)
    prim, comb, rej, xs, rows = [], [], [], [], []
    for _ in range(frames):
        s.set_pins(pidx, scene.stretch(dt))
        h = s.step()
        p, c, r = np.zeros(100), np.zeros(100), np.zeros(100)
        p[:len(h)], c[:len(h)] = h[:, 1], h[:, 2]
        if h.shape[1] > 3:
            r[:len(h)] = h[:, 3]
        prim.append(p), comb.append(c), rej.append(r), rows.append(len(h)), xs.append(s.x())
    return dict(prim=np.array(prim), comb=np.array(comb), rej=np.array(rej), rows=np.array(rows), x=np.array(xs),
                dims=np.array(dims), m=m, accel=int(accel), n_beams=n_beams)


def main():
    assert R.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    for variant in ("hard", "xzu"):
        np.savez_compressed(os.path.join(HERE, "%s_beam_12x3x3_m5.npz" % variant), **solver_case(variant, (12, 3, 3), 5, True))
        np.savez_compressed(os.path.join(HERE, "%s_beam_12x3x3_noacc.npz" % variant), **solver_case(variant, (12, 3, 3), 0, False))
        np.savez_compressed(os.path.join(HERE, "%s_beam_8x2x2_m3.npz" % variant), **solver_case(variant, (8, 2, 2), 3, True))
    np.savez_compressed(os.path.join(HERE, "hard_3beams_6x2x2_m5.npz"), **solver_case("hard", (6, 2, 2), 5, True, n_beams=3))
    # scene generator
    g = {}
    for dims in [(12, 3, 3), (5, 2, 4), (7, 7, 1)]:
        v, t, m = R.ref_make_beam(*dims, 1.75)
        k = "%dx%dx%d" % dims
        g["verts_" + k], g["tets_" + k], g["masses_" + k] = v, t, m
    np.savez_compressed(os.path.join(HERE, "beam_scene.npz"), **g)
    # per-element kernels
    rng = np.random.default_rng(0)
    F = np.eye(3).reshape(1, 9) + 0.3 * rng.standard_normal((512, 9))
    F[:32] *= -1.0
    F[32:64, 6:] = 0.0
    F[64] = 0.0
    tv = rng.standard_normal((64, 4, 3)) * 0.1 + np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    consts = np.array([[*R.ref_tet_constants(tv[i])[:2], *R.ref_tet_constants(tv[i])[2]] for i in range(64)])
    np.savez_compressed(os.path.join(HERE, "tet_element.npz"), F=F, prox=R.ref_tet_prox(F), fmuvt=R.ref_tet_F_minus_UVt(F),
                        tet_verts=tv, tet_consts=consts)
    # COD solves
    Ms, rhss, sols, ranks = [], [], [], []
    for m in (2, 3, 5, 6):
        rng = np.random.default_rng(m)
        for trial in range(12):
            B = rng.standard_normal((m + 3, m))
            if trial % 3 == 1:
                B[:, -1] = B[:, 0]
            if trial % 3 == 2:
                B[:, 1] = B[:, 0] * (1 + 1e-9)
            M = np.zeros((6, 6))
            M[:m, :m] = B.T @ B
            rhs = np.zeros(6)
            rhs[:m] = rng.standard_normal(m)
            sol = np.zeros(6)
            sol[:m] = R.ref_cod_solve(M[:m, :m], rhs[:m])
            Ms.append(M), rhss.append(rhs), sols.append(sol), ranks.append([m, R.ref_cod_rank(M[:m, :m])])
    np.savez_compressed(os.path.join(HERE, "cod.npz"), M=np.array(Ms), rhs=np.array(rhss), sol=np.array(sols), m_rank=np.array(ranks))
    # Anderson streams (variant H incl. reset/replace; variant X)
    out = {}
    for tag, (m, n, ne) in {"a": (5, 600, 600), "b": (3, 517, 300), "c": (1, 64, 64)}.items():
        rng = np.random.default_rng(7)
        Aop = rng.standard_normal((32, 32)) * 0.08
        u0 = rng.standard_normal(n)
        r = R.RefAndersonH(m, n, ne)
        r.init(u0)
        u, G, U = u0.copy(), [], []
        for it in range(10):
            if it == 6:
                r.reset(u)
            if it == 8:
                u = u * 0.5
                r.replace(u)
            g = 0.5 * u + 1.0
            g[:32] = Aop @ u[:32] + 1.0
            u = r.compute(g)
            G.append(g), U.append(u)
        out["H%s_u0" % tag], out["H%s_G" % tag], out["H%s_U" % tag], out["H%s_dims" % tag] = u0, np.array(G), np.array(U), np.array([m, n, ne])
    rng = np.random.default_rng(9)
    Aop = rng.standard_normal((40, 40)) * 0.08
    u0 = rng.standard_normal(200)
    r = R.RefAndersonX()
    r.init(4, 200, u0)
    u, G, U = u0.copy(), [], []
    for it in range(9):
        if it == 5:
            u = u * 0.9
            r.replace(u)
        g = 0.3 * u + 2.0
        g[:40] = Aop @ u[:40] - 1.0
        u = r.compute(g)
        G.append(g), U.append(u)
    out["X_u0"], out["X_G"], out["X_U"], out["X_dims"] = u0, np.array(G), np.array(U), np.array([4, 200, 200])
    np.savez_compressed(os.path.join(HERE, "anderson.npz"), **out)
    print("golden written to", HERE)


if __name__ == "__main__":
    main()
