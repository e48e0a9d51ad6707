"""CPU suite, part 1: the oracle. The C restatement (oracle/port) is pinned against the golden
vectors produced by the unmodified reference (tests/golden/make_golden.py) and, when oracle/_ref is
present, against the reference itself on fresh seeded inputs."""
import os

import numpy as np
import pytest

from oracle import refbind as R
from scenes import assert_iterations_to_tolerance, beam_arrays

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name))


def test_port_prox_and_gradient_vs_golden():
    g = load("tet_element.npz")
    assert np.abs(R.port_tet_prox(g["F"]) - g["prox"]).max() < 1e-15
    assert np.abs(R.port_tet_F_minus_UVt(g["F"]) - g["fmuvt"]).max() < 1e-15


def test_port_cod_vs_golden():
    g = load("cod.npz")
    for M, rhs, sol, (m, rank) in zip(g["M"], g["rhs"], g["sol"], g["m_rank"]):
        x, rk = R.port_cod_solve(M[:m, :m], rhs[:m])
        assert rk == rank
        assert np.abs(x - sol[:m]).max() <= 1e-9 * max(1.0, np.abs(sol).max())


def test_port_anderson_vs_golden():
    g = load("anderson.npz")
    for tag in ("a", "b", "c"):
        m, n, ne = g["H%s_dims" % tag]
        a = R.PortAnderson(int(m), int(n), int(ne))
        a.init(g["H%s_u0" % tag])
        u = g["H%s_u0" % tag].copy()
        for it, (gi, ui) in enumerate(zip(g["H%s_G" % tag], g["H%s_U" % tag])):
            if it == 6:
                a.reset(u)
            if it == 8:
                u = u * 0.5
                a.replace(u)
            u = a.compute(gi)
            assert np.abs(u - ui).max() <= 1e-9 * np.abs(ui).max(), (tag, it)
            u = ui.copy()  # follow the reference stream so that later iterations stay comparable
    m, n, _ = g["X_dims"]
    a = R.PortAnderson(int(m), int(n))
    a.init(g["X_u0"])
    u = g["X_u0"].copy()
    for it, (gi, ui) in enumerate(zip(g["X_G"], g["X_U"])):
        if it == 5:
            u = u * 0.9
            a.replace(u)
        u = a.compute(gi)
        assert np.abs(u - ui).max() <= 1e-9 * np.abs(ui).max(), it
        u = ui.copy()


def _run_port(A, variant, dims, m, accel, frames=2, n_beams=1):
    scene = beam_arrays(A, *dims, n_beams=n_beams)
    verts, tets, masses, pidx, ppts, pside = scene.arrays()
    s = R.PortSolver(variant)
    s.add_tetmesh(verts, tets, masses)
    dt = 1.0 / 30.0
    s.set_pins(pidx, scene.stretch(dt))
    s.initialize(dt, 100, -9.8, max(m, 1), accel, 1.0)
    H, X = [], []
    for _ in range(frames):
        s.set_pins(pidx, scene.stretch(dt))
        H.append(s.step())
        X.append(s.x())
    return H, X


@pytest.mark.parametrize("name,variant", [("hard_beam_12x3x3_m5", "hard"), ("hard_beam_12x3x3_noacc", "hard"),
                                          ("hard_beam_8x2x2_m3", "hard"), ("hard_3beams_6x2x2_m5", "hard"),
                                          ("xzu_beam_12x3x3_m5", "xzu"), ("xzu_beam_12x3x3_noacc", "xzu"),
                                          ("xzu_beam_8x2x2_m3", "xzu")])
def test_port_step_vs_golden_trajectory(A, name, variant):
    g = load(name + ".npz")
    H, X = _run_port(A, variant, tuple(int(d) for d in g["dims"]), int(g["m"]), bool(g["accel"]),
                     n_beams=int(g["n_beams"]))
    for f in range(2):
        rows = int(g["rows"][f])
        n = min(rows, len(H[f]))
        rel = np.abs(H[f][:n, 2] - g["comb"][f][:n]) / g["comb"][f][:n]
        floor = np.abs(H[f][:n, 2] - g["comb"][f][:n]) / g["comb"][f][0]
        assert rel[:8].max() < 1e-9
        assert floor.max() < 1e-9
        assert_iterations_to_tolerance(H[f][:, 2], g["comb"][f][:rows], (name, f))
        if not g["accel"]:
            assert len(H[f]) == rows
        assert np.abs(X[f] - g["x"][f]).max() / np.abs(g["x"][f]).max() < 1e-6


def test_port_vs_compiled_reference_fresh_inputs(ref):
    rng = np.random.default_rng(123)
    F = np.eye(3).reshape(1, 9) + 0.5 * rng.standard_normal((2000, 9))
    assert np.abs(R.port_tet_prox(F) - ref.ref_tet_prox(F)).max() < 1e-15
    for m in (2, 4, 6):
        B = rng.standard_normal((m + 2, m))
        M, rhs = B.T @ B, rng.standard_normal(m)
        x, rank = R.port_cod_solve(M, rhs)
        assert rank == ref.ref_cod_rank(M)
        assert np.abs(x - ref.ref_cod_solve(M, rhs)).max() < 1e-9 * max(1.0, np.abs(x).max())


# ---- C restatement of the Geometry path and of the triangle / collision terms vs the compiled reference -----
@pytest.fixture(scope="module")
def refgeo():
    if not R.have_ref_geo():
        pytest.skip("oracle/_ref/libref_geo.so not present")
    return R


def test_port_geo_projections_and_closest_point_vs_reference(refgeo):
    rng = np.random.default_rng(2)
    for kind, k, prm in ((0, 4, (0, 0, 0, 0)), (0, 6, (0, 0, 0, 0)), (1, 1, (0.7, 0, 0, 0)),
                         (2, 2, (0.6, 2.3, np.cos(0.6), np.cos(2.3)))):
        cols = rng.standard_normal((500, k, 3))
        if kind == 0:
            cols -= cols.mean(axis=1, keepdims=True)
        a = R.port_geo_project(kind, cols, prm)
        b = refgeo.ref_geo_project(kind, cols, *((0.7, 0.0) if kind == 1 else ((0.6, 2.3) if kind == 2 else ())))
        assert np.abs(a - b).max() < 1e-11, kind
    from geo_scenes import ref_surface
    V, F = ref_surface(9, 7)
    Q = rng.uniform(-1, 9, (400, 3))
    cr = refgeo.ref_geo_closest_points(V, F, Q)
    cr = cr[0] if isinstance(cr, tuple) else cr
    assert np.abs(R.port_geo_closest_points(V, F, Q) - cr).max() < 1e-12


@pytest.mark.parametrize("use_alm", [True, False])
@pytest.mark.parametrize("kind,rho", [("planarity", 1e5), ("wiremesh", 1e3)])
@pytest.mark.parametrize("m", [0, 4])
def test_port_geo_solve_vs_reference(refgeo, use_alm, kind, rho, m):
    """ALMGeometrySolver / GeometrySolver loops of the C restatement against the unmodified classes."""
    from geo_scenes import build_planarity, build_wiremesh, ref_surface, wavy_grid
    nx, ny = 9, 7
    P, quads, vid = wavy_grid(nx, ny)
    V, F = ref_surface(nx, ny)
    build = build_planarity if kind == "planarity" else build_wiremesh
    p = R.PortGeometrySolver(use_alm)
    build(p, P, quads, vid, V, F)
    p.setup(len(P), rho)
    hp, xp = p.solve(P, 40, m)
    r = refgeo.RefGeometrySolver(use_alm)
    build(r, P, quads, vid, V, F)
    r.setup(len(P), rho)
    hr, xr = r.solve(P, 40, m)
    assert len(hp) == len(hr) == 40
    rel = np.abs(hp - hr) / hr
    floor = np.abs(hp - hr) / hr[0]
    assert rel[:6].max() < 1e-8
    assert floor.max() < (1e-9 if m == 0 else 1e-5)
    assert np.abs(xp - xr).max() / np.abs(xr).max() < 1e-6


def test_port_tri_and_collision_prox_vs_reference(ref):
    rng = np.random.default_rng(4)
    n = 2000
    Q = np.linalg.qr(rng.standard_normal((n, 3, 3)))[0][:, :, :2]
    F = Q @ (np.eye(2) + 0.6 * rng.standard_normal((n, 2, 2)))
    F6 = np.concatenate([F[:, :, 0], F[:, :, 1]], axis=1)
    for variant in ("hard", "xzu"):
        for lim in ((-100.0, 100.0), (0.9, 1.1)):
            a = R.port_tri_prox(F6, variant, *lim)
            b = ref.ref_tri_prox(F6, variant, *lim)
            assert np.abs(a - b).max() < 1e-11 * max(1.0, np.abs(b).max())
    types = [0, 1, 2, 3, 4]
    prm = [[-0.8, 0, 0, 0, 0, 0, 0], [0.2, -0.5, 0.1, 0.3, 1.0, -0.2, 0], [0.5, 0.2, -0.3, 0, 0, 0, 0.6],
           [-0.6, -0.4, 0.5, 0, 0, 0, 0.5], [1.0, 0.8, 0.0, 0, 0, 0, 0.35]]
    pts = rng.uniform(-1.5, 1.5, (5000, 3))
    assert np.abs(R.port_collision_prox(types, prm, pts) - ref.ref_collision_prox(types, prm, pts)).max() < 1e-14


def test_port_hyper_prox_vs_reference(ref):
    """C restatement of the per-tet L-BFGS prox (Neo-Hookean, StVK) against the unmodified reference classes."""
    import ctypes as C
    L = ref._load("libref_xzu.so")
    dp = C.POINTER(C.c_double)
    L.ref_xzu_tet_prox_hyper.argtypes = [C.c_int, dp, C.c_double, C.c_double, dp, dp, C.c_int]
    verts = np.array([[0, 0, 0], [0.08, 0, 0], [0, 0.09, 0], [0, 0, 0.085]], float).reshape(-1)
    E, nu = 1e7, 0.399
    mu, lam = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
    _, vol, _ = ref.ref_tet_constants(verts.reshape(4, 3), E, nu)
    rng = np.random.default_rng(0)
    F = np.eye(3).reshape(1, 9) + 0.15 * rng.standard_normal((1000, 9))
    for mat in (1, 2):
        zr, gr = F.copy(), np.zeros_like(F)
        assert L.ref_xzu_tet_prox_hyper(mat, verts.ctypes.data_as(dp), E, nu, zr.ctypes.data_as(dp), gr.ctypes.data_as(dp), len(F)) == 0
        zp, gp = R.port_tet_prox_hyper(mat, mu, lam, vol, F)
        d = np.abs(zr - zp).max(axis=1)
        # the stopping rules make the iteration count round-off dependent: a few blocks stop one step apart
        assert np.median(d) < 1e-15 and d.max() < 1e-7
        assert np.abs(gr - gp).max() <= 1e-13 * np.abs(gr).max()


def test_numpy_beam_scene_vs_golden_reference_and_product(A):
    """oracle/ref_scene.py (the scene builder of the CPU arms of bench.py: no product library loaded there) is bitwise
    equal to the reference's own generator (golden vectors; oracle/_ref when present) and to the product's host builder,
    pins and stretched pin targets included."""
    from oracle.ref_scene import RefBeamScene, make_beam
    g = load("beam_scene.npz")
    for dims in [(12, 3, 3), (5, 2, 4), (7, 7, 1)]:
        k = "%dx%dx%d" % dims
        v, t, m = make_beam(*dims, 1.75)
        assert np.array_equal(v, g["verts_" + k]) and np.array_equal(t, g["tets_" + k]) and np.array_equal(m, g["masses_" + k])
        if R.have_ref():
            rv, rt, rm = R.ref_make_beam(*dims, 1.75)
            assert np.array_equal(v, rv) and np.array_equal(t, rt) and np.array_equal(m, rm)
    for dims, shifts in (((6, 5, 4), (0.0,)), ((4, 2, 3), (1.75, 0.0, -1.75))):
        a, b = A.BeamScene(), RefBeamScene()
        for sh in shifts:
            a.add(*dims, sh)
            b.add(*dims, sh)
        for u, w in zip(a.arrays(), b.arrays()):
            assert u.dtype == w.dtype and np.array_equal(u, w)
        for _ in range(3):
            assert np.array_equal(a.stretch(1.0 / 30.0), b.stretch(1.0 / 30.0))
