"""Developer tool (GPU box, oracle/_ref): where the product and the reference part on cfg 2 (PlanarityOpt costa2k).
(1) plane projections of the real faces (initial and converged mesh) product vs reference class, (2) closest points of the
real points on the real reference surface, (3) residual histories product vs reference solver class on the same recipe with
and without Anderson acceleration."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import aa_admm_b200 as A  # noqa: E402
from geo_recipes import planarity_recipe  # noqa: E402
from oracle import refbind  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "geo_costa2k.npz"))
faces = [[int(v) for v in f if v >= 0] for f in g["faces"]]
for name, X in (("initial mesh", g["P"]), ("converged mesh (golden solution)", g["solution"])):
    for k in sorted(set(len(f) for f in faces)):
        idx = np.array([f for f in faces if len(f) == k])
        Af = X[idx]
        Af = Af - Af.mean(axis=1, keepdims=True)
        got = A.geo_project(0, Af)
        exp = refbind.ref_geo_project(0, Af)
        scale = np.abs(Af).max()
        print("plane projection, %s, %d faces of %d corners: max |product - reference| / scale = %.2e (out-of-plane part of the input / scale: max %.2e)"
              % (name, len(idx), k, np.abs(got - exp).max() / scale, np.abs(Af - exp).max() / scale))
C, tri = A.geo_closest_points(g["Vref"], g["Fref"], g["P"])
Cr, Ir, dr = refbind.ref_geo_closest_points(g["Vref"], g["Fref"], g["P"])
print("closest points of the %d mesh points: max |C - C_ref| = %.2e, other triangle picked for %d points (max point difference among them %.2e)"
      % (len(C), np.abs(C - Cr).max(), int((tri != Ir).sum()), np.abs(C - Cr)[tri != Ir].max(initial=0.0)))
for m in (0, 5):
    s = A.GeometrySolver()
    planarity_recipe(s, g["P"], faces, g["Vref"], g["Fref"])
    s.setup(len(g["P"]), 1e5)
    hg, xg = s.solve(g["P"], 40, m)
    r = refbind.RefGeometrySolver(True)
    planarity_recipe(r, g["P"], faces, g["Vref"], g["Fref"])
    r.setup(len(g["P"]), 1e5)
    hr, xr = r.solve(g["P"], 40, m)
    n = min(len(hg), len(hr))
    rel = np.abs(hg[:n] - hr[:n]) / hr[:n]
    print("m = %d: iterations %d / %d, rel. difference of the residual per iteration:" % (m, len(hg), len(hr)))
    print("   " + " ".join("%.1e" % v for v in rel[:24]))
    print("   rejects product", s.info()["rejects"])

# cfg 3 (WireMeshOpt MaleTorso, 230,400 points): the same comparison of the histories, fewer iterations (the reference
# needs about 0.5 s per iteration)
p3 = os.path.join(ROOT, "tests", "golden_large", "geo_maletorso.npz")
if os.path.exists(p3):
    from geo_recipes import wiremesh_recipe  # noqa: E402
    g = np.load(p3)
    C, tri = A.geo_closest_points(g["Vref"], g["Fref"], g["P"])
    Cr, Ir, dr = refbind.ref_geo_closest_points(g["Vref"], g["Fref"], g["P"])
    print("cfg 3 closest points of the %d mesh points: max |C - C_ref| = %.2e, other triangle picked for %d points (max point difference among them %.2e)"
          % (len(C), np.abs(C - Cr).max(), int((tri != Ir).sum()), np.abs(C - Cr)[tri != Ir].max(initial=0.0)))
    for m, iters in ((0, 10), (5, 14)):
        s = A.GeometrySolver()
        wiremesh_recipe(s, g["P"], g["quads"], g["edges"], g["Vref"], g["Fref"], float(g["edge_length"]))
        s.setup(len(g["P"]), 1e3)
        hg, xg = s.solve(g["P"], iters, m)
        r = refbind.RefGeometrySolver(True)
        wiremesh_recipe(r, g["P"], g["quads"], g["edges"], g["Vref"], g["Fref"], float(g["edge_length"]))
        r.setup(len(g["P"]), 1e3)
        hr, xr = r.solve(g["P"], iters, m)
        n = min(len(hg), len(hr))
        rel = np.abs(hg[:n] - hr[:n]) / hr[:n]
        print("cfg 3, m = %d: iterations %d / %d, rel. difference of the residual per iteration:" % (m, len(hg), len(hr)))
        print("   " + " ".join("%.1e" % v for v in rel[:24]))
        print("   rejects product", s.info()["rejects"], " final positions: max |x - x_ref| / max |x_ref| = %.2e" % (np.abs(xg - xr).max() / np.abs(xr).max()))
