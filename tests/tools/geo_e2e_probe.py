"""Developer tool: where the wall time of one GeoApp.solve (cfg 3) goes: Python wall, C++ wall of aaadmm_geo_solve, device loop."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import aa_admm_b200 as A
import bench
g = bench._golden("tests/golden_large/geo_maletorso.npz")
tmp = tempfile.mkdtemp()
files = (os.path.join(tmp, "quad.obj"), os.path.join(tmp, "target.obj"))
bench._write_obj(files[0], g["P0"], g["quads0"]); bench._write_obj(files[1], g["Vref"], g["Fref"])
coarse, ref = A.PolyMesh.load(files[0]), A.PolyMesh.load(files[1])
el = 0.5 * coarse.counts()["average_edge_length"]
mesh = coarse.subdivide_and_smooth()
app = A.GeoApp("wiremesh", mesh, ref, [1e3, 0.25 * np.pi, 0.75 * np.pi, el, 1.0, -1.0])
for want in (False, True, False, True):
    t0 = time.perf_counter()
    hist, x, info = app.solve(100, 5, want)
    print("want_solution", want, "python wall ms %.1f" % (1e3 * (time.perf_counter() - t0)), "c++ wall of aaadmm_geo_solve ms %.1f" % info["wall_ms"], "device loop ms %.1f" % info["loop_ms"], "iters", len(hist))
