"""Small driver for compute-sanitizer (one tool per gpurun call, B200_PROFILING.md): one frame of the hard_zxu loop with
Anderson acceleration on a 12x3x3 beam (device-side numeric factorisation, blocked inverse, both sweep kernels with their
ticket / arrival-counter protocol, local step, Anderson passes, graph WHILE node off so that every launch is checked), an
incremental re-initialisation with another material, and one small Geometry solve."""
import os
import sys

os.environ.setdefault("AAADMM_NO_GRAPH", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import aa_admm_b200 as A  # noqa: E402
from scenes import beam_arrays  # noqa: E402

scene = beam_arrays(A, 12, 3, 3)
verts, tets, masses, pidx, ppts, pside = scene.arrays()
s = A.Solver()
s.add_tetmesh(verts, tets, masses, 1e7, 0.399, 0)
dt = 1.0 / 30.0
s.set_pins(pidx, scene.stretch(dt))
s.initialize(dt, 12, -9.8, 5, True, 1.0)
s.set_pins(pidx, scene.stretch(dt))
h = s.step()
print("frame 1:", len(h), "iterations, residual", h[0, 1], "->", h[-1, 1])
s.set_material(3e6, 0.33)
s.initialize(dt, 12, -9.8, 5, True, 1.0)
assert s.was_incremental()
s.set_pins(pidx, scene.stretch(dt))
h = s.step()
print("frame 2 (another material, incremental):", len(h), "iterations, residual", h[0, 1], "->", h[-1, 1])
from test_gpu_geometry import wavy_grid, ref_surface, build_wiremesh  # noqa: E402
P, quads, vid = wavy_grid(7, 5)
V, F = ref_surface(7, 5)
g = A.GeometrySolver()
build_wiremesh(g, P, quads, vid, V, F)
g.setup(len(P), 1e3)
hist, x = g.solve(P, 8, 3)
print("geometry:", len(hist), "iterations, residual", hist[0], "->", hist[-1])
