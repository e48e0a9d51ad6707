"""Developer tool: per-phase device times of the hard_zxu loop on a large cloth (triangle terms only, optional
collision terms on every free vertex) - the HBM figures of k_tri_update_z_hard / k_tri_update_u_hard / k_pt_*.
    python tests/tools/cloth_bench.py [n_cells=1000] [collisions=0|1]
Prints one JSON line (phases: ms, algorithmic GB, GB/s)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import aa_admm_b200 as A  # noqa: E402


def cloth(n, size=10.0):
    g = np.linspace(0.0, size, n + 1)
    X, Z = np.meshgrid(g, g, indexing="ij")
    rng = np.random.default_rng(1)
    verts = np.stack([X.ravel(), 1.0 + 0.001 * rng.standard_normal(X.size), Z.ravel()], 1).astype(np.float32)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    a = (i * (n + 1) + j).ravel()
    b, c, d = a + (n + 1), a + (n + 1) + 1, a + 1
    tris = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)]).astype(np.int32)
    masses = np.full(len(verts), 0.5 / len(verts), np.float32)
    pins = np.array([0, n], np.int32)
    return verts, tris, masses, pins


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    col = len(sys.argv) > 2 and sys.argv[2] == "1"
    verts, tris, masses, pins = cloth(n)
    s = A.Solver()
    t0 = time.perf_counter()
    s.add_trimesh(verts, tris, masses, 1e5, 0.3, 0.9, 1.1)
    s.set_pins(pins, verts[pins].astype(np.float64))
    if col:
        s.add_obstacle(2, (5.0, 0.2, 5.0, 0, 0, 0, 1.0))
        s.add_obstacle(0, (-0.5, 0, 0, 0, 0, 0, 0))
        free = np.setdiff1d(np.arange(len(verts), dtype=np.int32), pins)
        s.set_collisions(free)
    s.initialize(1.0 / 30.0, 30, -9.8, 5, True, 1.0)
    setup = time.perf_counter() - t0
    for _ in range(2):
        h = s.step()
    info = s.info()
    prof = s.profile(20, 5, True)
    peak = 6537.3
    try:
        peak = json.load(open(os.path.join(os.path.dirname(A.__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = {"workload": "cloth %dx%d cells: %d triangles, %d vertices, collision terms: %s" % (n, n, len(tris), len(verts), col),
           "setup_s": round(setup, 2), "iterations_per_s": info["iter_num"] / (info["loop_ms"] * 1e-3),
           "rejects": info["rejects"], "comb_first_last": [float(h[0, 1]), float(h[-1, 1])], "hbm_peak_GBps": peak,
           "phases": {k: {"ms": round(v["ms"], 4), "algo_GB": round(v["bytes"] / 1e9, 4),
                          "GBps": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1),
                          "frac_of_peak": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6 / peak, 3)}
                      for k, v in prof.items() if v["ms"] > 0}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
