"""Regenerates tests/golden_large/hard_cfg5_scene_*.npz: one frame of two scenes of BASELINE configs[4] (cfg 5: beam
88x22x22 = 212,960 tets, material sweep of aa_admm_b200.ensemble.scene_material, scenes s = 0 and s = 63) from the
UNMODIFIED reference compiled into oracle/_ref (admm_anderson_hard_zxu Solver::step, m = 5, 100 iterations).
Run in the build container, where /root/reference exists:  python tests/golden/make_golden_large.py
(Eigen's AMD + simplicial LDL^T of the 3n system needs about a minute per scene.)
tests/golden_large/ is git-ignored (size) but travels to the GPU box with the snapshot, like geo_maletorso.npz."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(os.path.dirname(HERE), "golden_large")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import aa_admm_b200 as A  # noqa: E402  (host-side scene builder only)
from aa_admm_b200 import ensemble as E  # noqa: E402
from oracle import refbind as R  # noqa: E402

DIMS = (88, 22, 22)


def scene_case(s, accel=True, m=5, iters=100, fma=False):
    youngs, poisson = E.scene_material(s)
    scene = A.BeamScene().add(*DIMS, 0.0)
    verts, tets, masses, pidx, ppts, pside = scene.arrays()
    r = R.RefSolver("hard", fma=fma)
    r.add_tetmesh(verts, tets, masses, youngs, poisson, 0)
    dt = 1.0 / 30.0
    r.set_pins(pidx, scene.stretch(dt))
    t0 = time.perf_counter()
    r.initialize(dt, iters, -9.8, m, accel, 1.0)
    t1 = time.perf_counter()
    r.set_pins(pidx, scene.stretch(dt))
    h = r.step()
    t2 = time.perf_counter()
    print("scene %d: E=%.4g nu=%.2f setup %.1f s, frame %.1f s, %d rows, %d rejects" %
          (s, youngs, poisson, t1 - t0, t2 - t1, len(h), int(h[:, 3].sum())), flush=True)
    return dict(scene=s, dims=np.array(DIMS), youngs=youngs, poisson=poisson, m=m, accel=int(accel), iters=iters,
                prim=h[:, 1], comb=h[:, 2], rej=h[:, 3], x=r.x())


def main():
    assert R.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    os.makedirs(OUT, exist_ok=True)
    for s in (0, 63):
        path = os.path.join(OUT, "hard_cfg5_scene_%d.npz" % s)
        d = dict(np.load(path)) if (os.path.exists(path) and "--fma-only" in sys.argv) else scene_case(s)
        if R.have_ref_fma():
            # the same frame on the reference's FMA flavour: the reference's own round-off noise floor (never the oracle)
            f = scene_case(s, fma=True)
            d["comb_fma"], d["x_fma_err"] = f["comb"], np.abs(f["x"] - d["x"]).max() / np.abs(d["x"]).max()
        np.savez_compressed(path, **d)


if __name__ == "__main__":
    main()
