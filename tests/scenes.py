"""Shared scene drivers for the parity tests: the same beam scene fed to the product
(aa_admm_b200.Solver) and to the compiled reference (oracle.refbind.RefSolver)."""
import numpy as np


def beam_arrays(A, cx, cy, cz, n_beams=1):
    scene = A.BeamScene()
    shifts = {1: [0.0], 3: [1.75, 0.0, -1.75]}[n_beams]
    for s in shifts:
        scene.add(cx, cy, cz, s)
    return scene


def run_product(A, scene, frames, dt=1.0 / 30.0, iters=100, m=5, accel=True, penalty=1.0, ordering=0,
                youngs=1e7, poisson=0.399):
    verts, tets, masses, pidx, ppts, pside = scene.arrays()
    s = A.Solver()
    s.add_tetmesh(verts, tets, masses, youngs, poisson, 0)
    s.set_pins(pidx, scene.stretch(dt))
    s.initialize(dt, iters, -9.8, m, accel, penalty, ordering)
    hist, xs = [], []
    for _ in range(frames):
        s.set_pins(pidx, scene.stretch(dt))
        hist.append(s.step())
        xs.append(s.x())
    return s, hist, xs


def run_reference(refbind, A, scene, frames, dt=1.0 / 30.0, iters=100, m=5, accel=True, penalty=1.0,
                  variant="hard", youngs=1e7, poisson=0.399):
    verts, tets, masses, pidx, ppts, pside = scene.arrays()
    r = refbind.RefSolver(variant)
    r.add_tetmesh(verts, tets, masses, youngs, poisson, 0)
    r.set_pins(pidx, scene.stretch(dt))
    r.initialize(dt, iters, -9.8, m, accel, penalty)
    hist, xs = [], []
    for _ in range(frames):
        r.set_pins(pidx, scene.stretch(dt))
        hist.append(r.step())
        xs.append(r.x())
    return r, hist, xs
