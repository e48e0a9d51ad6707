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


def cfg1_meshes(A, dims=(12, 3, 3)):
    """BASELINE configs[0]: three beams (LINEAR / NEOHOOKEAN / STVK) at y = +1.75, 0, -1.75
    (admm_anderson_xzu/samples/Asia2019/beams.cpp:94-160). Returns per-beam arrays + merged pins."""
    beams = []
    scene = A.BeamScene()
    off = 0
    for shift in (1.75, 0.0, -1.75):
        one = A.BeamScene().add(*dims, shift)
        v, t, m, _, _, _ = one.arrays()
        beams.append((v, t, m))
        scene.add(*dims, shift)
        off += len(v)
    return beams, scene


def run_cfg1(make_solver, A, frames=1, dims=(12, 3, 3), m=5, accel=True, ordering=1, iters=100):
    """make_solver() -> object with add_tetmesh/set_pins/initialize/step/x (product or reference)."""
    beams, scene = cfg1_meshes(A, dims)
    s = make_solver()
    for (v, t, mm), mat in zip(beams, (0, 1, 2)):
        s.add_tetmesh(v, t, mm, 1e7, 0.399, mat)
    _, _, _, pidx, _, _ = scene.arrays()
    dt = 1.0 / 30.0
    s.set_pins(pidx, scene.stretch(dt))
    if ordering is None:
        s.initialize(dt, iters, -9.8, max(m, 1), accel, 1.0)
    else:
        s.initialize(dt, iters, -9.8, max(m, 1), accel, 1.0, ordering)
    hist, xs = [], []
    for _ in range(frames):
        s.set_pins(pidx, scene.stretch(dt))
        hist.append(s.step())
        xs.append(s.x())
    return hist, xs
