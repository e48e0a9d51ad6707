"""Synthetic quad-mesh problems with the constraint recipes of the reference apps:
planarity (Geometry/PlanarityOpt.cpp:147-246) and wire mesh (Geometry/WireMeshOpt.cpp:253-289)."""
import numpy as np


def wavy_grid(nx, ny, amp=0.15, seed=0):
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.arange(nx, dtype=float), np.arange(ny, dtype=float), indexing="ij")
    z = amp * nx * 0.1 * (np.sin(0.6 * xs) * np.cos(0.45 * ys)) + 0.02 * rng.standard_normal(xs.shape)
    P = np.stack([xs, ys, z], -1).reshape(-1, 3)
    vid = np.arange(nx * ny).reshape(nx, ny)
    quads = np.stack([vid[:-1, :-1], vid[1:, :-1], vid[1:, 1:], vid[:-1, 1:]], -1).reshape(-1, 4)
    return P, quads, vid


def ref_surface(nx, ny, amp=0.15, sub=2):
    """A finer triangle mesh of the smooth part of the same height field."""
    mx, my = (nx - 1) * sub + 1, (ny - 1) * sub + 1
    xs, ys = np.meshgrid(np.linspace(0, nx - 1, mx), np.linspace(0, ny - 1, my), indexing="ij")
    z = amp * nx * 0.1 * (np.sin(0.6 * xs) * np.cos(0.45 * ys))
    V = np.stack([xs, ys, z], -1).reshape(-1, 3)
    vid = np.arange(mx * my).reshape(mx, my)
    a, b, c, d = vid[:-1, :-1].ravel(), vid[1:, :-1].ravel(), vid[1:, 1:].ravel(), vid[:-1, 1:].ravel()
    F = np.concatenate([np.stack([a, b, c], -1), np.stack([a, c, d], -1)], 0)
    return V, F.astype(np.int32)


def build_planarity(solver, P, quads, vid, V, F, closeness=1.0, rel_lap=0.1):
    n = len(P)
    solver.add_ref_surface(n, closeness, V, F)
    nx, ny = vid.shape
    for i in range(nx):
        for j in range(ny):
            interior = 0 < i < nx - 1 and 0 < j < ny - 1
            if interior:  # valence 4: two opposite pairs
                solver.add_relative_uniform_laplacian([vid[i, j], vid[i - 1, j], vid[i + 1, j]], rel_lap, P)
                solver.add_relative_uniform_laplacian([vid[i, j], vid[i, j - 1], vid[i, j + 1]], rel_lap, P)
            else:
                nb = []
                if (i in (0, nx - 1)) and 0 < j < ny - 1:
                    nb = [vid[i, j - 1], vid[i, j + 1]]
                if (j in (0, ny - 1)) and 0 < i < nx - 1:
                    nb = [vid[i - 1, j], vid[i + 1, j]]
                if nb:
                    solver.add_relative_uniform_laplacian([vid[i, j]] + nb, rel_lap, P)
    for q in quads:
        solver.add_plane(q, 1.0)


def build_wiremesh(solver, P, quads, vid, V, F, closeness=1.0, amin=np.pi * 0.25, amax=np.pi * 0.75):
    n = len(P)
    solver.add_ref_surface(n, closeness, V, F)
    for q in quads:
        for i in range(4):
            solver.add_angle(q[i], q[(i + 1) % 4], q[(i + 3) % 4], 1.0, amin, amax)
    nx, ny = vid.shape
    edges = [(vid[i, j], vid[i + 1, j]) for i in range(nx - 1) for j in range(ny)] + \
            [(vid[i, j], vid[i, j + 1]) for i in range(nx) for j in range(ny - 1)]
    L = float(np.mean([np.linalg.norm(P[a] - P[b]) for a, b in edges]))
    for a, b in edges:
        solver.add_edge(a, b, 1.0, L)
