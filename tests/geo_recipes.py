"""The constraint recipes of the reference's Geometry applications restated without OpenMesh:
Geometry/PlanarityOpt.cpp:147-246 (planar quads) and Geometry/WireMeshOpt.cpp:253-289 (wire mesh).
Used by the parity tests and by tests/golden/make_golden_geo.py."""
import numpy as np


def read_obj(path):
    V, F = [], []
    with open(path) as f:
        for line in f:
            if line.startswith("v "):
                V.append([float(t) for t in line.split()[1:4]])
            elif line.startswith("f "):
                F.append([int(t.split("/")[0]) - 1 for t in line.split()[1:]])
    return np.array(V, np.float64), F


def _halfedges(faces):
    he = {}
    for fi, f in enumerate(faces):
        n = len(f)
        for k in range(n):
            he[(f[k], f[(k + 1) % n])] = (fi, k)
    return he


def planarity_recipe(solver, P, faces, Vref, Fref, closeness=1.0, rel_lap=0.1):
    n = len(P)
    solver.add_ref_surface(n, closeness, Vref, Fref)
    he = _halfedges(faces)
    out = [[] for _ in range(n)]
    for (a, b) in he:
        out[a].append(b)
    boundary_v = np.zeros(n, bool)
    for (a, b) in he:
        if (b, a) not in he:
            boundary_v[a] = boundary_v[b] = True
    # vertices that only appear as the head of boundary halfedges still have them listed by the tail loop above
    nbrs_all = [set() for _ in range(n)]
    for (a, b) in he:
        nbrs_all[a].add(b)
        nbrs_all[b].add(a)
    for v in range(n):
        if not nbrs_all[v]:
            continue
        if not boundary_v[v]:
            # one-ring in rotation order: the halfedge before (v->a) in its face ends at v; its tail is next
            start = out[v][0]
            ring, cur = [], start
            for _ in range(len(out[v]) + 1):
                ring.append(cur)
                fi, k = he[(v, cur)]
                f = faces[fi]
                prev = f[(k - 1) % len(f)]   # halfedge prev->v precedes v->cur in the face
                cur = prev
                if cur == start:
                    break
            vhs = [v] + ring
            if len(vhs) == 5:
                solver.add_relative_uniform_laplacian([vhs[0], vhs[1], vhs[3]], rel_lap, P)
                solver.add_relative_uniform_laplacian([vhs[0], vhs[2], vhs[4]], rel_lap, P)
            else:
                solver.add_relative_uniform_laplacian(vhs, rel_lap, P)
        else:
            vhs, fhs = [v], []
            for b in sorted(nbrs_all[v]):
                e_boundary = ((v, b) not in he) or ((b, v) not in he)
                if e_boundary:
                    vhs.append(b)
                    fhs.append(he[(v, b)][0] if (v, b) in he else he[(b, v)][0])
            if len(fhs) == 2 and fhs[0] != fhs[1]:
                solver.add_relative_uniform_laplacian(vhs, rel_lap, P)
    for f in faces:
        if len(f) > 3:
            solver.add_plane(f, 1.0)


def wiremesh_recipe(solver, P, quads, edges, Vref, Fref, edge_length, closeness=1.0,
                    amin=np.pi * 0.25, amax=np.pi * 0.75):
    solver.add_ref_surface(len(P), closeness, Vref, Fref)
    for q in quads:
        for i in range(4):
            solver.add_angle(q[i], q[(i + 1) % 4], q[(i + 3) % 4], 1.0, amin, amax)
    for a, b in edges:
        solver.add_edge(a, b, 1.0, edge_length)
