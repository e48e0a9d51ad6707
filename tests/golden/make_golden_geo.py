"""Golden fixtures of the Geometry configs from the reference's OWN applications (unmodified
PlanarityOpt.cpp / WireMeshOpt.cpp compiled into oracle/_ref, run in-process on the shipped meshes).
cfg 2 (costa2k) is small and committed as tests/golden/geo_costa2k.npz; cfg 3 (MaleTorso, 230k points)
is written to tests/golden_large/ (git-ignored; travels to the GPU box with gpurun)."""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from geo_recipes import read_obj  # noqa: E402

GEO = "/root/reference/Geometry"


def read_obj_as_the_apps_do(path):
    """OpenMesh's OBJ reader parses vertex coordinates as float (Core/IO/reader/OBJReader.cc:294) and the applications'
    meshes then hold them as double: the arrays of the fixtures must be these float-rounded values, or every comparison
    with the applications' histories starts from inputs that differ by 6e-8 relative."""
    V, F = read_obj(path)
    return V.astype(np.float32).astype(np.float64), F


def run_app(lib, args):
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", lib))
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "result"))
        out = os.path.join(d, "out.obj")
        argv = [b"app"] + [a.encode() for a in args] + [out.encode()]
        arr = (C.c_char_p * len(argv))(*argv)
        cwd = os.getcwd()
        os.chdir(d)
        try:
            rc = L.ref_app_main(len(argv), arr)
        finally:
            os.chdir(cwd)
        assert rc == 0
        hist = np.loadtxt(os.path.join(d, "result", "residual-5.txt"))
        V, F = read_obj(out)
    return hist, V, L


def main():
    opts = os.path.join(GEO, "Options.txt")
    poly = os.path.join(GEO, "Geometry_model/PQMeshData/polymesh/costa2k_poly.obj")
    tri = os.path.join(GEO, "Geometry_model/PQMeshData/trimesh/costa2k_tri.obj")
    hist, Vsol, _ = run_app("libref_planarity.so", [poly, tri, opts])
    P, faces = read_obj_as_the_apps_do(poly)
    Vr, Fr = read_obj_as_the_apps_do(tri)
    fl = np.full((len(faces), max(len(f) for f in faces)), -1, np.int32)
    for i, f in enumerate(faces):
        fl[i, :len(f)] = f
    np.savez_compressed(os.path.join(HERE, "geo_costa2k.npz"), P=P, faces=fl, Vref=Vr, Fref=np.array(Fr, np.int32),
                        hist=hist[:, 1], secs=hist[:, 0], solution=Vsol)
    print("cfg2 costa2k:", len(P), "points", len(faces), "faces", "residual", hist[0, 1], "->", hist[-1, 1], "ref secs", hist[-1, 0])
    if "--large" in sys.argv:
        quad = os.path.join(GEO, "Geometry_model/WireMeshData/MaleTorso.obj")
        tgt = os.path.join(GEO, "Geometry_model/WireMeshData/MaleTorso_target.obj")
        hist, Vsol, L = run_app("libref_wiremesh.so", [quad, tgt, opts])
        nv, nq, ne, el = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        L.ref_app_subdivided(quad.encode(), None, None, C.byref(nv), C.byref(nq), C.byref(el))
        P = np.zeros((nv.value, 3))
        Q = np.zeros((nq.value, 4), np.int32)
        L.ref_app_subdivided(quad.encode(), P.ctypes.data_as(C.c_void_p), Q.ctypes.data_as(C.c_void_p), C.byref(nv), C.byref(nq), C.byref(el))
        L.ref_app_edges(quad.encode(), None, C.byref(ne))
        E = np.zeros((ne.value, 2), np.int32)
        L.ref_app_edges(quad.encode(), E.ctypes.data_as(C.c_void_p), C.byref(ne))
        Vr, Fr = read_obj_as_the_apps_do(tgt)
        P0, F0 = read_obj_as_the_apps_do(quad)  # the coarse input mesh itself: the product's own front-end subdivides it in the tests
        os.makedirs(os.path.join(ROOT, "tests", "golden_large"), exist_ok=True)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden_large", "geo_maletorso.npz"), P=P, quads=Q, edges=E, edge_length=el.value,
                            P0=P0, quads0=np.array(F0, np.int32),
                            Vref=Vr, Fref=np.array(Fr, np.int32), hist=hist[:, 1], secs=hist[:, 0], solution=Vsol)
        print("cfg3 MaleTorso:", len(P), "points", len(Q), "quads", len(E), "edges", "residual", hist[0, 1], "->", hist[-1, 1], "ref secs", hist[-1, 0])


if __name__ == "__main__":
    main()
