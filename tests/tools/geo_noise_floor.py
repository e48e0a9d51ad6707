"""Developer tool (CPU, needs /root/reference): the reference's own round-off noise floor on cfg 2 / cfg 3. Runs the FMA
flavour of the UNMODIFIED Geometry applications (oracle/_ref_fma, oracle/Makefile) on the shipped meshes and compares its
residual history with the golden history of the parity flavour (oracle/_ref, -ffp-contract=off; tests/golden/geo_costa2k.npz,
tests/golden_large/geo_maletorso.npz). Prints a markdown table: per-iteration relative difference reference(FMA) vs
reference - the bar any other arithmetic (the GPU) can be held to on these configs."""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
GEO = "/root/reference/Geometry"


def run_app(lib, args):
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref_fma", lib))
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "result"))
        argv = [b"app"] + [a.encode() for a in args] + [os.path.join(d, "out.obj").encode()]
        arr = (C.c_char_p * len(argv))(*argv)
        cwd = os.getcwd()
        os.chdir(d)
        try:
            assert L.ref_app_main(len(argv), arr) == 0
        finally:
            os.chdir(cwd)
        return np.loadtxt(os.path.join(d, "result", "residual-5.txt"))[:, 1]


def report(name, h_fma, h_ref):
    n = min(len(h_fma), len(h_ref))
    rel = np.abs(h_fma[:n] - h_ref[:n]) / h_ref[:n]
    print("### %s\n" % name)
    print("iterations %d / %d; reference(FMA) vs reference, relative difference of the logged residual:\n" % (len(h_fma), len(h_ref)))
    print("| iterations | max rel. difference |")
    print("|---|---:|")
    for a, b in ((0, 2), (0, 8), (0, 10), (0, 20), (0, 50), (0, n)):
        print("| %d - %d | %.2e |" % (a + 1, min(b, n), rel[a:min(b, n)].max()))
    print("\nfinal residual %.6e (FMA) vs %.6e\n" % (h_fma[n - 1], h_ref[n - 1]))


def main():
    opts = os.path.join(GEO, "Options.txt")
    g2 = np.load(os.path.join(ROOT, "tests", "golden", "geo_costa2k.npz"))
    h = run_app("libref_planarity.so", [os.path.join(GEO, "Geometry_model/PQMeshData/polymesh/costa2k_poly.obj"),
                                        os.path.join(GEO, "Geometry_model/PQMeshData/trimesh/costa2k_tri.obj"), opts])
    report("cfg 2: PlanarityOpt costa2k", h, g2["hist"])
    p3 = os.path.join(ROOT, "tests", "golden_large", "geo_maletorso.npz")
    if os.path.exists(p3):
        g3 = np.load(p3)
        h = run_app("libref_wiremesh.so", [os.path.join(GEO, "Geometry_model/WireMeshData/MaleTorso.obj"),
                                           os.path.join(GEO, "Geometry_model/WireMeshData/MaleTorso_target.obj"), opts])
        report("cfg 3: WireMeshOpt MaleTorso", h, g3["hist"])


if __name__ == "__main__":
    main()
