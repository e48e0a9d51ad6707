"""Regenerates tests/golden/mesh_io.npz: the synthetic files of tests/mesh_files.py parsed by the UNMODIFIED reference
reader (mcl::meshio through oracle/_ref) with the masses binding::add_tetmesh / add_trimesh give the nodes.
Run in the build container:  python tests/golden/make_golden_meshio.py"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refbind as R  # noqa: E402
from mesh_files import write_all  # noqa: E402

if __name__ == "__main__":
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for name, (path, kind) in write_all(d).items():
            v, e, m = R.ref_load_mesh(path, kind)
            out[name + "_verts"], out[name + "_elems"], out[name + "_masses"] = v, e, m
            print(name, v.shape, e.shape, m.sum())
    np.savez_compressed(os.path.join(HERE, "mesh_io.npz"), **out)
