"""Developer tool: per-task trace of the LDL^T applies of cfg 3 (WireMeshOpt MaleTorso through the product's front-end),
dumped after a short solve: AAADMM_LDLT_TRACE=1 python tests/tools/trace_cfg3.py && python tests/tools/trace_critical_path.py"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import aa_admm_b200 as A  # noqa: E402
import bench  # noqa: E402

g = bench._golden("tests/golden_large/geo_maletorso.npz")
tmp = tempfile.mkdtemp()
files = (os.path.join(tmp, "quad.obj"), os.path.join(tmp, "target.obj"))
bench._write_obj(files[0], g["P0"], g["quads0"])
bench._write_obj(files[1], g["Vref"], g["Fref"])
coarse, ref = A.PolyMesh.load(files[0]), A.PolyMesh.load(files[1])
el = 0.5 * coarse.counts()["average_edge_length"]
app = A.GeoApp("wiremesh", coarse.subdivide_and_smooth(), ref, [1e3, 0.25 * np.pi, 0.75 * np.pi, el, 1.0, -1.0])
hist, x, info = app.solve(6, 5, False)
print(app.stats(), info)
H = A.host_lib()
H.aaadmm_host_geoapp_device_factor.restype = C.c_void_p
H.aaadmm_host_geoapp_device_factor.argtypes = [C.c_void_p]
f = C.c_void_p(H.aaadmm_host_geoapp_device_factor(app.h))
L = A.cuda_lib()
L.aaadmm_ldlt_dump_trace.argtypes = [C.c_void_p, C.c_char_p]
os.makedirs("gpurun_out", exist_ok=True)
print("trace rc", L.aaadmm_ldlt_dump_trace(f, b"gpurun_out/ldlt_trace.csv"))
