"""Developer tool: factor the cfg4 (or given) beam system and run a few device LDLT applies
(for ncu launch lists of the triangular-solve kernels)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import aa_admm_b200 as A
cx, cy, cz = [int(a) for a in sys.argv[1:4]]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
s, scene = A.make_beam_solver(cx, cy, cz, iters=1)
print(s.ldlt_stats())
import ctypes as C
L = A.cuda_lib()
f = C.c_void_p(s.H.aaadmm_host_solver_device_factor(s.h))
n = s.info()["n_free"]
b = np.random.default_rng(0).standard_normal(3 * n)
x = np.zeros_like(b)
for r in range(reps):
    t0 = time.perf_counter()
    L.aaadmm_ldlt_solve(f, b.ctypes.data_as(A.c_dp), x.ctypes.data_as(A.c_dp))
    print("solve wall ms", 1e3 * (time.perf_counter() - t0))

if os.environ.get("AAADMM_LDLT_TRACE"):
    os.makedirs("gpurun_out", exist_ok=True)
    L.aaadmm_ldlt_dump_trace.argtypes = [C.c_void_p, C.c_char_p]
    print("trace rc", L.aaadmm_ldlt_dump_trace(f, b"gpurun_out/ldlt_trace.csv"))
