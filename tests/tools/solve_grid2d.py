"""Developer tool: time the device LDLT apply on a 2-D grid (surface-mesh-like fill) and optionally dump the task trace."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import scipy.sparse as sp
import aa_admm_b200 as A
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 480
n = nx * nx
idx = np.arange(n).reshape(nx, nx)
rng = np.random.default_rng(0)
r, c, v = [], [], []
diag = rng.uniform(0.1, 1.0, n)
for off in [(1, 0), (0, 1), (1, 1)]:
    a = idx[: nx - off[0], : nx - off[1]].ravel(); b = idx[off[0]:, off[1]:].ravel()
    w = rng.uniform(0.5, 2.0, a.size)
    r += [a, b]; c += [b, a]; v += [-w, -w]
    np.add.at(diag, a, w); np.add.at(diag, b, w)
r.append(np.arange(n)); c.append(np.arange(n)); v.append(diag)
Am = sp.csc_matrix((np.concatenate(v), (np.concatenate(r), np.concatenate(c))), shape=(n, n))
L = sp.tril(Am, format="csc")
coords = np.stack(np.meshgrid(np.arange(nx), np.arange(nx), indexing="ij"), -1).reshape(-1, 2).astype(float)
coords = np.concatenate([coords, np.zeros((n, 1))], axis=1)
hf = A.HostFactor(n, L.indptr, L.indices, L.data, coords, leaf_size=64)
Lp, Li, Lx, D, perm = hf.arrays()
dev = A.Ldlt(n, Lp, Li, Lx, D, perm, 3)
b = rng.standard_normal(3 * n)
import ctypes as C
lib = A.cuda_lib()
d_b = dev  # host-path solve includes copies; time several and report the minimum
ts = []
for _ in range(6):
    t0 = time.perf_counter(); x = dev.solve(b); ts.append(1e3 * (time.perf_counter() - t0))
res = Am @ x.reshape(n, 3) - b.reshape(n, 3)
print("n", n, "nnz(L)", Lp[-1], "solve wall ms (incl. copies) min %.3f" % min(ts), "residual %.2e" % np.abs(res).max(), dev.stats())
if os.environ.get("AAADMM_LDLT_TRACE"):
    os.makedirs("gpurun_out", exist_ok=True)
    lib.aaadmm_ldlt_dump_trace.argtypes = [C.c_void_p, C.c_char_p]
    print("trace rc", lib.aaadmm_ldlt_dump_trace(dev.h, b"gpurun_out/ldlt_trace.csv"))
