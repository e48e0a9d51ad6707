"""CPU suite, part 2: host-side logic of the product (no compute calls: there is no GPU here) and
the C-ABI surface."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def test_beam_scene_matches_reference_generator_bitwise(A):
    g = np.load(os.path.join(G, "beam_scene.npz"))
    for dims in [(12, 3, 3), (5, 2, 4), (7, 7, 1)]:
        k = "%dx%dx%d" % dims
        v, t, m, pidx, ppts, pside = A.BeamScene().add(*dims, 1.75).arrays()
        assert np.array_equal(v, g["verts_" + k])
        assert np.array_equal(t, g["tets_" + k])
        assert np.array_equal(m, g["masses_" + k])
        lo, hi = v[:, 0].min(), v[:, 0].max()
        assert set(pidx) == set(np.nonzero((v[:, 0] < lo + np.float32(1e-2)) | (v[:, 0] > hi - np.float32(1e-2)))[0])


def test_beam_stretch_moves_pins_like_beams_cpp(A):
    sc = A.BeamScene().add(4, 2, 2, 0.0)
    _, _, _, pidx, p0, side = sc.arrays()
    dt = 1.0 / 30.0
    p1 = sc.stretch(dt)
    assert np.allclose(p1[:, 0] - p0[:, 0], np.where(side == 0, -dt, dt), rtol=0, atol=1e-15)
    assert np.array_equal(p1[:, 1:], p0[:, 1:])


def test_tri_constants_match_reference(A):
    """TriEnergyTerm constructor (rest pose, area, weight) against the compiled reference class."""
    from oracle import refbind
    if not refbind.have_ref():
        pytest.skip("oracle/_ref not present")
    H = A.host_lib()
    rng = np.random.default_rng(21)
    for _ in range(200):
        v = rng.standard_normal((3, 3))
        rp, area, w = np.zeros(4), C.c_double(), C.c_double()
        rc = H.aaadmm_host_tri_constants(v.ctypes.data_as(A.c_dp), 3e6, 0.33, rp.ctypes.data_as(A.c_dp), C.byref(area), C.byref(w))
        rpr, ar, wr = refbind.ref_tri_constants(v, 3e6, 0.33)
        assert rc == 0
        assert np.abs(rp.reshape(2, 2).T - rpr).max() <= 1e-13 * np.abs(rpr).max()
        assert abs(area.value - ar) <= 1e-14 * ar and abs(w.value - wr) <= 1e-14 * wr


def test_tet_constants_match_reference_bitwise(A):
    g = np.load(os.path.join(G, "tet_element.npz"))
    H = A.host_lib()
    for tv, c in zip(g["tet_verts"], g["tet_consts"]):
        binv, vol, w = np.zeros(9), C.c_double(), C.c_double()
        r12 = np.ascontiguousarray(tv.reshape(-1))
        assert H.aaadmm_host_tet_constants(r12.ctypes.data_as(A.c_dp), 1e7, 0.399, binv.ctypes.data_as(A.c_dp),
                                           C.byref(vol), C.byref(w)) == 0
        assert w.value == c[0] and vol.value == c[1]
        assert np.array_equal(binv, c[2:])
    inv = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 1], [0, 1, 0.0]]).reshape(-1)  # inverted rest tet
    assert H.aaadmm_host_tet_constants(inv.ctypes.data_as(A.c_dp), 1e7, 0.399, binv.ctypes.data_as(A.c_dp),
                                       C.byref(vol), C.byref(w)) != 0


def test_factor_cache_round_trip(A, tmp_path):
    """ldlt_save / ldlt_load: the cached factor comes back bit-identical, and only for the matrix it belongs to."""
    rng = np.random.default_rng(5)
    Afull, coords, Ap, Ai, Ax = _grid_spd(6, 5, 4, rng)
    n = Afull.shape[0]
    Ap, Ai, Ax = np.array(Ap, np.int64), np.array(Ai, np.int32), np.array(Ax, np.float64)
    hf = A.HostFactor(n, Ap, Ai, Ax, coords, leaf_size=8)
    H = A.host_lib()
    H.aaadmm_host_factor_save.argtypes = [C.c_void_p, C.c_int, A.c_lp, A.c_ip, A.c_dp, C.c_char_p]
    H.aaadmm_host_factor_load.argtypes = [C.c_int, A.c_lp, A.c_ip, A.c_dp, C.c_char_p]
    H.aaadmm_host_factor_load.restype = C.c_void_p
    path = str(tmp_path / "factor.bin").encode()
    assert H.aaadmm_host_factor_save(hf.h, n, Ap.ctypes.data_as(A.c_lp), Ai.ctypes.data_as(A.c_ip), Ax.ctypes.data_as(A.c_dp), path) == 0
    h2 = H.aaadmm_host_factor_load(n, Ap.ctypes.data_as(A.c_lp), Ai.ctypes.data_as(A.c_ip), Ax.ctypes.data_as(A.c_dp), path)
    assert h2
    g = A.HostFactor.__new__(A.HostFactor)
    g.H, g.n, g.h = H, n, C.c_void_p(h2)
    for a, b in zip(hf.arrays(), g.arrays()):
        assert np.array_equal(a, b)
    Ax2 = Ax.copy()
    Ax2[3] *= 1.0000001  # another matrix: the cache must not be used
    assert not H.aaadmm_host_factor_load(n, Ap.ctypes.data_as(A.c_lp), Ai.ctypes.data_as(A.c_ip), Ax2.ctypes.data_as(A.c_dp), path)
    # a file whose header is intact but whose arrays are damaged (row index out of range, broken permutation,
    # truncated) is a cache miss, not an out-of-bounds index later on
    good = open(path.decode(), "rb").read()
    nnz = int(np.frombuffer(good[20:28], np.int64)[0])
    off_perm, off_li = 28, 28 + 4 * n + 8 * (n + 1)
    for lo, hi, val in ((off_li, off_li + 4, np.int32(n + 5).tobytes()), (off_perm, off_perm + 4, good[off_perm + 4:off_perm + 8])):
        open(path.decode(), "wb").write(good[:lo] + val + good[hi:])
        assert not H.aaadmm_host_factor_load(n, Ap.ctypes.data_as(A.c_lp), Ai.ctypes.data_as(A.c_ip), Ax.ctypes.data_as(A.c_dp), path)
    open(path.decode(), "wb").write(good[:-16])
    assert not H.aaadmm_host_factor_load(n, Ap.ctypes.data_as(A.c_lp), Ai.ctypes.data_as(A.c_ip), Ax.ctypes.data_as(A.c_dp), path)
    open(path.decode(), "wb").write(good)
    assert H.aaadmm_host_factor_load(n, Ap.ctypes.data_as(A.c_lp), Ai.ctypes.data_as(A.c_ip), Ax.ctypes.data_as(A.c_dp), path)
    assert nnz > 0
    open(path.decode(), "r+b").write(b"XXXX")  # damaged header
    assert not H.aaadmm_host_factor_load(n, Ap.ctypes.data_as(A.c_lp), Ai.ctypes.data_as(A.c_ip), Ax.ctypes.data_as(A.c_dp), path)


def _grid_spd(nx, ny, nz, rng):
    n = nx * ny * nz
    idx = np.arange(n).reshape(nx, ny, nz)
    Afull = np.zeros((n, n))
    for d in range(3):
        a = np.take(idx, np.arange(idx.shape[d] - 1), axis=d).ravel()
        b = np.take(idx, np.arange(1, idx.shape[d]), axis=d).ravel()
        w = rng.uniform(0.5, 2.0, a.size)
        Afull[a, b] -= w
        Afull[b, a] -= w
        Afull[a, a] += w
        Afull[b, b] += w
    Afull += np.diag(rng.uniform(0.1, 1.0, n))
    coords = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij"), -1).reshape(-1, 3).astype(float)
    L = np.tril(Afull)
    Ap, Ai, Ax = [0], [], []
    for j in range(n):
        r = np.nonzero(L[:, j])[0]
        Ai += list(r)
        Ax += list(L[r, j])
        Ap.append(len(Ai))
    return Afull, coords, Ap, Ai, Ax


@pytest.mark.parametrize("with_coords", [True, False])
def test_host_nested_dissection_ldlt(A, with_coords):
    rng = np.random.default_rng(5)
    Afull, coords, Ap, Ai, Ax = _grid_spd(8, 7, 6, rng)
    n = Afull.shape[0]
    hf = A.HostFactor(n, Ap, Ai, Ax, coords if with_coords else None, leaf_size=12)
    Lp, Li, Lx, D, perm = hf.arrays()
    assert sorted(perm) == list(range(n))
    # L D L^T == P A P^T
    Ld = np.eye(n)
    for j in range(n):
        Ld[Li[Lp[j]:Lp[j + 1]], j] = Lx[Lp[j]:Lp[j + 1]]
        assert np.all(np.diff(Li[Lp[j]:Lp[j + 1]]) > 0) and np.all(Li[Lp[j]:Lp[j + 1]] > j)
    assert np.abs(Ld @ np.diag(D) @ Ld.T - Afull[np.ix_(perm, perm)]).max() < 1e-11
    for nrhs in (1, 3):
        b = rng.standard_normal(n * nrhs)
        x = hf.solve(b, nrhs)
        assert np.abs(x.reshape(n, nrhs) - np.linalg.solve(Afull, b.reshape(n, nrhs))).max() < 1e-10


def test_host_ordering_is_a_postorder_of_the_elimination_tree(A):
    """host/sparse_ldlt.cpp: etree_postorder - every subtree of the elimination tree of the returned factor is a run of
    consecutive columns (what the device-side front builder needs to merge a small subtree into one front), with and
    without coordinates (geometric / BFS bisection)."""
    rng = np.random.default_rng(7)
    for dims, with_coords in (((9, 8, 7), True), ((23, 19, 1), False), ((23, 19, 1), True)):
        Afull, coords, Ap, Ai, Ax = _grid_spd(*dims, rng)
        n = Afull.shape[0]
        hf = A.HostFactor(n, Ap, Ai, Ax, coords if with_coords else None, leaf_size=16)
        Lp, Li, Lx, D, perm = hf.arrays()
        assert sorted(perm) == list(range(n))
        parent = np.array([Li[Lp[j]] if Lp[j + 1] > Lp[j] else -1 for j in range(n)])
        assert np.all((parent == -1) | (parent > np.arange(n)))
        size = np.ones(n, int)
        lo = np.arange(n)
        for j in range(n):  # children precede their parents
            if parent[j] >= 0:
                size[parent[j]] += size[j]
                lo[parent[j]] = min(lo[parent[j]], lo[j])
        assert np.all(np.arange(n) - lo + 1 == size), dims
        b = rng.standard_normal(n)
        assert np.abs(hf.solve(b, 1) - np.linalg.solve(Afull, b)).max() < 1e-10


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(aaadmm_[a-z0-9_]+)\s*\(", txt)))


def test_c_abi_exports_every_declared_symbol(A):
    for header, lib in (("aaadmm.h", A.LIB_CUDA), ("aaadmm_host.h", A.LIB_HOST)):
        out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
        exported = set(l.split()[-1] for l in out.splitlines() if " T " in l)
        names = _declared(header)
        assert len(names) > 10
        missing = [n for n in names if n not in exported]
        assert not missing, missing


def test_no_cpu_fallback_without_a_gpu(A):
    if A.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(A.AaadmmError, match="no CUDA device"):
        A.AndersonAcceleration(3, 10, 10)
    with pytest.raises(A.AaadmmError, match="no CUDA device"):
        A.tet_prox_linear(np.eye(3).reshape(1, 9))
    with pytest.raises(A.AaadmmError, match="no CUDA device"):
        A.make_beam_solver(4, 2, 2)


def test_product_never_imports_the_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "aa-admm_b200")):
        for f in files:
            # build.py only BUILDS the checker (make -C oracle); nothing in the package loads or calls it
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h")) and f != "build.py":
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "oracle" not in txt and "refbind" not in txt, os.path.join(d, f)


def _ensemble_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from aa_admm_b200 import ensemble as E
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = E.scenes_of_rank(7, rank, world)
    recs = [E.make_record(s, 10 + s, s % 2, 1e-3 * s, 1e-6 * s, 2.0 * s, 3.0 * s, rank) for s in mine]
    table = E.gather_records(np.array(recs).reshape(-1, 8), dist)
    q.put((rank, table, E.scenes_by_cost(table, rank, world)))
    dist.destroy_process_group()


def test_ensemble_sharding_and_gather_gloo_world2(A):
    import socket
    import torch.multiprocessing as mp
    from aa_admm_b200 import ensemble as E
    # static, balanced scene -> GPU map: a partition of the ensemble in which every GPU gets each stiffness group (s // 8)
    # and each Poisson ratio (s mod 8) equally often
    for world in (1, 2, 4, 8):
        parts = [E.scenes_of_rank(64, r, world) for r in range(world)]
        assert sorted(s for p in parts for s in p) == list(range(64)) and all(len(p) == 64 // world for p in parts)
        for p in parts:
            assert sorted(np.bincount([s // 8 for s in p], minlength=8)) == [8 // world] * 8
            assert sorted(np.bincount([s % 8 for s in p], minlength=8)) == [8 // world] * 8
    e0, n0 = E.scene_material(0)
    e63, n63 = E.scene_material(63)
    assert e0 == 1e6 and abs(n0 - 0.30) < 1e-15 and abs(e63 - 1e8) < 1e-3 and abs(n63 - 0.44) < 1e-12
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_ensemble_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = [q.get(timeout=120) for _ in ps]
    res = {g[0]: g[1] for g in got}
    nxt = {g[0]: g[2] for g in got}
    for p in ps:
        p.join(timeout=60)
    # the cost-balanced map of the next pass: a partition, the same on every rank, heavier scenes first, loads within one
    # scene of each other (loop_ms of scene s is 2 s in the worker: 0, 2, ..., 12 -> 6+5+... split 22 / 20)
    assert sorted(nxt[0] + nxt[1]) == list(range(7)) and nxt[0][0] == 6 and nxt[1][0] == 5
    assert abs(sum(2.0 * s for s in nxt[0]) - sum(2.0 * s for s in nxt[1])) <= 2.0 * 6
    assert nxt[0] == sorted(nxt[0], reverse=True) and nxt[1] == sorted(nxt[1], reverse=True)
    for r in (0, 1):
        t = res[r]
        assert t.shape == (7, 8)
        assert list(t[:, 0]) == list(range(7))
        assert list(t[:, 7]) == [(s + s // 8) % 2 for s in range(7)]
        assert np.allclose(t[:, 1], 10 + np.arange(7))


def write_elenode(path, verts, tets):
    """TetGen ASCII pair (0-based), float32 vertices with 9 significant digits."""
    with open(path + ".ele", "w") as f:
        f.write("%d 4 0\n" % len(tets))
        for i, t in enumerate(tets):
            f.write("%d %d %d %d %d\n" % (i, *t))
    with open(path + ".node", "w") as f:
        f.write("%d 3 0 0\n" % len(verts))
        for i, v in enumerate(verts):
            f.write("%d %.9g %.9g %.9g\n" % (i, *v))


def _build_sample(tmp_path, name="beams"):
    exe = str(tmp_path / name)
    cmd = ["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "aa-admm_b200", "host"),
           os.path.join(ROOT, "samples", name + ".cpp"), "-L" + os.path.join(ROOT, "aa-admm_b200"), "-laaadmm_host", "-laaadmm_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "aa-admm_b200"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_reference_style_sample_compiles_against_the_dropin_classes(A, tmp_path):
    """samples/beams.cpp is the reference's beam sample with only the include paths changed: it must build and link
    against the host mirror (source-level drop-in). Without a GPU it stops in initialize()."""
    exe = _build_sample(tmp_path)
    r = subprocess.run([exe, "-it", "5", "-frames", "1"], capture_output=True, text=True)
    if A.device_count() <= 0:
        assert r.returncode != 0 and ("CUDA" in (r.stderr + r.stdout) or "aaadmm" in (r.stderr + r.stdout))


def test_windyflag_sample_compiles_against_the_dropin_classes(A, tmp_path):
    """samples/windyflag.cpp: TriEnergyTerm + strain limits + WindForce + obstacle / set_collisions through the
    reference's own calls; builds against the host mirror. Without a GPU it stops in initialize()."""
    exe = _build_sample(tmp_path, "windyflag")
    r = subprocess.run([exe, "-it", "5", "-frames", "1", "-n", "6", "-sphere"], capture_output=True, text=True)
    if A.device_count() <= 0:
        assert r.returncode != 0 and ("CUDA" in (r.stderr + r.stdout) or "aaadmm" in (r.stderr + r.stdout))


def test_plinko_sample_compiles_and_reads_its_mesh(A, tmp_path):
    """samples/plinko.cpp: mcl::meshio::load_elenode + binding::add_tetmesh + add_obstacle + set_collisions through the
    reference's own calls. Without a GPU it reads the mesh, builds the operators and stops in initialize()."""
    exe = _build_sample(tmp_path, "plinko")
    v, t, m, _, _, _ = A.BeamScene().add(4, 2, 2, 0.0).arrays()
    write_elenode(str(tmp_path / "beam"), v, t)
    r = subprocess.run([exe, "-mesh", str(tmp_path / "beam"), "-frames", "1"], capture_output=True, text=True)
    if A.device_count() <= 0:
        assert r.returncode != 0 and ("CUDA" in (r.stderr + r.stdout) or "aaadmm" in (r.stderr + r.stdout))
    r = subprocess.run([exe, "-mesh", str(tmp_path / "nothing"), "-frames", "1"], capture_output=True, text=True)
    assert r.returncode != 0 and "Could not load" in r.stderr


@pytest.mark.parametrize("with_beam", [False, True])
def test_system_matrix_with_triangle_terms_matches_reference(A, with_beam):
    """Host setup of a cloth (TriEnergyTerm) scene, alone and mixed with a tet beam: the scalar matrix Ahat built
    matrix-free (host/tet_system.cpp) against the reference's solver_termA = M + rho dt^2 D^T W^2 D
    (hard/src/Solver.cpp:462-467), which must be Ahat (x) I3 with the pinned vertices eliminated."""
    from oracle import refbind
    if not refbind.have_ref():
        pytest.skip("oracle/_ref (compiled reference) not present")
    from scenes import cloth_arrays
    verts, tris, masses, pins = cloth_arrays(5)
    tets = np.zeros((0, 4), np.int32)
    r = refbind.RefSolver("hard")
    r.add_trimesh(verts, tris, masses, 1e5, 0.3)
    allv, allm, pidx = verts, masses, list(pins)
    if with_beam:
        bv, bt, bm, bp, _, _ = A.BeamScene().add(4, 2, 2, -1.75).arrays()
        r.add_tetmesh(bv, bt, bm, 1e5, 0.3, 0)
        tets = bt + len(verts)
        allv = np.concatenate([verts, bv])
        allm = np.concatenate([masses, bm])
        pidx += [int(p) + len(verts) for p in bp]
    pidx = np.array(pidx, np.int32)
    r.set_pins(pidx, allv[pidx].astype(np.float64))
    dt, rho = 1.0 / 30.0, 2.5
    r.initialize(dt, 5, -9.8, 5, True, rho)
    n, rp, ci, v = r.termA()
    Aref = np.zeros((n, n))
    for i in range(n):
        Aref[i, ci[rp[i]:rp[i + 1]]] = v[rp[i]:rp[i + 1]]
    Ahat, d2v = A.host_system_matrix(allv, tets, tris, allm, pidx, rho * dt * dt, 1e5, 0.3)
    nf = Ahat.shape[0]
    assert n == 3 * nf
    # free vertices keep their ascending order in both
    free = [i for i in range(len(allv)) if i not in set(pidx.tolist())]
    assert list(d2v[:nf]) == free
    scale = np.abs(Aref).max()
    for c in range(3):
        assert np.abs(Aref[c::3, c::3] - Ahat).max() < 1e-12 * scale
    for c in range(3):
        for e in range(3):
            if c != e:
                assert np.abs(Aref[c::3, e::3]).max() == 0.0


def test_wind_force_matches_reference(A):
    """Host WindForce::project against the reference class (src/ExplicitForce.cpp:47-105, one thread)."""
    from oracle import refbind
    if not refbind.have_ref():
        pytest.skip("oracle/_ref (compiled reference) not present")
    from scenes import cloth_arrays
    verts, tris, _, _ = cloth_arrays(7)
    rng = np.random.default_rng(5)
    x = verts.astype(np.float64) + 0.02 * rng.standard_normal(verts.shape)
    v = rng.standard_normal(verts.shape)
    d = np.array([25.0, 0.0, 5.0])
    got = A.wind_project(tris, d, 1.0 / 30.0, x, v)
    want = refbind.ref_wind_project(tris, d, 1.0 / 30.0, x, v)
    assert np.abs(got - v).max() > 1e-3
    assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()


def test_system_matrix_with_collision_terms_matches_reference(A):
    """Collision energy terms (Solver::set_collisions, one 3-row term per vertex with the weight of the reference's
    Collision ctor) add rho dt^2 w^2 to the diagonal: host matrix against the reference's solver_termA, no pins."""
    from oracle import refbind
    if not refbind.have_ref():
        pytest.skip("oracle/_ref (compiled reference) not present")
    bv, bt, bm, _, _, _ = A.BeamScene().add(4, 2, 2, 0.0).arrays()
    col = np.arange(0, len(bv), 2, dtype=np.int32)
    r = refbind.RefSolver("hard")
    r.add_tetmesh(bv, bt, bm, 1e6, 0.399, 0)
    r.add_obstacle(0, (-0.45, 0, 0, 0, 0, 0, 0))
    r.set_collisions(col)
    dt, rho = 1.0 / 30.0, 1.5
    r.initialize(dt, 5, -9.8, 5, True, rho)
    n, rp, ci, v = r.termA()
    Aref = np.zeros((n, n))
    for i in range(n):
        Aref[i, ci[rp[i]:rp[i + 1]]] = v[rp[i]:rp[i + 1]]
    Ahat, d2v = A.host_system_matrix(bv, bt, np.zeros((0, 3), np.int32), bm, [], rho * dt * dt, 1e6, 0.399, collisions=col)
    assert n == 3 * Ahat.shape[0] == 3 * len(bv)
    scale = np.abs(Aref).max()
    for c in range(3):
        assert np.abs(Aref[c::3, c::3] - Ahat).max() < 1e-12 * scale


# ---- mesh files: mcl::meshio readers + the masses of binding::add_tetmesh / add_trimesh ---------
def test_mesh_readers_vs_golden(A, tmp_path):
    """Hand-written .ele/.node (1-based with shuffled rows; 0-based) and .obj files parsed by the host mirror against
    the arrays the reference's reader produced for the same files (bitwise)."""
    from mesh_files import write_all
    g = np.load(os.path.join(ROOT, "tests", "golden", "mesh_io.npz"))
    for name, (path, kind) in write_all(str(tmp_path)).items():
        v, e, m = A.load_mesh(path, kind)
        assert np.array_equal(v, g[name + "_verts"]), name
        assert np.array_equal(e, g[name + "_elems"]), name
        assert np.array_equal(m, g[name + "_masses"]), name
    with pytest.raises(A.AaadmmError):
        A.load_mesh(os.path.join(str(tmp_path), "missing"), "elenode")


def test_mesh_writers_round_trip(A, tmp_path):
    from mesh_files import write_all
    for name, (path, kind) in write_all(str(tmp_path)).items():
        out = os.path.join(str(tmp_path), name + "_copy" + (".obj" if kind == "obj" else ""))
        v, e, m = A.load_mesh(path, kind, save_as=out)
        v2, e2, m2 = A.load_mesh(out, kind)
        assert np.array_equal(e, e2)
        if kind == "elenode":
            assert np.array_equal(v, v2) and np.array_equal(m, m2)  # 9 significant digits restore float32
        else:
            assert np.allclose(v, v2, rtol=1e-5)  # save_obj writes 6 significant digits like the reference


def test_mesh_readers_vs_reference_on_its_sample_data(A):
    """The reference's own sample meshes (plinko horse / box, windyflag cloth, pole) through both readers."""
    from oracle import refbind
    data = "/root/reference/admm_anderson_hard_zxu/samples/data"
    if not (refbind.have_ref() and os.path.isdir(data)):
        pytest.skip("reference sample data not present")
    for name, kind in (("horse759", "elenode"), ("box768", "elenode"), ("cloth.obj", "obj"), ("cloth_small.obj", "obj"),
                       ("pole.obj", "obj")):
        got = A.load_mesh(os.path.join(data, name), kind)
        want = refbind.ref_load_mesh(os.path.join(data, name), kind)
        for a, b in zip(got, want):
            assert a.shape == b.shape and np.array_equal(a, b), name
        assert got[1].shape[0] > 100


def test_triangle_and_collision_error_paths(A):
    """Error behaviour of the reference's constructors and setters, kept by the host mirror: bad strain limits
    (TriEnergyTerm.cpp:32-33), collision terms on a pinned vertex (not supported here), bad collision index."""
    from scenes import cloth_arrays
    verts, tris, masses, pins = cloth_arrays(3)
    s = A.Solver()
    with pytest.raises(A.AaadmmError, match="Strain limit min"):
        s.add_trimesh(verts, tris, masses, 1e5, 0.3, 1.5, 100.0)
    s = A.Solver()
    with pytest.raises(A.AaadmmError, match="Strain limit max"):
        s.add_trimesh(verts, tris, masses, 1e5, 0.3, -100.0, 0.5)
    # a collision term on a pinned vertex / out of range is refused by the operator setup
    with pytest.raises(A.AaadmmError, match="pinned"):
        A.host_system_matrix(verts, np.zeros((0, 4), np.int32), tris, masses, pins, 1e-3, 1e5, 0.3, collisions=[int(pins[0])])
    with pytest.raises(A.AaadmmError, match="out of range"):
        A.host_system_matrix(verts, np.zeros((0, 4), np.int32), tris, masses, pins, 1e-3, 1e5, 0.3, collisions=[10 ** 6])
    # degenerate (zero-area) rest triangle: weight 0 is refused like EnergyTerm::get_reduction does
    flat = verts.copy()
    flat[tris[0]] = flat[tris[0][0]]
    with pytest.raises(A.AaadmmError):
        A.host_system_matrix(flat, np.zeros((0, 4), np.int32), tris, masses, pins, 1e-3, 1e5, 0.3)


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the reference's own CPU path on a bounded sample): exactly one line on stdout, JSON,
    with the keys the driver reads; everything the reference itself prints goes to stderr."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-dims", "8", "6", "6"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, OMP_NUM_THREADS="1"))  # as torchrun exports it: the arm must override it
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.split("\n") if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "admm_anderson_iterations_per_sec_1M_tets" and d["unit"] == "iterations/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # a measurement at the stated size, on all host cores, with none of the product's libraries in the process
    assert d["config"]["sample_tets"] == 8 * 6 * 6 * 5 and "8x6x6" in d["config"]["workload"]
    assert d["config"]["same_config_as_gpu_arm"] is False
    if d["cpu_baseline"]["kind"] == "reference":
        assert d["cpu_baseline"]["cores"] == os.cpu_count()
    assert d["native_so_loaded"] and all(l.startswith("oracle/") for l in d["native_so_loaded"]), d["native_so_loaded"]


def test_material_update_of_the_system_matrix_is_bit_identical(A):
    """update_tet_system_materials (the values-only path a re-initialisation of a Solver takes for the next member of a
    material sweep) gives bit for bit the matrix a fresh build_tet_system gives for that material: tets + triangles +
    collision terms, another material AND another rho dt^2."""
    from scenes import cloth_arrays
    H = A.host_lib()
    H.aaadmm_host_system_update.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
    verts, tris, masses, pins = cloth_arrays(5)
    bv, bt, bm, bp, _, _ = A.BeamScene().add(5, 2, 3, -1.75).arrays()
    tets = (bt + len(verts)).astype(np.int32)
    allv = np.concatenate([verts, bv]).astype(np.float32)
    allm = np.concatenate([masses, bm]).astype(np.float32)
    pidx = np.array(list(pins) + [int(p) + len(verts) for p in bp], np.int32)
    col = np.array([len(verts) + 7, len(verts) + 9], np.int32)
    col = np.array([c for c in col if c not in set(pidx.tolist())], np.int32)

    def build(youngs, poisson, rho_dt2):
        h = H.aaadmm_host_system_new(allv.ctypes.data_as(A.c_fp), len(allv), tets.ctypes.data_as(A.c_ip), len(tets),
                                     tris.ctypes.data_as(A.c_ip), len(tris), allm.ctypes.data_as(A.c_fp), youngs, poisson,
                                     pidx.ctypes.data_as(A.c_ip), len(pidx), rho_dt2, col.ctypes.data_as(A.c_ip), len(col))
        assert h
        return C.c_void_p(h)

    def values(h):
        nf, nnz = C.c_int(0), C.c_int64(0)
        H.aaadmm_host_system_counts(h, C.byref(nf), C.byref(nnz))
        Ap, Ai, Ax = np.zeros(nf.value + 1, np.int64), np.zeros(nnz.value, np.int32), np.zeros(nnz.value)
        d2v = np.zeros(len(allv), np.int32)
        H.aaadmm_host_system_copy(h, Ap.ctypes.data_as(A.c_lp), Ai.ctypes.data_as(A.c_ip), Ax.ctypes.data_as(A.c_dp),
                                  d2v.ctypes.data_as(A.c_ip))
        return Ap, Ai, Ax

    h0 = build(1e7, 0.399, 1.0 / 900.0)
    for youngs, poisson, rho_dt2 in ((1e6, 0.30, 1.0 / 900.0), (3.7e8, 0.44, 2.5 / 576.0), (1e7, 0.399, 1.0 / 900.0)):
        assert H.aaadmm_host_system_update(h0, youngs, poisson, rho_dt2) == 0
        h1 = build(youngs, poisson, rho_dt2)
        for a, b in zip(values(h0), values(h1)):
            assert np.array_equal(a, b)
        H.aaadmm_host_system_free(h1)
    H.aaadmm_host_system_free(h0)


def _quad_patch_obj(path, nx, ny, closed=False, seed=2):
    """Wavy quad patch (or, closed = True, a quad torus without boundary) as .obj with a/b/c face tokens in some rows."""
    rng = np.random.default_rng(seed)
    if closed:
        V = [[(2 + np.cos(2 * np.pi * j / ny)) * np.cos(2 * np.pi * i / nx), (2 + np.cos(2 * np.pi * j / ny)) * np.sin(2 * np.pi * i / nx),
              np.sin(2 * np.pi * j / ny)] for i in range(nx) for j in range(ny)]
        vid = lambda i, j: (i % nx) * ny + (j % ny)
        F = [[vid(i, j), vid(i + 1, j), vid(i + 1, j + 1), vid(i, j + 1)] for i in range(nx) for j in range(ny)]
    else:
        V = [[i + 0.1 * rng.standard_normal(), j + 0.1 * rng.standard_normal(), 0.3 * np.sin(0.7 * i) * np.cos(0.5 * j)]
             for i in range(nx + 1) for j in range(ny + 1)]
        vid = lambda i, j: i * (ny + 1) + j
        F = [[vid(i, j), vid(i + 1, j), vid(i + 1, j + 1), vid(i, j + 1)] for i in range(nx) for j in range(ny)]
    with open(path, "w") as f:
        f.write("# test mesh\n")
        for v in V:
            f.write("v %.9f %.9f %.9f\n" % tuple(v))
        for k, q in enumerate(F):
            f.write("f " + " ".join(("%d/%d/%d" % (i + 1, i + 1, i + 1)) if k % 3 == 0 else str(i + 1) for i in q) + "\n")
    return np.array([[float("%.9f" % c) for c in v] for v in V]), F   # the coordinates as the file holds them


@pytest.mark.parametrize("closed", [False, True])
def test_geometry_front_end_vs_openmesh(A, tmp_path, closed):
    """host/GeometryApps: the OBJ reader (single-precision coordinates as OpenMesh's), the edge numbering, the
    average edge length and subdivide_and_smooth_mesh against the reference's own code on OpenMesh
    (oracle/_ref/libref_wiremesh.so: MeshTypes.h:147-161, 214-342): connectivity identical, positions to round-off."""
    from oracle import refbind
    if not refbind.have_ref_geo() or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_wiremesh.so")):
        pytest.skip("oracle/_ref (compiled reference) not present")
    path = str(tmp_path / "quads.obj")
    V0, F0 = _quad_patch_obj(path, 7, 5, closed)
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_wiremesh.so"))
    nv, nq, ne, el = C.c_int(), C.c_int(), C.c_int(), C.c_double()
    assert L.ref_app_subdivided(path.encode(), None, None, C.byref(nv), C.byref(nq), C.byref(el)) == 0
    P, Q = np.zeros((nv.value, 3)), np.zeros((nq.value, 4), np.int32)
    L.ref_app_subdivided(path.encode(), P.ctypes.data_as(C.c_void_p), Q.ctypes.data_as(C.c_void_p), C.byref(nv), C.byref(nq), C.byref(el))
    L.ref_app_edges(path.encode(), None, C.byref(ne))
    E = np.zeros((ne.value, 2), np.int32)
    L.ref_app_edges(path.encode(), E.ctypes.data_as(C.c_void_p), C.byref(ne))
    m = A.PolyMesh.load(path)
    V, F, _ = m.arrays()
    assert np.array_equal(V, V0.astype(np.float32).astype(np.float64)) and F == F0
    assert 0.5 * m.counts()["average_edge_length"] == el.value
    s = m.subdivide_and_smooth()
    Vs, Fs, Es = s.arrays()
    assert np.array_equal(np.array(Fs), Q) and np.array_equal(Es, E)
    assert np.abs(Vs - P).max() < 1e-10 * np.abs(P).max()
    # write / read round trip (16 significant digits; the reader keeps single precision like OpenMesh's)
    s.save(tmp_path / "sub.obj")
    V2, F2, E2 = A.PolyMesh.load(tmp_path / "sub.obj").arrays()
    assert F2 == Fs and np.array_equal(E2, Es) and np.array_equal(V2, Vs.astype(np.float32).astype(np.float64))
    # a face that re-uses a directed edge is refused like OpenMesh's add_face does (complex edge)
    with pytest.raises(A.AaadmmError):
        A.PolyMesh.from_arrays(V0, F0 + [F0[0]]).counts()


@pytest.mark.parametrize("name", ["planarity", "wiremesh"])
def test_geometry_samples_compile_and_report_missing_files(A, tmp_path, name):
    exe = _build_sample(tmp_path, name)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage" in r.stdout
    r = subprocess.run([exe, "nothing.obj", "nothing.obj", "nothing.txt", "out.obj"], capture_output=True, text=True)
    assert r.returncode == 1 and "unable to read" in r.stderr


def _build_harness(tmp_path, name):
    exe = str(tmp_path / name)
    cmd = ["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "aa-admm_b200", "host"),
           os.path.join(ROOT, "tests", "tools", name + ".cpp"), "-L" + os.path.join(ROOT, "aa-admm_b200"), "-laaadmm_host", "-laaadmm_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "aa-admm_b200"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_solver_public_members_beyond_the_hot_path(A, tmp_path):
    """surface_inds (binding::add_tetmesh), Eigen-style m_x access, set_pins with a foreign 3-vector type,
    add_dynamic_collider (throws): tests/tools/solver_surface_harness.cpp. With a GPU the same program also checks
    save_matrix and the per-iteration device time stamps (tests/test_gpu_parity.py)."""
    exe = _build_harness(tmp_path, "solver_surface_harness")
    r = subprocess.run([exe, "5", str(tmp_path / "A.mtx")], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "surface_inds" in r.stdout
