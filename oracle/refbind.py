"""TEST INFRASTRUCTURE ONLY: ctypes bindings for oracle/_ref (the unmodified reference
compiled by oracle/Makefile) and for the C restatement oracle/liboracle_port.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module. The product (aa-admm_b200/) never does.
"""
import ctypes as C
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_FMA_DIR = os.path.join(HERE, "_ref_fma")

c_dp = C.POINTER(C.c_double)
c_fp = C.POINTER(C.c_float)
c_ip = C.POINTER(C.c_int)


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _fp(a):
    return a.ctypes.data_as(c_fp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


def have_ref():
    return all(os.path.exists(os.path.join(REF_DIR, f))
               for f in ("libref_hard.so", "libref_xzu.so", "libref_aa.so"))


_libs = {}


def _load(name, fma=False):
    key = (name, bool(fma))
    if key not in _libs:
        _libs[key] = C.CDLL(os.path.join(REF_FMA_DIR if fma else REF_DIR, name))
    return _libs[key]


def have_ref_fma():
    """The second flavour of the compiled reference (FMA contraction allowed; oracle/Makefile: _ref_fma): noise-floor
    measurements only, never the parity oracle."""
    return os.path.exists(os.path.join(REF_FMA_DIR, "libref_hard.so"))


class RefSolver:
    """The reference admm::Solver (hard_zxu or xzu ordering), driven headless."""

    def __init__(self, variant="hard", workdir=None, fma=False):
        assert variant in ("hard", "xzu")
        self.variant = variant
        self.lib = _load("libref_%s.so" % variant, fma)
        self.p = "ref_%s_" % variant
        self._tmp = None
        if workdir is None:
            self._tmp = tempfile.TemporaryDirectory(prefix="refsolver_")
            workdir = self._tmp.name
        f = self._f
        f("new").restype = C.c_void_p
        f("new").argtypes = [C.c_char_p]
        self.h = C.c_void_p(f("new")(workdir.encode()))
        f("free").argtypes = [C.c_void_p]
        f("add_tetmesh").argtypes = [C.c_void_p, c_fp, C.c_int, c_ip, C.c_int, c_fp, C.c_double, C.c_double, C.c_int]
        f("add_trimesh").argtypes = [C.c_void_p, c_fp, C.c_int, c_ip, C.c_int, c_fp, C.c_double, C.c_double,
                                     C.c_double, C.c_double]
        f("add_wind").argtypes = [C.c_void_p, c_ip, C.c_int, c_dp]
        if variant == "hard":
            f("set_collisions").argtypes = [C.c_void_p, c_ip, C.c_int]
            f("add_obstacle").argtypes = [C.c_void_p, C.c_int, c_dp]
        f("set_threads").argtypes = [C.c_int]
        f("set_pins").argtypes = [C.c_void_p, c_ip, c_dp, C.c_int]
        f("initialize").argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double]
        f("step").argtypes = [C.c_void_p]
        f("step_wall_ms").argtypes = [C.c_void_p]
        f("step_wall_ms").restype = C.c_double
        f("runtime").argtypes = [C.c_void_p, c_dp]
        for n in ("hist_rows", "hist_cols", "n_dof", "termA_rows"):
            f(n).argtypes = [C.c_void_p]
        f("hist_copy").argtypes = [C.c_void_p, c_dp]
        f("get_x").argtypes = [C.c_void_p, c_dp]
        f("get_v").argtypes = [C.c_void_p, c_dp]
        f("termA_nnz").argtypes = [C.c_void_p]
        f("termA_nnz").restype = C.c_long
        f("termA_copy").argtypes = [C.c_void_p, c_ip, c_ip, c_dp]
        f("factor_nnz").argtypes = [C.c_void_p]
        f("factor_nnz").restype = C.c_long
        f("factor_copy").argtypes = [C.c_void_p, c_ip, c_ip, c_dp, c_dp, c_ip, C.c_long]
        f("solve").argtypes = [C.c_void_p, c_dp, c_dp]

    def _f(self, name):
        return getattr(self.lib, self.p + name)

    def __del__(self):
        try:
            if self.h:
                self._f("free")(self.h)
                self.h = None
        except Exception:
            pass

    def add_tetmesh(self, verts, tets, masses, youngs=1e7, poisson=0.399, material=0):
        verts = np.ascontiguousarray(verts, np.float32)
        tets = np.ascontiguousarray(tets, np.int32)
        masses = np.ascontiguousarray(masses, np.float32)
        r = self._f("add_tetmesh")(self.h, _fp(verts), len(verts), _ip(tets), len(tets), _fp(masses),
                                   youngs, poisson, material)
        if r < 0:
            raise RuntimeError("reference add_tetmesh failed")
        return r

    def add_trimesh(self, verts, tris, masses, youngs=1e7, poisson=0.399, limit_min=-100.0, limit_max=100.0):
        verts = np.ascontiguousarray(verts, np.float32)
        tris = np.ascontiguousarray(tris, np.int32)
        masses = np.ascontiguousarray(masses, np.float32)
        r = self._f("add_trimesh")(self.h, _fp(verts), len(verts), _ip(tris), len(tris), _fp(masses),
                                   youngs, poisson, limit_min, limit_max)
        if r < 0:
            raise RuntimeError("reference add_trimesh failed")
        return r

    def set_collisions(self, idx):
        idx = np.ascontiguousarray(idx, np.int32)
        if self._f("set_collisions")(self.h, _ip(idx), len(idx)) != 0:
            raise RuntimeError("reference set_collisions failed")

    def add_obstacle(self, kind, prm7):
        p = np.ascontiguousarray(prm7, np.float64)
        if self._f("add_obstacle")(self.h, int(kind), _dp(p)) != 0:
            raise RuntimeError("reference add_obstacle failed")

    def add_wind(self, tris, direction):
        """WindForce over `tris` (global vertex ids); forces one OpenMP thread (the reference's loop races otherwise)."""
        tris = np.ascontiguousarray(tris, np.int32)
        d = np.ascontiguousarray(direction, np.float64)
        self._f("set_threads")(1)
        self._f("add_wind")(self.h, _ip(tris), len(tris), _dp(d))

    def set_pins(self, idx, pts):
        idx = np.ascontiguousarray(idx, np.int32)
        pts = np.ascontiguousarray(pts, np.float64)
        if self._f("set_pins")(self.h, _ip(idx), _dp(pts), len(idx)) != 0:
            raise RuntimeError("reference set_pins failed")

    def initialize(self, dt=1.0 / 30.0, iters=100, gravity=-9.8, anderson_m=5, accel=True, penalty=1.0):
        r = self._f("initialize")(self.h, dt, iters, gravity, anderson_m, int(bool(accel)), penalty)
        if r != 0:
            raise RuntimeError("reference initialize failed (%d)" % r)

    def step(self):
        """One Solver::step(); returns the logged trajectory, columns
        cumulative_ms, prim_residual, comb_residual[, is_reject]."""
        r = self._f("step")(self.h)
        if r != 0:
            raise RuntimeError("reference step failed (%d)" % r)
        rows, cols = self._f("hist_rows")(self.h), self._f("hist_cols")(self.h)
        out = np.zeros((rows, cols))
        if rows:
            self._f("hist_copy")(self.h, _dp(out))
        return out

    def step_wall_ms(self):
        return self._f("step_wall_ms")(self.h)

    def runtime(self):
        out = np.zeros(4)
        self._f("runtime")(self.h, _dp(out))
        return dict(local_ms=out[0], global_ms=out[1], acceleration_ms=out[2], initialization_ms=out[3])

    def x(self):
        out = np.zeros(self._f("n_dof")(self.h))
        self._f("get_x")(self.h, _dp(out))
        return out

    def v(self):
        out = np.zeros(self._f("n_dof")(self.h))
        self._f("get_v")(self.h, _dp(out))
        return out

    def termA(self):
        n = self._f("termA_rows")(self.h)
        nnz = self._f("termA_nnz")(self.h)
        rp, ci, v = np.zeros(n + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
        self._f("termA_copy")(self.h, _ip(rp), _ip(ci), _dp(v))
        return n, rp, ci, v

    def factor(self):
        """Eigen's LDL^T factor: (n, colptr, rowidx, Lx) strictly lower CSC, D, perm[new]=old."""
        n = self._f("termA_rows")(self.h)
        cap = self._f("factor_nnz")(self.h)
        cp, ri, v = np.zeros(n + 1, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
        D, perm = np.zeros(n), np.zeros(n, np.int32)
        k = self._f("factor_copy")(self.h, _ip(cp), _ip(ri), _dp(v), _dp(D), _ip(perm), cap)
        assert k >= 0
        return n, cp, ri[:k].copy(), v[:k].copy(), D, perm

    def solve(self, b):
        b = np.ascontiguousarray(b, np.float64)
        x = np.zeros_like(b)
        self._f("solve")(self.h, _dp(b), _dp(x))
        return x


def ref_make_beam(cx, cy, cz, y_shift=0.0, density=1522.0):
    """The reference's own make_tet_blocks + centre/scale + weighted_masses."""
    lib = _load("libref_hard.so")
    fn = lib.ref_hard_make_beam
    fn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, c_fp, c_ip, c_fp, c_ip, c_ip]
    nv, nt = C.c_int(0), C.c_int(0)
    fn(cx, cy, cz, y_shift, density, None, None, None, C.byref(nv), C.byref(nt))
    verts = np.zeros((nv.value, 3), np.float32)
    tets = np.zeros((nt.value, 4), np.int32)
    masses = np.zeros(nv.value, np.float32)
    fn(cx, cy, cz, y_shift, density, _fp(verts), _ip(tets), _fp(masses), C.byref(nv), C.byref(nt))
    return verts, tets, masses


def ref_tet_prox(z):
    """TetEnergyTerm::prox on (n,9) column-major 3x3 blocks."""
    lib = _load("libref_hard.so")
    lib.ref_hard_tet_prox.argtypes = [c_dp, C.c_int]
    out = np.ascontiguousarray(z, np.float64).copy()
    lib.ref_hard_tet_prox(_dp(out), out.shape[0])
    return out


def ref_tet_F_minus_UVt(z):
    lib = _load("libref_xzu.so")
    lib.ref_xzu_tet_F_minus_UVt.argtypes = [c_dp, c_dp, C.c_int]
    z = np.ascontiguousarray(z, np.float64)
    out = np.zeros_like(z)
    lib.ref_xzu_tet_F_minus_UVt(_dp(z), _dp(out), z.shape[0])
    return out


def ref_tet_constants(verts4, youngs=1e7, poisson=0.399):
    lib = _load("libref_hard.so")
    lib.ref_hard_tet_constants.argtypes = [c_dp, C.c_double, C.c_double, c_dp, c_dp, c_dp]
    v = np.ascontiguousarray(verts4, np.float64)
    w, vol, binv = C.c_double(), C.c_double(), np.zeros(9)
    if lib.ref_hard_tet_constants(_dp(v), youngs, poisson, C.byref(w), C.byref(vol), _dp(binv)) != 0:
        raise RuntimeError("inverted tet")
    return w.value, vol.value, binv


def ref_load_mesh(path, kind="elenode"):
    """The reference's mcl::meshio reader + weighted_masses (densities of binding::add_tetmesh / add_trimesh)."""
    lib = _load("libref_hard.so")
    k = {"elenode": 0, "obj": 1}[kind]
    f = lib.ref_hard_mesh_load
    f.argtypes = [C.c_char_p, C.c_int, c_ip, c_ip, c_fp, c_ip, c_fp]
    nv, ne = C.c_int(0), C.c_int(0)
    if f(str(path).encode(), k, C.byref(nv), C.byref(ne), None, None, None) != 0:
        raise RuntimeError("reference mesh_load failed for %s" % path)
    verts = np.zeros((nv.value, 3), np.float32)
    elems = np.zeros((ne.value, 4 if k == 0 else 3), np.int32)
    masses = np.zeros(nv.value, np.float32)
    assert f(str(path).encode(), k, C.byref(nv), C.byref(ne), _fp(verts), _ip(elems), _fp(masses)) == 0
    return verts, elems, masses


def ref_wind_project(tris, direction, dt, x, v):
    """Reference WindForce::project (hard_zxu/src/ExplicitForce.cpp:47-105), one thread; returns the new v."""
    lib = _load("libref_hard.so")
    tris = np.ascontiguousarray(tris, np.int32)
    d = np.ascontiguousarray(direction, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    v = np.array(v, np.float64).copy()
    lib.ref_hard_wind_project.argtypes = [c_ip, C.c_int, c_dp, C.c_double, c_dp, c_dp, C.c_int]
    lib.ref_hard_wind_project(_ip(tris), len(tris), _dp(d), dt, _dp(x), _dp(v), x.size // 3)
    return v


def ref_tri_prox(F, variant="hard", limit_min=-100.0, limit_max=100.0):
    """The reference's TriEnergyTerm::prox (hard_zxu or xzu) on (n, 6) column-major 3x2 blocks."""
    lib = _load("libref_hard.so" if variant == "hard" else "libref_xzu.so")
    f = getattr(lib, "ref_%s_tri_prox" % variant)
    f.argtypes = [c_dp, C.c_int, C.c_double, C.c_double]
    f.restype = None
    z = np.ascontiguousarray(F, np.float64).reshape(-1, 6).copy()
    f(_dp(z), len(z), limit_min, limit_max)
    return z


def ref_tri_constants(verts3, youngs=1e7, poisson=0.399, variant="hard"):
    lib = _load("libref_hard.so" if variant == "hard" else "libref_xzu.so")
    f = getattr(lib, "ref_%s_tri_constants" % variant)
    f.argtypes = [c_dp, C.c_double, C.c_double, c_dp, c_dp, c_dp]
    v = np.ascontiguousarray(verts3, np.float64)
    rp, area, w = np.zeros(4), C.c_double(), C.c_double()
    if f(_dp(v), youngs, poisson, _dp(rp), C.byref(area), C.byref(w)) != 0:
        raise RuntimeError("inverted triangle")
    return rp.reshape(2, 2).T.copy(), area.value, w.value  # rest_pose is stored column-major


def ref_collision_prox(types, prm, pts):
    lib = _load("libref_hard.so")
    lib.ref_hard_collision_prox.argtypes = [C.c_int, c_ip, c_dp, c_dp, C.c_int]
    types = np.ascontiguousarray(types, np.int32)
    prm = np.ascontiguousarray(prm, np.float64).reshape(-1, 7)
    z = np.ascontiguousarray(pts, np.float64).reshape(-1, 3).copy()
    if lib.ref_hard_collision_prox(len(types), _ip(types), _dp(prm), _dp(z), len(z)) != 0:
        raise RuntimeError("unknown passive object")
    return z


def ref_spring_prox(pts, pins, active):
    lib = _load("libref_hard.so")
    lib.ref_hard_spring_prox.argtypes = [c_dp, c_dp, c_ip, C.c_int]
    lib.ref_hard_spring_prox.restype = None
    z = np.ascontiguousarray(pts, np.float64).reshape(-1, 3).copy()
    lib.ref_hard_spring_prox(_dp(z), _dp(np.ascontiguousarray(pins, np.float64)), _ip(np.ascontiguousarray(active, np.int32)), len(z))
    return z


class RefAndersonH:
    """hard/src/AndersonAcceleration.h (== Geometry/AndersonAcceleration.h)."""

    def __init__(self, m, total_dim, effective_dim):
        self.lib = _load("libref_aa.so")
        L = self.lib
        L.ref_aa_h_new.restype = C.c_void_p
        L.ref_aa_h_new.argtypes = [C.c_int, C.c_int, C.c_int]
        L.ref_aa_h_free.argtypes = [C.c_void_p]
        for n in ("init", "reset", "replace"):
            getattr(L, "ref_aa_h_" + n).argtypes = [C.c_void_p, c_dp, C.c_int]
        L.ref_aa_h_compute.argtypes = [C.c_void_p, c_dp, c_dp, C.c_int]
        self.n = total_dim
        self.h = C.c_void_p(L.ref_aa_h_new(m, total_dim, effective_dim))

    def __del__(self):
        try:
            self.lib.ref_aa_h_free(self.h)
        except Exception:
            pass

    def init(self, u):
        u = np.ascontiguousarray(u, np.float64)
        self.lib.ref_aa_h_init(self.h, _dp(u), u.size)

    def reset(self, u):
        u = np.ascontiguousarray(u, np.float64)
        self.lib.ref_aa_h_reset(self.h, _dp(u), u.size)

    def replace(self, u):
        u = np.ascontiguousarray(u, np.float64)
        self.lib.ref_aa_h_replace(self.h, _dp(u), u.size)

    def compute(self, g):
        g = np.ascontiguousarray(g, np.float64)
        out = np.zeros_like(g)
        self.lib.ref_aa_h_compute(self.h, _dp(g), _dp(out), g.size)
        return out


class RefAndersonX:
    """xzu/src/AndersonAcceleration.h."""

    def __init__(self):
        self.lib = _load("libref_aa.so")
        L = self.lib
        L.ref_aa_x_new.restype = C.c_void_p
        L.ref_aa_x_free.argtypes = [C.c_void_p]
        L.ref_aa_x_init.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp]
        L.ref_aa_x_replace.argtypes = [C.c_void_p, c_dp, C.c_int]
        L.ref_aa_x_compute.argtypes = [C.c_void_p, c_dp, c_dp, C.c_int]
        self.h = C.c_void_p(L.ref_aa_x_new())

    def __del__(self):
        try:
            self.lib.ref_aa_x_free(self.h)
        except Exception:
            pass

    def init(self, m, d, g0):
        g0 = np.ascontiguousarray(g0, np.float64)
        self.lib.ref_aa_x_init(self.h, m, d, _dp(g0))

    def replace(self, g):
        g = np.ascontiguousarray(g, np.float64)
        self.lib.ref_aa_x_replace(self.h, _dp(g), g.size)

    def compute(self, g):
        g = np.ascontiguousarray(g, np.float64)
        out = np.zeros_like(g)
        self.lib.ref_aa_x_compute(self.h, _dp(out), _dp(g), g.size)
        return out


def ref_cod_solve(M, rhs):
    lib = _load("libref_aa.so")
    lib.ref_cod_solve.argtypes = [C.c_int, c_dp, c_dp, c_dp]
    M = np.asfortranarray(M, np.float64)
    rhs = np.ascontiguousarray(rhs, np.float64)
    out = np.zeros_like(rhs)
    lib.ref_cod_solve(M.shape[0], M.ctypes.data_as(c_dp), _dp(rhs), _dp(out))
    return out


def ref_cod_rank(M):
    lib = _load("libref_aa.so")
    lib.ref_cod_rank.argtypes = [C.c_int, c_dp]
    M = np.asfortranarray(M, np.float64)
    return lib.ref_cod_rank(M.shape[0], M.ctypes.data_as(c_dp))


# ---------------------------------------------------------------------------------------------
# The C restatement (oracle/port/aaadmm_port.c -> oracle/liboracle_port.so)
# ---------------------------------------------------------------------------------------------
PORT_LIB = os.path.join(HERE, "liboracle_port.so")
_port = None


def port_lib():
    global _port
    if _port is None:
        P = C.CDLL(PORT_LIB)
        P.port_tet_prox.argtypes = [c_dp, C.c_int]
        P.port_tet_fmuvt.argtypes = [c_dp, c_dp, C.c_int]
        P.port_cod_solve.argtypes = [C.c_int, c_dp, c_dp, c_dp]
        P.port_aa_new.restype = C.c_void_p
        P.port_aa_new.argtypes = [C.c_int, C.c_int, C.c_int]
        P.port_aa_free.argtypes = [C.c_void_p]
        for n in ("init", "reset", "replace"):
            getattr(P, "port_aa_" + n).argtypes = [C.c_void_p, c_dp]
        P.port_aa_compute.argtypes = [C.c_void_p, c_dp, c_dp]
        P.port_scene_new.restype = C.c_void_p
        P.port_scene_new.argtypes = [C.c_int, c_dp, C.c_int, c_ip, c_dp, C.c_double, C.c_double, C.c_int, c_ip,
                                     C.c_double, C.c_double, C.c_int]
        P.port_scene_free.argtypes = [C.c_void_p]
        P.port_scene_nfree.argtypes = [C.c_void_p]
        P.port_scene_step.argtypes = [C.c_void_p, c_dp, c_dp, c_dp, C.c_int, C.c_int, C.c_int, C.c_double, c_dp, c_dp, c_ip]
        _port = P
    return _port


def port_tet_prox(z):
    out = np.ascontiguousarray(z, np.float64).copy()
    port_lib().port_tet_prox(_dp(out), out.shape[0])
    return out


def port_tet_F_minus_UVt(z):
    z = np.ascontiguousarray(z, np.float64)
    out = np.zeros_like(z)
    port_lib().port_tet_fmuvt(_dp(z), _dp(out), z.shape[0])
    return out


def port_cod_solve(M, rhs):
    M = np.asfortranarray(M, np.float64)
    rhs = np.ascontiguousarray(rhs, np.float64)
    out = np.zeros_like(rhs)
    rank = port_lib().port_cod_solve(M.shape[0], M.ctypes.data_as(c_dp), _dp(rhs), _dp(out))
    return out, rank


class PortAnderson:
    def __init__(self, m, total_dim, effective_dim=None):
        self.P = port_lib()
        self.n = total_dim
        self.h = C.c_void_p(self.P.port_aa_new(m, total_dim, total_dim if effective_dim is None else effective_dim))

    def __del__(self):
        try:
            self.P.port_aa_free(self.h)
        except Exception:
            pass

    def init(self, u):
        u = np.ascontiguousarray(u, np.float64)
        self.P.port_aa_init(self.h, _dp(u))

    def reset(self, u):
        u = np.ascontiguousarray(u, np.float64)
        self.P.port_aa_reset(self.h, _dp(u))

    def replace(self, u):
        u = np.ascontiguousarray(u, np.float64)
        self.P.port_aa_replace(self.h, _dp(u))

    def compute(self, g):
        g = np.ascontiguousarray(g, np.float64)
        out = np.zeros_like(g)
        self.P.port_aa_compute(self.h, _dp(g), _dp(out))
        return out


class PortSolver:
    """The C restatement of admm::Solver (LINEAR tets), ordering 'hard' or 'xzu'."""

    def __init__(self, variant="hard"):
        self.P = port_lib()
        self.ordering = 0 if variant == "hard" else 1
        self.h = None
        self.verts = None

    def __del__(self):
        try:
            if self.h:
                self.P.port_scene_free(self.h)
        except Exception:
            pass

    def add_tetmesh(self, verts, tets, masses, youngs=1e7, poisson=0.399, material=0):
        assert material == 0 and self.verts is None
        self.verts = np.ascontiguousarray(verts, np.float32).astype(np.float64)
        self.tets = np.ascontiguousarray(tets, np.int32)
        self.masses = np.ascontiguousarray(masses, np.float32).astype(np.float64)
        self.youngs, self.poisson = youngs, poisson
        self._x = self.verts.reshape(-1).copy()
        self._v = np.zeros_like(self._x)

    def set_pins(self, idx, pts):
        self.pin_idx = np.ascontiguousarray(idx, np.int32)
        order = np.argsort(self.pin_idx, kind="stable")
        # the reference pairs the k-th pinned vertex (ascending id) with the k-th point handed over
        self.pin_pts = np.ascontiguousarray(pts, np.float64).copy()
        self._pin_sorted = self.pin_idx[order]

    def initialize(self, dt=1.0 / 30.0, iters=100, gravity=-9.8, anderson_m=5, accel=True, penalty=1.0):
        self.dt, self.iters, self.gravity, self.m, self.accel = dt, iters, gravity, anderson_m, accel
        h = self.P.port_scene_new(len(self.verts), _dp(self.verts), len(self.tets), _ip(self.tets), _dp(self.masses),
                                  self.youngs, self.poisson, len(self.pin_idx), _ip(self.pin_idx), dt, penalty, self.ordering)
        if not h:
            raise RuntimeError("port: inverted rest tet")
        self.h = C.c_void_p(h)
        self._v[:] = 0.0

    def step(self):
        n = max(1, self.iters)
        prim, comb, rej = np.zeros(n), np.zeros(n), np.zeros(n, np.int32)
        rows = self.P.port_scene_step(self.h, _dp(self._x), _dp(self._v), _dp(self.pin_pts), self.iters, self.m,
                                      int(bool(self.accel)), self.gravity, _dp(prim), _dp(comb), _ip(rej))
        return np.stack([np.zeros(rows), prim[:rows], comb[:rows], rej[:rows].astype(np.float64)], axis=1)

    def x(self):
        return self._x.copy()

    def v(self):
        return self._v.copy()


# ---------------------------------------------------------------------------------------------
# C restatement of the Geometry path and of the triangle / collision terms (oracle/port/aaadmm_port_geo.c)
# ---------------------------------------------------------------------------------------------
class PortGeometrySolver:
    """Plain-C restatement of ALMGeometrySolver<3> (use_alm=True) / GeometrySolver<3>; same building interface as
    RefGeometrySolver; dense global solve (small scenes only)."""

    def __init__(self, use_alm=True):
        self.P = port_lib()
        P = self.P
        vp = C.c_void_p
        P.port_geo_new.restype = vp
        P.port_geo_new.argtypes = [C.c_int]
        P.port_geo_free.argtypes = [vp]
        P.port_geo_add_plane.argtypes = [vp, c_ip, C.c_int, C.c_double]
        P.port_geo_add_edge.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double]
        P.port_geo_add_angle.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        P.port_geo_add_ref_surface.argtypes = [vp, C.c_int, C.c_double, c_dp, C.c_int, c_ip, C.c_int]
        P.port_geo_add_uniform_laplacian.argtypes = [vp, c_ip, C.c_int, C.c_double, c_dp]
        P.port_geo_add_closeness.argtypes = [vp, C.c_int, C.c_double, c_dp]
        P.port_geo_setup.argtypes = [vp, C.c_int, C.c_double]
        P.port_geo_solve.argtypes = [vp, c_dp, C.c_int, C.c_int, c_dp, c_dp]
        self.h = vp(P.port_geo_new(int(use_alm)))
        self.n_points = 0

    def __del__(self):
        try:
            self.P.port_geo_free(self.h)
        except Exception:
            pass

    def add_plane(self, idx, weight=1.0):
        idx = np.ascontiguousarray(idx, np.int32)
        self.P.port_geo_add_plane(self.h, _ip(idx), len(idx), weight)

    def add_edge(self, i0, i1, weight, length):
        self.P.port_geo_add_edge(self.h, int(i0), int(i1), weight, length)

    def add_angle(self, tip, s1, s2, weight, amin, amax):
        self.P.port_geo_add_angle(self.h, int(tip), int(s1), int(s2), weight, amin, amax)

    def add_ref_surface(self, n_points, weight, V, F):
        V = np.ascontiguousarray(V, np.float64)
        F = np.ascontiguousarray(F, np.int32)
        self.P.port_geo_add_ref_surface(self.h, n_points, weight, _dp(V), len(V), _ip(F), len(F))

    def add_relative_uniform_laplacian(self, idx, weight, ref_pts):
        idx = np.ascontiguousarray(idx, np.int32)
        ref_pts = np.ascontiguousarray(ref_pts, np.float64)
        self.P.port_geo_add_uniform_laplacian(self.h, _ip(idx), len(idx), weight, _dp(ref_pts))

    def add_uniform_laplacian(self, idx, weight):
        idx = np.ascontiguousarray(idx, np.int32)
        self.P.port_geo_add_uniform_laplacian(self.h, _ip(idx), len(idx), weight, None)

    def add_closeness(self, idx, weight, target):
        target = np.ascontiguousarray(target, np.float64)
        self.P.port_geo_add_closeness(self.h, int(idx), weight, _dp(target))

    def setup(self, n_points, rho):
        self.n_points = n_points
        if self.P.port_geo_setup(self.h, n_points, rho) != 0:
            raise RuntimeError("port geometry: system matrix not positive definite")

    def solve(self, init_x, max_iter, anderson_m):
        x0 = np.ascontiguousarray(init_x, np.float64)
        x = np.zeros((self.n_points, 3))
        hist = np.zeros(max(1, max_iter))
        n = self.P.port_geo_solve(self.h, _dp(x0), max_iter, anderson_m, _dp(x), _dp(hist))
        return hist[:n], x


def port_geo_project(kind, cols, params=(0.0, 0.0, 0.0, 0.0)):
    """cols: (n, k, 3) transformed columns; returns the projections (n, k, 3)."""
    P = port_lib()
    P.port_geo_project.argtypes = [C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp]
    cols = np.ascontiguousarray(cols, np.float64)
    n, k = cols.shape[0], cols.shape[1]
    prm = np.ascontiguousarray(np.broadcast_to(np.asarray(params, np.float64), (n, 4)))
    out = np.zeros_like(cols)
    P.port_geo_project(kind, n, k, _dp(cols), _dp(prm), _dp(out))
    return out


def port_geo_closest_points(V, F, Q):
    P = port_lib()
    P.port_geo_closest_points.argtypes = [c_dp, C.c_int, c_ip, C.c_int, c_dp, C.c_int, c_dp]
    V = np.ascontiguousarray(V, np.float64)
    F = np.ascontiguousarray(F, np.int32)
    Q = np.ascontiguousarray(Q, np.float64)
    out = np.zeros_like(Q)
    P.port_geo_closest_points(_dp(V), len(V), _ip(F), len(F), _dp(Q), len(Q), _dp(out))
    return out


def port_tet_prox_hyper(material, mu, lam, vol, F):
    """C restatement of HyperElasticTet::prox on (n, 9) column-major blocks: returns (prox, gradient of the inputs)."""
    P = port_lib()
    P.port_tet_prox_hyper.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, c_dp, c_dp, C.c_int]
    P.port_tet_prox_hyper.restype = None
    z = np.ascontiguousarray(F, np.float64).reshape(-1, 9).copy()
    g = np.zeros_like(z)
    P.port_tet_prox_hyper(int(material), mu, lam, vol, _dp(z), _dp(g), len(z))
    return z, g


def port_tri_prox(F, variant="hard", limit_min=-100.0, limit_max=100.0):
    P = port_lib()
    P.port_tri_prox.argtypes = [C.c_int, c_dp, C.c_int, C.c_double, C.c_double]
    P.port_tri_prox.restype = None
    z = np.ascontiguousarray(F, np.float64).reshape(-1, 6).copy()
    P.port_tri_prox(0 if variant == "xzu" else 1, _dp(z), len(z), limit_min, limit_max)
    return z


def port_collision_prox(types, prm, pts):
    P = port_lib()
    P.port_collision_prox.argtypes = [C.c_int, c_ip, c_dp, c_dp, C.c_int]
    P.port_collision_prox.restype = None
    types = np.ascontiguousarray(types, np.int32)
    prm = np.ascontiguousarray(prm, np.float64).reshape(-1, 7)
    z = np.ascontiguousarray(pts, np.float64).reshape(-1, 3).copy()
    P.port_collision_prox(len(types), _ip(types), _dp(prm), _dp(z), len(z))
    return z


# ---------------------------------------------------------------------------------------------
# Reference Geometry solver (oracle/_ref/libref_geo.so)
# ---------------------------------------------------------------------------------------------
def have_ref_geo():
    return os.path.exists(os.path.join(REF_DIR, "libref_geo.so"))


class RefGeometrySolver:
    """The reference ALMGeometrySolver<3> (use_alm=True) or GeometrySolver<3> with its own Constraint
    classes, built from plain arrays."""

    def __init__(self, use_alm=True):
        self.L = _load("libref_geo.so")
        L = self.L
        vp = C.c_void_p
        L.ref_geo_new.restype = vp
        L.ref_geo_new.argtypes = [C.c_int]
        L.ref_geo_free.argtypes = [vp]
        L.ref_geo_add_plane.argtypes = [vp, c_ip, C.c_int, C.c_double, C.c_int]
        L.ref_geo_add_edge.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
        L.ref_geo_add_angle.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int]
        L.ref_geo_add_ref_surface.argtypes = [vp, C.c_int, C.c_double, c_dp, C.c_int, c_ip, C.c_int, C.c_int]
        L.ref_geo_add_relative_uniform_laplacian.argtypes = [vp, c_ip, C.c_int, C.c_double, c_dp, C.c_int]
        L.ref_geo_add_uniform_laplacian.argtypes = [vp, c_ip, C.c_int, C.c_double]
        L.ref_geo_add_closeness.argtypes = [vp, C.c_int, C.c_double, c_dp]
        L.ref_geo_setup.argtypes = [vp, C.c_int, C.c_double]
        L.ref_geo_solve.argtypes = [vp, c_dp, C.c_int, C.c_int, C.c_int]
        L.ref_geo_history.argtypes = [vp, c_dp, c_dp]
        L.ref_geo_solution.argtypes = [vp, c_dp, C.c_int]
        self.h = vp(L.ref_geo_new(int(use_alm)))
        self.n_points = 0

    def __del__(self):
        try:
            self.L.ref_geo_free(self.h)
        except Exception:
            pass

    def add_plane(self, idx, weight=1.0):
        idx = np.ascontiguousarray(idx, np.int32)
        self.L.ref_geo_add_plane(self.h, _ip(idx), len(idx), weight, 0)

    def add_edge(self, i0, i1, weight, length):
        self.L.ref_geo_add_edge(self.h, int(i0), int(i1), weight, length, 0)

    def add_angle(self, tip, s1, s2, weight, amin, amax):
        self.L.ref_geo_add_angle(self.h, int(tip), int(s1), int(s2), weight, amin, amax, 0)

    def add_ref_surface(self, n_points, weight, V, F):
        V = np.ascontiguousarray(V, np.float64)
        F = np.ascontiguousarray(F, np.int32)
        self.L.ref_geo_add_ref_surface(self.h, n_points, weight, _dp(V), len(V), _ip(F), len(F), 1)

    def add_relative_uniform_laplacian(self, idx, weight, ref_pts):
        idx = np.ascontiguousarray(idx, np.int32)
        ref_pts = np.ascontiguousarray(ref_pts, np.float64)
        self.L.ref_geo_add_relative_uniform_laplacian(self.h, _ip(idx), len(idx), weight, _dp(ref_pts), len(ref_pts))

    def add_uniform_laplacian(self, idx, weight):
        idx = np.ascontiguousarray(idx, np.int32)
        self.L.ref_geo_add_uniform_laplacian(self.h, _ip(idx), len(idx), weight)

    def add_closeness(self, idx, weight, target):
        target = np.ascontiguousarray(target, np.float64)
        self.L.ref_geo_add_closeness(self.h, int(idx), weight, _dp(target))

    def setup(self, n_points, rho):
        self.n_points = n_points
        if self.L.ref_geo_setup(self.h, n_points, rho) != 0:
            raise RuntimeError("reference setup_ADMM failed")

    def solve(self, init_x, max_iter, anderson_m):
        x0 = np.ascontiguousarray(init_x, np.float64)
        n = self.L.ref_geo_solve(self.h, _dp(x0), self.n_points, max_iter, anderson_m)
        hist, secs = np.zeros(max(n, 1)), np.zeros(max(n, 1))
        self.L.ref_geo_history(self.h, _dp(hist), _dp(secs))
        x = np.zeros((self.n_points, 3))
        self.L.ref_geo_solution(self.h, _dp(x), self.n_points)
        self.secs = secs[:n]
        return hist[:n], x


def ref_geo_project(kind, cols, a=0.0, b=0.0):
    """cols: (n, kc, 3) transformed columns; kind 0 plane, 1 edge (a = length), 2 angle (a, b = min, max)."""
    L = _load("libref_geo.so")
    L.ref_geo_project.argtypes = [C.c_int, C.c_int, c_dp, c_dp, C.c_double, C.c_double]
    cols = np.ascontiguousarray(cols, np.float64)
    out = np.zeros_like(cols)
    for i in range(cols.shape[0]):
        ci = np.ascontiguousarray(cols[i])
        oi = np.zeros_like(ci)
        L.ref_geo_project(kind, cols.shape[1], _dp(ci), _dp(oi), a, b)
        out[i] = oi
    return out


def ref_geo_closest_points(V, F, Q):
    L = _load("libref_geo.so")
    L.ref_geo_closest_points.argtypes = [c_dp, C.c_int, c_ip, C.c_int, c_dp, C.c_int, c_dp, c_ip, c_dp]
    V = np.ascontiguousarray(V, np.float64)
    F = np.ascontiguousarray(F, np.int32)
    Q = np.ascontiguousarray(Q, np.float64)
    Cp, I, d = np.zeros_like(Q), np.zeros(len(Q), np.int32), np.zeros(len(Q))
    L.ref_geo_closest_points(_dp(V), len(V), _ip(F), len(F), _dp(Q), len(Q), _dp(Cp), _ip(I), _dp(d))
    return Cp, I, d
