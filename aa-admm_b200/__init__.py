"""aa-admm_b200: ctypes plumbing over the two native libraries of this repo.

  libaaadmm_b200.so  the C ABI of include/aaadmm.h: sm_100a CUDA kernels for the AA-ADMM hot path
  libaaadmm_host.so  the host-side C++ mirror of the reference's scene/solver classes

Nothing here computes: every call lands in native code, and the compute entry points fail
loudly when no CUDA device is present (there is no CPU path in the product).
Import name: `aa_admm_b200` (the directory name has a dash; aa_admm_b200/__init__.py aliases it).
"""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_CUDA = os.path.join(PKG, "libaaadmm_b200.so")
LIB_HOST = os.path.join(PKG, "libaaadmm_host.so")

c_dp = C.POINTER(C.c_double)
c_fp = C.POINTER(C.c_float)
c_ip = C.POINTER(C.c_int)
c_lp = C.POINTER(C.c_int64)

ORDER_HARD_ZXU = 0
ORDER_XZU = 1
NPROF = 8
PROF_NAMES = ["update_z", "rhs_gather", "ldlt_apply", "update_u_resid", "aa_pass1", "aa_pass2", "safeguard", "total"]


class AaadmmError(RuntimeError):
    pass


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _fp(a):
    return a.ctypes.data_as(c_fp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


def _lp(a):
    return a.ctypes.data_as(c_lp)


class StepOpts(C.Structure):
    _fields_ = [("ordering", C.c_int), ("admm_iters", C.c_int), ("anderson_m", C.c_int), ("accel", C.c_int),
                ("eps", C.c_double), ("log_comb_xzu", C.c_int)]


class StepResult(C.Structure):
    _fields_ = [("iters_logged", C.c_int), ("rejects", C.c_int), ("broke_early", C.c_int), ("loop_ms", C.c_float),
                ("step_ms", C.c_float), ("kernel_launches", C.c_int)]


_cuda = None
_host = None


def build(force=False):
    import importlib.util
    spec = importlib.util.spec_from_file_location("aaadmm_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build_cuda(force)
    mod.build_host(force)
    return mod


def cuda_lib():
    """The C ABI library. Raises if it has not been built (no fallback)."""
    global _cuda
    if _cuda is None:
        if not os.path.exists(LIB_CUDA):
            raise AaadmmError("libaaadmm_b200.so is missing: run __graft_entry__.build() (no CPU fallback exists)")
        L = C.CDLL(LIB_CUDA, mode=C.RTLD_GLOBAL)
        L.aaadmm_last_error.restype = C.c_char_p
        vp = C.c_void_p
        L.aaadmm_aa_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int64, C.c_int64]
        L.aaadmm_aa_destroy.argtypes = [vp]
        for n in ("init", "reset", "replace"):
            getattr(L, "aaadmm_aa_" + n).argtypes = [vp, c_dp, C.c_int64]
        L.aaadmm_aa_compute.argtypes = [vp, c_dp, c_dp, C.c_int64]
        L.aaadmm_aa_state.argtypes = [vp, c_ip, c_ip]
        L.aaadmm_ldlt_create.argtypes = [C.POINTER(vp), C.c_int, c_lp, c_ip, c_dp, c_dp, c_ip, C.c_int]
        L.aaadmm_ldlt_create_from_matrix.argtypes = [C.POINTER(vp), C.c_int, c_lp, c_ip, c_dp, c_lp, c_ip, c_ip, C.c_int]
        L.aaadmm_ldlt_refactor.argtypes = [vp, c_dp]
        L.aaadmm_ldlt_destroy.argtypes = [vp]
        L.aaadmm_ldlt_solve.argtypes = [vp, c_dp, c_dp]
        L.aaadmm_ldlt_stats.argtypes = [vp, c_dp]
        L.aaadmm_tetscene_step_resident.argtypes = [vp, C.POINTER(StepOpts), C.POINTER(StepResult)]
        L.aaadmm_tetscene_profile.argtypes = [vp, C.POINTER(StepOpts), C.c_int, c_fp]
        L.aaadmm_tetscene_algo_bytes.argtypes = [vp, C.c_int, c_dp]
        L.aaadmm_tetscene_read_zu.argtypes = [vp, c_dp, c_dp]
        L.aaadmm_tet_prox_linear.argtypes = [c_dp, C.c_int64]
        L.aaadmm_tet_f_minus_uvt.argtypes = [c_dp, c_dp, C.c_int64]
        L.aaadmm_cod_solve.argtypes = [C.c_int, c_dp, c_dp, c_dp, c_ip]
        _cuda = L
    return _cuda


def host_lib():
    global _host
    if _host is None:
        cuda_lib()
        if not os.path.exists(LIB_HOST):
            raise AaadmmError("libaaadmm_host.so is missing: run __graft_entry__.build()")
        H = C.CDLL(LIB_HOST)
        vp = C.c_void_p
        H.aaadmm_host_last_error.restype = C.c_char_p
        H.aaadmm_host_beam_new.restype = vp
        H.aaadmm_host_beam_free.argtypes = [vp]
        H.aaadmm_host_beam_add.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
        H.aaadmm_host_beam_counts.argtypes = [vp, c_ip, c_ip, c_ip]
        H.aaadmm_host_beam_copy.argtypes = [vp, c_fp, c_ip, c_fp, c_ip, c_dp, c_ip]
        H.aaadmm_host_beam_stretch.argtypes = [vp, C.c_double]
        H.aaadmm_host_factor_new.restype = vp
        H.aaadmm_host_factor_new.argtypes = [C.c_int, c_lp, c_ip, c_dp, c_dp, C.c_int, C.c_int]
        H.aaadmm_host_factor_free.argtypes = [vp]
        H.aaadmm_host_factor_nnz.restype = C.c_int64
        H.aaadmm_host_factor_nnz.argtypes = [vp]
        H.aaadmm_host_factor_copy.argtypes = [vp, c_lp, c_ip, c_dp, c_dp, c_ip]
        H.aaadmm_host_factor_solve.argtypes = [vp, c_dp, c_dp, C.c_int]
        H.aaadmm_host_factor_stats.argtypes = [vp, c_dp]
        H.aaadmm_host_solver_new.restype = vp
        H.aaadmm_host_solver_free.argtypes = [vp]
        H.aaadmm_host_solver_add_tetmesh.argtypes = [vp, c_fp, C.c_int, c_ip, C.c_int, c_fp, C.c_double, C.c_double, C.c_int]
        H.aaadmm_host_solver_add_trimesh.argtypes = [vp, c_fp, C.c_int, c_ip, C.c_int, c_fp, C.c_double, C.c_double,
                                                     C.c_double, C.c_double]
        H.aaadmm_host_solver_add_wind.argtypes = [vp, c_ip, C.c_int, c_dp]
        H.aaadmm_host_solver_set_collisions.argtypes = [vp, c_ip, C.c_int]
        H.aaadmm_host_solver_add_obstacle.argtypes = [vp, C.c_int, c_dp]
        H.aaadmm_host_wind_project.argtypes = [c_ip, C.c_int, c_dp, C.c_double, c_dp, c_dp, C.c_int]
        H.aaadmm_host_solver_set_pins.argtypes = [vp, c_ip, c_dp, C.c_int]
        H.aaadmm_host_solver_initialize.argtypes = [vp, C.c_double, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double,
                                                    C.c_int, C.c_int]
        H.aaadmm_host_solver_set_factor.argtypes = [vp, C.c_int, c_lp, c_ip, c_dp, c_dp, c_ip]
        H.aaadmm_host_solver_set_material.argtypes = [vp, C.c_double, C.c_double]
        H.aaadmm_host_solver_set_x.argtypes = [vp, c_dp]
        H.aaadmm_host_solver_was_incremental.argtypes = [vp]
        H.aaadmm_host_solver_step.argtypes = [vp]
        H.aaadmm_host_solver_set_iters.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        H.aaadmm_host_solver_n_dof.argtypes = [vp]
        H.aaadmm_host_solver_get_x.argtypes = [vp, c_dp]
        H.aaadmm_host_solver_get_v.argtypes = [vp, c_dp]
        H.aaadmm_host_solver_hist_rows.argtypes = [vp]
        H.aaadmm_host_solver_hist_copy.argtypes = [vp, c_dp, c_dp, c_ip]
        H.aaadmm_host_solver_info.argtypes = [vp, c_dp]
        H.aaadmm_host_solver_factor_info.argtypes = [vp, c_dp]
        H.aaadmm_host_solver_device_scene.restype = vp
        H.aaadmm_host_solver_device_scene.argtypes = [vp]
        H.aaadmm_host_solver_device_factor.restype = vp
        H.aaadmm_host_solver_device_factor.argtypes = [vp]
        H.aaadmm_host_tet_constants.argtypes = [c_dp, C.c_double, C.c_double, c_dp, c_dp, c_dp]
        H.aaadmm_host_tri_constants.argtypes = [c_dp, C.c_double, C.c_double, c_dp, c_dp, c_dp]
        H.aaadmm_host_mesh_load.restype = vp
        H.aaadmm_host_mesh_load.argtypes = [C.c_char_p, C.c_int]
        H.aaadmm_host_mesh_free.argtypes = [vp]
        H.aaadmm_host_mesh_counts.argtypes = [vp, c_ip, c_ip]
        H.aaadmm_host_mesh_copy.argtypes = [vp, c_fp, c_ip, c_fp]
        H.aaadmm_host_mesh_save.argtypes = [vp, C.c_char_p]
        H.aaadmm_host_system_new.restype = vp
        H.aaadmm_host_system_new.argtypes = [c_fp, C.c_int, c_ip, C.c_int, c_ip, C.c_int, c_fp, C.c_double, C.c_double,
                                             c_ip, C.c_int, C.c_double, c_ip, C.c_int]
        H.aaadmm_host_system_free.argtypes = [vp]
        H.aaadmm_host_system_counts.argtypes = [vp, c_ip, C.POINTER(C.c_int64)]
        H.aaadmm_host_system_copy.argtypes = [vp, C.POINTER(C.c_int64), c_ip, c_dp, c_ip]
        _host = H
    return _host


def _ck(rc):
    if rc != 0:
        raise AaadmmError(cuda_lib().aaadmm_last_error().decode() or "aaadmm call failed (%d)" % rc)


def _hk(rc):
    if rc != 0:
        msg = host_lib().aaadmm_host_last_error().decode()
        raise AaadmmError(msg or cuda_lib().aaadmm_last_error().decode() or "host call failed (%d)" % rc)


def device_count():
    return cuda_lib().aaadmm_device_count()


def set_device(i):
    _ck(cuda_lib().aaadmm_set_device(int(i)))


# ---------------------------------------------------------------------------------------------
class AndersonAcceleration:
    """Device Anderson accelerator with the reference's variant-H interface
    (hard/src/AndersonAcceleration.h): ctor(m,total_dim,effective_dim), init/reset/replace/compute.
    Variant X (xzu) is the same object with effective_dim == total_dim."""

    def __init__(self, m, total_dim, effective_dim=None):
        self.L = cuda_lib()
        self.n = int(total_dim)
        self.h = C.c_void_p()
        _ck(self.L.aaadmm_aa_create(C.byref(self.h), int(m), self.n,
                                    int(total_dim if effective_dim is None else effective_dim)))

    def close(self):
        if self.h:
            self.L.aaadmm_aa_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _vec(self, u):
        return np.ascontiguousarray(u, np.float64).reshape(-1)

    def init(self, u):
        u = self._vec(u)
        _ck(self.L.aaadmm_aa_init(self.h, _dp(u), u.size))

    def reset(self, u):
        u = self._vec(u)
        _ck(self.L.aaadmm_aa_reset(self.h, _dp(u), u.size))

    def replace(self, u):
        u = self._vec(u)
        _ck(self.L.aaadmm_aa_replace(self.h, _dp(u), u.size))

    def compute(self, g):
        g = self._vec(g)
        out = np.empty_like(g)
        _ck(self.L.aaadmm_aa_compute(self.h, _dp(g), _dp(out), g.size))
        return out

    def state(self):
        it, col = C.c_int(), C.c_int()
        _ck(self.L.aaadmm_aa_state(self.h, C.byref(it), C.byref(col)))
        return it.value, col.value


class Ldlt:
    """Device-resident LDL^T factor (strictly-lower CSC L, D, perm[new]=old)."""

    def __init__(self, n, Lp, Li, Lx, D, perm, nrhs=3):
        self.L = cuda_lib()
        self.n, self.nrhs = int(n), int(nrhs)
        Lp = np.ascontiguousarray(Lp, np.int64)
        Li = np.ascontiguousarray(Li, np.int32)
        Lx = np.ascontiguousarray(Lx, np.float64)
        D = np.ascontiguousarray(D, np.float64)
        perm = np.ascontiguousarray(perm, np.int32)
        self.h = C.c_void_p()
        _ck(self.L.aaadmm_ldlt_create(C.byref(self.h), self.n, _lp(Lp), _ip(Li), _dp(Lx), _dp(D), _ip(perm), self.nrhs))

    @classmethod
    def from_matrix(cls, n, Ap, Ai, Ax, Lp, Li, perm, nrhs=3):
        """Numeric factorisation ON THE DEVICE: A lower CSC (incl. diagonal), pattern of L (Lp, Li) and the ordering
        from a symbolic analysis; aaadmm_ldlt_create_from_matrix."""
        self = cls.__new__(cls)
        self.L = cuda_lib()
        self.n, self.nrhs = int(n), int(nrhs)
        Ap = np.ascontiguousarray(Ap, np.int64)
        Ai = np.ascontiguousarray(Ai, np.int32)
        Ax = np.ascontiguousarray(Ax, np.float64)
        Lp = np.ascontiguousarray(Lp, np.int64)
        Li = np.ascontiguousarray(Li, np.int32)
        perm = np.ascontiguousarray(perm, np.int32)
        self.nnz_a = int(Ap[n])
        self.h = C.c_void_p()
        _ck(self.L.aaadmm_ldlt_create_from_matrix(C.byref(self.h), self.n, _lp(Ap), _ip(Ai), _dp(Ax), _lp(Lp), _ip(Li),
                                                  _ip(perm), self.nrhs))
        return self

    def refactor(self, Ax):
        """Other values, same pattern: aaadmm_ldlt_refactor (no allocation, no analysis)."""
        Ax = np.ascontiguousarray(Ax, np.float64)
        assert Ax.size == self.nnz_a
        _ck(self.L.aaadmm_ldlt_refactor(self.h, _dp(Ax)))

    def close(self):
        if self.h:
            self.L.aaadmm_ldlt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, b):
        b = np.ascontiguousarray(b, np.float64).reshape(-1)
        assert b.size == self.n * self.nrhs
        x = np.empty_like(b)
        _ck(self.L.aaadmm_ldlt_solve(self.h, _dp(b), _dp(x)))
        return x

    def stats(self):
        s = np.zeros(8)
        _ck(self.L.aaadmm_ldlt_stats(self.h, _dp(s)))
        return dict(n=int(s[0]), blocks=int(s[1]), levels=int(s[2]), max_block=int(s[3]), nnz_L=int(s[4]),
                    nnz_offblock=int(s[5]), dense_diag_entries=int(s[6]), bytes_per_solve=float(s[7]))


def tet_prox_linear(z):
    z = np.ascontiguousarray(z, np.float64).copy()
    _ck(cuda_lib().aaadmm_tet_prox_linear(_dp(z), z.shape[0]))
    return z


def tet_f_minus_uvt(z):
    z = np.ascontiguousarray(z, np.float64)
    out = np.empty_like(z)
    _ck(cuda_lib().aaadmm_tet_f_minus_uvt(_dp(z), _dp(out), z.shape[0]))
    return out


def tet_prox_hyper(material, mu, lam, vol, z):
    """Returns (prox(z), vol * dPsi/dF(z)) for NeoHookean (1) / StVK (2) blocks."""
    L = cuda_lib()
    L.aaadmm_tet_prox_hyper.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, c_dp, c_dp, C.c_int64]
    z = np.ascontiguousarray(z, np.float64).copy()
    g = np.zeros_like(z)
    _ck(L.aaadmm_tet_prox_hyper(material, mu, lam, vol, _dp(z), _dp(g), z.shape[0]))
    return z, g


def cod_solve(M, rhs):
    M = np.asfortranarray(M, np.float64)
    rhs = np.ascontiguousarray(rhs, np.float64)
    x = np.zeros_like(rhs)
    rank = C.c_int()
    _ck(cuda_lib().aaadmm_cod_solve(M.shape[0], M.ctypes.data_as(c_dp), _dp(rhs), _dp(x), C.byref(rank)))
    return x, rank.value


# ---------------------------------------------------------------------------------------------
class HostFactor:
    """Nested-dissection + multifrontal LDL^T computed on the host (setup)."""

    def __init__(self, n, Ap, Ai, Ax, coords=None, leaf_size=96, n_threads=0):
        self.H = host_lib()
        Ap = np.ascontiguousarray(Ap, np.int64)
        Ai = np.ascontiguousarray(Ai, np.int32)
        Ax = np.ascontiguousarray(Ax, np.float64)
        cp = None
        if coords is not None:
            coords = np.ascontiguousarray(coords, np.float64)
            cp = _dp(coords)
        self.n = int(n)
        self.h = self.H.aaadmm_host_factor_new(self.n, _lp(Ap), _ip(Ai), _dp(Ax), cp, leaf_size, n_threads)
        if not self.h:
            raise AaadmmError(self.H.aaadmm_host_last_error().decode())
        self.h = C.c_void_p(self.h)

    def __del__(self):
        try:
            if self.h:
                self.H.aaadmm_host_factor_free(self.h)
                self.h = None
        except Exception:
            pass

    def arrays(self):
        nnz = self.H.aaadmm_host_factor_nnz(self.h)
        Lp = np.zeros(self.n + 1, np.int64)
        Li = np.zeros(nnz, np.int32)
        Lx = np.zeros(nnz)
        D = np.zeros(self.n)
        perm = np.zeros(self.n, np.int32)
        self.H.aaadmm_host_factor_copy(self.h, _lp(Lp), _ip(Li), _dp(Lx), _dp(D), _ip(perm))
        return Lp, Li, Lx, D, perm

    def solve(self, b, nrhs=1):
        b = np.ascontiguousarray(b, np.float64).reshape(-1)
        x = np.zeros_like(b)
        self.H.aaadmm_host_factor_solve(self.h, _dp(b), _dp(x), nrhs)
        return x

    def stats(self):
        s = np.zeros(6)
        self.H.aaadmm_host_factor_stats(self.h, _dp(s))
        return dict(supernodes=int(s[0]), flops=s[1], seconds_symbolic=s[3], seconds_numeric=s[4], nnz_L=int(s[5]))


class BeamScene:
    """Beams of the reference's samples (make_tet_blocks + centre/scale + pins), O(n)."""

    def __init__(self):
        self.H = host_lib()
        self.h = C.c_void_p(self.H.aaadmm_host_beam_new())

    def __del__(self):
        try:
            if self.h:
                self.H.aaadmm_host_beam_free(self.h)
                self.h = None
        except Exception:
            pass

    def add(self, cx, cy, cz, y_shift=0.0, density=1522.0):
        _hk(0 if self.H.aaadmm_host_beam_add(self.h, cx, cy, cz, y_shift, density) >= 0 else -1)
        return self

    def counts(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.H.aaadmm_host_beam_counts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def arrays(self):
        nv, nt, npin = self.counts()
        verts = np.zeros((nv, 3), np.float32)
        tets = np.zeros((nt, 4), np.int32)
        masses = np.zeros(nv, np.float32)
        pidx = np.zeros(npin, np.int32)
        ppts = np.zeros((npin, 3))
        pside = np.zeros(npin, np.int32)
        self.H.aaadmm_host_beam_copy(self.h, _fp(verts), _ip(tets), _fp(masses), _ip(pidx), _dp(ppts), _ip(pside))
        return verts, tets, masses, pidx, ppts, pside

    def stretch(self, dt):
        """stretch_beams(): returns the new pin targets."""
        self.H.aaadmm_host_beam_stretch(self.h, dt)
        npin = self.counts()[2]
        ppts = np.zeros((npin, 3))
        self.H.aaadmm_host_beam_copy(self.h, None, None, None, None, _dp(ppts), None)  # only the pin targets move
        return ppts


class Solver:
    """admm::Solver mirror (host C++ class over the C ABI)."""

    def __init__(self):
        self.H = host_lib()
        self.h = C.c_void_p(self.H.aaadmm_host_solver_new())

    def close(self):
        if self.h:
            self.H.aaadmm_host_solver_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_tetmesh(self, verts, tets, masses, youngs=1e7, poisson=0.399, material=0):
        verts = np.ascontiguousarray(verts, np.float32)
        tets = np.ascontiguousarray(tets, np.int32)
        masses = np.ascontiguousarray(masses, np.float32)
        r = self.H.aaadmm_host_solver_add_tetmesh(self.h, _fp(verts), len(verts), _ip(tets), len(tets), _fp(masses),
                                                  youngs, poisson, material)
        if r < 0:
            _hk(r)
        return r

    def add_trimesh(self, verts, tris, masses, youngs=1e7, poisson=0.399, limit_min=-100.0, limit_max=100.0):
        """binding::add_trimesh: nodes + TriEnergyTerm per triangle (hard_zxu ordering only)."""
        verts = np.ascontiguousarray(verts, np.float32)
        tris = np.ascontiguousarray(tris, np.int32)
        masses = np.ascontiguousarray(masses, np.float32)
        r = self.H.aaadmm_host_solver_add_trimesh(self.h, _fp(verts), len(verts), _ip(tris), len(tris), _fp(masses),
                                                  youngs, poisson, limit_min, limit_max)
        if r < 0:
            _hk(r)
        return r

    def set_collisions(self, idx):
        """Solver::set_collisions (in place): one Collision energy term per listed vertex at initialize()."""
        idx = np.ascontiguousarray(idx, np.int32)
        _hk(self.H.aaadmm_host_solver_set_collisions(self.h, _ip(idx), len(idx)))

    def add_obstacle(self, kind, prm7):
        """Solver::add_obstacle with an analytic obstacle: kind = PASSIVE_* tag, prm7 = {cx,cy,cz,nx,ny,nz,radius}."""
        p = np.ascontiguousarray(prm7, np.float64)
        assert p.size == 7
        _hk(self.H.aaadmm_host_solver_add_obstacle(self.h, int(kind), _dp(p)))

    def add_wind(self, tris, direction):
        """WindForce over `tris` (global vertex ids), appended to Solver::ext_forces."""
        tris = np.ascontiguousarray(tris, np.int32)
        d = np.ascontiguousarray(direction, np.float64)
        _hk(self.H.aaadmm_host_solver_add_wind(self.h, _ip(tris), len(tris), _dp(d)))

    def set_pins(self, idx, pts):
        idx = np.ascontiguousarray(idx, np.int32)
        pts = np.ascontiguousarray(pts, np.float64)
        _hk(self.H.aaadmm_host_solver_set_pins(self.h, _ip(idx), _dp(pts), len(idx)))

    def initialize(self, dt=1.0 / 30.0, iters=100, gravity=-9.8, anderson_m=5, accel=True, penalty=1.0,
                   ordering=ORDER_HARD_ZXU, nd_leaf=0):
        _hk(self.H.aaadmm_host_solver_initialize(self.h, dt, iters, gravity, anderson_m, int(bool(accel)), penalty,
                                                 ordering, nd_leaf))

    def set_external_factor(self, n, Lp, Li, Lx, D, perm):
        """Solver::set_external_factor: the next initialize() uses this (L, D, perm) instead of factoring itself;
        n = free vertices (factor of Ahat) or 3 x that (factor of the full system, e.g. Eigen's own)."""
        Lp = np.ascontiguousarray(Lp, np.int64)
        Li = np.ascontiguousarray(Li, np.int32)
        Lx = np.ascontiguousarray(Lx, np.float64)
        D = np.ascontiguousarray(D, np.float64)
        perm = np.ascontiguousarray(perm, np.int32)
        _hk(self.H.aaadmm_host_solver_set_factor(self.h, int(n), _lp(Lp), _ip(Li), _dp(Lx), _dp(D), _ip(perm)))

    def set_material(self, youngs, poisson):
        """Solver::set_material: every energy term gets these Lame parameters; the next initialize() on the unchanged
        scene is incremental (numeric part only)."""
        _hk(self.H.aaadmm_host_solver_set_material(self.h, float(youngs), float(poisson)))

    def set_x(self, x):
        x = np.ascontiguousarray(x, np.float64).reshape(-1)
        assert x.size == self.H.aaadmm_host_solver_n_dof(self.h)
        self.H.aaadmm_host_solver_set_x(self.h, _dp(x))

    def was_incremental(self):
        return bool(self.H.aaadmm_host_solver_was_incremental(self.h))

    def set_iters(self, iters, anderson_m, accel):
        self.H.aaadmm_host_solver_set_iters(self.h, iters, anderson_m, int(bool(accel)))

    def step(self):
        """One Solver::step(); returns rows (prim_residual, comb_residual, is_reject)."""
        _hk(self.H.aaadmm_host_solver_step(self.h))
        n = self.H.aaadmm_host_solver_hist_rows(self.h)
        prim, comb, rej = np.zeros(max(n, 1)), np.zeros(max(n, 1)), np.zeros(max(n, 1), np.int32)
        if n:
            self.H.aaadmm_host_solver_hist_copy(self.h, _dp(prim), _dp(comb), _ip(rej))
        return np.stack([prim[:n], comb[:n], rej[:n].astype(np.float64)], axis=1)

    def x(self):
        out = np.zeros(self.H.aaadmm_host_solver_n_dof(self.h))
        self.H.aaadmm_host_solver_get_x(self.h, _dp(out))
        return out

    def v(self):
        out = np.zeros(self.H.aaadmm_host_solver_n_dof(self.h))
        self.H.aaadmm_host_solver_get_v(self.h, _dp(out))
        return out

    def info(self):
        s = np.zeros(8)
        self.H.aaadmm_host_solver_info(self.h, _dp(s))
        return dict(loop_ms=s[0], step_ms=s[1], kernel_launches=int(s[2]), init_ms=s[3], iter_num=int(s[4]),
                    rejects=int(s[5]), n_free=int(s[6]), n_tets=int(s[7]))

    def factor_info(self):
        s = np.zeros(6)
        self.H.aaadmm_host_solver_factor_info(self.h, _dp(s))
        return dict(nnz_L=int(s[0]), supernodes=int(s[1]), flops=s[2], seconds_symbolic=s[3], seconds_numeric=s[4],
                    nnz_A_lower=int(s[5]))

    def ldlt_stats(self):
        L = cuda_lib()
        f = C.c_void_p(self.H.aaadmm_host_solver_device_factor(self.h))
        s = np.zeros(8)
        _ck(L.aaadmm_ldlt_stats(f, _dp(s)))
        return dict(n=int(s[0]), blocks=int(s[1]), levels=int(s[2]), max_block=int(s[3]), nnz_L=int(s[4]),
                    nnz_offblock=int(s[5]), dense_diag_entries=int(s[6]), bytes_per_solve=float(s[7]))

    def solve(self, b):
        """x = A^-1 b with the device factor of this solver (b: 3 per free vertex, the reference's dof order):
        LDLTSolver::solve (LinearSolver.hpp:87-90) on an arbitrary right-hand side."""
        L = cuda_lib()
        f = C.c_void_p(self.H.aaadmm_host_solver_device_factor(self.h))
        b = np.ascontiguousarray(b, np.float64)
        assert b.size == 3 * self.info()["n_free"]
        x = np.zeros_like(b)
        _ck(L.aaadmm_ldlt_solve(f, _dp(b), _dp(x)))
        return x

    # --- measurement helpers over the device scene of this solver ---
    def _scene(self):
        return C.c_void_p(self.H.aaadmm_host_solver_device_scene(self.h))

    def step_resident(self, iters, anderson_m=5, accel=True, ordering=ORDER_HARD_ZXU, eps=1e-20):
        o = StepOpts(ordering, iters, anderson_m, int(bool(accel)), eps, 1)
        r = StepResult()
        _ck(cuda_lib().aaadmm_tetscene_step_resident(self._scene(), C.byref(o), C.byref(r)))
        return dict(iters_logged=r.iters_logged, rejects=r.rejects, broke_early=r.broke_early, loop_ms=r.loop_ms,
                    step_ms=r.step_ms, kernel_launches=r.kernel_launches)

    def profile(self, iters, anderson_m=5, accel=True, ordering=ORDER_HARD_ZXU):
        o = StepOpts(ordering, iters, anderson_m, int(bool(accel)), 1e-20, 1)
        ms = np.zeros(NPROF, np.float32)
        _ck(cuda_lib().aaadmm_tetscene_profile(self._scene(), C.byref(o), iters, _fp(ms)))
        by = np.zeros(NPROF)
        _ck(cuda_lib().aaadmm_tetscene_algo_bytes(self._scene(), anderson_m, _dp(by)))
        return {n: dict(ms=float(ms[i]), bytes=float(by[i])) for i, n in enumerate(PROF_NAMES)}


def load_mesh(path, kind="elenode", save_as=None):
    """mcl::meshio::load_elenode (`path` without extension) / load_obj through the host mirror. Returns float32
    vertices, int32 elements (4 or 3 per row) and the lumped float32 masses of binding::add_tetmesh / add_trimesh.
    save_as: also write the mesh back with save_elenode / save_obj."""
    H = host_lib()
    k = {"elenode": 0, "obj": 1}[kind]
    h = H.aaadmm_host_mesh_load(str(path).encode(), k)
    if not h:
        raise AaadmmError(H.aaadmm_host_last_error().decode())
    try:
        nv, ne = C.c_int(0), C.c_int(0)
        H.aaadmm_host_mesh_counts(h, C.byref(nv), C.byref(ne))
        verts = np.zeros((nv.value, 3), np.float32)
        elems = np.zeros((ne.value, 4 if k == 0 else 3), np.int32)
        masses = np.zeros(nv.value, np.float32)
        _hk(H.aaadmm_host_mesh_copy(h, _fp(verts), _ip(elems), _fp(masses)))
        if save_as is not None:
            _hk(H.aaadmm_host_mesh_save(h, str(save_as).encode()))
    finally:
        H.aaadmm_host_mesh_free(h)
    return verts, elems, masses


def host_system_matrix(verts, tets, tris, masses, pins, rho_dt2, youngs=1e7, poisson=0.399, collisions=()):
    """Host setup only: scalar system matrix Ahat (A = M + rho dt^2 D^T W^2 D = Ahat (x) I3) of a scene of tets and
    triangles with the pinned vertices eliminated. Returns (dense lower-filled symmetric n_free x n_free array,
    dev_to_vert); for tests on small scenes."""
    H = host_lib()
    verts = np.ascontiguousarray(verts, np.float32)
    tets = np.ascontiguousarray(tets, np.int32).reshape(-1, 4)
    tris = np.ascontiguousarray(tris, np.int32).reshape(-1, 3)
    masses = np.ascontiguousarray(masses, np.float32)
    pins = np.ascontiguousarray(pins, np.int32)
    col = np.ascontiguousarray(collisions, np.int32)
    h = H.aaadmm_host_system_new(_fp(verts), len(verts), _ip(tets), len(tets), _ip(tris), len(tris), _fp(masses),
                                 youngs, poisson, _ip(pins), len(pins), rho_dt2, _ip(col), len(col))
    if not h:
        raise AaadmmError(H.aaadmm_host_last_error().decode())
    try:
        nf, nnz = C.c_int(0), C.c_int64(0)
        H.aaadmm_host_system_counts(h, C.byref(nf), C.byref(nnz))
        Ap = np.zeros(nf.value + 1, np.int64)
        Ai = np.zeros(nnz.value, np.int32)
        Ax = np.zeros(nnz.value)
        d2v = np.zeros(len(verts), np.int32)
        H.aaadmm_host_system_copy(h, _lp(Ap), _ip(Ai), _dp(Ax), _ip(d2v))
    finally:
        H.aaadmm_host_system_free(h)
    n = nf.value
    M = np.zeros((n, n))
    for j in range(n):
        for p in range(Ap[j], Ap[j + 1]):
            M[Ai[p], j] = Ax[p]
            M[j, Ai[p]] = Ax[p]
    return M, d2v


def wind_project(tris, direction, dt, x, v):
    """Host WindForce::project (explicit force of the windyflag scene); returns the new velocities."""
    H = host_lib()
    tris = np.ascontiguousarray(tris, np.int32)
    d = np.ascontiguousarray(direction, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    v = np.array(v, np.float64).copy()
    _hk(H.aaadmm_host_wind_project(_ip(tris), len(tris), _dp(d), dt, _dp(x), _dp(v), x.size // 3))
    return v


def tri_prox(F, variant="hard", limit_min=-100.0, limit_max=100.0):
    """TriEnergyTerm::prox on (n, 6) column-major 3x2 blocks; variant 'hard' (hard_zxu) or 'xzu'."""
    L = cuda_lib()
    L.aaadmm_tri_prox.argtypes = [C.c_int, c_dp, C.c_int64, C.c_double, C.c_double]
    z = np.ascontiguousarray(F, np.float64).reshape(-1, 6).copy()
    _ck(L.aaadmm_tri_prox(ORDER_XZU if variant == "xzu" else ORDER_HARD_ZXU, _dp(z), len(z), limit_min, limit_max))
    return z


PASSIVE_TYPES = {"floor": 0, "slide_floor": 1, "sphere": 2, "plane_half_sphere": 3, "cylinder": 4}


def collision_prox(objects, pts):
    """Collision::prox of (n, 3) points against analytic passive objects: a list of
    (type name, 7 parameters {cx, cy, cz, nx, ny, nz, radius}); Floor uses cx as its height."""
    L = cuda_lib()
    L.aaadmm_collision_prox.argtypes = [C.c_int, c_ip, c_dp, c_dp, C.c_int64]
    types = np.ascontiguousarray([PASSIVE_TYPES[o[0]] for o in objects], np.int32)
    prm = np.ascontiguousarray([o[1] for o in objects], np.float64).reshape(-1, 7)
    z = np.ascontiguousarray(pts, np.float64).reshape(-1, 3).copy()
    _ck(L.aaadmm_collision_prox(len(types), _ip(types), _dp(prm), _dp(z), len(z)))
    return z


def spring_prox(pts, pins, active):
    L = cuda_lib()
    L.aaadmm_spring_prox.argtypes = [c_dp, c_dp, c_ip, C.c_int64]
    z = np.ascontiguousarray(pts, np.float64).reshape(-1, 3).copy()
    pins = np.ascontiguousarray(pins, np.float64).reshape(-1, 3)
    act = np.ascontiguousarray(active, np.int32)
    _ck(L.aaadmm_spring_prox(_dp(z), _dp(pins), _ip(act), len(z)))
    return z


def make_beam_solver(cx, cy, cz, n_beams=1, dt=1.0 / 30.0, iters=100, anderson_m=5, accel=True, penalty=1.0,
                     ordering=ORDER_HARD_ZXU, youngs=1e7, poisson=0.399, gravity=-9.8):
    """The beams.cpp scene with LINEAR tets: returns (solver, scene). Pins are stretched once
    before initialize, as beams.cpp:126 does."""
    scene = BeamScene()
    shifts = {1: [0.0], 3: [1.75, 0.0, -1.75]}.get(n_beams, [1.75 * (n_beams // 2 - i) for i in range(n_beams)])
    for s in shifts:
        scene.add(cx, cy, cz, s)
    verts, tets, masses, pidx, ppts, pside = scene.arrays()
    solver = Solver()
    solver.add_tetmesh(verts, tets, masses, youngs, poisson, 0)
    solver.set_pins(pidx, scene.stretch(dt))
    solver.initialize(dt, iters, gravity, anderson_m, accel, penalty, ordering)
    return solver, scene


# ---------------------------------------------------------------------------------------------
# Geometry (ALMGeometrySolver<3> mirror)
# ---------------------------------------------------------------------------------------------
def _geo_host():
    H = host_lib()
    if not hasattr(H, "_geo_ready"):
        vp = C.c_void_p
        H.aaadmm_host_geo_new.restype = vp
        H.aaadmm_host_geo_new_variant.restype = vp
        H.aaadmm_host_geo_new_variant.argtypes = [C.c_int]
        H.aaadmm_host_geo_free.argtypes = [vp]
        H.aaadmm_host_geo_add_plane.argtypes = [vp, c_ip, C.c_int, C.c_double]
        H.aaadmm_host_geo_add_edge.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double]
        H.aaadmm_host_geo_add_angle.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        H.aaadmm_host_geo_add_ref_surface.argtypes = [vp, C.c_int, C.c_double, c_dp, C.c_int, c_ip, C.c_int]
        H.aaadmm_host_geo_add_relative_uniform_laplacian.argtypes = [vp, c_ip, C.c_int, C.c_double, c_dp, C.c_int]
        H.aaadmm_host_geo_add_uniform_laplacian.argtypes = [vp, c_ip, C.c_int, C.c_double]
        H.aaadmm_host_geo_add_closeness.argtypes = [vp, C.c_int, C.c_double, c_dp]
        H.aaadmm_host_geo_setup.argtypes = [vp, C.c_int, C.c_double]
        H.aaadmm_host_geo_solve.argtypes = [vp, c_dp, C.c_int, C.c_int, C.c_int]
        H.aaadmm_host_geo_history.argtypes = [vp, c_dp]
        H.aaadmm_host_geo_solution.argtypes = [vp, c_dp, C.c_int]
        H.aaadmm_host_geo_info.argtypes = [vp, c_dp]
        H._geo_ready = True
    return H


class GeometrySolver:
    """ALMGeometrySolver<3> mirror: hard plane / edge / angle constraints, soft closest-point-to-
    reference-surface constraint, Laplacian / closeness regularisation.
    variant="gs" selects the older GeometrySolver<3> loop (Geometry/GeometrySolver.h:156-263)."""

    def __init__(self, variant="alm"):
        self.H = _geo_host()
        if variant not in ("alm", "gs"):
            raise ValueError("variant must be 'alm' or 'gs'")
        self.variant = variant
        self.h = C.c_void_p(self.H.aaadmm_host_geo_new_variant(1 if variant == "gs" else 0))
        self.n_points = 0

    def __del__(self):
        try:
            if self.h:
                self.H.aaadmm_host_geo_free(self.h)
                self.h = None
        except Exception:
            pass

    def add_plane(self, idx, weight=1.0):
        idx = np.ascontiguousarray(idx, np.int32)
        self.H.aaadmm_host_geo_add_plane(self.h, _ip(idx), len(idx), weight)

    def add_edge(self, i0, i1, weight, length):
        self.H.aaadmm_host_geo_add_edge(self.h, int(i0), int(i1), weight, length)

    def add_angle(self, tip, s1, s2, weight, amin, amax):
        self.H.aaadmm_host_geo_add_angle(self.h, int(tip), int(s1), int(s2), weight, amin, amax)

    def add_ref_surface(self, n_points, weight, V, F):
        V = np.ascontiguousarray(V, np.float64)
        F = np.ascontiguousarray(F, np.int32)
        self.H.aaadmm_host_geo_add_ref_surface(self.h, n_points, weight, _dp(V), len(V), _ip(F), len(F))

    def add_relative_uniform_laplacian(self, idx, weight, ref_pts):
        idx = np.ascontiguousarray(idx, np.int32)
        ref_pts = np.ascontiguousarray(ref_pts, np.float64)
        self.H.aaadmm_host_geo_add_relative_uniform_laplacian(self.h, _ip(idx), len(idx), weight, _dp(ref_pts), len(ref_pts))

    def add_uniform_laplacian(self, idx, weight):
        idx = np.ascontiguousarray(idx, np.int32)
        self.H.aaadmm_host_geo_add_uniform_laplacian(self.h, _ip(idx), len(idx), weight)

    def add_closeness(self, idx, weight, target):
        target = np.ascontiguousarray(target, np.float64)
        self.H.aaadmm_host_geo_add_closeness(self.h, int(idx), weight, _dp(target))

    def setup(self, n_points, rho):
        self.n_points = n_points
        _hk(self.H.aaadmm_host_geo_setup(self.h, n_points, rho))

    def solve(self, init_x, max_iter, anderson_m):
        """init_x: (n_points, 3). Returns (combined-residual history of the accepted iterations, solution)."""
        x0 = np.ascontiguousarray(init_x, np.float64)
        n = self.H.aaadmm_host_geo_solve(self.h, _dp(x0), self.n_points, max_iter, anderson_m)
        if n < 0:
            _hk(n)
        hist = np.zeros(max(n, 1))
        self.H.aaadmm_host_geo_history(self.h, _dp(hist))
        x = np.zeros((self.n_points, 3))
        self.H.aaadmm_host_geo_solution(self.h, _dp(x), self.n_points)
        return hist[:n], x

    def elapsed(self, n):
        """elapsed_time_ of the last solve's n logged iterations (cumulative seconds, device time stamps)."""
        t = np.zeros(max(n, 1))
        self.H.aaadmm_host_geo_elapsed(self.h, _dp(t))
        return t[:n]

    def info(self):
        s = np.zeros(4)
        self.H.aaadmm_host_geo_info(self.h, _dp(s))
        return dict(loop_ms=s[0], kernel_launches=int(s[1]), rejects=int(s[2]), iters=int(s[3]))


def geo_project(kind, cols, params=(0.0, 0.0, 0.0, 0.0)):
    """cols: (n, k, 3) already transformed columns of n constraints of one kind (0 plane, 1 edge, 2 angle)."""
    L = cuda_lib()
    L.aaadmm_geo_project.argtypes = [C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp]
    cols = np.ascontiguousarray(cols, np.float64)
    n, kc = cols.shape[0], cols.shape[1]
    k = kc if kind == 0 else kc + 1
    prm = np.ascontiguousarray(params, np.float64)
    out = np.zeros_like(cols)
    _ck(L.aaadmm_geo_project(kind, n, k, _dp(cols), _dp(prm), _dp(out)))
    return out


def geo_closest_points(V, F, Q):
    L = cuda_lib()
    L.aaadmm_geo_closest_points.argtypes = [c_dp, C.c_int, c_ip, C.c_int, c_dp, C.c_int, c_dp, c_ip]
    V = np.ascontiguousarray(V, np.float64)
    F = np.ascontiguousarray(F, np.int32)
    Q = np.ascontiguousarray(Q, np.float64)
    Cp = np.zeros_like(Q)
    tri = np.zeros(len(Q), np.int32)
    _ck(L.aaadmm_geo_closest_points(_dp(V), len(V), _ip(F), len(F), _dp(Q), len(Q), _dp(Cp), _ip(tri)))
    return Cp, tri


# ---------------------------------------------------------------------------------------------
# Geometry front-end (host/GeometryApps.hpp)
# ---------------------------------------------------------------------------------------------
class PolyMesh:
    """Polygon mesh numbered as OpenMesh numbers it (vertices / faces in file order, edges by first appearance)."""

    def __init__(self, handle):
        self.H = host_lib()
        self.h = C.c_void_p(handle)

    @staticmethod
    def _lib():
        H = host_lib()
        if not getattr(H, "_poly_ready", False):
            vp = C.c_void_p
            H.aaadmm_host_polymesh_load.restype = vp
            H.aaadmm_host_polymesh_load.argtypes = [C.c_char_p]
            H.aaadmm_host_polymesh_new.restype = vp
            H.aaadmm_host_polymesh_new.argtypes = [c_dp, C.c_int, c_ip, c_ip, C.c_int]
            H.aaadmm_host_polymesh_free.argtypes = [vp]
            H.aaadmm_host_polymesh_save.argtypes = [vp, C.c_char_p]
            H.aaadmm_host_polymesh_counts.argtypes = [vp, c_ip, c_dp]
            H.aaadmm_host_polymesh_copy.argtypes = [vp, c_dp, c_ip, c_ip, c_ip]
            H.aaadmm_host_polymesh_subdivide_and_smooth.restype = vp
            H.aaadmm_host_polymesh_subdivide_and_smooth.argtypes = [vp]
            H.aaadmm_host_geoapp_optimize.argtypes = [C.c_int, vp, vp, C.c_int, C.c_int, c_dp, c_dp, c_ip, c_dp, c_dp]
            H._poly_ready = True
        return H

    @classmethod
    def load(cls, path):
        H = cls._lib()
        h = H.aaadmm_host_polymesh_load(str(path).encode())
        if not h:
            raise AaadmmError(H.aaadmm_host_last_error().decode())
        return cls(h)

    @classmethod
    def from_arrays(cls, verts, faces):
        """faces: list of vertex-id lists (or an (n, k) array; -1 pads shorter faces)."""
        H = cls._lib()
        verts = np.ascontiguousarray(verts, np.float64).reshape(-1, 3)
        fl = [[int(i) for i in f if i >= 0] for f in faces]
        ptr = np.zeros(len(fl) + 1, np.int32)
        ptr[1:] = np.cumsum([len(f) for f in fl])
        idx = np.array([i for f in fl for i in f], np.int32)
        h = H.aaadmm_host_polymesh_new(_dp(verts), len(verts), _ip(ptr), _ip(idx), len(fl))
        if not h:
            raise AaadmmError(H.aaadmm_host_last_error().decode())
        return cls(h)

    def __del__(self):
        try:
            if self.h:
                self.H.aaadmm_host_polymesh_free(self.h)
                self.h = None
        except Exception:
            pass

    def counts(self):
        c = np.zeros(4, np.int32)
        el = C.c_double(0.0)
        _hk(self.H.aaadmm_host_polymesh_counts(self.h, _ip(c), C.byref(el)))
        return dict(vertices=int(c[0]), faces=int(c[1]), corners=int(c[2]), edges=int(c[3]), average_edge_length=el.value)

    def arrays(self):
        """verts (n, 3), faces (list of lists), edges (ne, 2) = halfedge 0 of every edge."""
        c = self.counts()
        V = np.zeros((c["vertices"], 3))
        ptr = np.zeros(c["faces"] + 1, np.int32)
        idx = np.zeros(max(c["corners"], 1), np.int32)
        E = np.zeros((max(c["edges"], 1), 2), np.int32)
        _hk(self.H.aaadmm_host_polymesh_copy(self.h, _dp(V), _ip(ptr), _ip(idx), _ip(E)))
        faces = [idx[ptr[f]:ptr[f + 1]].tolist() for f in range(c["faces"])]
        return V, faces, E[:c["edges"]]

    def save(self, path):
        _hk(self.H.aaadmm_host_polymesh_save(self.h, str(path).encode()))

    def subdivide_and_smooth(self):
        h = self.H.aaadmm_host_polymesh_subdivide_and_smooth(self.h)
        if not h:
            raise AaadmmError(self.H.aaadmm_host_last_error().decode())
        return PolyMesh(h)


def geoapp_optimize(app, mesh, ref_mesh, max_iter, anderson_m, prm):
    """The reference's optimize_mesh of PlanarityOpt (app 'planarity', prm = penalty, closeness_w, laplacian_w,
    relative_laplacian_w) or WireMeshOpt (app 'wiremesh', prm = penalty, min_angle, max_angle, edge_length, closeness_w,
    laplacian_w) on the device-backed ALMGeometrySolver. Returns (residual history, solution (n, 3), info)."""
    H = PolyMesh._lib()
    prm = np.ascontiguousarray(prm, np.float64)
    hist = np.zeros(max(1, max_iter))
    n = C.c_int(0)
    sol = np.zeros((mesh.counts()["vertices"], 3))
    info = np.zeros(3)
    _hk(H.aaadmm_host_geoapp_optimize(0 if app == "planarity" else 1, mesh.h, ref_mesh.h, max_iter, anderson_m, _dp(prm),
                                      _dp(hist), C.byref(n), _dp(sol), _dp(info)))
    return hist[:n.value], sol, dict(loop_ms=info[0], resets=int(info[1]))


class GeoApp:
    """geoapp::GeoApp: one of the two Geometry applications with the setup kept (constraints, system matrix,
    factorisation, device upload once); solve() runs solve_ADMM from the mesh's own positions."""

    def __init__(self, app, mesh, ref_mesh, prm):
        H = PolyMesh._lib()
        H.aaadmm_host_geoapp_new.restype = C.c_void_p
        H.aaadmm_host_geoapp_new.argtypes = [C.c_int, C.c_void_p, C.c_void_p, c_dp]
        H.aaadmm_host_geoapp_free.argtypes = [C.c_void_p]
        H.aaadmm_host_geoapp_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, c_ip, c_dp, c_dp]
        H.aaadmm_host_geoapp_stats.argtypes = [C.c_void_p, c_dp]
        self.H = H
        prm = np.ascontiguousarray(prm, np.float64)
        self.n_points = mesh.counts()["vertices"]
        h = H.aaadmm_host_geoapp_new(0 if app == "planarity" else 1, mesh.h, ref_mesh.h, _dp(prm))
        if not h:
            raise AaadmmError(H.aaadmm_host_last_error().decode())
        self.h = C.c_void_p(h)

    def __del__(self):
        try:
            if self.h:
                self.H.aaadmm_host_geoapp_free(self.h)
                self.h = None
        except Exception:
            pass

    def solve(self, max_iter, anderson_m, want_solution=True):
        hist = np.zeros(max(1, max_iter))
        n = C.c_int(0)
        sol = np.zeros((self.n_points, 3)) if want_solution else None
        info = np.zeros(4)
        _hk(self.H.aaadmm_host_geoapp_solve(self.h, max_iter, anderson_m, _dp(hist), C.byref(n), _dp(sol) if want_solution else None,
                                            _dp(info)))
        return hist[:n.value], sol, dict(loop_ms=info[0], resets=int(info[1]), kernel_launches=int(info[2]), wall_ms=info[3])

    def stats(self):
        s = np.zeros(8)
        self.H.aaadmm_host_geoapp_stats(self.h, _dp(s))
        return dict(points=int(s[0]), hard_constraints=int(s[1]), z_columns=int(s[2]), soft_constraints=int(s[3]),
                    nnz_L=int(s[4]), fronts=int(s[5]), levels=int(s[6]), bytes_per_apply=float(s[7]))
