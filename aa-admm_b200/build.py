"""Build recipe: nvcc -> libaaadmm_b200.so (C ABI + sm_100a kernels), g++ -> libaaadmm_host.so
(host mirror classes). In-tree outputs (git-ignored, shipped to the GPU box by gpurun)."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIB_CUDA = os.path.join(PKG, "libaaadmm_b200.so")
LIB_HOST = os.path.join(PKG, "libaaadmm_host.so")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

CUDA_SOURCES = ["aaadmm_capi.cu", "ldlt_apply.cu", "ldlt_factor.cu", "tet_kernels.cu", "tri_kernels.cu", "geo_kernels.cu", "extra_terms.cu"]
# per-element local-step math follows the reference operation by operation: no FMA contraction
CUDA_NO_FMAD = {"tet_kernels.cu", "tri_kernels.cu", "geo_kernels.cu", "extra_terms.cu"}
CUDA_HEADERS = ["common.cuh", "svd3.cuh", "cod_small.cuh", "aa_kernels.cuh", "tet_kernels.cuh", "ldlt_apply.cuh", "geo_kernels.cuh", "extra_terms.cuh", "pipe.cuh", "tri_prox.cuh", "lbfgs_prox.cuh", "collision_prox.cuh", "ldlt_factor.cuh"]
HOST_SOURCES = ["beam_scene.cpp", "sparse_ldlt.cpp", "tet_system.cpp", "Solver.cpp", "MeshIO.cpp", "GeometrySolver.cpp", "GeometryApps.cpp", "host_capi.cpp"]
HOST_HEADERS = ["beam_scene.hpp", "sparse_ldlt.hpp", "tet_system.hpp", "Solver.hpp", "AndersonAcceleration.hpp", "GeometrySolver.hpp", "GeometryApps.hpp", "MeshIO.hpp"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout


def build_cuda(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in CUDA_HEADERS] + [os.path.join(ROOT, "include", "aaadmm.h")]
    if not force and not _stale(LIB_CUDA, deps):
        return LIB_CUDA
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    base = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
            "-ccbin", GXX, "-Xcompiler", "-fPIC", "-Xcompiler", "-fopenmp"]
    if verbose:
        base += ["-Xptxas", "-v"]
    objs, procs = [], []
    for s in CUDA_SOURCES:
        o = os.path.join(objdir, s.replace(".cu", ".o"))
        cmd = base + (["-fmad=false"] if s in CUDA_NO_FMAD else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("build failed: " + " ".join(cmd))
        if verbose:
            print(out)
    _run(["nvcc", "-ccbin", GXX, "-shared", "-Xcompiler", "-fopenmp", "-o", LIB_CUDA] + objs)
    return LIB_CUDA


def build_host(force=False):
    srcs = [os.path.join(HOST, s) for s in HOST_SOURCES]
    deps = srcs + [os.path.join(HOST, h) for h in HOST_HEADERS] + [
        os.path.join(ROOT, "include", "aaadmm.h"), os.path.join(ROOT, "include", "aaadmm_host.h"), LIB_CUDA]
    if not force and not _stale(LIB_HOST, deps):
        return LIB_HOST
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for s in HOST_SOURCES:
        # scene/operator setup restates float32/float64 arithmetic of the reference operation by
        # operation: no FMA contraction there. The dense factorisation kernels may contract.
        contract = "-ffp-contract=fast" if s == "sparse_ldlt.cpp" else "-ffp-contract=off"
        o = os.path.join(objdir, s.replace(".cpp", ".o"))
        _run([GXX, "-std=c++17", "-O3", "-march=x86-64-v3", contract, "-fopenmp", "-fPIC", "-c",
              os.path.join(HOST, s), "-o", o])
        objs.append(o)
    _run([GXX, "-shared", "-fopenmp", "-o", LIB_HOST] + objs + ["-L" + PKG, "-laaadmm_b200", "-Wl,-rpath,$ORIGIN"])
    return LIB_HOST


def build_oracle():
    _run(["make", "-C", os.path.join(ROOT, "oracle"), "all"])


def build_all(force=False):
    build_cuda(force)
    build_host(force)
    build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built", LIB_CUDA, LIB_HOST)
