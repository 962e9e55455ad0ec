"""ctypes wrapper of oracle/_ref/libm3d_ref_{seq,omp}.so -- the REFERENCE'S OWN ransac.h /
iterative_plane_segmentation.cpp / logging.cpp compiled from /root/reference against the
Eigen/Open3D stand-ins of oracle/shim/ (recipe: oracle/Makefile, target `_ref`).

TEST INFRASTRUCTURE ONLY: used by tests/ (to pin the restated oracle against the reference's own
code), tools/make_golden.py and bench.py's CPU legs.  /root/reference exists only in the build
container; on the GPU box the prebuilt .so files travel with the snapshot (git-ignored, not
gpurun-ignored) and `available()` says whether they are there.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "_ref")
REFERENCE_ROOT = "/root/reference"

PLANE, SPHERE, CYLINDER = 0, 1, 2
NPARAM = {PLANE: 4, SPHERE: 4, CYLINDER: 7}
_libs = {}


def _path(omp):
    return os.path.join(_DIR, "libm3d_ref_omp.so" if omp else "libm3d_ref_seq.so")


def build():
    """compiles the reference sources where they lie (needs /root/reference); no-op otherwise"""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "include", "misc3d")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_ref"])
    return available()


def available(omp=False):
    return os.path.exists(_path(omp))


def lib(omp=False):
    if omp not in _libs:
        if not available(omp):
            raise RuntimeError(f"{_path(omp)} is missing: run `make -C oracle _ref` where /root/reference exists")
        L = C.CDLL(_path(omp))
        L.m3dref_last_error.restype = C.c_char_p
        L.m3dref_segment_plane_iterative.restype = C.c_long
        _libs[omp] = L
    return _libs[omp]


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def ransac_fit(kind, xyz, nrm=None, thr=0.01, max_it=1000, prob=0.9999, seed=1, omp=False):
    """RANSAC<...>::FitModel of the reference.  Returns (ret, model, inliers, stats) where stats has
    `iterations_run` (the `count` of ransac.h:616-619) and `fitness` (fraction)."""
    xyz, nrm = _f64(xyz), _f64(nrm)
    n = len(xyz)
    model = np.zeros(8)
    inl = np.empty(max(n, 1), dtype=np.uint64)
    n_inl = C.c_size_t(0)
    fit = C.c_double(0)
    its = C.c_ulonglong(0)
    L = lib(omp)
    rc = L.m3dref_ransac_fit(int(kind), _p(xyz), _p(nrm), C.c_size_t(n), C.c_double(thr), C.c_size_t(max_it),
                             C.c_double(prob), C.c_uint32(seed & 0xFFFFFFFF), _p(model), _p(inl, C.c_size_t),
                             C.byref(n_inl), C.byref(fit), C.byref(its))
    if rc < 0:
        raise RuntimeError(L.m3dref_last_error().decode(errors="replace"))
    return rc, model[:NPARAM[kind]].copy(), inl[:n_inl.value].copy(), {
        "iterations_run": int(its.value), "fitness": fit.value / 100.0}


def minimal_fit(kind, pts, nrm=None):
    pts, nrm = _f64(pts), _f64(nrm)
    m = np.zeros(8)
    L = lib()
    rc = L.m3dref_minimal_fit(int(kind), _p(pts), _p(nrm), int(len(pts)), _p(m))
    if rc < 0:
        raise RuntimeError(L.m3dref_last_error().decode(errors="replace"))
    return bool(rc), m[:NPARAM[kind]].copy()


def distances(kind, model, xyz):
    xyz = _f64(xyz).reshape(-1, 3)
    m = np.zeros(8)
    m[:len(model)] = model
    out = np.empty(len(xyz))
    lib().m3dref_distances(int(kind), _p(m), _p(xyz), C.c_size_t(len(xyz)), _p(out))
    return out


def general_fit(kind, xyz, model0=None):
    xyz = _f64(xyz)
    m = np.zeros(8)
    if model0 is not None:
        m[:len(model0)] = model0
    rc = lib().m3dref_general_fit(int(kind), _p(xyz), C.c_size_t(len(xyz)), _p(m))
    return bool(rc), m[:NPARAM[kind]].copy()


def sample_table(seed, n, k, rows):
    out = np.empty((rows, k), dtype=np.uint64)
    lib().m3dref_sample_table(C.c_uint32(seed & 0xFFFFFFFF), C.c_size_t(n), int(k), C.c_size_t(rows),
                              _p(out, C.c_size_t))
    return out


def segment_plane_iterative(xyz, thr, max_it=100, min_ratio=0.05, seed=1, omp=False, cap=256):
    xyz = _f64(xyz)
    n = len(xyz)
    planes = np.zeros((cap, 4))
    labels = np.empty(max(n, 1), dtype=np.uint64)
    L = lib(omp)
    npl = L.m3dref_segment_plane_iterative(_p(xyz), C.c_size_t(n), C.c_double(thr), int(max_it),
                                           C.c_double(min_ratio), C.c_uint32(seed & 0xFFFFFFFF), _p(planes),
                                           C.c_size_t(cap), _p(labels, C.c_size_t))
    if npl == -1:
        raise RuntimeError(L.m3dref_last_error().decode(errors="replace"))
    if npl < 0:
        raise RuntimeError(f"m3dref_segment_plane_iterative failed ({npl})")
    return int(npl), planes[:npl].copy(), labels[:n].copy()


FLANN, ANNOY = 0, 1


def match_correspondence(src, dst, method=FLANN, n_trees=4, omp=False):
    """ANNMatcher(method, n_trees).Match(src, dst) of the reference; src/dst are (dim, n) arrays.
    FLANN: exact (Open3D's kd-tree is stood in by an exact brute-force search).  ANNOY: the reference's
    real vendored Annoy forest -- approximate, and its 4-thread build is racy (not reproducible)."""
    src = np.asarray(src, dtype=np.float64)
    dst = np.asarray(dst, dtype=np.float64)
    dim, ns = src.shape
    nd = dst.shape[1]
    s = np.ascontiguousarray(src.T)
    d = np.ascontiguousarray(dst.T)
    i0 = np.empty(max(ns, 1), dtype=np.uint64)
    i1 = np.empty(max(ns, 1), dtype=np.uint64)
    L = lib(omp)
    L.m3dref_match_correspondence.restype = C.c_long
    n = L.m3dref_match_correspondence(_p(s), C.c_size_t(ns), _p(d), C.c_size_t(nd), int(dim), int(method),
                                      int(n_trees), _p(i0, C.c_size_t), _p(i1, C.c_size_t))
    if n < 0:
        raise RuntimeError(L.m3dref_last_error().decode(errors="replace"))
    return i0[:n].copy(), i1[:n].copy()


def omp_threads():
    return int(lib(True).m3dref_openmp())


def use_all_cores():
    """OpenMP threads := the cores this process may run on (launchers like torchrun export OMP_NUM_THREADS=1)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib(True).m3dref_set_threads(int(n))
    return omp_threads()
